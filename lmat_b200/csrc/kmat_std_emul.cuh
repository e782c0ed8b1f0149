// kmat_std_emul.cuh -- single-thread device emulation of the two libstdc++ algorithms whose tie order is
// visible in read_label's output:
//   * std::sort  (bits/stl_algo.h: __introsort_loop with threshold 16, median-of-3 to first, unguarded
//     partition, heapsort fallback at depth 2*floor(log2 n), then __final_insertion_sort) -- used by the
//     reference at read_label.cpp:1074 (CmpDepth1), :893 (TCmp, not a strict weak order), :351/:379 (CmpDepth)
//   * std::priority_queue push/pop (bits/stl_heap.h: __push_heap / __adjust_heap / __pop_heap) -- used by
//     TaxNodeStat's run-time pruning (TaxNodeStat.hpp:151-152,169-172,214-217)
// The oracle (oracle/kmat_oracle.c) restates the same algorithms in C and tests/test_logf_stdsort.py pins
// that restatement against g++'s own std::sort / priority_queue.
#ifndef KMAT_STD_EMUL_CUH
#define KMAT_STD_EMUL_CUH

namespace kmstd {

template <typename T> __device__ __forceinline__ void swap_(T &a, T &b) { T t = a; a = b; b = t; }

template <typename T, typename Less>
__device__ void push_heap_(T *first, int hole, int top, T value, Less less) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
template <typename T, typename Less>
__device__ void adjust_heap_(T *first, int hole, int len, T value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, less);
}
// priority_queue::push: c.push_back(v); push_heap(begin, end)
template <typename T, typename Less>
__device__ void pq_push(T *heap, int &n, T v, Less less) {
    heap[n] = v;
    n++;
    push_heap_(heap, n - 1, 0, v, less);
}
// priority_queue::top + pop: pop_heap(begin, end); c.pop_back()
template <typename T, typename Less>
__device__ T pq_pop(T *heap, int &n, Less less) {
    T top = heap[0];
    if (n > 1) {
        T value = heap[n - 1];
        heap[n - 1] = heap[0];
        adjust_heap_(heap, 0, n - 1, value, less);
    }
    n--;
    return top;
}

template <typename T, typename Less>
__device__ void heap_sort_all_(T *first, int n, Less less) {   // __partial_sort(first, last, last)
    if (n >= 2) {
        int parent = (n - 2) / 2;
        for (;;) {
            T value = first[parent];
            adjust_heap_(first, parent, n, value, less);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = n;
    while (last > 1) {
        --last;
        T value = first[last];
        first[last] = first[0];
        adjust_heap_(first, 0, last, value, less);
    }
}
template <typename T, typename Less>
__device__ void unguarded_linear_insert_(T *last, Less less) {
    T val = *last;
    T *next = last - 1;
    while (less(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
template <typename T, typename Less>
__device__ void insertion_sort_(T *first, T *last, Less less) {
    if (first == last) return;
    for (T *i = first + 1; i != last; ++i) {
        if (less(*i, *first)) {
            T val = *i;
            for (T *j = i; j != first; --j) *j = *(j - 1);
            *first = val;
        } else unguarded_linear_insert_(i, less);
    }
}
template <typename T, typename Less>
__device__ void sort(T *first, int n, Less less) {
    if (n <= 1) return;
    if (n > 16) {
        // __introsort_loop, recursion on the right part replaced by an explicit stack
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        struct Frame { int lo, hi, depth; };
        Frame stack[64];
        int sp = 0;
        stack[sp++] = Frame{0, n, lg * 2};
        while (sp > 0) {
            Frame f = stack[--sp];
            int lo = f.lo, hi = f.hi, depth = f.depth;
            while (hi - lo > 16) {
                if (depth == 0) { heap_sort_all_(first + lo, hi - lo, less); break; }
                --depth;
                T *a = first + lo + 1, *b = first + lo + (hi - lo) / 2, *c = first + hi - 1, *res = first + lo;
                if (less(*a, *b)) {
                    if (less(*b, *c)) swap_(*res, *b);
                    else if (less(*a, *c)) swap_(*res, *c);
                    else swap_(*res, *a);
                } else if (less(*a, *c)) swap_(*res, *a);
                else if (less(*b, *c)) swap_(*res, *c);
                else swap_(*res, *b);
                T *pf = first + lo + 1, *pl = first + hi, *pivot = first + lo;
                for (;;) {
                    while (less(*pf, *pivot)) ++pf;
                    --pl;
                    while (less(*pivot, *pl)) --pl;
                    if (!(pf < pl)) break;
                    swap_(*pf, *pl);
                    ++pf;
                }
                int cut = (int)(pf - first);
                // reference recurses into [cut, hi) first, then loops on [lo, cut); the two ranges are disjoint,
                // so processing order does not change the result
                if (sp < 64) stack[sp++] = Frame{cut, hi, depth};
                hi = cut;
            }
        }
        insertion_sort_(first, first + 16, less);
        for (T *i = first + 16; i != first + n; ++i) unguarded_linear_insert_(i, less);
    } else insertion_sort_(first, first + n, less);
}

}  // namespace kmstd
#endif
