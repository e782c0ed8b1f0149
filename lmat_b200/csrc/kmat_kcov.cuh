// kmat_kcov.cuh -- content_summ's k-mer coverage on the GPU; included at the end of kmat_db.cu.
//
// Reference: src/content_summ.cpp.  For every classified read that passes its filters, retrieve_kmer_labels (:114-155)
// walks the read once per k of the -k list and inserts each DISTINCT canonical k-mer of the read into
// kmer_track[k][taxid][kmer] += 1 (std::map of std::map, one per OpenMP thread); compKmerCov (:527-571) later merges the
// threads and prints, per (taxid, k), the number of distinct k-mers, the sum of the counts and the histogram of the counts.
//
// Here the maps are a flat sorted array of 64-bit keys [k index:3 | group:21 | canonical k-mer:40] with a count each:
//   km_kcov_emit_kernel   warp per read, 32 bases per step (the packed chunk is shared by all k of the list): one
//                         (key, read) pair per valid k-mer window, appended behind a device cursor
//   cub radix sorts       by read, then (stable) by key: equal (key, read) pairs -- a k-mer repeated inside one read --
//                         become neighbours, which restates the per-read std::set no_dups (:131,146-147)
//   km_kcov_flag_kernel   1 for the first pair of every (key, read) run
//   cub ReduceByKey       key -> number of reads that contain it, appended to the accumulated (key, count) runs
// kmat_kcov_finish merges the runs of all batches (sort by key + ReduceByKey) and brings the table to the host, where
// kmat_kcov_query answers per (k, group) by binary search.  "group" is the caller's dense index of the taxid.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#define KC_MAX_K 8
#define KC_GROUP_BITS 21
#define KC_KMER_BITS 40

struct kmat_kcov {
    int device = 0, n_k = 0;
    int k[KC_MAX_K] = {};
    // accumulated (key, count) runs on the device; sorted and unique after compaction
    uint64_t *d_keys = nullptr; uint32_t *d_cnt = nullptr; uint64_t n = 0, cap = 0, n_sorted = 0;
    int runs = 0;
    std::vector<uint64_t> h_keys; std::vector<uint32_t> h_cnt; bool finished = false;
};

struct KcEmitParams {
    const char *bases; const uint64_t *offs; const uint32_t *groups; uint32_t n_reads;
    int n_k; int k[KC_MAX_K];
    uint64_t *keys; uint32_t *reads; unsigned long long *cursor;
};

__global__ void __launch_bounds__(256) km_kcov_emit_kernel(KcEmitParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    for (uint32_t r = warp_global; r < P.n_reads; r += n_warps) {
        const uint32_t grp = P.groups[r];
        if (grp == KMAT_NONE) continue;
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        uint64_t prev = 0; uint32_t pinv = 0xFFFFFFFFu;
        const int nchunks = (len + 31) >> 5;
        for (int c = 0; c < nchunks; c++) {
            const int j = (c << 5) + lane;
            const int code = j < len ? km_code((unsigned char)P.bases[off + j]) : -1;
            const uint32_t cinv = __ballot_sync(KM_FULL, code < 0);
            const uint32_t cc = code < 0 ? 0u : (uint32_t)code;
            const uint32_t hi = __reduce_or_sync(KM_FULL, lane < 16 ? cc << (30 - 2 * lane) : 0u);
            const uint32_t lo = __reduce_or_sync(KM_FULL, lane >= 16 ? cc << (62 - 2 * lane) : 0u);
            const uint64_t cur = ((uint64_t)hi << 32) | lo;
            const int s = 62 - 2 * lane;
            const uint64_t win = (cur >> s) | (s ? (prev << (64 - s)) : 0ull);         // the 32 bases ending at base j
            const uint64_t inv64 = ((uint64_t)cinv << 32) | pinv;
            for (int ki = 0; ki < P.n_k; ki++) {
                const int k = P.k[ki];
                const int wsh = 32 + lane - k + 1;
                const bool ok = ((inv64 >> wsh) & ((1ull << k) - 1)) == 0;              // no invalid base in the window [j-k+1, j]
                const uint32_t okm = __ballot_sync(KM_FULL, ok);
                if (!okm) continue;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.cursor, (unsigned long long)__popc(okm));
                base = ((unsigned long long)__shfl_sync(KM_FULL, (uint32_t)(base >> 32), 0) << 32) | __shfl_sync(KM_FULL, (uint32_t)base, 0);
                if (ok) {
                    const uint64_t fwd = win & ((1ull << (2 * k)) - 1);
                    const uint64_t rc = km_revcomp(fwd, 2 * k);
                    const uint64_t canon = fwd < rc ? fwd : rc;                        // content_summ.cpp:142
                    const unsigned long long at = base + __popc(okm & ((1u << lane) - 1));
                    P.keys[at] = ((uint64_t)ki << (KC_GROUP_BITS + KC_KMER_BITS)) | ((uint64_t)grp << KC_KMER_BITS) | canon;
                    P.reads[at] = r;
                }
            }
            prev = cur; pinv = cinv;
        }
    }
}
__global__ void __launch_bounds__(256) km_kcov_flag_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ reads, uint64_t n, uint32_t *flag) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || keys[i] != keys[i - 1] || reads[i] != reads[i - 1]) ? 1u : 0u;
}

extern "C" int kmat_kcov_create(int device, const int32_t *k_sizes, int n_k, kmat_kcov **out) {
    if (!out || !k_sizes || n_k < 1 || n_k > KC_MAX_K) { kmat_set_error("kmat_kcov_create: bad argument (1..%d k values)", KC_MAX_K); return KMAT_ERR_ARG; }
    for (int i = 0; i < n_k; i++)
        if (k_sizes[i] < 1 || 2 * k_sizes[i] > KC_KMER_BITS) { kmat_set_error("kmat_kcov_create: k = %d outside 1..%d", k_sizes[i], KC_KMER_BITS / 2); return KMAT_ERR_UNSUPPORTED; }
    if (kmat_device_count() <= device) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    kmat_kcov *c = new kmat_kcov();
    c->device = device; c->n_k = n_k;
    for (int i = 0; i < n_k; i++) c->k[i] = k_sizes[i];
    *out = c;
    return KMAT_OK;
}
extern "C" void kmat_kcov_free(kmat_kcov *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_keys); cudaFree(c->d_cnt);
    delete c;
}

// sort (key, count) pairs by key and add the counts of equal keys; in place on the accumulated arrays
static int kc_compact(kmat_kcov *c) {
    if (c->n == 0 || (c->runs <= 1 && c->n_sorted == c->n)) return KMAT_OK;
    uint64_t *k2 = nullptr, *k3 = nullptr; uint32_t *v2 = nullptr, *v3 = nullptr; void *tmp = nullptr; unsigned long long *d_runs = nullptr;
    auto cleanup = [&] { cudaFree(k2); cudaFree(k3); cudaFree(v2); cudaFree(v3); cudaFree(tmp); cudaFree(d_runs); };
    const uint64_t n = c->n;
    if (n >= (1ull << 31)) { kmat_set_error("kmat_kcov: %llu accumulated entries exceed the 2^31 limit of one merge", (unsigned long long)n); return KMAT_ERR_UNSUPPORTED; }
    int rc = KMAT_OK;
    do {
        if (cudaMalloc((void **)&k2, n * 8) != cudaSuccess || cudaMalloc((void **)&v2, n * 4) != cudaSuccess || cudaMalloc((void **)&k3, n * 8) != cudaSuccess ||
            cudaMalloc((void **)&v3, n * 4) != cudaSuccess || cudaMalloc((void **)&d_runs, 8) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; }
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, c->d_keys, k2, c->d_cnt, v2, (int)n);
        cub::DeviceReduce::ReduceByKey(nullptr, b2, k2, k3, v2, v3, d_runs, cub::Sum(), (int)n);
        if (cudaMalloc(&tmp, std::max(b1, b2) + 256) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; }
        size_t tb = std::max(b1, b2) + 256;
        cub::DeviceRadixSort::SortPairs(tmp, tb, c->d_keys, k2, c->d_cnt, v2, (int)n);
        tb = std::max(b1, b2) + 256;
        cub::DeviceReduce::ReduceByKey(tmp, tb, k2, k3, v2, v3, d_runs, cub::Sum(), (int)n);
        g_km_launches += 2;
        unsigned long long runs = 0;
        if (cudaMemcpy(&runs, d_runs, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        if (cudaMemcpy(c->d_keys, k3, runs * 8, cudaMemcpyDeviceToDevice) != cudaSuccess || cudaMemcpy(c->d_cnt, v3, runs * 4, cudaMemcpyDeviceToDevice) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        c->n = runs; c->n_sorted = runs; c->runs = 1;
    } while (0);
    if (rc == KMAT_ERR_NOMEM) { cudaGetLastError(); kmat_set_error("kmat_kcov: out of device memory while merging %llu entries", (unsigned long long)n); }
    if (rc == KMAT_ERR_CUDA) kmat_set_error("kmat_kcov: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return rc;
}

static int kc_add_chunk(kmat_kcov *c, const char *bases, const uint64_t *offs, const uint32_t *groups, uint32_t n_reads) {
    const uint64_t nb = offs[n_reads] - offs[0];
    const uint64_t cap = nb * (uint64_t)c->n_k + 32;
    char *d_b = nullptr; uint64_t *d_o = nullptr; uint32_t *d_g = nullptr;
    uint64_t *kA = nullptr, *kB = nullptr, *kU = nullptr; uint32_t *rA = nullptr, *rB = nullptr, *fl = nullptr, *cU = nullptr;
    unsigned long long *d_cur = nullptr; void *tmp = nullptr;
    auto cleanup = [&] { cudaFree(d_b); cudaFree(d_o); cudaFree(d_g); cudaFree(kA); cudaFree(kB); cudaFree(kU); cudaFree(rA); cudaFree(rB); cudaFree(fl); cudaFree(cU); cudaFree(d_cur); cudaFree(tmp); };
    int rc = KMAT_OK;
    do {
        std::vector<uint64_t> lo((size_t)n_reads + 1);
        for (uint32_t i = 0; i <= n_reads; i++) lo[i] = offs[i] - offs[0];
        if (cudaMalloc((void **)&d_b, nb + 1) != cudaSuccess || cudaMalloc((void **)&d_o, ((size_t)n_reads + 1) * 8) != cudaSuccess || cudaMalloc((void **)&d_g, (size_t)n_reads * 4) != cudaSuccess ||
            cudaMalloc((void **)&kA, cap * 8) != cudaSuccess || cudaMalloc((void **)&kB, cap * 8) != cudaSuccess || cudaMalloc((void **)&rA, cap * 4) != cudaSuccess ||
            cudaMalloc((void **)&rB, cap * 4) != cudaSuccess || cudaMalloc((void **)&d_cur, 16) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; }
        if (cudaMemcpy(d_b, bases + offs[0], nb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(d_o, lo.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d_g, groups, (size_t)n_reads * 4, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemset(d_cur, 0, 16) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        KcEmitParams P;
        P.bases = d_b; P.offs = d_o; P.groups = d_g; P.n_reads = n_reads; P.n_k = c->n_k;
        for (int i = 0; i < KC_MAX_K; i++) P.k[i] = i < c->n_k ? c->k[i] : 0;
        P.keys = kA; P.reads = rA; P.cursor = d_cur;
        km_kcov_emit_kernel<<<(int)std::max<uint32_t>(1, std::min<uint32_t>((n_reads + 7) / 8, 148u * 8)), 256>>>(P);
        g_km_launches++;
        unsigned long long n = 0;
        if (cudaMemcpy(&n, d_cur, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        if (n == 0) break;
        if (n > cap || n >= (1ull << 31)) { kmat_set_error("kmat_kcov_add: internal error (%llu pairs, capacity %llu)", n, (unsigned long long)cap); rc = KMAT_ERR_UNSUPPORTED; break; }
        if (cudaMalloc((void **)&fl, n * 4) != cudaSuccess || cudaMalloc((void **)&kU, n * 8) != cudaSuccess || cudaMalloc((void **)&cU, n * 4) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; }
        int read_bits = 1;
        while ((1ull << read_bits) < n_reads) read_bits++;
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, rA, rB, kA, kB, (int)n, 0, read_bits);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, kB, kA, rB, rA, (int)n);
        cub::DeviceReduce::ReduceByKey(nullptr, b3, kA, kU, fl, cU, d_cur + 1, cub::Sum(), (int)n);
        const size_t tbytes = std::max(b1, std::max(b2, b3)) + 256;
        if (cudaMalloc(&tmp, tbytes) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; }
        size_t tb = tbytes;
        cub::DeviceRadixSort::SortPairs(tmp, tb, rA, rB, kA, kB, (int)n, 0, read_bits);       // by read ...
        tb = tbytes;
        cub::DeviceRadixSort::SortPairs(tmp, tb, kB, kA, rB, rA, (int)n);                      // ... then, stable, by key
        km_kcov_flag_kernel<<<(int)((n + 255) / 256), 256>>>(kA, rA, n, fl);
        tb = tbytes;
        cub::DeviceReduce::ReduceByKey(tmp, tb, kA, kU, fl, cU, d_cur + 1, cub::Sum(), (int)n);
        g_km_launches += 4;
        unsigned long long runs = 0;
        if (cudaMemcpy(&runs, d_cur + 1, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        // append to the accumulated runs
        if (c->n + runs > c->cap) {
            const uint64_t ncap = (c->n + runs) * 3 / 2 + 1024;
            uint64_t *nk = nullptr; uint32_t *nc = nullptr;
            if (cudaMalloc((void **)&nk, ncap * 8) != cudaSuccess || cudaMalloc((void **)&nc, ncap * 4) != cudaSuccess) { cudaFree(nk); rc = KMAT_ERR_NOMEM; break; }
            if (c->n) { cudaMemcpy(nk, c->d_keys, c->n * 8, cudaMemcpyDeviceToDevice); cudaMemcpy(nc, c->d_cnt, c->n * 4, cudaMemcpyDeviceToDevice); }
            cudaFree(c->d_keys); cudaFree(c->d_cnt);
            c->d_keys = nk; c->d_cnt = nc; c->cap = ncap;
        }
        if (cudaMemcpy(c->d_keys + c->n, kU, runs * 8, cudaMemcpyDeviceToDevice) != cudaSuccess || cudaMemcpy(c->d_cnt + c->n, cU, runs * 4, cudaMemcpyDeviceToDevice) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        c->n += runs; c->runs++;
    } while (0);
    if (rc == KMAT_ERR_NOMEM) { cudaGetLastError(); kmat_set_error("kmat_kcov_add: out of device memory"); }
    if (rc == KMAT_ERR_CUDA) kmat_set_error("kmat_kcov_add: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    if (rc == KMAT_OK && cudaGetLastError() != cudaSuccess) { kmat_set_error("kmat_kcov_add: kernel launch failed"); rc = KMAT_ERR_CUDA; }
    // keep the accumulated runs bounded: small k saturate quickly (4^k / 2 distinct k-mers per group)
    if (rc == KMAT_OK && c->runs >= 8 && c->n > (64u << 20)) rc = kc_compact(c);
    return rc;
}

extern "C" int kmat_kcov_add(kmat_kcov *c, const char *bases, const uint64_t *offs, const uint32_t *groups, uint32_t n_reads) {
    if (!c || !offs || !groups || (n_reads && !bases)) { kmat_set_error("kmat_kcov_add: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= c->device) { kmat_set_error("CUDA device %d not available", c->device); return KMAT_ERR_NO_DEVICE; }
    for (uint32_t r = 0; r < n_reads; r++)
        if (groups[r] != KMAT_NONE && groups[r] >= (1u << KC_GROUP_BITS)) { kmat_set_error("kmat_kcov_add: group %u of read %u exceeds %d bits", groups[r], r, KC_GROUP_BITS); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    c->finished = false;
    const uint64_t max_pairs = 96ull << 20;                  // per chunk: ~45 bytes of device memory per pair
    for (uint32_t r0 = 0; r0 < n_reads;) {
        uint32_t r1 = r0 + 1;
        while (r1 < n_reads && (offs[r1 + 1] - offs[r0]) * (uint64_t)c->n_k <= max_pairs) r1++;
        const int rc = kc_add_chunk(c, bases, offs + r0, groups + r0, r1 - r0);
        if (rc != KMAT_OK) return rc;
        r0 = r1;
    }
    return KMAT_OK;
}

extern "C" int kmat_kcov_finish(kmat_kcov *c) {
    if (!c) { kmat_set_error("kmat_kcov_finish: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= c->device) { kmat_set_error("CUDA device %d not available", c->device); return KMAT_ERR_NO_DEVICE; }
    KM_CUDA(cudaSetDevice(c->device));
    const int rc = kc_compact(c);
    if (rc != KMAT_OK) return rc;
    c->h_keys.resize(c->n); c->h_cnt.resize(c->n);
    if (c->n) {
        KM_CUDA(cudaMemcpy(c->h_keys.data(), c->d_keys, c->n * 8, cudaMemcpyDeviceToHost));
        KM_CUDA(cudaMemcpy(c->h_cnt.data(), c->d_cnt, c->n * 4, cudaMemcpyDeviceToHost));
    }
    c->finished = true;
    return KMAT_OK;
}

extern "C" int kmat_kcov_query(kmat_kcov *c, int k_index, uint32_t group, uint64_t *distinct, uint64_t *total, uint32_t *hist_count, uint64_t *hist_n,
                               uint32_t cap, uint32_t *n_hist) {
    if (!c || k_index < 0 || k_index >= c->n_k || group >= (1u << KC_GROUP_BITS)) { kmat_set_error("kmat_kcov_query: bad argument"); return KMAT_ERR_ARG; }
    if (!c->finished) { kmat_set_error("kmat_kcov_query: call kmat_kcov_finish first"); return KMAT_ERR_ARG; }
    const uint64_t lo = ((uint64_t)k_index << (KC_GROUP_BITS + KC_KMER_BITS)) | ((uint64_t)group << KC_KMER_BITS), hi = lo + (1ull << KC_KMER_BITS);
    const size_t a = std::lower_bound(c->h_keys.begin(), c->h_keys.end(), lo) - c->h_keys.begin();
    const size_t b = std::lower_bound(c->h_keys.begin(), c->h_keys.end(), hi) - c->h_keys.begin();
    std::map<uint32_t, uint64_t> hist;
    uint64_t tot = 0;
    for (size_t i = a; i < b; i++) { tot += c->h_cnt[i]; hist[c->h_cnt[i]]++; }
    if (distinct) *distinct = b - a;
    if (total) *total = tot;
    if (n_hist) *n_hist = (uint32_t)hist.size();
    if (hist.size() > cap) { if (cap) { kmat_set_error("kmat_kcov_query: %zu histogram entries, capacity %u", hist.size(), cap); return KMAT_ERR_OVERFLOW; } return KMAT_OK; }
    uint32_t i = 0;
    for (const auto &kv : hist) { if (hist_count) hist_count[i] = kv.first; if (hist_n) hist_n[i] = kv.second; i++; }
    return KMAT_OK;
}
