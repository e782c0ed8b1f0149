// kmat_label.cu -- kmat_ctx and the per-read scoring kernel (K3 candidate sets + K4 scoring / LCA).
//
// Restates, per read, retrieve_kmer_labels' list handling and post-pass (read_label.cpp:1031-1204),
// construct_labels (:692-941) and findReadLabelVer2 (:284-419).  One warp per read: lanes own k-mer
// positions for the data-parallel parts; the order-dependent float arithmetic and the std::sort / LCA walk run
// on lane 0 from shared memory, in the reference's own operation order (no FMA contraction: every float op is
// an explicit __f*_rn intrinsic; logf is the glibc algorithm, km_logf).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kmat_device.cuh"
#include "kmat_priv.h"
#include "kmat_std_emul.cuh"

#define KB_WARPS 8         // warps per CTA of the candidate kernel
#define KB_CMAX 64         // candidate taxids per read: candidate i lives in lane i & 31, register slot i >> 5
#define KB_LFAST 16        // list length resolved in registers (libstdc++ uses plain insertion sort up to 16)
#define KB_LIN 128         // lineage entries in findReadLabelVer2
#ifndef KS_THREADS
#define KS_THREADS 128     // threads per CTA of the scoring kernel (one read per thread)
#endif
#define KMAT_ST_PENDING 7  // internal: candidates built, scoring still to run
#define KMAT_ST_PENDING_BIG 8   // internal: candidates built by km_cand_big_kernel (more than KB_CMAX of them), scored by km_score_big_kernel
#define KMAT_ST_DEFERRED 9      // internal: queued for the slow candidate kernel
#define KB_CBIG 512        // candidate taxids per read the slow path holds (reads beyond it: KMAT_ERR_UNSUPPORTED)
#define KB_LBIG 1024       // lineage entries of the big scoring kernel
#define KB_BIGQ (1u << 20)  // reads per pass the slow path takes
#define KMAT_ST_PENDING_HUGE 10 // internal: candidates built by km_cand_huge_kernel (more than KB_CBIG of them), scored by km_score_huge_kernel
#define KB_CHUGE 16384     // candidate taxids per read of the last-resort path (all working arrays in global memory, linear in the count)
#define KB_HHASH_BITS 15   // its taxid -> index hash: 2 * KB_CHUGE slots
#define KB_HLIN (1u << 20) // lineage entries (candidate indices) one read's qualifying members may add up to
#define KB_HUGEQ (1u << 16) // reads per pass the last-resort path takes
#define KBH_WARPS 128      // warps (reads in flight) of km_cand_huge_kernel (2.8 MB of scratch each); KBH_THREADS reads in flight in
#define KBH_THREADS 256    // km_score_huge_kernel (0.55 MB each): ~0.5 GB per context, so that a sample rich in such reads does not crawl
#define KB_PASS_SHIFT 40   // per-pass cursor: pairs in the low 40 bits, reads queued for K4 above
#define KB_PASS_MASK ((1ull << KB_PASS_SHIFT) - 1)
#define KB_PASS_ONE_READ (1ull << KB_PASS_SHIFT)

// stored id -> node: low 30 bits nid, bit 31 = isHuman, bit 30 = dropped tid (flags folded in so that the common
// singleton hit needs no node-record load)
#define KB_SID_HUMAN 0x80000000u
#define KB_SID_DROP 0x40000000u
#define KB_SID_NIDMASK 0x3FFFFFFFu

// resolved list record (pool2, built once per ctx by km_resolve_kernel): header word, then node ids
#define KR_ERR_BAD 0xFFFFFFFFu    // a stored id is missing from the -f map (TaxNodeStat.hpp:235-238 asserts)
#define KR_ERR_BIG 0xFFFFFFFEu    // deferred to the big-list pass (internal, never left behind)

struct KmCtxDev {
    KmDbDev db;
    const KmNodeA *nodeA; const KmNodeB *nodeB; const uint32_t *paths; const uint32_t *prune_rank; const uint32_t *sid2nid;
    uint32_t n_sid, n_nodes, nid_human, nid_one;
    int nbins, n_models, n_classes;
    const int16_t *model_of_cand; const int32_t *mrow; const float *cut; const uint8_t *cls;
    int8_t class_ranknum[64];
    kmat_opts opt;
    const uint32_t *pool2;     // resolved lists: record of the list at pool word offset o starts at pool2[o * pool2_mul]
    int pool2_mul;
};

struct KmScoreParams {
    KmCtxDev C;
    const uint64_t *offs; uint32_t n_reads; const uint32_t *hit; const int2 *hdr;
    kmat_read_result *out;
    kmat_pair *cands; unsigned long long *cand_cursor; unsigned long long cand_cap;
    // cand_cursor: pairs handed out by the passes before this one (read-only here).  pass_cursor: this pass's packed
    // cursor -- low 40 bits pairs, high 24 bits reads queued in pend_q -- so that ONE atomic per read gives K3 both the
    // place of its pairs and a slot in the dense queue of reads K4 has to score (km_cursor_roll_kernel folds it into
    // cand_cursor afterwards).  pend_q == NULL: no queue, K4 walks all reads.
    unsigned long long *pass_cursor; uint32_t *pend_q;
    kmat_pair *lin; unsigned long long *lin_cursor; unsigned long long lin_cap;
    unsigned long long *long_masks; uint32_t long_cap;    // per-warp position-mask scratch for reads longer than the register path
    KmStatsDev *stats;
    // slow path for reads with more than KB_CMAX candidate taxids (or more than KB_LIN lineage entries): K3 queues them
    // in big_qa for km_cand_big_kernel, which (like K4 on a lineage overflow) queues them in big_qb for the big scoring kernel
    uint32_t *big_qa, *big_qb; unsigned int *big_cnt;     // big_cnt[0] / [1]: entries of big_qa / big_qb (may exceed KB_BIGQ: the excess is dropped)
    unsigned char *big_scratch3, *big_scratch4; uint32_t big_np_cap, big_threads3, big_threads4;
    // last resort for reads with more than KB_CBIG candidates: km_cand_big_kernel queues them in huge_qa for km_cand_huge_kernel,
    // which queues them in huge_qb for km_score_huge_kernel (big_cnt[2] / [3] count the entries)
    uint32_t *huge_qa, *huge_qb; unsigned char *huge_scratch3, *huge_scratch4;
};

struct KmRl { float score; uint32_t idx; };          // rank_label element: candidate index + (bias-adjusted) score

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ KmNodeA kb_nodeA(const KmCtxDev &C, uint32_t nid) {
    const uint4 v = __ldg((const uint4 *)(C.nodeA + nid));
    return KmNodeA{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ KmNodeB kb_nodeB(const KmCtxDev &C, uint32_t nid) {
    const uint4 v = __ldg((const uint4 *)(C.nodeB + nid));
    return KmNodeB{v.x, v.y, v.z, v.w};
}
// strict ancestor test on Euler intervals == "anc is on TaxTree::getPathToRoot(desc)" (read_label.cpp:138-150)
__device__ __forceinline__ bool kb_is_anc(uint32_t a_tin, uint32_t a_tout, uint32_t d_tin) { return a_tin < d_tin && d_tin <= a_tout; }
__device__ __forceinline__ unsigned long long kb_warp_min64(unsigned long long v) {
    const uint32_t hi = __reduce_min_sync(KM_FULL, (uint32_t)(v >> 32));
    const uint32_t lo = __reduce_min_sync(KM_FULL, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0xFFFFFFFFu);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long kb_shfl64(unsigned long long v, int src) {
    const uint32_t lo = __shfl_sync(KM_FULL, (uint32_t)v, src), hi = __shfl_sync(KM_FULL, (uint32_t)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t kb_list_count(const KmDbDev &db, uint32_t off) {
    return db.tid_bytes == 2 ? (uint32_t)(*(const uint16_t *)(db.pool + off)) : db.pool[off];
}
__device__ __forceinline__ uint32_t kb_list_id(const KmDbDev &db, uint32_t off, uint32_t j) {
    return db.tid_bytes == 2 ? (uint32_t)((const uint16_t *)(db.pool + off))[1 + j] : db.pool[off + 1 + j];
}

struct KbDepthDesc {      // CmpDepth1 (read_label.cpp:169-177) over {nid, depth} pairs
    __device__ bool operator()(const uint2 &a, const uint2 &b) const { return (int)a.y > (int)b.y; }
};
struct KbRankLess {       // MyPair::operator< (SortedDb.hpp:133-135): by rank number only
    __device__ bool operator()(const uint2 &a, const uint2 &b) const { return a.y < b.y; }
};

// ---------------------------------------------------------------------------------------------
// List resolution, once per ctx (the result depends on the table, the taxonomy and the -g / -s options, not on
// the read).  For every stored taxid list the record in pool2 holds what retrieve_kmer_labels derives from it:
//   TaxNodeStat::begin + next (TaxNodeStat.hpp:60-256, run-time -g pruning included), the 16->32 conversion,
//   human collapse and dropped ids (read_label.cpp:1031-1038), the depth sort (:1073-1074) and
//   default mode : the leaf filter (:1103-1134)       -> [m | n_raw << 16][m node ids, insertion order]
//   permissive   : (:1050-1058, 1075-1102)            -> [a | n_raw << 16][b][a ids in list order][b ids in depth
//                  order whose root paths are added: every id before the first depth-0 one]
// seq = n entries {node id, -} in next() order.  Returns nothing; writes the record.
// ---------------------------------------------------------------------------------------------
__device__ void kr_finish(const KmCtxDev &C, uint2 *seq, int n, uint32_t n_raw, uint32_t *rec) {
    bool seenHuman = false;
    int w = 0;
    for (int i = 0; i < n; i++) {                                         // :1031-1038
        uint32_t nid = seq[i].x;
        uint32_t meta = kb_nodeA(C, nid).meta;
        if ((meta & KM_META_HUMAN) && !C.opt.rkmer_mode) {              // no collapse in rkmer.hpp (:119-121)
            if (seenHuman) continue;
            nid = C.nid_human; meta = kb_nodeA(C, nid).meta; seenHuman = true;
        }
        if (meta & KM_META_DROP) continue;
        seq[w++] = make_uint2(nid, meta & KM_META_DEPTH_MASK);
    }
    if (C.opt.permissive) {
        rec[0] = (uint32_t)w | (n_raw << 16);
        for (int i = 0; i < w; i++) rec[2 + i] = seq[i].x;                // inserted in next() order (:1050-1058)
        kmstd::sort(seq, w, KbDepthDesc());                               // :1073-1074
        int b = 0;
        for (int i = 0; i < w; i++) {                                     // :1077-1101 (last_depth is never updated there)
            if (seq[i].y == 0) break;
            rec[2 + w + b] = seq[i].x; b++;
        }
        rec[1] = (uint32_t)b;
        return;
    }
    kmstd::sort(seq, w, KbDepthDesc());                                   // :1073-1074
    int m = 0;                                                            // leaf filter :1103-1134
    for (int i = 0; i < w; i++) {
        const KmNodeB tb = kb_nodeB(C, seq[i].x);
        bool anc = false;
        for (int j = 0; j < m && !anc; j++) anc = kb_is_anc(tb.tin, tb.tout, kb_nodeB(C, seq[j].x).tin);
        if (!anc) seq[m++] = seq[i];
    }
    rec[0] = (uint32_t)m | (n_raw << 16);
    for (int j = 0; j < m; j++) rec[1 + j] = seq[j].x;
}

// One list.  scratch: 2 * count entries when the pruning heap is needed, else count entries.
__device__ void kr_resolve_list(const KmCtxDev &C, uint32_t lo, uint2 *scratch, uint32_t *rec) {
    int count = (int)kb_list_count(C.db, lo);
    const uint32_t n_raw = (uint32_t)count;
    uint2 *seq = scratch;                  // ids in next() order
    int n = 0;
    const int tid_cut = C.opt.max_count;
    if (tid_cut > 0 && count > tid_cut) {
        if (!C.prune_rank) {               // p_map.size() == 0: count forced to 1, next() reads the first stored id (TaxNodeStat.hpp:78-81)
            const uint32_t sid = kb_list_id(C.db, lo, 0);
            const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
            if (e == KMAT_NONE) { rec[0] = KR_ERR_BAD; return; }
            seq[n++] = make_uint2(e & KB_SID_NIDMASK, 0);
        } else {                           // TaxNodeStat.hpp:118-201
            uint2 *heap = scratch + count;
            int hn = 0;
            for (int i = 0; i < count; i++) {
                const uint32_t sid = kb_list_id(C.db, lo, i);
                const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
                if (e == KMAT_NONE) { rec[0] = KR_ERR_BAD; return; }
                const uint32_t nid = e & KB_SID_NIDMASK;
                kmstd::pq_push(heap, hn, make_uint2(nid, C.prune_rank[nid]), KbRankLess());
            }
            int newcount = count;
            while (hn > 0) {
                const uint32_t pr = heap[0].y;
                while (heap[0].y == pr) { kmstd::pq_pop(heap, hn, KbRankLess()); if (hn == 0) break; }
                if (hn <= tid_cut) { newcount = hn; break; }
            }
            if (hn == 0) { newcount = 1; kmstd::pq_push(heap, hn, make_uint2(C.nid_one, 1u), KbRankLess()); }
            for (int i = 0; i < newcount; i++) seq[n++] = kmstd::pq_pop(heap, hn, KbRankLess());
        }
    } else {
        for (int i = 0; i < count; i++) {
            const uint32_t sid = kb_list_id(C.db, lo, i);
            const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
            if (e == KMAT_NONE) { rec[0] = KR_ERR_BAD; return; }
            seq[n++] = make_uint2(e & KB_SID_NIDMASK, 0);
        }
    }
    kr_finish(C, seq, n, n_raw, rec);
}

struct KmResolveParams {
    KmCtxDev C;
    uint32_t *pool2;
    unsigned long long n_slots, n_table_slots, n_line_slots;   // n_slots = first-level slots + second-level slots + stash entries
    uint32_t *big_queue; uint32_t big_cap;       // pool offsets of the lists left to the big pass
    unsigned int *counters;                      // [0] lists queued, [1] longest list (entries), [2] lists seen
    uint2 *scratch; uint32_t scratch_entries;    // big pass: per-thread scratch
};
// pass 1: one thread per table slot; short lists are resolved in local memory, the others queued
__global__ void __launch_bounds__(256) km_resolve_kernel(KmResolveParams R) {
    const KmCtxDev &C = R.C;
    uint2 local[2 * KB_LFAST];
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < R.n_slots; s += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t lo;
        if (s < R.n_line_slots + R.n_table_slots) {       // both levels use bit 63 = occupied, bit 62 = list, payload below
            const uint64_t v = s < R.n_line_slots ? C.db.lines[s] : C.db.slots[s - R.n_line_slots];
            if (!(v >> 63) || !((v >> 62) & 1)) continue;
            lo = (uint32_t)v & 0x7FFFFFFFu;
        } else {                                          // the overflow stash holds hit words
            const uint32_t hw = C.db.stash_hit[s - R.n_line_slots - R.n_table_slots];
            if (!(hw & KM_HIT_LIST)) continue;
            lo = hw & 0x7FFFFFFFu;
        }
        const uint32_t count = kb_list_count(C.db, lo);
        atomicMax(&R.counters[1], count);
        atomicAdd(&R.counters[2], 1u);
        uint32_t *rec = R.pool2 + (size_t)lo * C.pool2_mul;
        if (count <= KB_LFAST) kr_resolve_list(C, lo, local, rec);
        else {
            rec[0] = KR_ERR_BIG;
            const unsigned int q = atomicAdd(&R.counters[0], 1u);
            if (q < R.big_cap) R.big_queue[q] = lo;
        }
    }
}
// pass 2: the queued long lists, one thread each over a per-thread global scratch
__global__ void __launch_bounds__(128) km_resolve_big_kernel(KmResolveParams R, uint32_t n_big) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint2 *scratch = R.scratch + (size_t)t * R.scratch_entries;
    for (uint32_t q = t; q < n_big; q += gridDim.x * blockDim.x) {
        const uint32_t lo = R.big_queue[q];
        kr_resolve_list(R.C, lo, scratch, R.pool2 + (size_t)lo * R.C.pool2_mul);
    }
}

// ---------------------------------------------------------------------------------------------
// K3: candidate sets.  One warp per read, all per-read state in registers:
//   * candidate i (taxid_lst entry) lives in lane i & 31, slot i >> 5: node id, first-appearance key, leaf count ...
//   * position p lives in lane p & 31, chunk p >> 5: a 64-bit set of candidate indices (label_vec[pos].second)
// Restates the list handling and the post-pass of retrieve_kmer_labels (read_label.cpp:1031-1204) and the
// per-taxid position counts of construct_labels (:748-759).  Leaves, per read, the candidates in taxid_lst order
// as (node id, hits) pairs in the cands buffer for the scoring kernel.
// NCH = chunks of 32 positions kept in registers (0: per-warp global scratch, any length).
// ---------------------------------------------------------------------------------------------
struct KbCand {            // per lane, two candidate slots
    uint32_t nid[2], key[2], leaf[2];
};

// find-or-append `v` (warp-uniform) among the candidates; returns its index or -1 when KB_CMAX is exceeded
__device__ __forceinline__ int kb_find_or_add(KbCand &K, int &C, uint32_t v, int lane) {
    const uint32_t f0 = __ballot_sync(KM_FULL, K.nid[0] == v);
    if (f0) return __ffs(f0) - 1;
    if (C > 32) {
        const uint32_t f1 = __ballot_sync(KM_FULL, K.nid[1] == v);
        if (f1) return 32 + __ffs(f1) - 1;
    }
    if (C >= KB_CMAX) return -1;
    const int idx = C++;
    if (lane == (idx & 31)) {
        if (idx < 32) { K.nid[0] = v; K.key[0] = 0xFFFFFFFFu; K.leaf[0] = 0; }
        else { K.nid[1] = v; K.key[1] = 0xFFFFFFFFu; K.leaf[1] = 0; }
    }
    return idx;
}

// One chunk of 32 positions (lane = position): insert the members of every position into the candidate set and
// return this lane's position mask.  All lanes of the warp call it together.
// the two dependent loads of a position, split off so that a caller can issue them for every chunk of a read up front:
// the hit word, and then the word it points at (stored id -> node entry for a singleton, record header for a list)
// resolved record of a list hit word
__device__ __forceinline__ const uint32_t *kb_rec_of(const KmCtxDev &X, uint32_t hw) {
    return X.pool2 + (size_t)(hw & 0x7FFFFFFFu) * X.pool2_mul;
}
__device__ __forceinline__ uint32_t kb_load_hit(const KmScoreParams &P, int c, uint64_t off, int np, int lane) {
    const int p = (c << 5) + lane;
    return p < np ? __ldg(P.hit + off + p) : KM_HIT_INVALID;
}
__device__ __forceinline__ uint32_t kb_load_aux(const KmCtxDev &X, uint32_t hw) {
    if (hw == KM_HIT_INVALID || hw == KM_HIT_MISS) return 0;
    if (!(hw & KM_HIT_LIST)) return hw < X.n_sid ? __ldg(X.sid2nid + hw) : KMAT_NONE;
    return __ldg(kb_rec_of(X, hw));
}

__device__ __forceinline__ unsigned long long kb_chunk(const KmScoreParams &P, KbCand &K, int &C, int c, uint32_t hw, uint32_t aux, int lane,
                                                        int &cand_cnt, int &fnd_cnt, int &err, bool &overflow,
                                                        unsigned long long &st_list_ids, unsigned long long &st_list_sectors) {
    const KmCtxDev &X = P.C;
    const bool permissive = X.opt.permissive != 0;
    // member iterator state: `a` ids first (a single id in v0, or a resolved record at rec), then (permissive) the
    // root paths of `b` ids
    uint32_t a = 0, b = 0, v0 = KMAT_NONE;
    const uint32_t *rec = nullptr;
    if (hw != KM_HIT_INVALID) {
        cand_cnt++;                                                   // label_vec[pos].first >= 0 (:1015, :702)
        if (hw != KM_HIT_MISS) {
            if (!(hw & KM_HIT_LIST)) {
                // singleton: one stored id.  16->32 conversion, human collapse, dropped tids (:1031-1038)
                const uint32_t e = aux;
                if (e == KMAT_NONE) err = KMAT_ERR_BAD_TAXID;          // "bad taxid" assert (TaxNodeStat.hpp:235-238)
                else if (!(e & KB_SID_DROP)) {
                    v0 = ((e & KB_SID_HUMAN) && !X.opt.rkmer_mode) ? X.nid_human : (e & KB_SID_NIDMASK); a = 1;
                    if (permissive) b = (kb_nodeA(X, v0).meta & KM_META_DEPTH_MASK) ? 1 : 0;
                }
            } else {
                rec = kb_rec_of(X, hw);
                const uint32_t h = aux;
                if (h == KR_ERR_BAD) { err = KMAT_ERR_BAD_TAXID; rec = nullptr; }
                else {
                    a = h & 0xFFFFu;
                    const uint32_t n_raw = h >> 16;
                    st_list_ids += n_raw;
                    st_list_sectors += (2 + n_raw * X.db.tid_bytes + 31) / 32;
                    if (permissive) { b = rec[1]; rec += 2; } else rec += 1;
                }
            }
        }
    }
    if (a) fnd_cnt++;
    unsigned long long mymask = 0;
    uint32_t seqno = 0;                               // index of the next member in this position's insertion order
    uint32_t bi = 0, pq = 0, poff = 0, plen = 0;      // permissive: current path owner / cursor
    for (;;) {
        uint32_t val = KMAT_NONE;                     // next member of this lane
        if (seqno < a) val = rec ? rec[seqno] : v0;
        else if (permissive) {
            while (bi < b && pq >= plen) {            // open the next path
                const uint32_t owner = rec ? rec[a + bi] : v0;
                const KmNodeB nb = kb_nodeB(X, owner);
                poff = nb.path_off; plen = nb.path_len; pq = 0; bi++;
            }
            if (pq < plen) val = X.paths[poff + pq++];
        }
        uint32_t pending = __ballot_sync(KM_FULL, val != KMAT_NONE);
        if (!pending) break;
        while (pending) {                             // one round per distinct taxid among the lanes
            const int leader = __ffs(pending) - 1;
            const uint32_t v = __shfl_sync(KM_FULL, val, leader);
            const uint32_t grp = __ballot_sync(KM_FULL, val == v);
            const int idx = kb_find_or_add(K, C, v, lane);
            if (idx < 0) { overflow = true; break; }
            if (lane == (idx & 31)) {
                // first appearance in taxid_lst order: lowest position, then insertion order (:1111-1122)
                // (leader = lowest lane of the group = its lowest position; seqno is the same in every lane)
                const uint32_t key = ((uint32_t)((c << 5) + leader) << 16) | min(seqno, 0xFFFFu);
                if (idx < 32) { K.key[0] = min(K.key[0], key); K.leaf[0] += __popc(grp); }
                else { K.key[1] = min(K.key[1], key); K.leaf[1] += __popc(grp); }
            }
            if (val == v) mymask |= 1ull << idx;
            pending &= ~grp;
        }
        if (overflow) break;
        seqno++;
    }
    return mymask;
}

// A read K3's register-resident candidate set cannot hold: hand it to km_cand_big_kernel (all lanes call; lane 0 acts)
__device__ __forceinline__ void kb_defer_big(const KmScoreParams &P, uint32_t r, kmat_read_result &res, int lane) {
    if (lane != 0) return;
    res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED;
    if (P.big_qa) {
        const unsigned int q = atomicAdd(P.big_cnt, 1u);
        if (q < KB_BIGQ) { P.big_qa[q] = r; res.status = KMAT_ST_DEFERRED; res.err = 0; }
    }
    P.out[r] = res;
}

template <int NCH>
__global__ void __launch_bounds__(KB_WARPS * 32, NCH == 5 ? 4 : 1) km_cand_kernel(KmScoreParams P) {
    const KmCtxDev &X = P.C;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = blockIdx.x * KB_WARPS + (threadIdx.x >> 5), n_warps = gridDim.x * KB_WARPS;
    const int k = X.db.kmer_len;
    const bool permissive = X.opt.permissive != 0;
    unsigned long long *gmask = NCH == 0 ? P.long_masks + (size_t)warp_global * P.long_cap : nullptr;
    const unsigned long long cand_base = *P.cand_cursor;
    unsigned long long st_list_ids = 0, st_list_sectors = 0, st_fast = 0, st_err = 0;

    for (uint32_t r = warp_global; r < P.n_reads; r += n_warps) {
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        const int2 hd = P.hdr[r];
        kmat_read_result res;
        memset(&res, 0, sizeof res);
        res.valid_kmers = hd.x; res.bin_sel = hd.y; res.match = KMAT_NOMATCH;
        bool done = false;
        if (len < k) { res.status = KMAT_ST_SHORT_LEN; res.n1 = len; res.n2 = k; res.valid_kmers = 0; done = true; }             // :1217-1218
        else if (hd.x < X.opt.min_kmer) { res.status = KMAT_ST_SHORT_VALID; res.n1 = hd.x; res.n2 = X.opt.min_kmer; done = true; }   // :1232-1233
        else if (NCH ? np > NCH * 32 : ((uint32_t)np > P.long_cap || np > 0xFFFF)) {      // first-appearance keys hold 16-bit positions: reads beyond 65 kbp are refused, not mis-ordered
            res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; done = true; st_err++;
        }
        if (done) { if (lane == 0) P.out[r] = res; continue; }

        KbCand K;
        K.nid[0] = K.nid[1] = KMAT_NONE; K.key[0] = K.key[1] = 0xFFFFFFFFu; K.leaf[0] = K.leaf[1] = 0;
        int C = 0;                                   // candidates so far (warp-uniform)
        unsigned long long pm[NCH ? NCH : 1];        // position masks of this lane
        int cand_cnt = 0, fnd_cnt = 0, err = 0;
        bool overflow = false;
        const int nch = (np + 31) >> 5;
        // ---- per position: members of label_vec[pos].second in insertion order (read_label.cpp:1019-1134)
        if (NCH) {
            // every hit word of the read, then every word they point at, before the first chunk is walked
            uint32_t hws[NCH ? NCH : 1], auxs[NCH ? NCH : 1];
#pragma unroll
            for (int c = 0; c < (NCH ? NCH : 1); c++) hws[c] = c < nch ? kb_load_hit(P, c, off, np, lane) : KM_HIT_INVALID;
#pragma unroll
            for (int c = 0; c < (NCH ? NCH : 1); c++) auxs[c] = kb_load_aux(X, hws[c]);
#pragma unroll
            for (int c = 0; c < (NCH ? NCH : 1); c++) pm[c] = 0ull;
            // one copy of the chunk walk in the instruction stream (it is the bulk of the kernel's code): the chunk's
            // words are selected into scalars and its mask selected back, so pm[] / hws[] stay in registers
#pragma unroll 1
            for (int c = 0; c < nch && !overflow; c++) {
                uint32_t hw1 = hws[0], aux1 = auxs[0];
#pragma unroll
                for (int q = 1; q < (NCH ? NCH : 1); q++) if (c == q) { hw1 = hws[q]; aux1 = auxs[q]; }
                const unsigned long long m = kb_chunk(P, K, C, c, hw1, aux1, lane, cand_cnt, fnd_cnt, err, overflow, st_list_ids, st_list_sectors);
#pragma unroll
                for (int q = 0; q < (NCH ? NCH : 1); q++) if (c == q) pm[q] = m;
            }
        } else {
            for (int c = 0; c < nch && !overflow; c++) {
                const uint32_t hw1 = kb_load_hit(P, c, off, np, lane);
                const unsigned long long m = kb_chunk(P, K, C, c, hw1, kb_load_aux(X, hw1), lane, cand_cnt, fnd_cnt, err, overflow, st_list_ids, st_list_sectors);
                if ((c << 5) + lane < np) gmask[(c << 5) + lane] = m;
            }
        }
        cand_cnt = km_warp_sum(cand_cnt); fnd_cnt = km_warp_sum(fnd_cnt);
        err = __reduce_max_sync(KM_FULL, err < 0 ? -err : 0);
        overflow = __any_sync(KM_FULL, overflow);
        if (err) { res.status = KMAT_ST_ERROR; res.err = -err; if (lane == 0) P.out[r] = res; st_err++; continue; }
        if (overflow) { kb_defer_big(P, r, res, lane); continue; }
        const int C1 = C;
        if (C1 == 0) {                                                       // taxid_lst.empty() -> NoDbHits (:1270-1271)
            res.status = KMAT_ST_NODBHITS; res.n1 = len; res.n2 = k;
            if (lane == 0) P.out[r] = res;
            st_fast++;
            continue;
        }
        const uint16_t cand16 = (uint16_t)cand_cnt;
        res.cand_kmer_cnt = cand16;
        if (fnd_cnt < X.opt.min_fnd_kmer || (int)cand16 < X.opt.min_kmer) {       // construct_labels :727-733: silent NoMatch
            res.status = KMAT_ST_SILENT; res.match = KMAT_NOMATCH; res.tid = 0; res.score = -1.0f;
            if (lane == 0) P.out[r] = res;
            st_fast++;
            continue;
        }
        // ---- per candidate: node data; taxid_lst index = rank of the first-appearance key; representative strain
        //      per species (:1143-1177) -> which members get their lineage added
        uint32_t c_tid[2], c_meta[2], c_spec[2], c_poff[2], c_plen[2], c_ord[2];
        unsigned long long c_anc[2] = {0ull, 0ull};
        bool c_qual[2] = {false, false};
#pragma unroll
        for (int s = 0; s < 2; s++) {
            c_tid[s] = c_meta[s] = 0; c_spec[s] = KMAT_NONE; c_poff[s] = c_plen[s] = 0; c_ord[s] = 0;
            if (lane + 32 * s < C1) {
                const KmNodeA na = kb_nodeA(X, K.nid[s]); const KmNodeB nb = kb_nodeB(X, K.nid[s]);
                c_tid[s] = na.tid; c_meta[s] = na.meta; c_spec[s] = na.species_anc; c_poff[s] = nb.path_off; c_plen[s] = nb.path_len;
            }
        }
        if (!permissive) {
            bool beaten[2] = {false, false};
            for (int j = 0; j < C1; j++) {
                const int src = j & 31;
                uint32_t kj, mj, sj, lj, tj;
                if (j < 32) { kj = __shfl_sync(KM_FULL, K.key[0], src); mj = __shfl_sync(KM_FULL, c_meta[0], src); sj = __shfl_sync(KM_FULL, c_spec[0], src); lj = __shfl_sync(KM_FULL, K.leaf[0], src); tj = __shfl_sync(KM_FULL, c_tid[0], src); }
                else { kj = __shfl_sync(KM_FULL, K.key[1], src); mj = __shfl_sync(KM_FULL, c_meta[1], src); sj = __shfl_sync(KM_FULL, c_spec[1], src); lj = __shfl_sync(KM_FULL, K.leaf[1], src); tj = __shfl_sync(KM_FULL, c_tid[1], src); }
                const bool strain_j = ((mj >> KM_META_RANK_SHIFT) & 3) == 1;
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    c_ord[s] += kj < K.key[s];
                    // another strain of the same species with more leaf hits, or as many and a smaller taxid (:1159)
                    if (strain_j && sj == c_spec[s] && (lj > K.leaf[s] || (lj == K.leaf[s] && tj < c_tid[s]))) beaten[s] = true;
                }
            }
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const bool strain = ((c_meta[s] >> KM_META_RANK_SHIFT) & 3) == 1;      // gRank_table[tid] == "strain"
                c_qual[s] = lane + 32 * s < C1 && (!strain || (c_spec[s] != KMAT_NONE && !beaten[s]));
            }
            // ---- lineage expansion (:1178-1203): qualifying members in (first position, taxid) order; the ancestors
            //      appended to taxid_lst get the next candidate indices
            unsigned long long key0 = c_qual[0] ? (((unsigned long long)(K.key[0] >> 16) << 32) | c_tid[0]) : ~0ull;
            unsigned long long key1 = c_qual[1] ? (((unsigned long long)(K.key[1] >> 16) << 32) | c_tid[1]) : ~0ull;
            for (;;) {
                const unsigned long long mine = key0 < key1 ? key0 : key1;
                const unsigned long long best = kb_warp_min64(mine);
                if (best == ~0ull) break;
                const int src = __ffs(__ballot_sync(KM_FULL, mine == best)) - 1;
                int slot = 0;
                if (lane == src) { if (key0 == best) { slot = 0; key0 = ~0ull; } else { slot = 1; key1 = ~0ull; } }
                slot = __shfl_sync(KM_FULL, slot, src);
                const uint32_t poff = __shfl_sync(KM_FULL, slot ? c_poff[1] : c_poff[0], src), plen = __shfl_sync(KM_FULL, slot ? c_plen[1] : c_plen[0], src);
                unsigned long long anc = 0;
                for (uint32_t c0 = 0; c0 < plen && !overflow; c0 += 32) {
                    const uint32_t av = c0 + lane < plen ? X.paths[poff + c0 + lane] : KMAT_NONE;
                    const int cnt = min(32u, plen - c0);
                    for (int q = 0; q < cnt; q++) {                            // path order = the order the reference appends them
                        const int idx = kb_find_or_add(K, C, __shfl_sync(KM_FULL, av, q), lane);
                        if (idx < 0) { overflow = true; break; }
                        anc |= 1ull << idx;
                    }
                }
                if (overflow) break;
                if (lane == src) c_anc[slot] = anc;
            }
            if (overflow) { kb_defer_big(P, r, res, lane); continue; }
            // ---- expanded position sets: every qualifying member brings its ancestors
            if (NCH) {
                for (int j = 0; j < C1; j++) {
                    const unsigned long long aj = kb_shfl64(j < 32 ? c_anc[0] : c_anc[1], j & 31);
                    if (!aj) continue;
#pragma unroll
                    for (int c = 0; c < (NCH ? NCH : 1); c++) if ((pm[c] >> j) & 1) pm[c] |= aj;
                }
            }                                   // NCH == 0: done in the counting sweep below, one pass over the global masks
        } else {
            for (int j = 0; j < C1; j++) {
                const uint32_t kj = __shfl_sync(KM_FULL, j < 32 ? K.key[0] : K.key[1], j & 31);
#pragma unroll
                for (int s = 0; s < 2; s++) c_ord[s] += kj < K.key[s];
            }
        }
        // ---- hits per candidate = number of positions whose set holds it (:748-759)
        uint32_t c_hits[2] = {0, 0};
        if (NCH) {
            // bit-sliced: the (up to) five masks of a lane are added per candidate bit with carry-save logic into three
            // bit planes (counts 0..5 for all 64 candidates at once); four candidates' counts are then spread into the
            // bytes of one word and summed over the warp by a single redux.add (32 lanes x 5 <= 160 fits a byte)
#pragma unroll
            for (int g5 = 0; g5 < (NCH ? NCH : 1); g5 += 5) {
                const unsigned long long a = pm[g5], b = g5 + 1 < NCH ? pm[g5 + 1] : 0ull, d = g5 + 2 < NCH ? pm[g5 + 2] : 0ull,
                                         e = g5 + 3 < NCH ? pm[g5 + 3] : 0ull, f = g5 + 4 < NCH ? pm[g5 + 4] : 0ull;
                const unsigned long long s1 = a ^ b ^ d, c1 = (a & b) | (a & d) | (b & d);
                const unsigned long long p0 = s1 ^ e ^ f, c2 = (s1 & e) | (s1 & f) | (e & f);
                const unsigned long long p1 = c1 ^ c2, p2 = c1 & c2;
                for (int g = 0; g * 4 < C; g++) {
                    const uint32_t n0 = (uint32_t)(p0 >> (4 * g)) & 15u, n1 = (uint32_t)(p1 >> (4 * g)) & 15u, n2 = (uint32_t)(p2 >> (4 * g)) & 15u;
                    // (n * 0x00204081) & 0x01010101 moves bit q of the nibble to bit 0 of byte q
                    const uint32_t w = ((n0 * 0x00204081u) & 0x01010101u) + 2u * ((n1 * 0x00204081u) & 0x01010101u) + 4u * ((n2 * 0x00204081u) & 0x01010101u);
                    const uint32_t tot = __reduce_add_sync(KM_FULL, w);
                    const uint32_t mine = (lane >> 2) == (g & 7) ? (tot >> (8 * (lane & 3))) & 0xFFu : 0u;
                    if (g < 8) c_hits[0] += mine; else c_hits[1] += mine;
                }
            }
        } else {
            // long reads: ONE sweep over the position masks in global scratch, 32 positions per step.  Expansion first (a
            // closure: an ancestor's lineage is part of its descendant's, so the member order does not matter), then one
            // ballot per candidate bit present in any of the 32 sets (neighbouring positions carry nearly the same set).
            __syncwarp();
            const unsigned long long qualmask = permissive ? 0ull
                : ((unsigned long long)__ballot_sync(KM_FULL, c_anc[0] != 0ull) | ((unsigned long long)__ballot_sync(KM_FULL, c_anc[1] != 0ull) << 32));
            for (int p0 = 0; p0 < np; p0 += 32) {
                unsigned long long m = p0 + lane < np ? gmask[p0 + lane] : 0ull;
                unsigned long long todo = qualmask & km_warp_or64(m);
                while (todo) {
                    const int j = __ffsll((long long)todo) - 1; todo &= todo - 1;
                    const unsigned long long aj = kb_shfl64(j < 32 ? c_anc[0] : c_anc[1], j & 31);
                    if ((m >> j) & 1) m |= aj;
                }
                unsigned long long u = km_warp_or64(m);
                while (u) {
                    const int i = __ffsll((long long)u) - 1; u &= u - 1;
                    const uint32_t cnt = __popc(__ballot_sync(KM_FULL, (m >> i) & 1));
                    if (lane == (i & 31)) { if (i < 32) c_hits[0] += cnt; else c_hits[1] += cnt; }
                }
            }
        }
        // ---- hand over to the scoring kernel: (nid, hits) in taxid_lst order
        unsigned long long co = 0;
        if (lane == 0) co = atomicAdd(P.pass_cursor, (unsigned long long)C | (P.pend_q ? KB_PASS_ONE_READ : 0ull));
        co = __shfl_sync(KM_FULL, co, 0);
        const uint32_t qi = (uint32_t)(co >> KB_PASS_SHIFT);
        co = cand_base + (co & KB_PASS_MASK);
        res.status = KMAT_ST_PENDING; res.n_cand = (uint32_t)C; res.cand_off = co;
        if (P.cands && co + C <= P.cand_cap) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int i = lane + 32 * s;
                if (i < C) P.cands[co + (i < C1 ? c_ord[s] : (uint32_t)i)] = kmat_pair{K.nid[s], __uint_as_float(c_hits[s])};
            }
        } else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_OVERFLOW; st_err++; }       // candidate buffer too small: the host re-runs with the size asked for
        if (lane == 0) { P.out[r] = res; if (P.pend_q) P.pend_q[qi] = r; }
        st_fast++;
    }
    if (P.stats && lane == 0) {
        atomicAdd(&P.stats->reads_fast, st_fast); atomicAdd(&P.stats->reads_error, st_err);
    }
    if (P.stats) {
        st_list_ids = (unsigned long long)km_warp_sum((int)st_list_ids); st_list_sectors = (unsigned long long)km_warp_sum((int)st_list_sectors);
        if (lane == 0) { atomicAdd(&P.stats->list_ids, st_list_ids); atomicAdd(&P.stats->list_sectors, st_list_sectors); }
    }
}

// ---------------------------------------------------------------------------------------------
// K4: scoring and LCA.  One THREAD per read (the arithmetic is order dependent and cannot be spread over
// lanes; running 32 reads per warp keeps the lanes busy instead).  Restates construct_labels from the null-model
// lookup on (read_label.cpp:735-941) and findReadLabelVer2 (:284-419) in the reference's float operation order.
// ---------------------------------------------------------------------------------------------
template <int CMAX, int LIN, typename PermT>
struct KsLocalT {
    uint32_t nid[CMAX], tid[CMAX], tin[CMAX], tout[CMAX];
    float score[CMAX];
    uint16_t depth[CMAX];
    uint8_t flags[CMAX], cls[CMAX];              // flags: 1 human, 2 phix, 4 plasmid
    KmRl rl[CMAX];
    uint32_t l_tid[LIN], l_tin[LIN], l_tout[LIN];
    float l_score[LIN];
    uint16_t l_depth[LIN];
    uint8_t l_nogood[LIN];
    PermT l_perm[LIN];
    float track_val[64];
    uint8_t track_has[64];
};
typedef KsLocalT<KB_CMAX, KB_LIN, uint8_t> KsLocal;          // per-thread local arrays of the regular kernel
typedef KsLocalT<KB_CBIG, KB_LBIG, uint16_t> KsLocalBig;     // global scratch slot of the big kernel
// The sorted rank_label element carries the candidate's depth in the upper half of idx, so TCmp does not load depth[a.idx] /
// depth[b.idx] from local memory twice per comparison (12.5 % of the scoring kernel's stall samples sat on that line,
// profiles/r01r_hot_lines_k3_k4.txt).  Candidate indices are < 512, depths fit 16 bits.
#define KS_RL_IDX(v) ((v) & 0xFFFFu)
#define KS_RL_PACK(f, depth) ((uint32_t)(f) | ((uint32_t)(depth) << 16))
struct KsTCmp {           // TCmp, read_label.cpp:475-485: |a-b| < 0.001 (double compare) -> shallower first, else by score
    __device__ bool operator()(const KmRl &a, const KmRl &b) const {
        if ((double)fabsf(__fsub_rn(a.score, b.score)) < 0.001) return (int)(a.idx >> 16) < (int)(b.idx >> 16);
        return a.score < b.score;
    }
};
template <typename PermT>
struct KsLinDepthDesc {   // CmpDepth over lineage entries (:159-167), sorting a permutation
    const uint16_t *l_depth;
    __device__ bool operator()(const PermT &a, const PermT &b) const { return (int)l_depth[a] > (int)l_depth[b]; }
};

// One read.  T: the thread's working arrays (local memory in the regular kernel, a global scratch slot in the big one).
template <int LIN, typename PermT, int BIG /* 0: regular kernel, 1: big, 2: huge */, typename TT>
__device__ __forceinline__ void ks_score_one(const KmScoreParams &P, const uint32_t r, TT &T) {
    kmat_read_result res = P.out[r];
    if (res.status != KMAT_ST_PENDING && !(BIG == 1 && res.status == KMAT_ST_PENDING_BIG) && !(BIG == 2 && res.status == KMAT_ST_PENDING_HUGE)) return;
    const KmCtxDev &X = P.C;
    const int C = (int)res.n_cand;
    const unsigned long long co = res.cand_off;
    const uint16_t cand16 = (uint16_t)res.cand_kmer_cnt;
    const int bin = res.bin_sel;
    const int model = X.n_models ? (int)X.model_of_cand[cand16] : -1;          // getReadLen(cand_kmer_cnt) -> model (:736-742)
    const bool useRandMod = model >= 0;
    bool hasHuman = false, bad_model = false;
    // ---- :748-802 per-taxid hit fraction, null-model cut-off, class maxima (order dependent)
    if (useRandMod) for (int c = 0; c < X.n_classes; c++) T.track_has[c] = 0;
    for (int f = 0; f < C; f++) {
        const kmat_pair e = P.cands[co + f];
        const uint32_t nid = e.tid, hits = __float_as_uint(e.score);
        const KmNodeA na = kb_nodeA(X, nid); const KmNodeB nb = kb_nodeB(X, nid);
        T.nid[f] = nid; T.tid[f] = na.tid; T.tin[f] = nb.tin; T.tout[f] = nb.tout;
        T.depth[f] = (uint16_t)(na.meta & KM_META_DEPTH_MASK);
        T.flags[f] = (uint8_t)(((na.meta & KM_META_HUMAN) ? 1 : 0) | ((na.meta & KM_META_PHIX) ? 2 : 0) | ((na.meta & KM_META_PLASMID) ? 4 : 0));
        hasHuman |= (na.meta & KM_META_HUMAN) != 0;
        T.score[f] = __fdiv_rn((float)hits, (float)cand16);                     // label_prob (:761)
        if (useRandMod) {
            const int32_t row = X.mrow[(size_t)model * X.n_nodes + nid];
            if (row < 0) { bad_model = true; continue; }                        // :773-778: the reference asserts here
            // val_vec[bin_sel]: bin_sel == nbins (GC 100 %) reads past the vector in the reference (:770); 0 here
            const float val = bin >= 0 && bin < X.nbins ? X.cut[(size_t)row * X.nbins + bin] : 0.0f;
            const float rp = __double2float_rn(__dadd_rn((double)val, 0.0001)); // :771
            const int cid = X.cls[row];
            T.cls[f] = (uint8_t)cid;
            if (!T.track_has[cid]) { T.track_has[cid] = 1; T.track_val[cid] = rp; }
            else T.track_val[cid] = rp < T.track_val[cid] ? T.track_val[cid] : rp;          // std::max(random_prob, track[cval])
            for (int ti = (int)X.class_ranknum[cid] - 1; ti >= 0; ti--) {                     // :787-790 / :795-798
                if (!T.track_has[ti]) { T.track_has[ti] = 1; T.track_val[ti] = 0.0f; }      // operator[] default-inserts 0
                T.track_val[cid] = T.track_val[cid] < T.track_val[ti] ? T.track_val[ti] : T.track_val[cid];
            }
        }
    }
    if (bad_model) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_FORMAT; res.n_cand = 0; P.out[r] = res; return; }
    // ---- :807-837 log-odds and sums in taxid_lst order
    bool fndPhiX = false;
    float log_sum = 0.0f, pos_log_sum = 0.0f, top_score = 0.0f, phiXscore = 0.0f;
    unsigned sig_hits = 0, pos_sig_hits = 0;
    for (int f = 0; f < C; f++) {
        float lo = T.score[f];
        if (useRandMod) {
            const float random_prob = T.track_val[T.cls[f]];
            const float denom = random_prob <= 0 ? 0.00001f : random_prob;      // :687
            lo = km_logf(__fdiv_rn(lo, denom));                                 // :688
            T.score[f] = lo;
        }
        log_sum = __fadd_rn(log_sum, lo);
        sig_hits++;
        if (lo > 0) { pos_sig_hits++; pos_log_sum = __fadd_rn(pos_log_sum, lo); }
        if (X.opt.phix_screen && (T.flags[f] & 2)) { phiXscore = lo; fndPhiX = true; }
        if (f == 0 || lo > top_score) top_score = lo;
    }
    if (X.opt.phix_screen && phiXscore >= top_score && fndPhiX) {              // :841-848
        res.status = KMAT_ST_PHIX; res.match = KMAT_DIRECT; res.tid = 32630u; res.score = phiXscore; res.n_cand = 0;
        P.out[r] = res;
        return;
    }
    float log_avg; unsigned use_sig_hits;
    const unsigned min_pos_examples = 3;
    if (pos_sig_hits > min_pos_examples) { use_sig_hits = pos_sig_hits; log_avg = __fdiv_rn(pos_log_sum, (float)pos_sig_hits); }
    else { use_sig_hits = sig_hits; log_avg = sig_hits > 0 ? __fdiv_rn(log_sum, (float)sig_hits) : 0.0f; }
    float log_std = 0.0f;
    for (int f = 0; f < C; f++) {                                              // :865-880
        const float sc = T.score[f];
        if (sc > 0 && pos_sig_hits > min_pos_examples) { const float v = __fsub_rn(log_avg, sc); log_std = __fadd_rn(log_std, __fmul_rn(v, v)); }
        if (pos_sig_hits <= min_pos_examples) { const float v = __fsub_rn(log_avg, sc); log_std = __fadd_rn(log_std, __fmul_rn(v, v)); }
    }
    const float stdev1 = use_sig_hits > 1 ? __fsqrt_rn(__fdiv_rn(log_std, (float)(use_sig_hits - 1))) : 0.0f;   // :881
    res.status = KMAT_ST_LABELED; res.log_avg = log_avg; res.stdev = stdev1;
    // rank_label = (tid, score [+ hbias*stdev for human tids]) in taxid_lst order, then sort(TCmp)   :882-893
    for (int f = 0; f < C; f++) {
        float sc = T.score[f];
        if (hasHuman && (T.flags[f] & 1)) sc = __fadd_rn(sc, __fmul_rn(X.opt.hbias, stdev1));
        T.rl[f].score = sc; T.rl[f].idx = KS_RL_PACK(f, T.depth[f]);
    }
    kmstd::sort(T.rl, C, KsTCmp());
    const float diff_thresh = __fmul_rn(stdev1, X.opt.sdiff);                  // :895
    // ---- findReadLabelVer2 (:284-419)
    int nlin = 0;
    bool plasmidTopHit = false; int savePlasmid = -1;
    unsigned lowest_depth = 0, highest_depth = 0;
    int lowest = -1, highest = -1; float lowest_score = 0;
    int lidx = -1; bool linDone = false, lin_overflow = false;
    for (int i = C - 1; i >= 0; --i) {                                          // :295-325
        const int ci = (int)KS_RL_IDX(T.rl[i].idx); const float sc = T.rl[i].score;
        const unsigned cdepth = T.depth[ci];
        if (sc >= top_score && (T.flags[ci] & 4)) { plasmidTopHit = true; savePlasmid = ci; }
        bool added = false;
        if (!linDone) {                                                         // addToCandLineage :225-262
            bool addNode = true;
            for (int q = 0; q < nlin; q++) {
                const unsigned chk = T.l_depth[q];
                if (chk > cdepth && !kb_is_anc(T.tin[ci], T.tout[ci], T.l_tin[q])) { addNode = false; break; }
                else if (chk < cdepth && !kb_is_anc(T.l_tin[q], T.l_tout[q], T.tin[ci])) { addNode = false; break; }
                else if (chk == cdepth) { addNode = false; break; }
            }
            if (addNode) {
                if (nlin >= LIN) lin_overflow = true;
                else {
                    T.l_tid[nlin] = T.tid[ci]; T.l_tin[nlin] = T.tin[ci]; T.l_tout[nlin] = T.tout[ci]; T.l_score[nlin] = sc;
                    T.l_depth[nlin] = (uint16_t)cdepth; T.l_nogood[nlin] = 0; nlin++;
                }
                added = true;
            }
        }
        if (!linDone && !added) { lidx = i; linDone = true; }
        else if (!linDone) {
            if (cdepth > lowest_depth || i == C - 1) { lowest = ci; lowest_score = sc; lowest_depth = cdepth; }
            if (cdepth < highest_depth || i == C - 1) { highest = ci; highest_depth = cdepth; }
        }
        if (linDone && sc < top_score) break;
    }
    const int add_lo = nlin; int add_hi = nlin;                                 // add_set = lineage entries [add_lo, add_hi)
    if (highest_depth != 0 && highest >= 0) {                                   // :327-343
        const KmNodeB hb = kb_nodeB(X, T.nid[highest]);
        for (uint32_t q = 0; q < hb.path_len; q++) {
            const uint32_t a = X.paths[hb.path_off + q];
            if (nlin >= LIN) { lin_overflow = true; break; }
            int fc = -1;
            for (int c2 = 0; c2 < C; c2++) if (T.nid[c2] == a) { fc = c2; break; }
            if (fc >= 0) {                                                      // all_cand_set holds the un-biased score
                T.l_tid[nlin] = T.tid[fc]; T.l_tin[nlin] = T.tin[fc]; T.l_tout[nlin] = T.tout[fc]; T.l_score[nlin] = T.score[fc];
                T.l_depth[nlin] = T.depth[fc];
            } else {
                const KmNodeA na = kb_nodeA(X, a); const KmNodeB nb = kb_nodeB(X, a);
                T.l_tid[nlin] = na.tid; T.l_tin[nlin] = nb.tin; T.l_tout[nlin] = nb.tout; T.l_score[nlin] = -10000.0f;
                T.l_depth[nlin] = (uint16_t)(na.meta & KM_META_DEPTH_MASK);
            }
            T.l_nogood[nlin] = 0; nlin++;
        }
        add_hi = nlin;
    }
    if (lin_overflow) {
        if (!BIG && P.big_qb) {                      // more lineage entries than the local arrays hold: the big kernel redoes this read
            const unsigned int q = atomicAdd(P.big_cnt + 1, 1u);
            if (q < KB_BIGQ) { P.big_qb[q] = r; return; }
        }
        res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; res.n_cand = 0; P.out[r] = res; return;
    }
    for (int q = 0; q < nlin; q++) T.l_perm[q] = (PermT)q;
    kmstd::sort(T.l_perm, nlin, KsLinDepthDesc<PermT>{T.l_depth});                     // cand_lin_vec sorted by depth desc :350-351
    bool any_nogood = false;
    for (int i = lidx; i >= 0; --i) {                                           // :355-362
        const int ci = (int)KS_RL_IDX(T.rl[i].idx); const float sc = T.rl[i].score;
        bool in_add = false;
        for (int q = add_lo; q < add_hi && !in_add; q++) in_add = T.l_tid[q] == T.tid[ci];
        if (in_add) continue;
        bool keep_going = true;                                                 // cmpCompLineage :264-282
        for (int z = 0; z < nlin; z++) {
            const int q = T.l_perm[z];
            if (kb_is_anc(T.l_tin[q], T.l_tout[q], T.tin[ci])) break;
            const float dlt = __fsub_rn(T.l_score[q], sc);
            if (T.l_score[q] != -10000.0f && dlt > diff_thresh) { keep_going = false; break; }
            if (dlt <= diff_thresh) { T.l_nogood[q] = 1; any_nogood = true; }
        }
        if (!keep_going) break;
    }
    uint32_t call_tid = 0; float call_score = 0; int match = KMAT_NOMATCH;
    uint32_t call_tin = 0, call_tout = 0; bool call_has_node = false;
    if (nlin == 0 && !any_nogood) match = KMAT_NOMATCH;
    else if (nlin != 0 && !any_nogood) {                                        // :366-368
        call_tid = T.tid[lowest]; call_score = lowest_score; match = KMAT_DIRECT;
        call_tin = T.tin[lowest]; call_tout = T.tout[lowest]; call_has_node = true;
    } else {                                                                    // :369-409
        float max_val = -10000.0f; int root = -1;
        for (int z = 0; z < nlin; z++) {
            const int q = T.l_perm[z];
            max_val = T.l_score[q] < max_val ? max_val : T.l_score[q];         // std::max(cand, max_val)
            bool ng = false;                                                    // no_good is a set of taxids
            for (int y = 0; y < nlin && !ng; y++) ng = T.l_nogood[y] && T.l_tid[y] == T.l_tid[q];
            if (!ng) { root = q; break; }
        }
        if (root < 0) { call_tid = 0; call_score = -1.0f; match = KMAT_LCA_ERROR; }
        else {
            match = KMAT_MULTI;
            bool in_all = false;
            for (int c2 = 0; c2 < C && !in_all; c2++) in_all = T.tid[c2] == T.l_tid[root];
            if (in_all && max_val < T.l_score[root]) { match = KMAT_PARTIAL; max_val = T.l_score[root]; }   // :400-406 (unreachable in practice)
            call_tid = T.l_tid[root]; call_score = max_val; call_tin = T.l_tin[root]; call_tout = T.l_tout[root]; call_has_node = true;
        }
    }
    if (plasmidTopHit && call_has_node && kb_is_anc(call_tin, call_tout, T.tin[savePlasmid])) call_tid = T.tid[savePlasmid];   // :410-416
    res.match = match;
    if (match == KMAT_DIRECT || match == KMAT_MULTI || match == KMAT_PARTIAL) { res.tid = call_tid; res.score = call_score; }
    else { res.tid = 0; res.score = 0; }                                        // best_guess stays (0,0), :839
    // ---- outputs: sorted rank_label overwrites the (nid, hits) hand-over in place; lineage on request
    if (P.cands && co + C <= P.cand_cap) for (int i = 0; i < C; i++) P.cands[co + i] = kmat_pair{T.tid[KS_RL_IDX(T.rl[i].idx)], T.rl[i].score};
    if (X.opt.want_lineage) {
        res.n_lin = (uint32_t)nlin;
        const unsigned long long lo2 = atomicAdd(P.lin_cursor, (unsigned long long)nlin);
        res.lin_off = lo2;
        if (P.lin && lo2 + nlin <= P.lin_cap) for (int q = 0; q < nlin; q++) P.lin[lo2 + q] = kmat_pair{T.l_tid[q], T.l_score[q]};
    }
    P.out[r] = res;
}

// A CTA takes KS_THREADS * KS_SORT_ROUNDS entries of the pending queue, counting-sorts them by candidate count in shared memory
// and scores them in that order, so the 32 reads of a warp have similar loop lengths (the insertion sort of rank_label and
// the ancestor scans ran with 7-16 of 32 lanes active before: 13.9 -> 9.4 ms per 10 M reads, profiles/r02a_variants.txt).
// Results do not depend on which thread scores a read.
// KS_SORT_ROUNDS: 4 for large passes; the ~1 M-read chunks of the host-buffer pipeline take 2 -- their grid is only a couple of
// waves of CTAs, and halving the window halves what the last, partly filled wave costs (e2e 128 -> 132 M reads/s,
// profiles/r02s_k4_window.txt) while the 10 M-read launch keeps the better sort (9.45 against 9.51 ms).
template <int KS_SORT_ROUNDS>
__global__ void __launch_bounds__(KS_THREADS) km_score_kernel(KmScoreParams P) {
    constexpr int KS_READS_PER_CTA = KS_THREADS * KS_SORT_ROUNDS;
    __shared__ uint32_t s_bin[KB_CMAX + 2];
    __shared__ uint32_t s_r[KS_READS_PER_CTA];
    const uint32_t n_q = P.pend_q ? (uint32_t)(*P.pass_cursor >> KB_PASS_SHIFT) : P.n_reads;
    const uint32_t base = blockIdx.x * KS_READS_PER_CTA;
    if (base >= n_q) return;                                   // the whole CTA leaves together
    for (int i = threadIdx.x; i < KB_CMAX + 2; i += KS_THREADS) s_bin[i] = 0;
    __syncthreads();
    uint32_t rr[KS_SORT_ROUNDS], cc[KS_SORT_ROUNDS], pos[KS_SORT_ROUNDS];
#pragma unroll
    for (int k = 0; k < KS_SORT_ROUNDS; k++) {
        const uint32_t q = base + k * KS_THREADS + threadIdx.x;
        rr[k] = KMAT_NONE; cc[k] = 0; pos[k] = 0;
        if (q < n_q) {
            const uint32_t r = P.pend_q ? P.pend_q[q] : q;
            if (P.out[r].status == KMAT_ST_PENDING) {
                rr[k] = r; cc[k] = min(P.out[r].n_cand, (uint32_t)KB_CMAX);
                pos[k] = atomicAdd(&s_bin[cc[k]], 1u);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                    // exclusive scan, most candidates first
        uint32_t run = 0;
        for (int i = KB_CMAX; i >= 0; i--) { const uint32_t t = s_bin[i]; s_bin[i] = run; run += t; }
        s_bin[KB_CMAX + 1] = run;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KS_SORT_ROUNDS; k++) if (rr[k] != KMAT_NONE) s_r[s_bin[cc[k]] + pos[k]] = rr[k];
    __syncthreads();
    const uint32_t total = s_bin[KB_CMAX + 1];
    KsLocal T;
    for (uint32_t i = threadIdx.x; i < total; i += KS_THREADS) ks_score_one<KB_LIN, uint8_t, 0>(P, s_r[i], T);
}
// The reads of big_qb (more than KB_CMAX candidates, or a lineage the regular kernel could not hold): same code, working
// arrays in a global scratch slot per thread.
__global__ void __launch_bounds__(32) km_score_big_kernel(KmScoreParams P) {
    const uint32_t slot = blockIdx.x * 32 + threadIdx.x, n_threads = gridDim.x * 32;
    if (slot >= P.big_threads4) return;
    KsLocalBig &T = *(KsLocalBig *)(P.big_scratch4 + (size_t)slot * sizeof(KsLocalBig));
    const uint32_t n = min(P.big_cnt[1], (unsigned int)KB_BIGQ);
    for (uint32_t q = slot; q < n; q += min(n_threads, P.big_threads4)) ks_score_one<KB_LBIG, uint16_t, 1>(P, P.big_qb[q], T);
}

__global__ void km_cursor_roll_kernel(unsigned long long *cand_cursor, unsigned long long *pass_cursor) {
    *cand_cursor += *pass_cursor & KB_PASS_MASK;
    *pass_cursor = 0ull;
}

// ---------------------------------------------------------------------------------------------
// K3 for the reads of big_qa (more than KB_CMAX candidate taxids): one WARP per read, the candidate set (up to KB_CBIG)
// in shared memory, the per-position candidate sets (KB_CBIG bits each) in a global scratch slot of the warp.  Same
// steps as km_cand_kernel.  Not rare for long reads (a 10 kbp read with no relative in the table still collects ~17
// chance hits all over the taxonomy) and for the low-complexity reads of the null-model generator.
// ---------------------------------------------------------------------------------------------
#define KB_BIGW (KB_CBIG / 64)
#define KBG_WARPS 8
struct KbBigW {                                   // one warp's candidate set: 12 KB of shared memory (2 CTAs of 8 warps per SM)
    uint32_t nid[KB_CBIG], leaf[KB_CBIG], key[KB_CBIG], hits[KB_CBIG], tid[KB_CBIG], spec[KB_CBIG];   // leaf: count | rank code << 30 once the node data is in
    unsigned long long *anc;                      // [KB_CBIG][KB_BIGW] lineage sets of the qualifying members, in the warp's global scratch slot
};
// find-or-append `v` (warp-uniform); returns the index or -1 when KB_CBIG is exceeded
__device__ __forceinline__ int kbg_find_or_add(KbBigW &W, int &C, uint32_t v, int lane) {
    for (int t = 0; t < C; t += 32) {
        const uint32_t f = __ballot_sync(KM_FULL, t + lane < C && W.nid[t + lane] == v);
        if (f) return t + __ffs(f) - 1;
    }
    if (C >= KB_CBIG) return -1;
    const int idx = C++;
    if (lane == 0) { W.nid[idx] = v; W.leaf[idx] = 0; W.key[idx] = 0xFFFFFFFFu; W.hits[idx] = 0; }
    if (lane < KB_BIGW) W.anc[(size_t)idx * KB_BIGW + lane] = 0ull;
    __syncwarp();
    return idx;
}
// more than KB_CBIG candidates: on to km_cand_huge_kernel (all lanes call; lane 0 acts)
__device__ __forceinline__ void kbg_defer_huge(const KmScoreParams &P, uint32_t r, kmat_read_result &res, int lane) {
    if (lane != 0) return;
    res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED;
    if (P.huge_qa) {
        const unsigned int q = atomicAdd(P.big_cnt + 2, 1u);
        if (q < KB_HUGEQ) { P.huge_qa[q] = r; res.status = KMAT_ST_DEFERRED; res.err = 0; }
    }
    P.out[r] = res;
}
__global__ void __launch_bounds__(KBG_WARPS * 32) km_cand_big_kernel(KmScoreParams P) {
    extern __shared__ __align__(16) unsigned char kbg_smem[];
    const KmCtxDev &X = P.C;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.x * KBG_WARPS + wib, n_warps = min((uint32_t)gridDim.x * KBG_WARPS, P.big_threads3);
    if (warp_global >= P.big_threads3) return;
    KbBigW &W = *(KbBigW *)(kbg_smem + (size_t)wib * sizeof(KbBigW));
    unsigned long long *gmask = (unsigned long long *)P.big_scratch3 + (size_t)warp_global * ((size_t)P.big_np_cap + KB_CBIG) * KB_BIGW;
    if (lane == 0) W.anc = gmask + (size_t)P.big_np_cap * KB_BIGW;
    __syncwarp();
    const int k = X.db.kmer_len;
    const bool permissive = X.opt.permissive != 0;
    const uint32_t nq = min(P.big_cnt[0], (unsigned int)KB_BIGQ);
    const unsigned long long cand_base = *P.cand_cursor;
    for (uint32_t q = warp_global; q < nq; q += n_warps) {
        const uint32_t r = P.big_qa[q];
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        const int2 hd = P.hdr[r];
        kmat_read_result res;
        memset(&res, 0, sizeof res);
        res.valid_kmers = hd.x; res.bin_sel = hd.y; res.match = KMAT_NOMATCH;
        if (np <= 0 || (uint32_t)np > P.big_np_cap || np > 0xFFFF) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; continue; }
        int C = 0, cand_cnt = 0, fnd_cnt = 0, err = 0;
        bool overflow = false;
        // ---- per position: members in insertion order (kb_chunk with the set in shared memory)
        const int nch = (np + 31) >> 5;
        for (int c = 0; c < nch && !overflow; c++) {
            const int p = (c << 5) + lane;
            const uint32_t hw = p < np ? __ldg(P.hit + off + p) : KM_HIT_INVALID;
            unsigned long long *mym = gmask + (size_t)p * KB_BIGW;
            if (p < np) for (int w = 0; w < KB_BIGW; w++) mym[w] = 0ull;
            uint32_t a = 0, b = 0, v0 = KMAT_NONE;
            const uint32_t *rec = nullptr;
            if (hw != KM_HIT_INVALID) {
                cand_cnt++;
                if (hw != KM_HIT_MISS) {
                    if (!(hw & KM_HIT_LIST)) {
                        const uint32_t e = hw < X.n_sid ? __ldg(X.sid2nid + hw) : KMAT_NONE;
                        if (e == KMAT_NONE) err = KMAT_ERR_BAD_TAXID;
                        else if (!(e & KB_SID_DROP)) {
                            v0 = ((e & KB_SID_HUMAN) && !X.opt.rkmer_mode) ? X.nid_human : (e & KB_SID_NIDMASK); a = 1;
                            if (permissive) b = (kb_nodeA(X, v0).meta & KM_META_DEPTH_MASK) ? 1 : 0;
                        }
                    } else {
                        rec = kb_rec_of(X, hw);
                        const uint32_t h = __ldg(rec);
                        if (h == KR_ERR_BAD) { err = KMAT_ERR_BAD_TAXID; rec = nullptr; }
                        else { a = h & 0xFFFFu; if (permissive) { b = rec[1]; rec += 2; } else rec += 1; }
                    }
                }
            }
            if (a) fnd_cnt++;
            uint32_t seqno = 0, bi = 0, pq = 0, poff = 0, plen = 0;
            for (;;) {
                uint32_t val = KMAT_NONE;
                if (seqno < a) val = rec ? rec[seqno] : v0;
                else if (permissive) {
                    while (bi < b && pq >= plen) {
                        const KmNodeB nb = kb_nodeB(X, rec ? rec[a + bi] : v0);
                        poff = nb.path_off; plen = nb.path_len; pq = 0; bi++;
                    }
                    if (pq < plen) val = X.paths[poff + pq++];
                }
                uint32_t pending = __ballot_sync(KM_FULL, val != KMAT_NONE);
                if (!pending) break;
                while (pending) {
                    const int leader = __ffs(pending) - 1;
                    const uint32_t v = __shfl_sync(KM_FULL, val, leader);
                    const uint32_t grp = __ballot_sync(KM_FULL, val == v);
                    const int idx = kbg_find_or_add(W, C, v, lane);
                    if (idx < 0) { overflow = true; break; }
                    if (lane == 0) {
                        const uint32_t key = ((uint32_t)((c << 5) + leader) << 16) | min(seqno, 0xFFFFu);
                        W.key[idx] = min(W.key[idx], key); W.leaf[idx] += __popc(grp);
                    }
                    if (val == v) mym[idx >> 6] |= 1ull << (idx & 63);
                    pending &= ~grp;
                }
                if (overflow) break;
                seqno++;
            }
            __syncwarp();
        }
        cand_cnt = km_warp_sum(cand_cnt); fnd_cnt = km_warp_sum(fnd_cnt);
        err = __reduce_max_sync(KM_FULL, err < 0 ? -err : 0);
        overflow = __any_sync(KM_FULL, overflow);
        if (err) { res.status = KMAT_ST_ERROR; res.err = -err; if (lane == 0) P.out[r] = res; continue; }
        if (overflow) { kbg_defer_huge(P, r, res, lane); continue; }
        const int C1 = C;
        if (C1 == 0) { res.status = KMAT_ST_NODBHITS; res.n1 = len; res.n2 = k; if (lane == 0) P.out[r] = res; continue; }
        const uint16_t cand16 = (uint16_t)cand_cnt;
        res.cand_kmer_cnt = cand16;
        if (fnd_cnt < X.opt.min_fnd_kmer || (int)cand16 < X.opt.min_kmer) {
            res.status = KMAT_ST_SILENT; res.match = KMAT_NOMATCH; res.tid = 0; res.score = -1.0f;
            if (lane == 0) P.out[r] = res;
            continue;
        }
        if (!permissive) {
            // ---- node data of the members; representative strain per species -> which members bring their lineage
            for (int i = lane; i < C1; i += 32) {
                const KmNodeA na = kb_nodeA(X, W.nid[i]);
                W.tid[i] = na.tid; W.spec[i] = na.species_anc; W.leaf[i] = (W.leaf[i] & 0x3FFFFFFFu) | (((na.meta >> KM_META_RANK_SHIFT) & 3u) << 30);
            }
            __syncwarp();
            for (int i = lane; i < C1; i += 32) {
                const bool strain = (W.leaf[i] >> 30) == 1;
                bool qual = !strain;
                if (strain && W.spec[i] != KMAT_NONE) {
                    bool beaten = false;                    // same rank code in the top bits of both: comparing the words compares the counts
                    for (int j = 0; j < C1 && !beaten; j++)
                        beaten = (W.leaf[j] >> 30) == 1 && W.spec[j] == W.spec[i] && (W.leaf[j] > W.leaf[i] || (W.leaf[j] == W.leaf[i] && W.tid[j] < W.tid[i]));
                    qual = !beaten;
                }
                W.hits[i] = qual ? 1u : 0u;                 // "still to expand" marker until the counts are taken
            }
            __syncwarp();
            // ---- lineage expansion in (first position, taxid) order: the appended ancestors take the next indices
            for (;;) {
                unsigned long long mine = ~0ull;
                for (int i = lane; i < C1; i += 32)
                    if (W.hits[i]) { const unsigned long long kk = ((unsigned long long)(W.key[i] >> 16) << 32) | W.tid[i]; if (kk < mine) mine = kk; }
                const unsigned long long best = kb_warp_min64(mine);
                if (best == ~0ull) break;
                int s2 = -1;
                for (int i = lane; i < C1; i += 32)
                    if (W.hits[i] && ((((unsigned long long)(W.key[i] >> 16) << 32) | W.tid[i]) == best)) s2 = i;
                s2 = __reduce_max_sync(KM_FULL, s2);
                if (lane == 0) W.hits[s2] = 0;
                const KmNodeB nb = kb_nodeB(X, W.nid[s2]);
                for (uint32_t c0 = 0; c0 < nb.path_len && !overflow; c0 += 32) {
                    const uint32_t av = c0 + lane < nb.path_len ? X.paths[nb.path_off + c0 + lane] : KMAT_NONE;
                    const int cnt = min(32u, nb.path_len - c0);
                    for (int z = 0; z < cnt; z++) {
                        const int idx = kbg_find_or_add(W, C, __shfl_sync(KM_FULL, av, z), lane);
                        if (idx < 0) { overflow = true; break; }
                        if (lane == 0) W.anc[(size_t)s2 * KB_BIGW + (idx >> 6)] |= 1ull << (idx & 63);
                    }
                }
                __syncwarp();
                if (overflow) break;
            }
            if (overflow) { kbg_defer_huge(P, r, res, lane); continue; }
        }
        for (int i = lane; i < C; i += 32) W.hits[i] = 0;
        __syncwarp();
        // ---- expanded position sets and hits per candidate in one sweep: lane = position, 32 positions per step.  The
        //      expansion is a closure (an ancestor's lineage is part of its descendant's), so the order of the members does
        //      not matter.  Counting: for every candidate bit present in ANY of the 32 sets one ballot gives its count --
        //      the sets of neighbouring positions are nearly identical, shared-memory atomics would serialise 32-fold.
        for (int p0 = 0; p0 < np; p0 += 32) {
            const int p = p0 + lane;
            unsigned long long m[KB_BIGW];
            const unsigned long long *mym = gmask + (size_t)p * KB_BIGW;
#pragma unroll
            for (int w = 0; w < KB_BIGW; w++) m[w] = p < np ? mym[w] : 0ull;
            if (!permissive) {
                unsigned long long o[KB_BIGW];
#pragma unroll
                for (int w = 0; w < KB_BIGW; w++) o[w] = m[w];
#pragma unroll
                for (int w = 0; w < KB_BIGW; w++) {
                    unsigned long long v = o[w];
                    while (v) {
                        const int j = w * 64 + __ffsll((long long)v) - 1; v &= v - 1;
                        if (j < C1) {
#pragma unroll
                            for (int w2 = 0; w2 < KB_BIGW; w2++) m[w2] |= W.anc[(size_t)j * KB_BIGW + w2];
                        }
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < KB_BIGW; w++) {
                unsigned long long u = km_warp_or64(m[w]);
                while (u) {
                    const int bit = __ffsll((long long)u) - 1; u &= u - 1;
                    const uint32_t cnt = __popc(__ballot_sync(KM_FULL, (m[w] >> bit) & 1));
                    if (lane == 0) W.hits[w * 64 + bit] += cnt;
                }
            }
        }
        __syncwarp();
        // ---- hand over in taxid_lst order: members by first appearance (rank of the key), then the appended ancestors
        unsigned long long co = 0;
        if (lane == 0) co = atomicAdd(P.pass_cursor, (unsigned long long)C);
        co = cand_base + (kb_shfl64(co, 0) & KB_PASS_MASK);
        res.status = KMAT_ST_PENDING_BIG; res.n_cand = (uint32_t)C; res.cand_off = co;
        if (P.cands && co + C <= P.cand_cap) {
            for (int i = lane; i < C; i += 32) {
                uint32_t ord = (uint32_t)i;
                if (i < C1) { ord = 0; const uint32_t ki = W.key[i]; for (int j = 0; j < C1; j++) ord += W.key[j] < ki; }
                P.cands[co + ord] = kmat_pair{W.nid[i], __uint_as_float(W.hits[i])};
            }
        } else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_OVERFLOW; }
        if (lane == 0) {
            if (res.status == KMAT_ST_PENDING_BIG) {
                const unsigned int q2 = atomicAdd(P.big_cnt + 1, 1u);
                if (q2 < KB_BIGQ) P.big_qb[q2] = r; else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; }
            }
            P.out[r] = res;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Last resort: reads with more than KB_CBIG candidate taxids (the reference has no bound: std::set / std::map per read,
// read_label.cpp:698-726).  One warp per read, every working array in a global scratch slot, memory LINEAR in the candidate
// count: a taxid -> index hash instead of a scan, the lineage of a qualifying member as a list of candidate indices instead of
// a bit set, and the per-candidate position counts by walking the positions once more with a "last position that counted me"
// stamp per candidate instead of per-position bit sets.  Same steps and the same results as km_cand_kernel.
// ---------------------------------------------------------------------------------------------
struct KbHugeW {
    uint32_t *nid, *leaf, *key, *hits, *tid, *spec, *stamp, *lin_off, *hash;
    uint16_t *lin_len, *lin;
};
#define KBH_SLOT3_BYTES ((size_t)KB_CHUGE * (8 * 4 + 2) + ((size_t)4 << KB_HHASH_BITS) + (size_t)KB_HLIN * 2)
__device__ __forceinline__ KbHugeW kbh_carve(unsigned char *base) {
    KbHugeW W;
    uint32_t *u = (uint32_t *)base;
    W.nid = u; W.leaf = u + KB_CHUGE; W.key = u + 2 * KB_CHUGE; W.hits = u + 3 * KB_CHUGE; W.tid = u + 4 * KB_CHUGE; W.spec = u + 5 * KB_CHUGE;
    W.stamp = u + 6 * KB_CHUGE; W.lin_off = u + 7 * KB_CHUGE; W.hash = u + 8 * KB_CHUGE;
    W.lin_len = (uint16_t *)(W.hash + ((size_t)1 << KB_HHASH_BITS));
    W.lin = W.lin_len + KB_CHUGE;
    return W;
}
__device__ __forceinline__ uint32_t kbh_hash(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - KB_HHASH_BITS); }
// index of `v` among the candidates or -1 (any lane, on its own)
__device__ __forceinline__ int kbh_find(const KbHugeW &W, uint32_t v) {
    uint32_t h = kbh_hash(v);
    for (;;) {
        const uint32_t s = W.hash[h];
        if (!s) return -1;
        if (W.nid[s - 1] == v) return (int)s - 1;
        h = (h + 1) & ((1u << KB_HHASH_BITS) - 1);
    }
}
// find-or-append `v` (warp-uniform; all lanes call, lane 0 works); -1 when KB_CHUGE is exceeded
__device__ __forceinline__ int kbh_find_or_add(const KbHugeW &W, int &C, uint32_t v, int lane) {
    int idx = -1;
    if (lane == 0) {
        uint32_t h = kbh_hash(v);
        for (;;) {
            const uint32_t s = W.hash[h];
            if (!s) break;
            if (W.nid[s - 1] == v) { idx = (int)s - 1; break; }
            h = (h + 1) & ((1u << KB_HHASH_BITS) - 1);
        }
        if (idx < 0 && C < KB_CHUGE) {
            idx = C; W.hash[h] = (uint32_t)C + 1u;
            W.nid[C] = v; W.leaf[C] = 0; W.key[C] = 0xFFFFFFFFu; W.hits[C] = 0; W.stamp[C] = 0; W.lin_len[C] = 0; W.lin_off[C] = 0;
        }
    }
    idx = __shfl_sync(KM_FULL, idx, 0);
    if (idx == C) C++;
    __syncwarp();
    return idx;
}
// position `stamp` - 1 contains candidate idx: count it once
__device__ __forceinline__ void kbh_mark(const KbHugeW &W, int idx, uint32_t stamp) {
    if (atomicExch(W.stamp + idx, stamp) != stamp) atomicAdd(W.hits + idx, 1u);
}
__global__ void __launch_bounds__(128) km_cand_huge_kernel(KmScoreParams P) {
    const KmCtxDev &X = P.C;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nq = min(P.big_cnt[2], (unsigned int)KB_HUGEQ);
    if (warp_global >= nq) return;
    const KbHugeW W = kbh_carve(P.huge_scratch3 + (size_t)warp_global * KBH_SLOT3_BYTES);
    const int k = X.db.kmer_len;
    const bool permissive = X.opt.permissive != 0;
    const unsigned long long cand_base = *P.cand_cursor;
    for (uint32_t q = warp_global; q < nq; q += n_warps) {
        const uint32_t r = P.huge_qa[q];
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        const int2 hd = P.hdr[r];
        kmat_read_result res;
        memset(&res, 0, sizeof res);
        res.valid_kmers = hd.x; res.bin_sel = hd.y; res.match = KMAT_NOMATCH;
        if (np <= 0 || np > 0xFFFF) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; continue; }
        for (uint32_t i = lane; i < (1u << KB_HHASH_BITS); i += 32) W.hash[i] = 0;
        __syncwarp();
        int C = 0, cand_cnt = 0, fnd_cnt = 0, err = 0;
        bool overflow = false;
        // ---- per position: members in insertion order (as kb_chunk; the position sets are not kept)
        const int nch = (np + 31) >> 5;
        for (int c = 0; c < nch && !overflow; c++) {
            const int p = (c << 5) + lane;
            const uint32_t hw = p < np ? __ldg(P.hit + off + p) : KM_HIT_INVALID;
            uint32_t a = 0, b = 0, v0 = KMAT_NONE;
            const uint32_t *rec = nullptr;
            if (hw != KM_HIT_INVALID) {
                cand_cnt++;
                if (hw != KM_HIT_MISS) {
                    if (!(hw & KM_HIT_LIST)) {
                        const uint32_t e = hw < X.n_sid ? __ldg(X.sid2nid + hw) : KMAT_NONE;
                        if (e == KMAT_NONE) err = KMAT_ERR_BAD_TAXID;
                        else if (!(e & KB_SID_DROP)) {
                            v0 = ((e & KB_SID_HUMAN) && !X.opt.rkmer_mode) ? X.nid_human : (e & KB_SID_NIDMASK); a = 1;
                            if (permissive) b = (kb_nodeA(X, v0).meta & KM_META_DEPTH_MASK) ? 1 : 0;
                        }
                    } else {
                        rec = kb_rec_of(X, hw);
                        const uint32_t h = __ldg(rec);
                        if (h == KR_ERR_BAD) { err = KMAT_ERR_BAD_TAXID; rec = nullptr; }
                        else { a = h & 0xFFFFu; if (permissive) { b = rec[1]; rec += 2; } else rec += 1; }
                    }
                }
            }
            if (a) fnd_cnt++;
            uint32_t seqno = 0, bi = 0, pq = 0, poff = 0, plen = 0;
            for (;;) {
                uint32_t val = KMAT_NONE;
                if (seqno < a) val = rec ? rec[seqno] : v0;
                else if (permissive) {
                    while (bi < b && pq >= plen) {
                        const KmNodeB nb = kb_nodeB(X, rec ? rec[a + bi] : v0);
                        poff = nb.path_off; plen = nb.path_len; pq = 0; bi++;
                    }
                    if (pq < plen) val = X.paths[poff + pq++];
                }
                uint32_t pending = __ballot_sync(KM_FULL, val != KMAT_NONE);
                if (!pending) break;
                while (pending) {
                    const int leader = __ffs(pending) - 1;
                    const uint32_t v = __shfl_sync(KM_FULL, val, leader);
                    const uint32_t grp = __ballot_sync(KM_FULL, val == v);
                    const int idx = kbh_find_or_add(W, C, v, lane);
                    if (idx < 0) { overflow = true; break; }
                    if (lane == 0) {
                        const uint32_t key = ((uint32_t)((c << 5) + leader) << 16) | min(seqno, 0xFFFFu);
                        W.key[idx] = min(W.key[idx], key); W.leaf[idx] += __popc(grp);
                    }
                    pending &= ~grp;
                }
                if (overflow) break;
                seqno++;
            }
            __syncwarp();
        }
        cand_cnt = km_warp_sum(cand_cnt); fnd_cnt = km_warp_sum(fnd_cnt);
        err = __reduce_max_sync(KM_FULL, err < 0 ? -err : 0);
        overflow = __any_sync(KM_FULL, overflow);
        if (err) { res.status = KMAT_ST_ERROR; res.err = -err; if (lane == 0) P.out[r] = res; continue; }
        if (overflow) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; continue; }
        const int C1 = C;
        if (C1 == 0) { res.status = KMAT_ST_NODBHITS; res.n1 = len; res.n2 = k; if (lane == 0) P.out[r] = res; continue; }
        const uint16_t cand16 = (uint16_t)cand_cnt;
        res.cand_kmer_cnt = cand16;
        if (fnd_cnt < X.opt.min_fnd_kmer || (int)cand16 < X.opt.min_kmer) {
            res.status = KMAT_ST_SILENT; res.match = KMAT_NOMATCH; res.tid = 0; res.score = -1.0f;
            if (lane == 0) P.out[r] = res;
            continue;
        }
        if (!permissive) {
            // ---- node data of the members; representative strain per species -> which members bring their lineage
            for (int i = lane; i < C1; i += 32) {
                const KmNodeA na = kb_nodeA(X, W.nid[i]);
                W.tid[i] = na.tid; W.spec[i] = na.species_anc; W.leaf[i] = (W.leaf[i] & 0x3FFFFFFFu) | (((na.meta >> KM_META_RANK_SHIFT) & 3u) << 30);
            }
            __syncwarp();
            for (int i = lane; i < C1; i += 32) {
                const bool strain = (W.leaf[i] >> 30) == 1;
                bool qual = !strain;
                if (strain && W.spec[i] != KMAT_NONE) {
                    bool beaten = false;
                    for (int j = 0; j < C1 && !beaten; j++)
                        beaten = (W.leaf[j] >> 30) == 1 && W.spec[j] == W.spec[i] && (W.leaf[j] > W.leaf[i] || (W.leaf[j] == W.leaf[i] && W.tid[j] < W.tid[i]));
                    qual = !beaten;
                }
                W.hits[i] = qual ? 1u : 0u;                 // "still to expand" marker until the counts are taken
            }
            __syncwarp();
            // ---- lineage expansion in (first position, taxid) order: the appended ancestors take the next indices; the
            //      lineage of a member is kept as the list of its ancestors' candidate indices
            uint32_t lin_used = 0;
            for (;;) {
                unsigned long long mine = ~0ull;
                for (int i = lane; i < C1; i += 32)
                    if (W.hits[i]) { const unsigned long long kk = ((unsigned long long)(W.key[i] >> 16) << 32) | W.tid[i]; if (kk < mine) mine = kk; }
                const unsigned long long best = kb_warp_min64(mine);
                if (best == ~0ull) break;
                int s2 = -1;
                for (int i = lane; i < C1; i += 32)
                    if (W.hits[i] && ((((unsigned long long)(W.key[i] >> 16) << 32) | W.tid[i]) == best)) s2 = i;
                s2 = __reduce_max_sync(KM_FULL, s2);
                const KmNodeB nb = kb_nodeB(X, W.nid[s2]);
                if (lin_used + nb.path_len > KB_HLIN || nb.path_len > 0xFFFFu) { overflow = true; break; }
                if (lane == 0) { W.hits[s2] = 0; W.lin_off[s2] = lin_used; W.lin_len[s2] = (uint16_t)nb.path_len; }
                for (uint32_t c0 = 0; c0 < nb.path_len && !overflow; c0 += 32) {
                    const uint32_t av = c0 + lane < nb.path_len ? X.paths[nb.path_off + c0 + lane] : KMAT_NONE;
                    const int cnt = min(32u, nb.path_len - c0);
                    for (int z = 0; z < cnt; z++) {
                        const int idx = kbh_find_or_add(W, C, __shfl_sync(KM_FULL, av, z), lane);
                        if (idx < 0) { overflow = true; break; }
                        if (lane == 0) W.lin[lin_used + c0 + z] = (uint16_t)idx;
                    }
                }
                lin_used += nb.path_len;
                __syncwarp();
                if (overflow) break;
            }
            if (overflow) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; continue; }
        }
        for (int i = lane; i < C; i += 32) { W.hits[i] = 0; W.stamp[i] = 0; }
        __syncwarp();
        // ---- hits per candidate = positions whose expanded set holds it: one position after the other, the lanes over its
        //      members (and their lineages); stamp[idx] = last position that counted idx
        for (int p = 0; p < np; p++) {
            const uint32_t hw = __ldg(P.hit + off + p), stamp = (uint32_t)p + 1u;
            if (hw == KM_HIT_INVALID || hw == KM_HIT_MISS) continue;                       // warp-uniform
            uint32_t a = 0, b = 0, v0 = KMAT_NONE;
            const uint32_t *rec = nullptr;
            if (!(hw & KM_HIT_LIST)) {
                const uint32_t e = __ldg(X.sid2nid + hw);                                  // checked in the first walk
                if (e & KB_SID_DROP) continue;
                v0 = ((e & KB_SID_HUMAN) && !X.opt.rkmer_mode) ? X.nid_human : (e & KB_SID_NIDMASK); a = 1;
                if (permissive) b = (kb_nodeA(X, v0).meta & KM_META_DEPTH_MASK) ? 1 : 0;
            } else {
                rec = kb_rec_of(X, hw);
                a = __ldg(rec) & 0xFFFFu;
                if (permissive) { b = rec[1]; rec += 2; } else rec += 1;
            }
            for (uint32_t m = lane; m < a; m += 32) {
                const int idx = kbh_find(W, rec ? rec[m] : v0);
                if (idx < 0) continue;                                                     // cannot happen: the first walk added every member
                kbh_mark(W, idx, stamp);
                if (!permissive && idx < C1) {
                    const uint32_t lo = W.lin_off[idx], ll = W.lin_len[idx];
                    for (uint32_t z = 0; z < ll; z++) kbh_mark(W, (int)W.lin[lo + z], stamp);
                }
            }
            for (uint32_t bi = 0; bi < b; bi++) {                                          // permissive: the root paths of the b owners
                const KmNodeB nb = kb_nodeB(X, rec ? rec[a + bi] : v0);
                for (uint32_t z = lane; z < nb.path_len; z += 32) { const int idx = kbh_find(W, X.paths[nb.path_off + z]); if (idx >= 0) kbh_mark(W, idx, stamp); }
            }
            __syncwarp();
        }
        __syncwarp();
        // ---- hand over in taxid_lst order: members by first appearance (rank of the key), then the appended ancestors
        unsigned long long co = 0;
        if (lane == 0) co = atomicAdd(P.pass_cursor, (unsigned long long)C);
        co = cand_base + (kb_shfl64(co, 0) & KB_PASS_MASK);
        res.status = KMAT_ST_PENDING_HUGE; res.n_cand = (uint32_t)C; res.cand_off = co;
        if (P.cands && co + C <= P.cand_cap) {
            for (int i = lane; i < C; i += 32) {
                uint32_t ord = (uint32_t)i;
                if (i < C1) { ord = 0; const uint32_t ki = W.key[i]; for (int j = 0; j < C1; j++) ord += W.key[j] < ki; }
                P.cands[co + ord] = kmat_pair{W.nid[i], __uint_as_float(W.hits[i])};
            }
        } else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_OVERFLOW; }
        if (lane == 0) {
            if (res.status == KMAT_ST_PENDING_HUGE) {
                const unsigned int q2 = atomicAdd(P.big_cnt + 3, 1u);
                if (q2 < KB_HUGEQ) P.huge_qb[q2] = r; else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; }
            }
            P.out[r] = res;
        }
        __syncwarp();
    }
}
// The working arrays of ks_score_one for a read of up to KB_CHUGE candidates: views into a global scratch slot
struct KsHuge {
    uint32_t *nid, *tid, *tin, *tout; float *score; uint16_t *depth; uint8_t *flags, *cls; KmRl *rl;
    uint32_t *l_tid, *l_tin, *l_tout; float *l_score; uint16_t *l_depth; uint8_t *l_nogood; uint16_t *l_perm;
    float *track_val; uint8_t *track_has;
};
#define KBH_SLOT4_BYTES ((size_t)KB_CHUGE * (4 * 5 + 8 + 2 + 1 + 1) + (size_t)KB_LBIG * (4 * 4 + 2 + 2 + 2) + 64 * 4 + 64)
__global__ void __launch_bounds__(32) km_score_huge_kernel(KmScoreParams P) {
    const uint32_t slot = blockIdx.x * 32 + threadIdx.x, n_threads = gridDim.x * 32;
    const uint32_t n = min(P.big_cnt[3], (unsigned int)KB_HUGEQ);
    if (slot >= n) return;
    unsigned char *b = P.huge_scratch4 + (size_t)slot * KBH_SLOT4_BYTES;
    KsHuge T;
    T.nid = (uint32_t *)b; T.tid = T.nid + KB_CHUGE; T.tin = T.tid + KB_CHUGE; T.tout = T.tin + KB_CHUGE; T.score = (float *)(T.tout + KB_CHUGE);
    T.rl = (KmRl *)(T.score + KB_CHUGE);
    T.l_tid = (uint32_t *)(T.rl + KB_CHUGE); T.l_tin = T.l_tid + KB_LBIG; T.l_tout = T.l_tin + KB_LBIG; T.l_score = (float *)(T.l_tout + KB_LBIG);
    T.track_val = T.l_score + KB_LBIG;
    T.depth = (uint16_t *)(T.track_val + 64); T.l_depth = T.depth + KB_CHUGE; T.l_perm = T.l_depth + KB_LBIG;
    T.flags = (uint8_t *)(T.l_perm + KB_LBIG); T.cls = T.flags + KB_CHUGE; T.l_nogood = T.cls + KB_CHUGE; T.track_has = T.l_nogood + KB_LBIG;
    for (uint32_t q = slot; q < n; q += n_threads) ks_score_one<KB_LBIG, uint16_t, 2>(P, P.huge_qb[q], T);
}

// ---------------------------------------------------------------------------------------------
// kmat_ctx
// ---------------------------------------------------------------------------------------------
struct KmShardState;
static void km_shard_free(KmShardState *);
struct kmat_ctx {
    const kmat_db *db = nullptr;
    KmShardState *shard = nullptr;           // DB-sharded mode buffers (kmat_shard.cuh), allocated on first use
    int device = 0;
    kmat_opts opt{};
    cudaStream_t stream = nullptr;
    // device tables
    KmNodeA *d_nodeA = nullptr; KmNodeB *d_nodeB = nullptr; uint32_t *d_paths = nullptr, *d_prune = nullptr, *d_sid2nid = nullptr;
    int16_t *d_model_of_cand = nullptr; int32_t *d_mrow = nullptr; float *d_cut = nullptr; uint8_t *d_cls = nullptr;
    KmHostCtx h;
    // batch buffers (grown on demand).  Host-buffer batches are cut into chunks that alternate between two slots so
    // that the H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i.
    struct Slot {
        char *d_bases = nullptr; uint64_t cap_bases = 0;
        uint64_t *d_offs = nullptr; kmat_read_result *d_out = nullptr; uint32_t cap_reads = 0;
        uint32_t *d_codes = nullptr; uint64_t cap_codes = 0;      // compact interface: the chunk's 2-bit code words ...
        uint64_t *d_inv = nullptr; uint64_t cap_inv = 0;          // ... and the positions of its non-ACGT bases
        kmat_read_result32 *d_out32 = nullptr;                    // ... and its 32-byte results (cap_reads entries)
        unsigned long long *d_ref = nullptr; uint64_t cap_ref = 0; // K5: where each read's tail went (kmat_label_batch_text)
        unsigned long long *h_cur = nullptr;       // pinned: cursors after this slot's chunk
        cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    } slot[2];
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    // intra-pass pipeline: the encode+probe kernel of sub-batch i+1 (request-rate bound, few warps) runs next to the
    // candidate and scoring kernels of sub-batch i (issue bound) on a second stream
    cudaStream_t st_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sub[16] = {};
    int pipeline = 1;                        // sub-batches per pass: 0/1 serial (default), -1 automatic
    int sms = 148;
    uint32_t *d_hit = nullptr; uint64_t cap_hit = 0;
    int2 *d_hdr = nullptr; uint32_t cap_hdr = 0;
    kmat_read_result *d_out_dev = nullptr; uint32_t cap_out_dev = 0;   // kmat_label_batch_device with d_out == NULL
    kmat_pair *d_cands = nullptr, *d_lin = nullptr; uint64_t cap_cands = 0, cap_lin = 0;
    unsigned long long *d_cursors = nullptr;     // [0] cands, [1] lineage, [2] text bytes (K5, kmat_format.cuh), [3] run-length list words
    uint32_t *d_plist = nullptr; uint64_t cap_plist = 0;     // the pair lists of a host-buffer call in run-length form (kmat_label_batch_packed_rl)
    uint32_t max_tid = 0;                                    // largest taxid of the node universe (the run-length form needs bit 31 of a taxid)
    char *d_text = nullptr; uint64_t cap_text = 0;    // the tails of a host-buffer call formatted on the device
    uint32_t *d_pool2 = nullptr; int pool2_mul = 1;            // resolved lists (km_resolve_kernel)
    int resolved_max_count = -1, resolved_permissive = -1, resolved_rkmer = -1;
    // rand_read_label accumulators (kmat_null.cuh): [n_nodes * KMAT_NULL_BUCKETS] max fraction (float bits) / read counts
    uint32_t *d_null_max = nullptr, *d_null_cnt = nullptr; unsigned long long *d_null_err = nullptr;
    uint64_t null_first = 0;                 // run index of read 0 of the pass being launched
    // direct sharded mode (kmat_ctx_peer_attach): where every shard's buckets / stash / resolved pool are mapped on this GPU
    // slow path for reads with more than KB_CMAX candidates: queues, counters and per-thread scratch slots
    unsigned long long *d_pass = nullptr;            // per-pass packed cursor (KmScoreParams::pass_cursor)
    uint32_t *d_pendq = nullptr; uint64_t cap_pendq = 0;
    uint32_t *d_bigq = nullptr; unsigned int *d_bigcnt = nullptr;
    unsigned char *d_huge3 = nullptr, *d_huge4 = nullptr;             // scratch slots of km_cand_huge_kernel / km_score_huge_kernel
    unsigned char *d_big3 = nullptr, *d_big4 = nullptr; uint64_t cap_big3 = 0; uint32_t big_np_cap = 0, big_threads3 = 0, big_threads4 = 0;
    KmPeer *d_peers = nullptr; uint32_t n_peers = 0; std::vector<void *> ipc_mapped;
    uint32_t *d_pool2_all = nullptr;                                  // every shard's resolved pool, concatenated (list hits stay local)
    bool pool2_all_alias = false;                                     //   ... or simply d_pool2 when every shard holds the same pool
    uint32_t *d_peer_recs = nullptr; uint64_t cap_peer_recs = 0;      // list records of the pass copied from their owners (km_peer_fetch_kernel)
    unsigned long long *d_peer_cur = nullptr;                         // [0] words used (per pass), [1] list hits dropped for lack of room (monotonic)
    unsigned long long peer_dropped_seen = 0; int peer_grow = 1;
    int cand_grow = 1;                                                // candidate buffer: factor on the per-read estimate
    char *d_null_bases = nullptr; uint64_t *d_null_offs = nullptr; uint64_t cap_null_bases = 0, cap_null_offs = 0;
    unsigned long long *d_long_masks = nullptr; uint32_t long_mask_cap = 0;
    unsigned long long *d_long_sets = nullptr; uint32_t long_slots = 0; int long_warps = 0;
    KmStatsDev *d_stats = nullptr;
    int collect_stats = 1;
    kmat_batch_stats last{};
    int cand_grid[3] = {0, 0, 0};   // persistent grids of km_cand_kernel<5>, <10>, <0> (warp per read)
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // around the three kernels of the last batch; [4]: before the direct-mode list fetch
};

template <typename T>
static int km_upload(T **dst, const std::vector<T> &src) {
    *dst = nullptr;
    if (src.empty()) return KMAT_OK;
    KM_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    KM_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return KMAT_OK;
}

static KmCtxDev km_ctx_dev(const kmat_ctx *c);

// (Re)build the resolved list pool for the ctx's current -g / -s options.
static int km_resolve_lists(kmat_ctx *c) {
    const kmat_db *db = c->db;
    if (!db->pool_words) return KMAT_OK;
    if (c->d_pool2 && c->resolved_max_count == c->opt.max_count && c->resolved_permissive == (c->opt.permissive != 0) &&
        c->resolved_rkmer == (c->opt.rkmer_mode != 0)) return KMAT_OK;
    if (c->pool2_all_alias) {
        // direct mode over a shared pool: the records of the other shards' lists were merged in at attach time and cannot be
        // rebuilt by this rank alone
        kmat_set_error("kmat_ctx_set_opts: -g / -s cannot change after kmat_ctx_peer_attach on a shared list pool; create the contexts with the new options and attach again");
        return KMAT_ERR_UNSUPPORTED;
    }
    KM_CUDA(cudaStreamSynchronize(c->stream));
    const int mul = (db->tid_bytes == 2 ? 2 : 1) * (c->opt.permissive ? 2 : 1);
    if (!c->d_pool2 || mul != c->pool2_mul) {
        cudaFree(c->d_pool2); c->d_pool2 = nullptr;
        KM_CUDA(cudaMalloc((void **)&c->d_pool2, ((size_t)db->pool_words * mul + 32) * 4));      // + padding: the direct-mode fetch reads two whole sectors from a record's start
        c->pool2_mul = mul;
    }
    // A shard built from the whole table's arrays resolves only the lists ITS k-mers point at (the kernel below walks this
    // shard's slots); the rest of the pool stays zero so that kmat_ctx_peer_attach can merge the shards' pools word by word
    if (db->pool_shared) KM_CUDA(cudaMemsetAsync(c->d_pool2, 0, ((size_t)db->pool_words * mul + 32) * 4, c->stream));
    KmResolveParams R;
    R.C = km_ctx_dev(c);
    R.pool2 = c->d_pool2;
    R.n_table_slots = db->d_slots ? db->n_buckets * KM_SLOTS_PER_BUCKET : 0;
    R.n_line_slots = db->n_lines * 16;
    R.n_slots = R.n_line_slots + R.n_table_slots + db->n_stash;
    R.big_cap = (uint32_t)(db->pool_words / 9 + 16);         // a list of > KB_LFAST 16-bit ids occupies >= 9 pool words
    R.scratch = nullptr; R.scratch_entries = 0;
    KM_CUDA(cudaMalloc((void **)&R.big_queue, (size_t)R.big_cap * 4));
    KM_CUDA(cudaMalloc((void **)&R.counters, 16));
    KM_CUDA(cudaMemsetAsync(R.counters, 0, 16, c->stream));
    const int grid = (int)std::min<unsigned long long>((R.n_slots + 255) / 256, 148ull * 32);
    km_resolve_kernel<<<grid, 256, 0, c->stream>>>(R);
    g_km_launches++;
    unsigned int cnt[4] = {0, 0, 0, 0};
    KM_CUDA(cudaMemcpyAsync(cnt, R.counters, 16, cudaMemcpyDeviceToHost, c->stream));
    KM_CUDA(cudaStreamSynchronize(c->stream));
    int rc = KMAT_OK;
    if (cnt[0] > R.big_cap) { kmat_set_error("resolve: %u long lists exceed the queue of %u", cnt[0], R.big_cap); rc = KMAT_ERR_UNSUPPORTED; }
    else if (cnt[0]) {
        const uint32_t per = 2 * cnt[1] + 2;                                  // ids in next() order + the pruning heap
        uint32_t threads = std::min<uint32_t>(cnt[0], 148u * 128);
        while (threads > 128 && (size_t)threads * per * sizeof(uint2) > ((size_t)512 << 20)) threads /= 2;
        threads = (threads + 127) / 128 * 128;
        R.scratch_entries = per;
        KM_CUDA(cudaMalloc((void **)&R.scratch, (size_t)threads * per * sizeof(uint2)));
        km_resolve_big_kernel<<<threads / 128, 128, 0, c->stream>>>(R, cnt[0]);
        g_km_launches++;
        KM_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(R.scratch);
    }
    cudaFree(R.big_queue); cudaFree(R.counters);
    KM_CUDA(cudaGetLastError());
    c->resolved_max_count = c->opt.max_count; c->resolved_permissive = c->opt.permissive != 0; c->resolved_rkmer = c->opt.rkmer_mode != 0;
    return rc;
}

extern "C" int kmat_ctx_create(const kmat_db *db, const kmat_inputs *in, const kmat_opts *opt, kmat_ctx **out) {
    if (!db || !in || !out) { kmat_set_error("kmat_ctx_create: bad argument"); return KMAT_ERR_ARG; }
    kmat_opts o;
    if (opt) o = *opt; else kmat_opts_default(&o);
    kmat_ctx *c = new kmat_ctx();
    if (o.rkmer_mode) { o.min_kmer = 0; o.min_fnd_kmer = 0; o.permissive = 0; o.want_lineage = 0; }   // rand_read_label has neither -j nor -z (and never sets gPERMISSIVE_MATCH)
    c->db = db; c->device = db->device; c->opt = o;
    int rc = kmat_build_host_ctx(*in, db->tid_bytes, db->stored_tids, c->h);
    if (rc != KMAT_OK) { delete c; return rc; }
    if (cudaSetDevice(c->device) != cudaSuccess) { delete c; kmat_set_error("cudaSetDevice(%d) failed", db->device); return KMAT_ERR_NO_DEVICE; }
#define UP(dst, src) do { rc = km_upload(&c->dst, c->h.src); if (rc != KMAT_OK) { kmat_ctx_destroy(c); return rc; } } while (0)
    UP(d_nodeA, nodeA); UP(d_nodeB, nodeB); UP(d_paths, paths); UP(d_prune, prune_rank); UP(d_sid2nid, sid2nid);
    UP(d_model_of_cand, model_of_cand); UP(d_mrow, mrow); UP(d_cut, cut); UP(d_cls, cls);
#undef UP
    KM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    KM_CUDA(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
    KM_CUDA(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
    KM_CUDA(cudaStreamCreateWithFlags(&c->st_aux, cudaStreamNonBlocking));
    KM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); KM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (auto &e : c->ev_sub) KM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int i = 0; i < 5; i++) KM_CUDA(cudaEventCreate(&c->ev[i]));
    for (auto &sl : c->slot) {
        KM_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
        KM_CUDA(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
        KM_CUDA(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
        KM_CUDA(cudaMallocHost((void **)&sl.h_cur, 32));
    }
    KM_CUDA(cudaMalloc((void **)&c->d_cursors, 32));
    KM_CUDA(cudaMalloc((void **)&c->d_pass, 8));
    KM_CUDA(cudaMemset(c->d_pass, 0, 8));
    KM_CUDA(cudaMalloc((void **)&c->d_stats, sizeof(KmStatsDev)));
    int per_sm[3] = {0, 0, 0}, sms = 148;
    KM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], km_cand_kernel<5>, KB_WARPS * 32, 0));
    KM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], km_cand_kernel<10>, KB_WARPS * 32, 0));
    KM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[2], km_cand_kernel<0>, KB_WARPS * 32, 0));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    c->sms = sms;
    for (int i = 0; i < 3; i++) c->cand_grid[i] = std::max(1, per_sm[i]) * sms;
    rc = km_resolve_lists(c);
    if (rc != KMAT_OK) { kmat_ctx_destroy(c); return rc; }
    for (const KmNodeA &na : c->h.nodeA) c->max_tid = std::max(c->max_tid, na.tid);
    *out = c;
    return KMAT_OK;
}
extern "C" int kmat_ctx_set_opts(kmat_ctx *c, const kmat_opts *o) {
    if (!c || !o) return KMAT_ERR_ARG;
    const kmat_opts before = c->opt;
    c->opt = *o;
    if (c->opt.rkmer_mode) { c->opt.min_kmer = 0; c->opt.min_fnd_kmer = 0; c->opt.permissive = 0; c->opt.want_lineage = 0; }
    KM_CUDA(cudaSetDevice(c->device));
    const int rc = km_resolve_lists(c);  // the resolved lists depend on -g and -s
    if (rc != KMAT_OK) c->opt = before;  // refused: the context keeps the options its lists were resolved for
    return rc;
}
extern "C" void kmat_ctx_destroy(kmat_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_nodeA); cudaFree(c->d_nodeB); cudaFree(c->d_paths); cudaFree(c->d_prune); cudaFree(c->d_sid2nid);
    cudaFree(c->d_model_of_cand); cudaFree(c->d_mrow); cudaFree(c->d_cut); cudaFree(c->d_cls);
    for (auto &sl : c->slot) {
        cudaFree(sl.d_bases); cudaFree(sl.d_offs); cudaFree(sl.d_out); cudaFree(sl.d_codes); cudaFree(sl.d_inv); cudaFree(sl.d_out32); cudaFree(sl.d_ref); cudaFreeHost(sl.h_cur);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
        if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
    }
    km_shard_free(c->shard);
    for (void *p : c->ipc_mapped) cudaIpcCloseMemHandle(p);
    cudaFree(c->d_peers); cudaFree(c->d_peer_recs); cudaFree(c->d_peer_cur); if (!c->pool2_all_alias) cudaFree(c->d_pool2_all);
    cudaFree(c->d_bigq); cudaFree(c->d_bigcnt); cudaFree(c->d_big3); cudaFree(c->d_big4); cudaFree(c->d_huge3); cudaFree(c->d_huge4); cudaFree(c->d_pass); cudaFree(c->d_pendq);
    cudaFree(c->d_null_max); cudaFree(c->d_null_cnt); cudaFree(c->d_null_err); cudaFree(c->d_null_bases); cudaFree(c->d_null_offs);
    cudaFree(c->d_hit); cudaFree(c->d_hdr); cudaFree(c->d_out_dev);
    if (c->st_h2d) cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
    if (c->st_aux) { cudaStreamSynchronize(c->st_aux); cudaStreamDestroy(c->st_aux); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto &e : c->ev_sub) if (e) cudaEventDestroy(e);
    cudaFree(c->d_cands); cudaFree(c->d_lin); cudaFree(c->d_text); cudaFree(c->d_plist); cudaFree(c->d_cursors); cudaFree(c->d_pool2); cudaFree(c->d_long_masks); cudaFree(c->d_long_sets); cudaFree(c->d_stats);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
}

static KmCtxDev km_ctx_dev(const kmat_ctx *c) {
    KmCtxDev X;
    X.db = km_db_dev(c->db);           // peers stay out of K3 / K4: direct mode copies the remote list records first (km_peer_fetch_kernel)
    X.nodeA = c->d_nodeA; X.nodeB = c->d_nodeB; X.paths = c->d_paths; X.prune_rank = c->d_prune; X.sid2nid = c->d_sid2nid;
    X.n_sid = (uint32_t)c->h.sid2nid.size(); X.n_nodes = (uint32_t)c->h.nodeA.size(); X.nid_human = c->h.nid_human; X.nid_one = c->h.nid_one;
    X.nbins = c->h.nbins; X.n_models = c->h.n_models; X.n_classes = c->h.n_classes;
    X.model_of_cand = c->d_model_of_cand; X.mrow = c->d_mrow; X.cut = c->d_cut; X.cls = c->d_cls;
    memset(X.class_ranknum, 0, sizeof X.class_ranknum);
    for (int i = 0; i < c->h.n_classes && i < 64; i++) X.class_ranknum[i] = (int8_t)c->h.class_ranknum[i];
    X.opt = c->opt;
    X.pool2 = c->d_pool2; X.pool2_mul = c->pool2_mul;
    return X;
}

template <typename T>
static int km_grow(T **p, uint64_t *cap, uint64_t want) {
    if (want <= *cap) return KMAT_OK;
    cudaFree(*p); *p = nullptr;
    const uint64_t ncap = want + want / 4 + 1024;
    KM_CUDA(cudaMalloc((void **)p, ncap * sizeof(T)));
    *cap = ncap;
    return KMAT_OK;
}

// One pass of the kernels over reads already on the device.  d_bases / the hit array are addressed as
// base[offs[r] + j], so a chunk of a larger batch passes pointers shifted by -offs[first read].
struct KmPass {
    const char *d_bases; const uint64_t *d_offs; uint32_t n_reads;
    uint64_t first_off, total_bases; uint32_t max_len;
    kmat_read_result *d_out;
    bool reset;                 // zero the candidate cursors and the statistics first
};

// Buffers of a pass (hit words, per-read headers, long-read scratch) and the reset of cursors / statistics.
// variant: which candidate-kernel instantiation the longest read needs (0: <= 160 positions, 1: <= 320, 2: any).
static int km_prepare_pass(kmat_ctx *c, const KmPass &L, cudaStream_t st, uint32_t **hit, int *variant) {
    int rc;
    if ((rc = km_grow(&c->d_hit, &c->cap_hit, L.total_bases + 1)) != KMAT_OK) return rc;
    { uint64_t cap = c->cap_hdr; if ((rc = km_grow(&c->d_hdr, &cap, L.n_reads)) != KMAT_OK) return rc; c->cap_hdr = (uint32_t)cap; }
    if (L.max_len > 256) {
        uint32_t slots = 1024; while (slots < 2 * L.max_len) slots <<= 1;
        const int warps = 148 * 6 * KM_PROBE_WARPS_HOST;
        if (slots > c->long_slots || warps > c->long_warps) {
            KM_CUDA(cudaStreamSynchronize(st));
            cudaFree(c->d_long_sets); c->d_long_sets = nullptr;
            KM_CUDA(cudaMalloc((void **)&c->d_long_sets, (size_t)warps * slots * 8));
            c->long_slots = slots; c->long_warps = warps;
        }
    }
    if (L.reset) {
        KM_CUDA(cudaMemsetAsync(c->d_cursors, 0, 32, st));
        if (c->collect_stats) KM_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(KmStatsDev), st));
    }
    *hit = c->d_hit - L.first_off;
    const int max_pos = (int)L.max_len - c->db->kmer_len + 1;
    *variant = max_pos <= 5 * 32 ? 0 : max_pos <= 10 * 32 ? 1 : 2;
    if (*variant == 2) {
        // long reads: the position masks live in a per-warp global scratch
        const uint32_t cap = ((uint32_t)max_pos + 31u) & ~31u;
        if (cap > c->long_mask_cap) {
            KM_CUDA(cudaStreamSynchronize(st)); KM_CUDA(cudaStreamSynchronize(c->st_aux));
            cudaFree(c->d_long_masks); c->d_long_masks = nullptr;
            KM_CUDA(cudaMalloc((void **)&c->d_long_masks, (size_t)c->cand_grid[2] * KB_WARPS * cap * 8));
            c->long_mask_cap = cap;
        }
    }
    return KMAT_OK;
}

static int km_launch_nullacc(kmat_ctx *c, const KmScoreParams &P, uint32_t r0, uint32_t n, cudaStream_t s2);   // kmat_null.cuh
// K3 + K4 over reads [r0, r0 + n) of a pass on stream s2.  pool2 / pool2_mul override the ctx's resolved list pool
// (DB-sharded mode: the records fetched from the owning shards for this batch).  ev_mid, if set, is recorded between
// the two kernels.
static int km_launch_cand_score(kmat_ctx *c, const KmPass &L, uint32_t r0, uint32_t n, uint32_t *hit, int variant, int cand_ctas_per_sm,
                                cudaStream_t s2, const uint32_t *pool2, int pool2_mul, cudaEvent_t ev_mid) {
    KmScoreParams P;
    P.C = km_ctx_dev(c);
    if (pool2) { P.C.pool2 = pool2; P.C.pool2_mul = pool2_mul; }
    P.offs = L.d_offs + r0; P.n_reads = n; P.hit = hit; P.hdr = c->d_hdr + r0; P.out = L.d_out + r0;
    P.cands = c->d_cands; P.cand_cursor = c->d_cursors; P.cand_cap = c->cap_cands;
    P.pass_cursor = c->d_pass; P.pend_q = nullptr;
    if (!c->opt.rkmer_mode && n < (1u << (64 - KB_PASS_SHIFT)) && !getenv("KMAT_NO_PEND_QUEUE")) {
        if ((uint64_t)n > c->cap_pendq) { KM_CUDA(cudaStreamSynchronize(s2)); int rcq = km_grow(&c->d_pendq, &c->cap_pendq, (uint64_t)n); if (rcq != KMAT_OK) return rcq; }
        P.pend_q = c->d_pendq;
    }
    P.lin = c->d_lin; P.lin_cursor = c->d_cursors + 1; P.lin_cap = c->cap_lin;
    P.stats = c->collect_stats ? c->d_stats : nullptr;
    P.long_masks = nullptr; P.long_cap = 0;
    // slow path buffers (first use / longer reads than before)
    // threads (= global scratch slots of 42 KB) of the big scoring kernel: few for short reads, where it is a rare path,
    // a GB worth of them for long reads, where nearly every read takes it
    const uint32_t KBIG_THREADS = L.max_len > 1024 ? 24576u : 2048u;
    const uint32_t np_need = (uint32_t)std::max<int>(1, (int)L.max_len - c->db->kmer_len + 1);
    if (!c->d_bigq) {
        KM_CUDA(cudaMalloc((void **)&c->d_bigq, ((size_t)2 * KB_BIGQ + 2 * KB_HUGEQ) * 4));
        KM_CUDA(cudaMalloc((void **)&c->d_bigcnt, 16));
        // the last-resort path (more than KB_CBIG candidates): KBH_WARPS + KBH_THREADS scratch slots, ~0.5 GB
        KM_CUDA(cudaMalloc((void **)&c->d_huge3, (size_t)KBH_WARPS * KBH_SLOT3_BYTES));
        KM_CUDA(cudaMalloc((void **)&c->d_huge4, (size_t)KBH_THREADS * KBH_SLOT4_BYTES));
    }
    if (KBIG_THREADS > c->big_threads4) {
        KM_CUDA(cudaStreamSynchronize(s2));
        cudaFree(c->d_big4); c->d_big4 = nullptr;
        KM_CUDA(cudaMalloc((void **)&c->d_big4, (size_t)KBIG_THREADS * sizeof(KsLocalBig)));
        c->big_threads4 = KBIG_THREADS;
    }
    if (np_need > c->big_np_cap) {
        const size_t per = ((size_t)np_need + KB_CBIG) * KB_BIGW * 8;                       // one warp's position sets + lineage sets
        uint32_t threads = (uint32_t)std::min<size_t>((size_t)c->sms * KBG_WARPS * 2, std::max<size_t>(1, ((size_t)1 << 30) / per));
        // per launch: function attributes belong to the current device's context (read_label drives every visible GPU from one process)
        KM_CUDA(cudaFuncSetAttribute(km_cand_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KBG_WARPS * sizeof(KbBigW))));
        KM_CUDA(cudaStreamSynchronize(s2));
        int rc3 = km_grow(&c->d_big3, &c->cap_big3, (uint64_t)per * threads);
        if (rc3 != KMAT_OK) return rc3;
        c->big_np_cap = np_need; c->big_threads3 = threads;
    }
    P.big_qa = c->d_bigq; P.big_qb = c->d_bigq + KB_BIGQ; P.big_cnt = c->d_bigcnt;
    P.big_scratch3 = c->d_big3; P.big_scratch4 = c->d_big4; P.big_np_cap = c->big_np_cap; P.big_threads3 = c->big_threads3; P.big_threads4 = c->big_threads4;
    P.huge_qa = c->d_bigq + 2 * KB_BIGQ; P.huge_qb = P.huge_qa + KB_HUGEQ; P.huge_scratch3 = c->d_huge3; P.huge_scratch4 = c->d_huge4;
    KM_CUDA(cudaMemsetAsync(c->d_bigcnt, 0, 16, s2));
    const int want_grid = (int)((n + KB_WARPS - 1) / KB_WARPS);
    const int g0 = cand_ctas_per_sm > 0 ? std::min(c->cand_grid[0], cand_ctas_per_sm * c->sms) : c->cand_grid[0];
    if (variant == 0) km_cand_kernel<5><<<std::max(1, std::min(g0, want_grid)), KB_WARPS * 32, 0, s2>>>(P);
    else if (variant == 1) km_cand_kernel<10><<<std::max(1, std::min(c->cand_grid[1], want_grid)), KB_WARPS * 32, 0, s2>>>(P);
    else {
        P.long_masks = c->d_long_masks; P.long_cap = c->long_mask_cap;
        km_cand_kernel<0><<<std::max(1, std::min(c->cand_grid[2], want_grid)), KB_WARPS * 32, 0, s2>>>(P);
    }
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    // the (rare) reads K3 could not hold in registers; an empty queue costs two tiny launches
    km_cand_big_kernel<<<(c->big_threads3 + KBG_WARPS - 1) / KBG_WARPS, KBG_WARPS * 32, KBG_WARPS * sizeof(KbBigW), s2>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    km_cand_huge_kernel<<<KBH_WARPS / 4, 128, 0, s2>>>(P);        // the reads even that one could not hold (an empty queue: every warp leaves at once)
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    if (ev_mid) KM_CUDA(cudaEventRecord(ev_mid, s2));
    if (c->opt.rkmer_mode) {                      // rand_read_label: accumulate instead of scoring
        const int rcn = km_launch_nullacc(c, P, r0, n, s2);
        if (rcn != KMAT_OK) return rcn;
        km_cursor_roll_kernel<<<1, 1, 0, s2>>>(c->d_cursors, c->d_pass);
        g_km_launches++;
        KM_CUDA(cudaGetLastError());
        return KMAT_OK;
    }
    if (n >= (4u << 20)) km_score_kernel<4><<<(n + KS_THREADS * 4 - 1) / (KS_THREADS * 4), KS_THREADS, 0, s2>>>(P);
    else km_score_kernel<2><<<(n + KS_THREADS * 2 - 1) / (KS_THREADS * 2), KS_THREADS, 0, s2>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    km_score_big_kernel<<<c->big_threads4 / 32, 32, 0, s2>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    km_score_huge_kernel<<<KBH_THREADS / 32, 32, 0, s2>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    km_cursor_roll_kernel<<<1, 1, 0, s2>>>(c->d_cursors, c->d_pass);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}

static int km_peer_prepare(kmat_ctx *c, const KmPass &L, cudaStream_t st);      // kmat_shard.cuh
static int km_peer_fetch(kmat_ctx *c, const KmPass &L, cudaStream_t st);
static int km_run_device(kmat_ctx *c, const KmPass &L, cudaStream_t st) {
    int rc, variant;
    uint32_t *hit;
    if ((rc = km_prepare_pass(c, L, st, &hit, &variant)) != KMAT_OK) return rc;
    // Sub-batches.  Serial: probe, candidates, scoring one after the other on `st` (per-kernel times through ev[]).
    // Pipelined (short reads, large passes): the probe kernel keeps ~1 CTA per SM (it is bound by the table request
    // rate and loses ~15 % at 8 warps per SM) and the other two kernels work on the previous sub-batch in the rest of
    // every SM, on st_aux.
    int S = c->pipeline;
    if (S < 0) S = (variant == 0 && L.n_reads >= (1u << 19)) ? 8 : 1;
    if (variant != 0 || S < 1 || c->d_peers) S = 1;
    if (c->d_peers && !c->d_pool2_all) {
        if ((rc = km_peer_prepare(c, L, st)) != KMAT_OK) return rc;
        // km_peer_fetch_kernel walks EVERY hit word of the pass; the last k - 1 positions of a read are never written by the
        // probe kernel and would otherwise hold stale words of an earlier pass (with owner tags of another shard count)
        KM_CUDA(cudaMemsetAsync(hit, 0xFF, (size_t)L.total_bases * 4, st));
    }
    if (S > 16) S = 16;
    const bool piped = S > 1;
    KM_CUDA(cudaEventRecord(c->ev[0], st));
    if (piped) { KM_CUDA(cudaEventRecord(c->ev_fork, st)); KM_CUDA(cudaStreamWaitEvent(c->st_aux, c->ev_fork, 0)); }
    for (int sb = 0; sb < S; sb++) {
        const uint32_t r0 = (uint32_t)((uint64_t)L.n_reads * sb / S), r1 = (uint32_t)((uint64_t)L.n_reads * (sb + 1) / S);
        if (r1 == r0) continue;
        const uint32_t n = r1 - r0;
        rc = km_launch_encode_probe(c->db, L.d_bases, L.d_offs + r0, n, L.max_len, hit, c->d_hdr + r0, nullptr, nullptr, c->d_long_sets, c->long_slots,
                                    km_probe_grid(n), c->collect_stats ? c->d_stats : nullptr, 1, st, piped ? 1 : 0, nullptr, c->d_peers, c->n_peers);
        if (rc != KMAT_OK) return rc;
        cudaStream_t s2 = st;
        if (piped) { KM_CUDA(cudaEventRecord(c->ev_sub[sb], st)); KM_CUDA(cudaStreamWaitEvent(c->st_aux, c->ev_sub[sb], 0)); s2 = c->st_aux; }
        else KM_CUDA(cudaEventRecord(c->ev[1], st));
        if (c->d_peers && c->d_pool2_all) {
            // direct sharded mode with a local copy of every shard's list pool: the hit words already point into it
            if (!piped) KM_CUDA(cudaEventRecord(c->ev[1], st));
            if ((rc = km_launch_cand_score(c, L, r0, n, hit, variant, 0, s2, c->d_pool2_all, c->pool2_mul, c->ev[3])) != KMAT_OK) return rc;
            continue;
        }
        if (c->d_peers) {
            // direct sharded mode: the list records the hits point at live in their owners' pools; copy them next to the
            // batch (many remote reads in flight) so that K3 only touches local memory
            KM_CUDA(cudaEventRecord(c->ev[4], st));
            if ((rc = km_peer_fetch(c, L, st)) != KMAT_OK) return rc;
            KM_CUDA(cudaEventRecord(c->ev[1], st));
            if (getenv("KMAT_DEBUG_TIMING")) {              // debugging aid: split of the probe time (synchronises)
                KM_CUDA(cudaEventSynchronize(c->ev[1]));
                float a = 0, b = 0; unsigned long long cur[2] = {0, 0};
                cudaEventElapsedTime(&a, c->ev[0], c->ev[4]); cudaEventElapsedTime(&b, c->ev[4], c->ev[1]);
                cudaMemcpy(cur, c->d_peer_cur, 16, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[kmat dev %d] probe %.2f ms, list fetch %.2f ms, %llu record words, %llu dropped\n", c->device, a, b, cur[0], cur[1]);
            }
            if ((rc = km_launch_cand_score(c, L, r0, n, hit, variant, 0, s2, c->d_peer_recs, 1, c->ev[3])) != KMAT_OK) return rc;
            continue;
        }
        if ((rc = km_launch_cand_score(c, L, r0, n, hit, variant, piped ? 2 : 0, s2, nullptr, 0, piped ? nullptr : c->ev[3])) != KMAT_OK) return rc;
    }
    if (piped) {
        KM_CUDA(cudaEventRecord(c->ev_join, c->st_aux)); KM_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
        KM_CUDA(cudaEventRecord(c->ev[1], st)); KM_CUDA(cudaEventRecord(c->ev[3], st));     // no per-kernel split in this mode
    }
    KM_CUDA(cudaEventRecord(c->ev[2], st));
    return KMAT_OK;
}

static int km_grow_cands(kmat_ctx *c, uint64_t want) { return km_grow(&c->d_cands, &c->cap_cands, want); }
// Candidate / lineage pair buffers for a pass of n_reads reads of at most max_len bases.  A 150 bp read of the synthetic
// workloads carries ~10 candidates; long reads collect chance hits all over the taxonomy (a 10 kbp read ~100-200), hence
// the length term.  cand_grow doubles after a pass overflowed (kmat_ctx_sync / kmat_label_batch report that).
static int km_reserve_cands(kmat_ctx *c, uint64_t n_reads, uint32_t max_len = 0) {
    int rc;
    const uint64_t per = std::min<uint64_t>(KB_CBIG, (uint64_t)(24 + max_len / 32) * (uint64_t)c->cand_grow);
    const uint64_t want = n_reads * per + 4096;
    if (!c->d_cands || c->cap_cands < want / 2) { if ((rc = km_grow_cands(c, want)) != KMAT_OK) return rc; }
    if (c->opt.want_lineage && (!c->d_lin || c->cap_lin < want / 2)) { if ((rc = km_grow(&c->d_lin, &c->cap_lin, want)) != KMAT_OK) return rc; }
    return KMAT_OK;
}

extern "C" int kmat_label_batch_device(kmat_ctx *c, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint64_t total_bases,
                                       uint32_t max_read_len, kmat_read_result *d_out, void *stream) {
    if (!c || !d_offs || (n_reads && !d_bases)) { kmat_set_error("kmat_label_batch_device: bad argument"); return KMAT_ERR_ARG; }
    if (c->opt.rkmer_mode) { kmat_set_error("kmat_label_batch_device: the ctx was created with rkmer_mode (kmat_null_* only)"); return KMAT_ERR_ARG; }
    if (c->db->shard_count > 1 && !c->d_peers) { kmat_set_error("kmat_label_batch_device: the table is shard %d of %d; attach the peers (kmat_ctx_peer_attach) or use the kmat_shard_* rounds", c->db->shard_index, c->db->shard_count); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    if (!n_reads) return KMAT_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc;
    if (!d_out) {
        uint64_t cap = c->cap_out_dev;
        if ((rc = km_grow(&c->d_out_dev, &cap, n_reads)) != KMAT_OK) return rc;
        c->cap_out_dev = (uint32_t)cap;
        d_out = c->d_out_dev;
    }
    if ((rc = km_reserve_cands(c, n_reads, max_read_len)) != KMAT_OK) return rc;
    KmPass L{d_bases, d_offs, n_reads, 0, total_bases, max_read_len, d_out, true};
    return km_run_device(c, L, st);
}
extern "C" int kmat_ctx_last_kernel_ms(kmat_ctx *c, float *probe_ms, float *cand_ms, float *score_ms) {
    if (!c) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(c->device));
    KM_CUDA(cudaEventSynchronize(c->ev[2]));
    float a = 0, b = 0, d = 0;
    KM_CUDA(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
    KM_CUDA(cudaEventElapsedTime(&b, c->ev[1], c->ev[3]));
    KM_CUDA(cudaEventElapsedTime(&d, c->ev[3], c->ev[2]));
    if (probe_ms) *probe_ms = a;
    if (cand_ms) *cand_ms = b;
    if (score_ms) *score_ms = d;
    return KMAT_OK;
}
extern "C" int kmat_ctx_set_pipeline(kmat_ctx *c, int sub_batches) {
    if (!c) return KMAT_ERR_ARG;
    c->pipeline = sub_batches;
    return KMAT_OK;
}
extern "C" int kmat_ctx_set_stats(kmat_ctx *c, int enable) {
    if (!c) return KMAT_ERR_ARG;
    c->collect_stats = enable ? 1 : 0;
    return KMAT_OK;
}
// direct sharded mode: list hits the record buffer had no room for were turned into misses (never into garbage); report
// that once and size the buffer larger for the next pass
static int km_peer_check(kmat_ctx *c) {
    if (!c->d_peers || !c->d_peer_cur) return KMAT_OK;
    unsigned long long cur[2] = {0, 0};
    KM_CUDA(cudaMemcpy(cur, c->d_peer_cur, 16, cudaMemcpyDeviceToHost));
    if (cur[1] == c->peer_dropped_seen) return KMAT_OK;
    const unsigned long long d = cur[1] - c->peer_dropped_seen;
    c->peer_dropped_seen = cur[1];
    c->peer_grow *= 2;
    kmat_set_error("direct sharded mode: the list-record buffer was too small for %llu list hits of the last pass; it is doubled now, run the batch again", d);
    return KMAT_ERR_OVERFLOW;
}
extern "C" int kmat_ctx_sync(kmat_ctx *c) {
    if (!c) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(c->device));
    KM_CUDA(cudaStreamSynchronize(c->stream));
    // a device-resident pass whose candidate buffer was too small left reads in KMAT_ST_ERROR / KMAT_ERR_OVERFLOW: say so
    unsigned long long cur[2] = {0, 0};
    KM_CUDA(cudaMemcpy(cur, c->d_cursors, 16, cudaMemcpyDeviceToHost));
    if (cur[0] > c->cap_cands || (c->opt.want_lineage && cur[1] > c->cap_lin)) {
        c->cand_grow *= 2;
        kmat_set_error("the candidate buffer (%llu pairs) was too small for the last pass (%llu needed): some reads are in KMAT_ST_ERROR; it grows now, run the batch again",
                       (unsigned long long)c->cap_cands, cur[0]);
        return KMAT_ERR_OVERFLOW;
    }
    return km_peer_check(c);
}

static int km_fetch_stats(kmat_ctx *c, cudaStream_t st) {
    if (!c->collect_stats) return KMAT_OK;
    KmStatsDev s;
    KM_CUDA(cudaMemcpyAsync(&s, c->d_stats, sizeof s, cudaMemcpyDeviceToHost, st));
    KM_CUDA(cudaStreamSynchronize(st));
    kmat_batch_stats &b = c->last;
    b.lookups = s.lookups; b.hits = s.hits; b.list_hits = s.list_hits; b.list_ids = s.list_ids; b.probe_extra_buckets = s.extra_buckets;
    // 1 sector for a prefix miss, 2 for any other lookup, plus ceil((2 + n*w)/32) per fetched list
    b.algorithmic_bytes = 32ull * (s.prefix_miss + 2ull * (s.lookups - s.prefix_miss) + s.list_sectors);
    b.reads_fast = s.reads_fast; b.reads_slow = s.reads_slow; b.reads_error = s.reads_error;
    return KMAT_OK;
}
extern "C" int kmat_ctx_last_stats(kmat_ctx *c, kmat_batch_stats *out) {
    if (!c || !out) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(c->device));
    int rc = km_fetch_stats(c, c->stream);
    if (rc != KMAT_OK) return rc;
    *out = c->last;
    return KMAT_OK;
}

// ---- compact interface: 2-bit packed reads in, 32-byte results out (kmat_label_batch_packed) --------------------------
// One thread per code word: 16 bases -> 16 ASCII bytes, stored as one 16-byte vector.  The chunk's byte buffer starts at a
// multiple of 16 bases, so word w of the chunk is bytes [16 w, 16 w + 16).
__global__ void __launch_bounds__(256) km_unpack_kernel(const uint32_t *__restrict__ codes, uint64_t n_words, uint4 *out) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint32_t c = codes[w];
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v |= ((0x54474341u >> (8 * ((c >> (2 * (4 * i + j))) & 3u))) & 0xFFu) << (8 * j);     // "ACGT"[code]
        q[i] = v;
    }
    out[w] = make_uint4(q[0], q[1], q[2], q[3]);
}
__global__ void __launch_bounds__(256) km_unpack_inv_kernel(const uint64_t *__restrict__ inv, uint64_t n_inv, uint64_t base0, char *out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n_inv) out[inv[i] - base0] = 'N';                      // any byte outside ACGTacgt resets the k-mer run (read_label.cpp:943-950)
}
// 64-byte internal results -> the 32-byte records of the compact interface
__global__ void __launch_bounds__(256) km_compact_kernel(const kmat_read_result *__restrict__ in, uint32_t n, int want_lineage, kmat_read_result32 *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const kmat_read_result r = in[i];
    kmat_read_result32 o;
    o.tid = r.tid; o.score = r.score; o.log_avg = r.log_avg; o.stdev = r.stdev;
    o.list_off = (uint32_t)(want_lineage ? r.lin_off : r.cand_off);
    o.n_list = (uint16_t)min(want_lineage ? r.n_lin : r.n_cand, 0xFFFFu);
    o.valid_kmers = (uint16_t)min(max(r.valid_kmers, 0), 0xFFFF);
    o.cand_kmer_cnt = (uint16_t)r.cand_kmer_cnt;
    o.flags = (uint16_t)((r.status & 7) | ((r.match & 7) << 3) | ((r.bin_sel & 15) << 6) | (((-r.err) & 63) << 10));
    o.n_cand = r.n_cand;
    out[i] = o;
}
extern "C" void kmat_result_expand(const kmat_read_result32 *in, uint32_t read_len, int kmer_length, int min_kmer, int want_lineage, kmat_read_result *out) {
    memset(out, 0, sizeof *out);
    out->status = in->flags & 7; out->match = (in->flags >> 3) & 7; out->bin_sel = (in->flags >> 6) & 15; out->err = -(int32_t)((in->flags >> 10) & 63);
    out->tid = in->tid; out->score = in->score; out->log_avg = in->log_avg; out->stdev = in->stdev;
    out->valid_kmers = in->valid_kmers; out->cand_kmer_cnt = in->cand_kmer_cnt; out->n_cand = in->n_cand;
    if (want_lineage) { out->n_lin = in->n_list; out->lin_off = in->list_off; } else { out->n_cand = in->n_list; out->cand_off = in->list_off; }
    // the two integers of the ReadTooShort / NoDbHits lines (read_label.cpp:1217-1218, 1232-1233, 1270-1271)
    if (out->status == KMAT_ST_SHORT_LEN) { out->n1 = (int32_t)read_len; out->n2 = kmer_length; out->valid_kmers = 0; }
    else if (out->status == KMAT_ST_SHORT_VALID) { out->n1 = out->valid_kmers; out->n2 = min_kmer; }
    else if (out->status == KMAT_ST_NODBHITS) { out->n1 = (int32_t)read_len; out->n2 = kmer_length; }
}

// The pair list of a read in run-length form: equal scores are adjacent in rank_label (sorted by score) and in valid_cand
// (a lineage chain), and a 150-base read's ~9 pairs carry ~2.5 distinct scores.  Word stream: a taxid with bit 31 set is followed
// by the score (float bits) that holds for it and for the unflagged taxids after it.  8 B per pair -> ~5 B; the copy out is what
// bounds eight GPUs behind one host (profiles/r03g_pipe_trace_n8.txt).  One thread per read; a warp takes one piece of the
// pass's word buffer.  list_off of the 32-byte record becomes the read's word offset.
__global__ void __launch_bounds__(256) km_pack_lists_kernel(kmat_read_result32 *out, uint32_t n, const kmat_pair *__restrict__ pairs,
                                                            uint32_t *words, unsigned long long cap, unsigned long long *cursor) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t nl = 0; const kmat_pair *src = nullptr;
    if (i < n) { nl = out[i].n_list; src = pairs + out[i].list_off; }
    uint32_t len = 0, prev = 0;
    for (uint32_t j = 0; j < nl; j++) { const uint32_t b = __float_as_uint(src[j].score); len += (j == 0 || b != prev) ? 2u : 1u; prev = b; }
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(KM_FULL, incl, d); if (lane >= d) incl += v; }
    const uint32_t total = __shfl_sync(KM_FULL, incl, 31);
    unsigned long long at = 0;
    if (lane == 0 && total) at = atomicAdd(cursor, (unsigned long long)total);
    at = kb_shfl64(at, 0) + (incl - len);
    if (i >= n) return;
    out[i].list_off = (uint32_t)at;
    if (at + len > cap) return;                                  // reported through the cursor
    uint32_t *w = words + at;
    for (uint32_t j = 0; j < nl; j++) {
        const kmat_pair p = src[j];
        const uint32_t b = __float_as_uint(p.score);
        if (j == 0 || b != prev) { *w++ = p.tid | 0x80000000u; *w++ = b; } else *w++ = p.tid;
        prev = b;
    }
}
extern "C" uint32_t kmat_list_decode(const uint32_t *words, uint32_t n_list, kmat_pair *out) {
    const uint32_t *w = words;
    float sc = 0.0f;
    for (uint32_t j = 0; j < n_list; j++) {
        const uint32_t t = *w++;
        if (t & 0x80000000u) { memcpy(&sc, w, 4); w++; }
        out[j].tid = t & 0x7FFFFFFFu; out[j].score = sc;
    }
    return (uint32_t)(w - words);
}

#include "kmat_format.cuh"

// Host buffers in, host buffers out.  The batch is cut into chunks; chunk i's kernels (one stream, in order, so the
// candidate records of a chunk are contiguous behind one running cursor) overlap the H2D copy of chunk i+1 and the
// D2H copy of chunk i-1 on two copy streams.  Pinned caller buffers (kmat_host_alloc) make those copies truly
// asynchronous; pageable ones are staged by the driver and still overlap the kernels.
// Two front ends share it: ASCII reads + 64-byte results (kmat_label_batch), and the compact interface -- 2-bit packed reads
// (unpacked on the device: the kernels read the same byte buffer either way) + 32-byte results carrying ONE pair list.
struct KmHostIO {
    const char *bases = nullptr;                                        // ASCII front end
    const uint32_t *codes = nullptr; const uint64_t *inv = nullptr; uint64_t n_inv = 0;      // compact front end
    kmat_read_result *out = nullptr; kmat_read_result32 *out32 = nullptr;
    kmat_pair *cands = nullptr; uint64_t cands_cap = 0; uint64_t *n_cands = nullptr;
    kmat_pair *lineage = nullptr; uint64_t lineage_cap = 0; uint64_t *n_lineage = nullptr;
    // K5 (kmat_label_batch_text): the tails of the output lines, formatted on the device
    char *text = nullptr; uint64_t text_cap = 0; uint64_t *n_text = nullptr; uint64_t *text_ref = nullptr; int prn_all = 0;
    // compact front end, run-length lists (kmat_label_batch_packed_rl): the one pair list goes out as words instead of pairs
    uint32_t *plist = nullptr; uint64_t plist_cap = 0; uint64_t *n_plist = nullptr;
};
static int km_label_host(kmat_ctx *c, const KmHostIO &io, const uint64_t *offs, uint32_t n_reads) {
    const char *bases = io.bases;
    kmat_read_result *out = io.out;
    kmat_pair *cands = io.cands, *lineage = io.lineage;
    const uint64_t cands_cap = io.cands_cap, lineage_cap = io.lineage_cap;
    uint64_t *n_cands = io.n_cands, *n_lineage = io.n_lineage;
    const bool compact = io.codes != nullptr;
    if (c->opt.rkmer_mode) { kmat_set_error("kmat_label_batch: the ctx was created with rkmer_mode (kmat_null_* only)"); return KMAT_ERR_ARG; }
    if (c->db->shard_count > 1 && !c->d_peers) { kmat_set_error("kmat_label_batch: the table is shard %d of %d; attach the peers (kmat_ctx_peer_attach) or use the kmat_shard_* rounds", c->db->shard_index, c->db->shard_count); return KMAT_ERR_ARG; }
    if (n_cands) *n_cands = 0;
    if (n_lineage) *n_lineage = 0;
    if (!n_reads) return KMAT_OK;
    KM_CUDA(cudaSetDevice(c->device));
    const uint32_t chunk_reads = [] { const char *e = getenv("KMAT_CHUNK_READS"); const long v = e ? atol(e) : 0; return (uint32_t)(v > 0 ? v : (1 << 20)); }();
    const uint64_t chunk_bases = (uint64_t)256 << 20;
    int rc;
    if (io.n_text) *io.n_text = 0;
    if (io.text_ref) {
        // room for the device-formatted tails: what the caller can take, at most ~256 bytes per read (a tail without room is
        // left to the host formatter, never an error)
        const uint64_t want = std::min<uint64_t>(io.text_cap, (uint64_t)n_reads * 256 + 4096);
        if (want > c->cap_text) { KM_CUDA(cudaStreamSynchronize(c->stream)); KM_CUDA(cudaStreamSynchronize(c->st_d2h)); if ((rc = km_grow(&c->d_text, &c->cap_text, want)) != KMAT_OK) return rc; }
    }
    const uint64_t text_room = io.text_ref ? std::min<uint64_t>(io.text_cap, c->cap_text) : 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        if ((rc = km_reserve_cands(c, n_reads)) != KMAT_OK) return rc;
        if (io.plist) {                                   // at worst two words per pair
            const uint64_t want = 2 * (c->opt.want_lineage ? c->cap_lin : c->cap_cands) + 64;
            if (want > c->cap_plist) { KM_CUDA(cudaStreamSynchronize(c->stream)); KM_CUDA(cudaStreamSynchronize(c->st_d2h)); if ((rc = km_grow(&c->d_plist, &c->cap_plist, want)) != KMAT_OK) return rc; }
        }
        unsigned long long done_c = 0, done_l = 0, done_t = 0, done_p = 0;      // candidate / lineage pairs, text bytes and list words already copied out
        struct Chunk { uint32_t r0 = 0, r1 = 0; int slot = 0; bool valid = false; } prev;
        auto drain = [&](const Chunk &ch) -> int {      // results of a finished chunk -> caller buffers
            kmat_ctx::Slot &sl = c->slot[ch.slot];
            KM_CUDA(cudaEventSynchronize(sl.ev_comp));
            const unsigned long long cur_c = std::min<unsigned long long>(sl.h_cur[0], c->cap_cands), cur_l = std::min<unsigned long long>(sl.h_cur[1], c->cap_lin);
            if (compact) KM_CUDA(cudaMemcpyAsync(io.out32 + ch.r0, sl.d_out32, (size_t)(ch.r1 - ch.r0) * sizeof(kmat_read_result32), cudaMemcpyDeviceToHost, c->st_d2h));
            else KM_CUDA(cudaMemcpyAsync(out + ch.r0, sl.d_out, (size_t)(ch.r1 - ch.r0) * sizeof(kmat_read_result), cudaMemcpyDeviceToHost, c->st_d2h));
            if (cands && cur_c > done_c && done_c < cands_cap) {
                const unsigned long long hi = std::min<unsigned long long>(cur_c, cands_cap);
                KM_CUDA(cudaMemcpyAsync(cands + done_c, c->d_cands + done_c, (size_t)(hi - done_c) * sizeof(kmat_pair), cudaMemcpyDeviceToHost, c->st_d2h));
            }
            if (lineage && c->opt.want_lineage && cur_l > done_l && done_l < lineage_cap) {
                const unsigned long long hi = std::min<unsigned long long>(cur_l, lineage_cap);
                KM_CUDA(cudaMemcpyAsync(lineage + done_l, c->d_lin + done_l, (size_t)(hi - done_l) * sizeof(kmat_pair), cudaMemcpyDeviceToHost, c->st_d2h));
            }
            done_c = std::max(done_c, cur_c); done_l = std::max(done_l, cur_l);
            if (io.plist) {
                const unsigned long long cur_p = std::min<unsigned long long>(std::min<unsigned long long>(sl.h_cur[3], c->cap_plist), io.plist_cap);
                if (cur_p > done_p) KM_CUDA(cudaMemcpyAsync(io.plist + done_p, c->d_plist + done_p, (size_t)(cur_p - done_p) * 4, cudaMemcpyDeviceToHost, c->st_d2h));
                done_p = std::max(done_p, cur_p);
            }
            if (io.text_ref) {
                const unsigned long long cur_t = std::min<unsigned long long>(sl.h_cur[2], text_room);
                if (cur_t > done_t) KM_CUDA(cudaMemcpyAsync(io.text + done_t, c->d_text + done_t, (size_t)(cur_t - done_t), cudaMemcpyDeviceToHost, c->st_d2h));
                KM_CUDA(cudaMemcpyAsync(io.text_ref + ch.r0, sl.d_ref, (size_t)(ch.r1 - ch.r0) * 8, cudaMemcpyDeviceToHost, c->st_d2h));
                done_t = std::max(done_t, cur_t);
            }
            KM_CUDA(cudaEventRecord(sl.ev_d2h, c->st_d2h));
            return KMAT_OK;
        };
        uint32_t r0 = 0;
        int ci = 0;
        unsigned long long total_c = 0, total_l = 0;
        // KMAT_PIPE_TRACE=1 (debugging aid): per chunk, when its copies and kernels started and ended on the device and when
        // the host queued them -- printed to stderr after the call
        static const bool trace = getenv("KMAT_PIPE_TRACE") != nullptr;
        struct Tr { cudaEvent_t h0, h1, k0, k1, d0, d1; double t_submit, t_launched, t_drain0, t_drain1; };
        std::vector<Tr> tr;
        cudaEvent_t tr_base = nullptr;
        auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t_host0 = now_ms();
        if (trace) { cudaEventCreate(&tr_base); cudaEventRecord(tr_base, c->stream); }
        auto tr_ev = [&](cudaEvent_t *e, cudaStream_t st) { if (trace) { cudaEventCreate(e); cudaEventRecord(*e, st); } };
        while (r0 < n_reads) {
            // chunk [r0, r1): bounded by reads and by bases (at least one read).  The first and the last chunk of a large batch
            // are a quarter of the regular size: nothing overlaps the first chunk's copy in and the last chunk's copy out
            const uint32_t left = n_reads - r0, quarter = std::max(1u, chunk_reads / 4);
            uint32_t take = chunk_reads;
            if (n_reads > chunk_reads) {
                if (ci == 0) take = quarter;
                else if (left <= quarter) take = left;
                else if (left <= chunk_reads + quarter) take = left - quarter;
            }
            uint32_t r1 = std::min<uint64_t>(n_reads, (uint64_t)r0 + take);
            if (offs[r1] - offs[r0] > chunk_bases) {
                uint32_t lo = r0 + 1, hi = r1;                      // largest r1 with offs[r1] - offs[r0] <= chunk_bases
                while (lo < hi) { const uint32_t mid = lo + (hi - lo + 1) / 2; if (offs[mid] - offs[r0] <= chunk_bases) lo = mid; else hi = mid - 1; }
                r1 = lo;
            }
            const uint32_t n = r1 - r0;
            const uint64_t nb = offs[r1] - offs[r0];
            // compact front end: the chunk's byte buffer covers whole code words, i.e. starts at base 16 * w0
            const uint64_t w0 = offs[r0] / 16, w1 = (offs[r1] + 15) / 16, base0 = compact ? w0 * 16 : offs[r0];
            const uint64_t need_b = compact ? (w1 - w0) * 16 : nb;
            uint32_t max_len = 0;
            for (uint32_t r = r0; r < r1; r++) max_len = std::max<uint32_t>(max_len, (uint32_t)(offs[r + 1] - offs[r]));
            kmat_ctx::Slot &sl = c->slot[ci & 1];
            if (need_b + 1 > sl.cap_bases || n + 1 > sl.cap_reads || (compact && ((w1 - w0) > sl.cap_codes || !sl.d_out32))) {
                // growing a slot: everything queued on it must have finished
                KM_CUDA(cudaStreamSynchronize(c->st_h2d)); KM_CUDA(cudaStreamSynchronize(c->stream)); KM_CUDA(cudaStreamSynchronize(c->st_d2h));
                if (need_b + 1 > sl.cap_bases) { if ((rc = km_grow(&sl.d_bases, &sl.cap_bases, need_b + 1)) != KMAT_OK) return rc; }
                if (compact && (w1 - w0) > sl.cap_codes) { if ((rc = km_grow(&sl.d_codes, &sl.cap_codes, w1 - w0)) != KMAT_OK) return rc; }
                if (n + 1 > sl.cap_reads || (compact && !sl.d_out32)) {
                    cudaFree(sl.d_offs); cudaFree(sl.d_out); cudaFree(sl.d_out32); sl.d_offs = nullptr; sl.d_out = nullptr; sl.d_out32 = nullptr;
                    const size_t cap = std::max<size_t>((size_t)n + n / 4 + 64, sl.cap_reads);
                    KM_CUDA(cudaMalloc((void **)&sl.d_offs, (cap + 1) * 8));
                    KM_CUDA(cudaMalloc((void **)&sl.d_out, cap * sizeof(kmat_read_result)));
                    if (compact) KM_CUDA(cudaMalloc((void **)&sl.d_out32, cap * sizeof(kmat_read_result32)));
                    sl.cap_reads = (uint32_t)cap;
                }
            }
            if (io.text_ref && (uint64_t)n > sl.cap_ref) {
                KM_CUDA(cudaStreamSynchronize(c->st_h2d)); KM_CUDA(cudaStreamSynchronize(c->stream)); KM_CUDA(cudaStreamSynchronize(c->st_d2h));
                if ((rc = km_grow(&sl.d_ref, &sl.cap_ref, (uint64_t)n + n / 4 + 64)) != KMAT_OK) return rc;
            }
            uint64_t i_lo = 0, i_hi = 0;
            if (compact && io.n_inv) {
                i_lo = (uint64_t)(std::lower_bound(io.inv, io.inv + io.n_inv, offs[r0]) - io.inv);
                i_hi = (uint64_t)(std::lower_bound(io.inv, io.inv + io.n_inv, offs[r1]) - io.inv);
                if (i_hi - i_lo > sl.cap_inv) {
                    KM_CUDA(cudaStreamSynchronize(c->st_h2d)); KM_CUDA(cudaStreamSynchronize(c->stream));
                    if ((rc = km_grow(&sl.d_inv, &sl.cap_inv, i_hi - i_lo)) != KMAT_OK) return rc;
                }
            }
            // H2D once the kernels that last read this slot's inputs are done
            if (ci >= 2) KM_CUDA(cudaStreamWaitEvent(c->st_h2d, sl.ev_comp, 0));
            if (trace) { tr.push_back(Tr{}); tr.back().t_submit = now_ms() - t_host0; tr_ev(&tr.back().h0, c->st_h2d); }
            if (compact) {
                KM_CUDA(cudaMemcpyAsync(sl.d_codes, io.codes + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, c->st_h2d));
                if (i_hi > i_lo) KM_CUDA(cudaMemcpyAsync(sl.d_inv, io.inv + i_lo, (size_t)(i_hi - i_lo) * 8, cudaMemcpyHostToDevice, c->st_h2d));
            } else KM_CUDA(cudaMemcpyAsync(sl.d_bases, bases + offs[r0], nb, cudaMemcpyHostToDevice, c->st_h2d));
            KM_CUDA(cudaMemcpyAsync(sl.d_offs, offs + r0, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->st_h2d));
            KM_CUDA(cudaEventRecord(sl.ev_h2d, c->st_h2d));
            if (trace) tr_ev(&tr.back().h1, c->st_h2d);
            // kernels once the inputs are in and the previous results of this slot have been copied out
            KM_CUDA(cudaStreamWaitEvent(c->stream, sl.ev_h2d, 0));
            if (ci >= 2) KM_CUDA(cudaStreamWaitEvent(c->stream, sl.ev_d2h, 0));
            if (trace) tr_ev(&tr.back().k0, c->stream);
            if (compact) {
                km_unpack_kernel<<<(unsigned)((w1 - w0 + 255) / 256), 256, 0, c->stream>>>(sl.d_codes, w1 - w0, (uint4 *)sl.d_bases);
                g_km_launches++;
                if (i_hi > i_lo) { km_unpack_inv_kernel<<<(unsigned)((i_hi - i_lo + 255) / 256), 256, 0, c->stream>>>(sl.d_inv, i_hi - i_lo, base0, sl.d_bases); g_km_launches++; }
                KM_CUDA(cudaGetLastError());
            }
            KmPass L{sl.d_bases - base0, sl.d_offs, n, offs[r0], nb, max_len, sl.d_out, ci == 0};
            if ((rc = km_run_device(c, L, c->stream)) != KMAT_OK) return rc;
            if (io.text_ref) {
                // K5 sees the chunk's final results; its text lands behind the previous chunks' in c->d_text
                if ((rc = km_launch_format(c, sl.d_out, n, io.prn_all, text_room, sl.d_ref, c->stream)) != KMAT_OK) return rc;
            }
            if (compact) {
                km_compact_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(sl.d_out, n, c->opt.want_lineage, sl.d_out32);
                g_km_launches++;
                KM_CUDA(cudaGetLastError());
                if (io.plist) {
                    km_pack_lists_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(sl.d_out32, n, c->opt.want_lineage ? c->d_lin : c->d_cands, c->d_plist, c->cap_plist, c->d_cursors + 3);
                    g_km_launches++;
                    KM_CUDA(cudaGetLastError());
                }
            }
            KM_CUDA(cudaMemcpyAsync(sl.h_cur, c->d_cursors, 32, cudaMemcpyDeviceToHost, c->stream));
            KM_CUDA(cudaEventRecord(sl.ev_comp, c->stream));
            if (trace) { tr_ev(&tr.back().k1, c->stream); tr.back().t_launched = now_ms() - t_host0; }
            if (prev.valid) {
                if (trace) { tr[ci - 1].t_drain0 = now_ms() - t_host0; tr_ev(&tr[ci - 1].d0, c->st_d2h); }
                if ((rc = drain(prev)) != KMAT_OK) return rc;
                if (trace) { tr[ci - 1].t_drain1 = now_ms() - t_host0; tr_ev(&tr[ci - 1].d1, c->st_d2h); }
            }
            prev.r0 = r0; prev.r1 = r1; prev.slot = ci & 1; prev.valid = true;
            r0 = r1; ci++;
        }
        if (trace) { tr[ci - 1].t_drain0 = now_ms() - t_host0; tr_ev(&tr[ci - 1].d0, c->st_d2h); }
        if ((rc = drain(prev)) != KMAT_OK) return rc;
        if (trace) {
            tr[ci - 1].t_drain1 = now_ms() - t_host0; tr_ev(&tr[ci - 1].d1, c->st_d2h);
            cudaStreamSynchronize(c->st_d2h); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->st_h2d);
            fprintf(stderr, "[kmat pipe trace] %d chunks, host total %.2f ms; device times relative to the call's first event, host times to its start\n", ci, now_ms() - t_host0);
            fprintf(stderr, "chunk   h2d0   h2d1 |  krn0   krn1 |  d2h0   d2h1 || submit launched drain0 drain1\n");
            for (int i = 0; i < ci; i++) {
                float v[6] = {0, 0, 0, 0, 0, 0};
                cudaEvent_t ev[6] = {tr[i].h0, tr[i].h1, tr[i].k0, tr[i].k1, tr[i].d0, tr[i].d1};
                for (int j = 0; j < 6; j++) { cudaEventElapsedTime(&v[j], tr_base, ev[j]); cudaEventDestroy(ev[j]); }
                fprintf(stderr, "%5d %6.2f %6.2f | %6.2f %6.2f | %6.2f %6.2f || %6.2f %6.2f %6.2f %6.2f\n", i, v[0], v[1], v[2], v[3], v[4], v[5], tr[i].t_submit, tr[i].t_launched, tr[i].t_drain0, tr[i].t_drain1);
            }
            cudaEventDestroy(tr_base);
        }
        total_c = c->slot[prev.slot].h_cur[0]; total_l = c->slot[prev.slot].h_cur[1];
        KM_CUDA(cudaStreamSynchronize(c->st_d2h));
        bool again = false;
        if (total_c > c->cap_cands) { KM_CUDA(cudaStreamSynchronize(c->stream)); if ((rc = km_grow_cands(c, total_c)) != KMAT_OK) return rc; again = true; }
        if (c->opt.want_lineage && total_l > c->cap_lin) { KM_CUDA(cudaStreamSynchronize(c->stream)); if ((rc = km_grow(&c->d_lin, &c->cap_lin, total_l)) != KMAT_OK) return rc; again = true; }
        if (c->d_peers) {                        // direct sharded mode: list-record buffer too small -> doubled by km_peer_check, re-run
            KM_CUDA(cudaStreamSynchronize(c->stream));
            const int prc = km_peer_check(c);
            if (prc == KMAT_ERR_OVERFLOW && attempt < 2) again = true; else if (prc != KMAT_OK) return prc;
        }
        if (again) continue;                     // candidate buffer was too small: re-run with the exact size
        if (n_cands) *n_cands = total_c;
        if (n_lineage) *n_lineage = c->opt.want_lineage ? total_l : 0;
        if (io.n_text) *io.n_text = done_t;
        if (io.plist) {
            const unsigned long long total_p = c->slot[prev.slot].h_cur[3];
            if (io.n_plist) *io.n_plist = total_p;
            if (total_p >= (1ull << 32)) { kmat_set_error("compact interface: more than 2^32 list words in one call; split the batch"); return KMAT_ERR_UNSUPPORTED; }
            if (total_p > io.plist_cap) { kmat_set_error("list buffer too small: need %llu words", total_p); return KMAT_ERR_OVERFLOW; }
        }
        if ((rc = km_fetch_stats(c, c->stream)) != KMAT_OK) return rc;
        if (compact && std::max(total_c, total_l) >= (1ull << 32)) { kmat_set_error("compact interface: more than 2^32 pairs in one call; split the batch"); return KMAT_ERR_UNSUPPORTED; }
        if ((cands && total_c > cands_cap) || (lineage && c->opt.want_lineage && total_l > lineage_cap)) {
            kmat_set_error("candidate buffer too small: need %llu candidate and %llu lineage pairs", total_c, total_l);
            return KMAT_ERR_OVERFLOW;
        }
        return KMAT_OK;
    }
    kmat_set_error("candidate buffer kept overflowing");
    return KMAT_ERR_CUDA;
}
extern "C" int kmat_label_batch(kmat_ctx *c, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                                kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands, kmat_pair *lineage, uint64_t lineage_cap,
                                uint64_t *n_lineage) {
    if (!c || !offs || !out || (n_reads && !bases)) { kmat_set_error("kmat_label_batch: bad argument"); return KMAT_ERR_ARG; }
    KmHostIO io;
    io.bases = bases; io.out = out; io.cands = cands; io.cands_cap = cands_cap; io.n_cands = n_cands;
    io.lineage = lineage; io.lineage_cap = lineage_cap; io.n_lineage = n_lineage;
    return km_label_host(c, io, offs, n_reads);
}
// kmat_label_batch + K5: the tail of every read's output line (what kmat_format_tail writes, byte for byte) formatted on the
// device.  text_ref[i] = offset << KMAT_TEXT_LEN_BITS | length into `text`, or KMAT_TEXT_ON_HOST for the reads the device
// formatter leaves to kmat_format_tail (a number in exponent notation, a very long tail, no room left in `text`, an error).
extern "C" int kmat_label_batch_text(kmat_ctx *c, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                                     kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands, kmat_pair *lineage, uint64_t lineage_cap,
                                     uint64_t *n_lineage, int prn_all, char *text, uint64_t text_cap, uint64_t *n_text, uint64_t *text_ref) {
    if (!c || !offs || !out || (n_reads && !bases) || (n_reads && (!text_ref || (!text && text_cap)))) { kmat_set_error("kmat_label_batch_text: bad argument"); return KMAT_ERR_ARG; }
    if (prn_all && !cands) { kmat_set_error("kmat_label_batch_text: prn_all needs the candidate buffer"); return KMAT_ERR_ARG; }
    KmHostIO io;
    io.bases = bases; io.out = out; io.cands = cands; io.cands_cap = cands_cap; io.n_cands = n_cands;
    io.lineage = lineage; io.lineage_cap = lineage_cap; io.n_lineage = n_lineage;
    io.text = text; io.text_cap = text_cap; io.n_text = n_text; io.text_ref = text_ref; io.prn_all = prn_all;
    return km_label_host(c, io, offs, n_reads);
}
// The compact interface: reads as 2-bit code words indexed by GLOBAL base offset (word w = bases 16 w .. 16 w + 15, base i in
// bits 2 (i % 16) .. of its word) + the ascending offsets of the non-ACGT bases (kmat_pack_reads writes both); results as
// 32-byte records with ONE pair list each -- rank_label when the ctx has want_lineage == 0 (the -p line), else valid_cand.
extern "C" int kmat_label_batch_packed(kmat_ctx *c, const uint32_t *codes, const uint64_t *inv_pos, uint64_t n_inv, const uint64_t *offs, uint32_t n_reads,
                                       kmat_read_result32 *out, kmat_pair *list, uint64_t list_cap, uint64_t *n_list) {
    if (!c || !offs || !out || (n_reads && !codes) || (n_inv && !inv_pos)) { kmat_set_error("kmat_label_batch_packed: bad argument"); return KMAT_ERR_ARG; }
    KmHostIO io;
    io.codes = codes; io.inv = inv_pos; io.n_inv = n_inv; io.out32 = out;
    if (c->opt.want_lineage) { io.lineage = list; io.lineage_cap = list_cap; io.n_lineage = n_list; }
    else { io.cands = list; io.cands_cap = list_cap; io.n_cands = n_list; }
    return km_label_host(c, io, offs, n_reads);
}

// kmat_label_batch_packed with the pair list in run-length form (km_pack_lists_kernel): list_off of a record is the offset of the
// read's words in `words`, n_list the number of pairs kmat_list_decode rebuilds from them.
extern "C" int kmat_label_batch_packed_rl(kmat_ctx *c, const uint32_t *codes, const uint64_t *inv_pos, uint64_t n_inv, const uint64_t *offs, uint32_t n_reads,
                                          kmat_read_result32 *out, uint32_t *words, uint64_t words_cap, uint64_t *n_words) {
    if (!c || !offs || !out || (n_reads && !codes) || (n_inv && !inv_pos) || (n_reads && !words)) { kmat_set_error("kmat_label_batch_packed_rl: bad argument"); return KMAT_ERR_ARG; }
    if (c->max_tid & 0x80000000u) { kmat_set_error("kmat_label_batch_packed_rl: the taxonomy holds a taxid of 2^31 or more (the run-length form uses bit 31); use kmat_label_batch_packed"); return KMAT_ERR_UNSUPPORTED; }
    KmHostIO io;
    io.codes = codes; io.inv = inv_pos; io.n_inv = n_inv; io.out32 = out;
    io.plist = words; io.plist_cap = words_cap; io.n_plist = n_words;
    return km_label_host(c, io, offs, n_reads);
}

// Page-locked host memory for the buffers handed to kmat_label_batch (optional: pageable buffers work, pinned
// ones make the copies asynchronous).
extern "C" void *kmat_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void kmat_host_free(void *p) { if (p) cudaFreeHost(p); }

#include "kmat_shard.cuh"
#include "kmat_comm.cuh"
#include "kmat_gene.cuh"
#include "kmat_null.cuh"
