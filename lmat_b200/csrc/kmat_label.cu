// kmat_label.cu -- kmat_ctx and the per-read scoring kernel (K3 candidate sets + K4 scoring / LCA).
//
// Restates, per read, retrieve_kmer_labels' list handling and post-pass (read_label.cpp:1031-1204),
// construct_labels (:692-941) and findReadLabelVer2 (:284-419).  One warp per read: lanes own k-mer
// positions for the data-parallel parts; the order-dependent float arithmetic and the std::sort / LCA walk run
// on lane 0 from shared memory, in the reference's own operation order (no FMA contraction: every float op is
// an explicit __f*_rn intrinsic; logf is the glibc algorithm, km_logf).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kmat_device.cuh"
#include "kmat_priv.h"
#include "kmat_std_emul.cuh"

#define KB_WARPS 8         // warps per CTA of the candidate kernel
#define KB_PMAX 320        // k-mer positions per read handled by the shared-memory path
#define KB_CMAX 64         // candidate taxids per read (one 64-bit mask)
#define KB_HSLOTS 128
#define KB_LFAST 16        // list length sorted in place per lane (libstdc++ uses plain insertion sort up to 16)
#define KB_LIN 128         // lineage entries in findReadLabelVer2
#define KB_BIGCAP 8192     // per-warp global scratch for longer lists / run-time pruning
#define KS_THREADS 128     // threads per CTA of the scoring kernel (one read per thread)
#define KMAT_ST_PENDING 7  // internal: candidates built, scoring still to run

// stored id -> node: low 30 bits nid, bit 31 = isHuman, bit 30 = dropped tid (flags folded in so that the common
// singleton hit needs no node-record load)
#define KB_SID_HUMAN 0x80000000u
#define KB_SID_DROP 0x40000000u
#define KB_SID_NIDMASK 0x3FFFFFFFu

struct KmCtxDev {
    KmDbDev db;
    const KmNodeA *nodeA; const KmNodeB *nodeB; const uint32_t *paths; const uint32_t *prune_rank; const uint32_t *sid2nid;
    uint32_t n_sid, n_nodes, nid_human, nid_one;
    int nbins, n_models, n_classes;
    const int16_t *model_of_cand; const int32_t *mrow; const float *cut; const uint8_t *cls;
    int8_t class_ranknum[64];
    kmat_opts opt;
};

struct KmScoreParams {
    KmCtxDev C;
    const uint64_t *offs; uint32_t n_reads; const uint32_t *hit; const int2 *hdr;
    kmat_read_result *out;
    kmat_pair *cands; unsigned long long *cand_cursor; unsigned long long cand_cap;
    kmat_pair *lin; unsigned long long *lin_cursor; unsigned long long lin_cap;
    uint2 *big_scratch;
    KmStatsDev *stats;
};

struct KmRl { float score; uint32_t idx; };          // rank_label element: candidate index + (bias-adjusted) score

struct __align__(16) KmWarpB {
    // candidate hash: nid -> slot; h_idx[slot] = candidate id (dense, in insertion order)
    uint32_t h_nid[KB_HSLOTS], h_seq[KB_HSLOTS], h_leaf[KB_HSLOTS];
    uint8_t h_idx[KB_HSLOTS];
    // candidates, indexed by candidate id.  order[f] = id of the f-th entry of the reference's taxid_lst.
    uint32_t c_nid[KB_CMAX], c_tid[KB_CMAX], c_meta[KB_CMAX], c_spec[KB_CMAX], c_poff[KB_CMAX], c_plen[KB_CMAX];
    uint32_t c_leaf[KB_CMAX], c_first[KB_CMAX], c_hits[KB_CMAX];
    unsigned long long c_anc[KB_CMAX];
    uint8_t c_qual[KB_CMAX], c_slot[KB_CMAX], order[KB_CMAX];
    // per position: bit set of the candidate ids kept there (label_vec[pos].second before the post-pass)
    unsigned long long posmask[KB_PMAX];
    uint32_t lst[32][KB_LFAST];          // position loop: per-lane list scratch
    uint16_t dep[32][KB_LFAST];
    uint32_t n_used;
};

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t kb_hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; return x & (KB_HSLOTS - 1); }
// insert-or-find; returns slot or -1 when the table would exceed KB_CMAX distinct keys.  The inserting lane
// assigns the dense candidate id; other lanes may read h_idx[slot] only after a __syncwarp.
__device__ int kb_cand_insert(KmWarpB &S, uint32_t nid) {
    uint32_t h = kb_hash(nid);
    for (int step = 0; step < KB_HSLOTS; step++) {
        const uint32_t cur = ((volatile uint32_t *)S.h_nid)[h];
        if (cur == nid) return (int)h;
        if (cur == KMAT_NONE) {
            if (((volatile uint32_t *)&S.n_used)[0] >= KB_CMAX) return -1;
            const uint32_t old = atomicCAS(&S.h_nid[h], KMAT_NONE, nid);
            if (old == KMAT_NONE) {
                const uint32_t id = atomicAdd(&S.n_used, 1u);
                if (id >= KB_CMAX) return -1;
                S.h_idx[h] = (uint8_t)id; S.c_slot[id] = (uint8_t)h;
                return (int)h;
            }
            if (old == nid) return (int)h;
        }
        h = (h + 1) & (KB_HSLOTS - 1);
    }
    return -1;
}
__device__ __forceinline__ int kb_cand_find(const KmWarpB &S, uint32_t nid) {
    uint32_t h = kb_hash(nid);
    for (int step = 0; step < KB_HSLOTS; step++) {
        const uint32_t cur = S.h_nid[h];
        if (cur == nid) return (int)h;
        if (cur == KMAT_NONE) return -1;
        h = (h + 1) & (KB_HSLOTS - 1);
    }
    return -1;
}
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
__device__ __forceinline__ uint32_t kb_list_count(const KmDbDev &db, uint32_t off) {
    return db.tid_bytes == 2 ? (uint32_t)(*(const uint16_t *)(db.pool + off)) : db.pool[off];
}
__device__ __forceinline__ uint32_t kb_list_id(const KmDbDev &db, uint32_t off, uint32_t j) {
    return db.tid_bytes == 2 ? (uint32_t)((const uint16_t *)(db.pool + off))[1 + j] : db.pool[off + 1 + j];
}

// Leaf filter (read_label.cpp:1103-1134): ids sorted by depth descending; keep an id unless it is a strict
// ancestor of an id kept before it.  In place; returns the number kept.
__device__ int kb_leaf_filter(const KmCtxDev &C, uint32_t *ids, int n) {
    if (n <= 1) return n;
    int m = 1;                                   // the deepest id is always kept
    uint32_t first_tin = kb_nodeB(C, ids[0]).tin;
    for (int i = 1; i < n; i++) {
        const uint32_t t = ids[i];
        const KmNodeB tb = kb_nodeB(C, t);
        bool anc = kb_is_anc(tb.tin, tb.tout, first_tin);
        for (int j = 1; j < m && !anc; j++) anc = kb_is_anc(tb.tin, tb.tout, kb_nodeB(C, ids[j]).tin);
        if (!anc) ids[m++] = t;
    }
    return m;
}

struct KbDepthDesc {      // CmpDepth1 (read_label.cpp:169-177) over {nid, depth} pairs
    __device__ bool operator()(const uint2 &a, const uint2 &b) const { return (int)a.y > (int)b.y; }
};
struct KbRankLess {       // MyPair::operator< (SortedDb.hpp:133-135): by rank number only
    __device__ bool operator()(const uint2 &a, const uint2 &b) const { return a.y < b.y; }
};

// Long list / run-time pruning path, one lane, per-warp global scratch.  Restates TaxNodeStat::begin + next
// (TaxNodeStat.hpp:60-256) followed by the filters of read_label.cpp:1031-1074.  Returns kept member count
// (members left in scratch[0..m).x) or <0 on error (-1 bad taxid, -2 list too long for the scratch).
__device__ int kb_big_list(const KmCtxDev &C, uint32_t hw, uint2 *scratch) {
    const uint32_t lo = hw & 0x7FFFFFFFu;
    int count = (int)kb_list_count(C.db, lo);
    if (count > KB_BIGCAP / 2) return -2;
    uint2 *seq = scratch;                  // ids in next() order
    uint2 *heap = scratch + KB_BIGCAP / 2;
    int n = 0;
    const int tid_cut = C.opt.max_count;
    if (tid_cut > 0 && count > tid_cut) {
        if (!C.prune_rank) {               // p_map.size() == 0: count forced to 1, next() reads the first stored id (:78-81)
            count = 1;
            const uint32_t sid = kb_list_id(C.db, lo, 0);
            const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
            if (e == KMAT_NONE) return -1;
            seq[n++] = make_uint2(e & KB_SID_NIDMASK, 0);
        } else {                           // :118-201
            int hn = 0;
            for (int i = 0; i < count; i++) {
                const uint32_t sid = kb_list_id(C.db, lo, i);
                const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
                if (e == KMAT_NONE) return -1;
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
            count = newcount;
            for (int i = 0; i < count; i++) seq[n++] = kmstd::pq_pop(heap, hn, KbRankLess());
        }
    } else {
        for (int i = 0; i < count; i++) {
            const uint32_t sid = kb_list_id(C.db, lo, i);
            const uint32_t e = sid < C.n_sid ? C.sid2nid[sid] : KMAT_NONE;
            if (e == KMAT_NONE) return -1;
            seq[n++] = make_uint2(e & KB_SID_NIDMASK, 0);
        }
    }
    // human collapse, dropped ids, depth lookup (read_label.cpp:1031-1066)
    bool seenHuman = false;
    int w = 0;
    for (int i = 0; i < n; i++) {
        uint32_t nid = seq[i].x;
        uint32_t meta = kb_nodeA(C, nid).meta;
        if (meta & KM_META_HUMAN) {
            if (seenHuman) continue;
            nid = C.nid_human; meta = kb_nodeA(C, nid).meta; seenHuman = true;
        }
        if (meta & KM_META_DROP) continue;
        seq[w++] = make_uint2(nid, meta & KM_META_DEPTH_MASK);
    }
    kmstd::sort(seq, w, KbDepthDesc());                                  // :1073-1074
    // leaf filter
    int m = 0;
    for (int i = 0; i < w; i++) {
        const uint32_t t = seq[i].x;
        const KmNodeB tb = kb_nodeB(C, t);
        bool anc = false;
        for (int j = 0; j < m && !anc; j++) anc = kb_is_anc(tb.tin, tb.tout, kb_nodeB(C, seq[j].x).tin);
        if (!anc) seq[m++] = seq[i];
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// K3: candidate sets.  One warp per read.  Restates the list handling and the post-pass of retrieve_kmer_labels
// (read_label.cpp:1031-1204) and the per-taxid position counts of construct_labels (:748-759).  Leaves, per
// read, the candidates in taxid_lst order as (nid, hits) pairs in the cands buffer for the scoring kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KB_WARPS * 32) km_cand_kernel(KmScoreParams P) {
    extern __shared__ __align__(16) unsigned char kb_smem[];
    KmWarpB &S = reinterpret_cast<KmWarpB *>(kb_smem)[threadIdx.x >> 5];
    const KmCtxDev &X = P.C;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = blockIdx.x * KB_WARPS + (threadIdx.x >> 5), n_warps = gridDim.x * KB_WARPS;
    const int k = X.db.kmer_len;
    uint2 *big = P.big_scratch + (size_t)warp_global * KB_BIGCAP;
    const uint32_t lt_mask = (1u << lane) - 1;
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
        else if (np > KB_PMAX) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; done = true; st_err++; }
        if (done) { if (lane == 0) P.out[r] = res; continue; }

        for (int i = lane; i < KB_HSLOTS; i += 32) { S.h_nid[i] = KMAT_NONE; S.h_seq[i] = 0xFFFFFFFFu; S.h_leaf[i] = 0; }
        if (lane == 0) S.n_used = 0;
        __syncwarp();
        int cand_cnt = 0, fnd_cnt = 0, err = 0;
        bool overflow = false;
        // ---- per position: list -> filtered, depth-sorted, leaf-filtered members (read_label.cpp:1019-1134)
        for (int p0 = 0; p0 < np; p0 += 32) {
            const int p = p0 + lane;
            const uint32_t hw = p < np ? P.hit[off + p] : KM_HIT_INVALID;
            uint32_t *L = S.lst[lane];
            uint16_t *D = S.dep[lane];
            int m = 0;
            bool bigl = false;
            uint32_t single = KMAT_NONE;                                      // the kept member when there is exactly one
            if (hw != KM_HIT_INVALID) {
                cand_cnt++;                                                   // label_vec[pos].first >= 0 (:1015, :702)
                if (hw != KM_HIT_MISS) {
                    if (!(hw & KM_HIT_LIST)) {
                        // singleton: one stored id.  16->32 conversion, human collapse, dropped tids (:1031-1038)
                        const uint32_t e = hw < X.n_sid ? X.sid2nid[hw] : KMAT_NONE;
                        if (e == KMAT_NONE) err = KMAT_ERR_BAD_TAXID;          // "bad taxid" assert (TaxNodeStat.hpp:235-238)
                        else if (!(e & KB_SID_DROP)) { single = (e & KB_SID_HUMAN) ? X.nid_human : (e & KB_SID_NIDMASK); m = 1; }
                    } else {
                        const uint32_t lo = hw & 0x7FFFFFFFu;
                        const int n_raw = (int)kb_list_count(X.db, lo);
                        st_list_ids += n_raw;
                        st_list_sectors += (2 + n_raw * X.db.tid_bytes + 31) / 32;
                        if (n_raw > KB_LFAST || n_raw > X.opt.max_count) bigl = true;
                        else {
                            bool seenHuman = false;
                            int n = 0;
                            for (int j = 0; j < n_raw; j++) {
                                const uint32_t sid = kb_list_id(X.db, lo, j);
                                const uint32_t e = sid < X.n_sid ? X.sid2nid[sid] : KMAT_NONE;
                                if (e == KMAT_NONE) { err = KMAT_ERR_BAD_TAXID; break; }
                                if (e & KB_SID_DROP) continue;                                       // :1038
                                uint32_t nid = e & KB_SID_NIDMASK;
                                if (e & KB_SID_HUMAN) {                                               // :1033-1037
                                    if (seenHuman) continue;
                                    nid = X.nid_human; seenHuman = true;
                                }
                                // insertion into depth-descending order == std::sort's insertion sort for n <= 16 (stable)
                                const uint16_t dpt = (uint16_t)(kb_nodeA(X, nid).meta & KM_META_DEPTH_MASK);
                                int q = n;
                                while (q > 0 && D[q - 1] < dpt) { L[q] = L[q - 1]; D[q] = D[q - 1]; q--; }
                                L[q] = nid; D[q] = dpt;
                                n++;
                            }
                            m = err ? 0 : kb_leaf_filter(X, L, n);
                            if (m == 1) single = L[0];
                        }
                    }
                }
            }
            // insert the kept members (taxid_lst / leaf_track bookkeeping, :1111-1122)
            // (a) positions with exactly one member: lanes holding the same taxid elect a leader
            {
                const uint32_t smask = __ballot_sync(KM_FULL, m == 1);
                if (m == 1) {
                    const uint32_t grp = __match_any_sync(smask, single);
                    const int leader = __ffs(grp) - 1;
                    int slot = 0;
                    if (lane == leader) {
                        slot = kb_cand_insert(S, single);
                        if (slot >= 0) {
                            atomicMin(&S.h_seq[slot], (uint32_t)p << 16);             // first appearance: lowest position of the group
                            atomicAdd(&S.h_leaf[slot], (uint32_t)__popc(grp));        // leaf_track
                        }
                    }
                    slot = __shfl_sync(grp, slot, leader);
                    if (slot < 0) { overflow = true; m = 0; } else L[0] = (uint32_t)slot;
                    fnd_cnt++;
                }
            }
            // (b) positions with several members
            if (m > 1) {
                fnd_cnt++;
                for (int j = 0; j < m; j++) {
                    const int slot = kb_cand_insert(S, L[j]);
                    if (slot < 0) { overflow = true; m = 0; break; }
                    atomicMin(&S.h_seq[slot], ((uint32_t)p << 16) | (uint32_t)j);   // first appearance in taxid_lst order
                    atomicAdd(&S.h_leaf[slot], 1u);
                    L[j] = (uint32_t)slot;
                }
            }
            __syncwarp();                                                     // candidate ids of this chunk's inserts are visible
            if (p < np) {
                unsigned long long mask = 0;
                for (int j = 0; j < m; j++) mask |= 1ull << S.h_idx[L[j]];
                S.posmask[p] = mask;
            }
            // (c) long lists / run-time pruning: one lane at a time through the per-warp global scratch
            uint32_t bigmask = __ballot_sync(KM_FULL, bigl);
            while (bigmask) {
                const int src = __ffs(bigmask) - 1;
                bigmask &= bigmask - 1;
                if (lane == src) {
                    const int mm = kb_big_list(X, hw, big);
                    if (mm == -1) err = KMAT_ERR_BAD_TAXID;
                    else if (mm < 0) overflow = true;
                    else if (mm > 0) {
                        unsigned long long mask = 0;
                        for (int j = 0; j < mm; j++) {
                            const int slot = kb_cand_insert(S, big[j].x);
                            if (slot < 0) { overflow = true; break; }
                            atomicMin(&S.h_seq[slot], ((uint32_t)p << 16) | (uint32_t)min(j, 0xFFFF));
                            atomicAdd(&S.h_leaf[slot], 1u);
                            mask |= 1ull << ((volatile uint8_t *)S.h_idx)[slot];   // inserted by this lane now, or before the last __syncwarp
                        }
                        S.posmask[p] = mask; fnd_cnt++;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        cand_cnt = km_warp_sum(cand_cnt); fnd_cnt = km_warp_sum(fnd_cnt);
        err = __reduce_max_sync(KM_FULL, err < 0 ? -err : 0);
        overflow = __any_sync(KM_FULL, overflow);
        if (err) { res.status = KMAT_ST_ERROR; res.err = -err; if (lane == 0) P.out[r] = res; st_err++; continue; }
        if (overflow) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; st_err++; continue; }
        const int C1 = (int)S.n_used;
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
        // ---- taxid_lst order of the first C1 candidates = order of first appearance (position, then list order)
        for (int i = lane; i < C1; i += 32) {
            const int s = S.c_slot[i];
            const uint32_t q = S.h_seq[s];
            int rank = 0;
            for (int j = 0; j < C1; j++) rank += S.h_seq[S.c_slot[j]] < q;
            S.order[rank] = (uint8_t)i;
            const uint32_t nid = S.h_nid[s];
            const KmNodeA na = kb_nodeA(X, nid); const KmNodeB nb = kb_nodeB(X, nid);
            S.c_nid[i] = nid; S.c_tid[i] = na.tid; S.c_meta[i] = na.meta; S.c_spec[i] = na.species_anc;
            S.c_poff[i] = nb.path_off; S.c_plen[i] = nb.path_len;
            S.c_leaf[i] = S.h_leaf[s]; S.c_first[i] = q >> 16; S.c_anc[i] = 0ull;
        }
        for (int i = lane; i < KB_CMAX; i += 32) S.c_hits[i] = 0;
        __syncwarp();
        // ---- representative strain per species (:1143-1177) -> which members get their lineage added
        for (int i = lane; i < C1; i += 32) {
            const uint32_t rk = (S.c_meta[i] >> KM_META_RANK_SHIFT) & 3;
            uint8_t qual = 1;
            if (rk == 1) {                                                    // gRank_table[tid] == "strain"
                bool rep = false;
                const uint32_t sp = S.c_spec[i];
                if (sp != KMAT_NONE) {
                    rep = true;
                    for (int j = 0; j < C1 && rep; j++) {
                        if (j == i || ((S.c_meta[j] >> KM_META_RANK_SHIFT) & 3) != 1 || S.c_spec[j] != sp) continue;
                        if (S.c_leaf[j] > S.c_leaf[i] || (S.c_leaf[j] == S.c_leaf[i] && S.c_tid[j] < S.c_tid[i])) rep = false;
                    }
                }
                qual = rep;
            }
            S.c_qual[i] = qual;
        }
        __syncwarp();
        // ---- lineage expansion (:1178-1203): qualifying members in (first position, taxid) order; the ancestors
        //      appended to taxid_lst get the next candidate ids, so for them id == taxid_lst index
        int C = C1;
        {
            unsigned long long key0 = ~0ull, key1 = ~0ull;
            if (lane < C1 && S.c_qual[lane]) key0 = ((unsigned long long)S.c_first[lane] << 32) | S.c_tid[lane];
            if (lane + 32 < C1 && S.c_qual[lane + 32]) key1 = ((unsigned long long)S.c_first[lane + 32] << 32) | S.c_tid[lane + 32];
            for (;;) {
                const unsigned long long mine = key0 < key1 ? key0 : key1;
                const unsigned long long best = kb_warp_min64(mine);
                if (best == ~0ull) break;
                const uint32_t who = __ballot_sync(KM_FULL, mine == best);
                const int src = __ffs(who) - 1;
                int ci = -1;
                if (lane == src) { if (key0 == best) { ci = lane; key0 = ~0ull; } else { ci = lane + 32; key1 = ~0ull; } }
                ci = __shfl_sync(KM_FULL, ci, src);
                const uint32_t poff = S.c_poff[ci], plen = S.c_plen[ci];
                unsigned long long anc = 0;
                for (uint32_t c0 = 0; c0 < plen; c0 += 32) {
                    const uint32_t a = c0 + lane < plen ? X.paths[poff + c0 + lane] : KMAT_NONE;
                    int slot = a != KMAT_NONE ? kb_cand_find(S, a) : -1;
                    const bool isnew = a != KMAT_NONE && slot < 0;
                    const uint32_t nm = __ballot_sync(KM_FULL, isnew);
                    const int nnew = __popc(nm);
                    if (C + nnew > KB_CMAX) { overflow = true; break; }
                    // claim ids C.. in path order (nearest ancestor first) = the order the reference appends them
                    if (lane == 0) S.n_used = (uint32_t)(C + nnew);
                    if (isnew) {
                        const int idx = C + __popc(nm & lt_mask);
                        uint32_t h = kb_hash(a);
                        for (;;) {                                            // distinct new keys: plain CAS insert
                            if (atomicCAS(&S.h_nid[h], KMAT_NONE, a) == KMAT_NONE) break;
                            h = (h + 1) & (KB_HSLOTS - 1);
                        }
                        slot = (int)h;
                        S.h_idx[h] = (uint8_t)idx; S.c_slot[idx] = (uint8_t)h; S.order[idx] = (uint8_t)idx;
                        S.c_nid[idx] = a; S.c_anc[idx] = 0ull; S.c_qual[idx] = 0;
                    }
                    C += nnew;
                    __syncwarp();
                    anc |= km_warp_or64(a != KMAT_NONE ? (1ull << S.h_idx[slot]) : 0ull);
                }
                if (overflow) break;
                if (lane == 0) S.c_anc[ci] = anc;
                __syncwarp();
            }
        }
        if (overflow) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; st_err++; continue; }
        // ---- hits per candidate = number of positions whose (expanded) set holds it (:748-759).  Positions of a
        //      chunk with the same set are counted once by a leader lane.
        for (int p0 = 0; p0 < np; p0 += 32) {
            const int p = p0 + lane;
            unsigned long long mask = 0;
            if (p < np) {
                unsigned long long mm = S.posmask[p];
                mask = mm;
                while (mm) {
                    const int idx = __ffsll((long long)mm) - 1;
                    mm &= mm - 1;
                    if (S.c_qual[idx]) mask |= S.c_anc[idx];
                }
            }
            const uint32_t act = __ballot_sync(KM_FULL, mask != 0);
            if (mask != 0) {
                const uint32_t grp = __match_any_sync(act, mask);
                if (lane == __ffs(grp) - 1) {
                    const uint32_t cnt = (uint32_t)__popc(grp);
                    while (mask) {
                        const int idx = __ffsll((long long)mask) - 1;
                        mask &= mask - 1;
                        atomicAdd(&S.c_hits[idx], cnt);
                    }
                }
            }
        }
        __syncwarp();
        // ---- hand over to the scoring kernel: (nid, hits) in taxid_lst order
        unsigned long long co = 0;
        if (lane == 0) co = atomicAdd(P.cand_cursor, (unsigned long long)C);
        co = __shfl_sync(KM_FULL, co, 0);
        res.status = KMAT_ST_PENDING; res.n_cand = (uint32_t)C; res.cand_off = co;
        if (P.cands && co + C <= P.cand_cap) {
            for (int f = lane; f < C; f += 32) {
                const int i = S.order[f];
                P.cands[co + f] = kmat_pair{S.c_nid[i], __uint_as_float(S.c_hits[i])};
            }
        } else { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_OVERFLOW; }       // candidate buffer too small: the host re-runs with the size asked for
        if (lane == 0) P.out[r] = res;
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
struct KsLocal {
    uint32_t nid[KB_CMAX], tid[KB_CMAX], tin[KB_CMAX], tout[KB_CMAX];
    float score[KB_CMAX], rp[KB_CMAX];
    uint16_t depth[KB_CMAX];
    uint8_t flags[KB_CMAX], cls[KB_CMAX];        // flags: 1 human, 2 phix, 4 plasmid
    KmRl rl[KB_CMAX];
    uint32_t l_tid[KB_LIN], l_tin[KB_LIN], l_tout[KB_LIN];
    float l_score[KB_LIN];
    uint16_t l_depth[KB_LIN];
    uint8_t l_nogood[KB_LIN], l_perm[KB_LIN];
    float track_val[64];
    uint8_t track_has[64];
};
struct KsTCmp {           // TCmp, read_label.cpp:475-485: |a-b| < 0.001 (double compare) -> shallower first, else by score
    const uint16_t *depth;
    __device__ bool operator()(const KmRl &a, const KmRl &b) const {
        if ((double)fabsf(__fsub_rn(a.score, b.score)) < 0.001) return (int)depth[a.idx] < (int)depth[b.idx];
        return a.score < b.score;
    }
};
struct KsLinDepthDesc {   // CmpDepth over lineage entries (:159-167), sorting a permutation
    const uint16_t *l_depth;
    __device__ bool operator()(const uint8_t &a, const uint8_t &b) const { return (int)l_depth[a] > (int)l_depth[b]; }
};

__global__ void __launch_bounds__(KS_THREADS) km_score_kernel(KmScoreParams P) {
    const uint32_t r = blockIdx.x * KS_THREADS + threadIdx.x;
    if (r >= P.n_reads) return;
    kmat_read_result res = P.out[r];
    if (res.status != KMAT_ST_PENDING) return;
    const KmCtxDev &X = P.C;
    KsLocal T;
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
        T.rl[f].score = sc; T.rl[f].idx = (uint32_t)f;
    }
    kmstd::sort(T.rl, C, KsTCmp{T.depth});
    const float diff_thresh = __fmul_rn(stdev1, X.opt.sdiff);                  // :895
    // ---- findReadLabelVer2 (:284-419)
    int nlin = 0;
    bool plasmidTopHit = false; int savePlasmid = -1;
    unsigned lowest_depth = 0, highest_depth = 0;
    int lowest = -1, highest = -1; float lowest_score = 0;
    int lidx = -1; bool linDone = false, lin_overflow = false;
    for (int i = C - 1; i >= 0; --i) {                                          // :295-325
        const int ci = (int)T.rl[i].idx; const float sc = T.rl[i].score;
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
                if (nlin >= KB_LIN) lin_overflow = true;
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
            if (nlin >= KB_LIN) { lin_overflow = true; break; }
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
    if (lin_overflow) { res.status = KMAT_ST_ERROR; res.err = KMAT_ERR_UNSUPPORTED; res.n_cand = 0; P.out[r] = res; return; }
    for (int q = 0; q < nlin; q++) T.l_perm[q] = (uint8_t)q;
    kmstd::sort(T.l_perm, nlin, KsLinDepthDesc{T.l_depth});                     // cand_lin_vec sorted by depth desc :350-351
    bool any_nogood = false;
    for (int i = lidx; i >= 0; --i) {                                           // :355-362
        const int ci = (int)T.rl[i].idx; const float sc = T.rl[i].score;
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
    if (P.cands && co + C <= P.cand_cap) for (int i = 0; i < C; i++) P.cands[co + i] = kmat_pair{T.tid[T.rl[i].idx], T.rl[i].score};
    if (X.opt.want_lineage) {
        res.n_lin = (uint32_t)nlin;
        const unsigned long long lo2 = atomicAdd(P.lin_cursor, (unsigned long long)nlin);
        res.lin_off = lo2;
        if (P.lin && lo2 + nlin <= P.lin_cap) for (int q = 0; q < nlin; q++) P.lin[lo2 + q] = kmat_pair{T.l_tid[q], T.l_score[q]};
    }
    P.out[r] = res;
}

// ---------------------------------------------------------------------------------------------
// kmat_ctx
// ---------------------------------------------------------------------------------------------
struct kmat_ctx {
    const kmat_db *db = nullptr;
    int device = 0;
    kmat_opts opt{};
    cudaStream_t stream = nullptr;
    // device tables
    KmNodeA *d_nodeA = nullptr; KmNodeB *d_nodeB = nullptr; uint32_t *d_paths = nullptr, *d_prune = nullptr, *d_sid2nid = nullptr;
    int16_t *d_model_of_cand = nullptr; int32_t *d_mrow = nullptr; float *d_cut = nullptr; uint8_t *d_cls = nullptr;
    KmHostCtx h;
    // batch buffers (grown on demand)
    char *d_bases = nullptr; uint64_t cap_bases = 0;
    uint64_t *d_offs = nullptr; uint32_t cap_reads = 0;
    uint32_t *d_hit = nullptr; uint64_t cap_hit = 0;
    int2 *d_hdr = nullptr; kmat_read_result *d_out = nullptr;
    kmat_pair *d_cands = nullptr, *d_lin = nullptr; uint64_t cap_cands = 0, cap_lin = 0;
    unsigned long long *d_cursors = nullptr;     // [0] cands, [1] lineage
    uint2 *d_big = nullptr; int big_warps = 0;
    unsigned long long *d_long_sets = nullptr; uint32_t long_slots = 0; int long_warps = 0;
    KmStatsDev *d_stats = nullptr;
    int collect_stats = 1;
    kmat_batch_stats last{};
    // pinned staging
    char *h_bases = nullptr; uint64_t hcap_bases = 0;
    uint64_t *h_offs = nullptr; uint32_t hcap_reads = 0;
    kmat_read_result *h_out = nullptr;
    int score_grid = 0;          // grid of the candidate kernel (persistent, warp per read)
    size_t score_smem = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // around the three kernels of the last batch
};

template <typename T>
static int km_upload(T **dst, const std::vector<T> &src) {
    *dst = nullptr;
    if (src.empty()) return KMAT_OK;
    KM_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    KM_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return KMAT_OK;
}

extern "C" int kmat_ctx_create(const kmat_db *db, const kmat_inputs *in, const kmat_opts *opt, kmat_ctx **out) {
    if (!db || !in || !out) { kmat_set_error("kmat_ctx_create: bad argument"); return KMAT_ERR_ARG; }
    kmat_opts o;
    if (opt) o = *opt; else kmat_opts_default(&o);
    if (o.permissive) { kmat_set_error("permissive matching (-s) is not implemented in this build"); return KMAT_ERR_UNSUPPORTED; }
    kmat_ctx *c = new kmat_ctx();
    c->db = db; c->device = db->device; c->opt = o;
    int rc = kmat_build_host_ctx(*in, db->tid_bytes, db->stored_tids, c->h);
    if (rc != KMAT_OK) { delete c; return rc; }
    if (cudaSetDevice(c->device) != cudaSuccess) { delete c; kmat_set_error("cudaSetDevice(%d) failed", db->device); return KMAT_ERR_NO_DEVICE; }
#define UP(dst, src) do { rc = km_upload(&c->dst, c->h.src); if (rc != KMAT_OK) { kmat_ctx_destroy(c); return rc; } } while (0)
    UP(d_nodeA, nodeA); UP(d_nodeB, nodeB); UP(d_paths, paths); UP(d_prune, prune_rank); UP(d_sid2nid, sid2nid);
    UP(d_model_of_cand, model_of_cand); UP(d_mrow, mrow); UP(d_cut, cut); UP(d_cls, cls);
#undef UP
    KM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) KM_CUDA(cudaEventCreate(&c->ev[i]));
    KM_CUDA(cudaMalloc((void **)&c->d_cursors, 16));
    KM_CUDA(cudaMalloc((void **)&c->d_stats, sizeof(KmStatsDev)));
    c->score_smem = sizeof(KmWarpB) * KB_WARPS;
    KM_CUDA(cudaFuncSetAttribute(km_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->score_smem));
    int per_sm = 0;
    KM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, km_cand_kernel, KB_WARPS * 32, c->score_smem));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    c->score_grid = std::max(1, per_sm) * sms;
    c->big_warps = c->score_grid * KB_WARPS;
    KM_CUDA(cudaMalloc((void **)&c->d_big, (size_t)c->big_warps * KB_BIGCAP * sizeof(uint2)));
    *out = c;
    return KMAT_OK;
}
extern "C" int kmat_ctx_set_opts(kmat_ctx *c, const kmat_opts *o) {
    if (!c || !o) return KMAT_ERR_ARG;
    if (o->permissive) { kmat_set_error("permissive matching (-s) is not implemented in this build"); return KMAT_ERR_UNSUPPORTED; }
    c->opt = *o;
    return KMAT_OK;
}
extern "C" void kmat_ctx_destroy(kmat_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_nodeA); cudaFree(c->d_nodeB); cudaFree(c->d_paths); cudaFree(c->d_prune); cudaFree(c->d_sid2nid);
    cudaFree(c->d_model_of_cand); cudaFree(c->d_mrow); cudaFree(c->d_cut); cudaFree(c->d_cls);
    cudaFree(c->d_bases); cudaFree(c->d_offs); cudaFree(c->d_hit); cudaFree(c->d_hdr); cudaFree(c->d_out);
    cudaFree(c->d_cands); cudaFree(c->d_lin); cudaFree(c->d_cursors); cudaFree(c->d_big); cudaFree(c->d_long_sets); cudaFree(c->d_stats);
    cudaFreeHost(c->h_bases); cudaFreeHost(c->h_offs); cudaFreeHost(c->h_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int i = 0; i < 4; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
}

static KmCtxDev km_ctx_dev(const kmat_ctx *c) {
    KmCtxDev X;
    X.db = km_db_dev(c->db);
    X.nodeA = c->d_nodeA; X.nodeB = c->d_nodeB; X.paths = c->d_paths; X.prune_rank = c->d_prune; X.sid2nid = c->d_sid2nid;
    X.n_sid = (uint32_t)c->h.sid2nid.size(); X.n_nodes = (uint32_t)c->h.nodeA.size(); X.nid_human = c->h.nid_human; X.nid_one = c->h.nid_one;
    X.nbins = c->h.nbins; X.n_models = c->h.n_models; X.n_classes = c->h.n_classes;
    X.model_of_cand = c->d_model_of_cand; X.mrow = c->d_mrow; X.cut = c->d_cut; X.cls = c->d_cls;
    memset(X.class_ranknum, 0, sizeof X.class_ranknum);
    for (int i = 0; i < c->h.n_classes && i < 64; i++) X.class_ranknum[i] = (int8_t)c->h.class_ranknum[i];
    X.opt = c->opt;
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

// per-read device buffers share one capacity
static int km_reserve_reads(kmat_ctx *c, uint32_t n_reads) {
    if (n_reads <= c->cap_reads && c->d_hdr) return KMAT_OK;
    cudaFree(c->d_hdr); cudaFree(c->d_offs); cudaFree(c->d_out);
    c->d_hdr = nullptr; c->d_offs = nullptr; c->d_out = nullptr; c->cap_reads = 0;
    const size_t cap = (size_t)n_reads + n_reads / 4 + 64;
    KM_CUDA(cudaMalloc((void **)&c->d_hdr, cap * sizeof(int2)));
    KM_CUDA(cudaMalloc((void **)&c->d_offs, (cap + 1) * 8));
    KM_CUDA(cudaMalloc((void **)&c->d_out, cap * sizeof(kmat_read_result)));
    c->cap_reads = (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFull);
    return KMAT_OK;
}

// Launch K1+K2 then K3+K4 on d_bases / d_offs (already on the device).  max_len bounds the longest read.
static int km_run_device(kmat_ctx *c, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint64_t total_bases,
                         uint32_t max_len, kmat_read_result *d_out, cudaStream_t st) {
    int rc;
    if ((rc = km_grow(&c->d_hit, &c->cap_hit, total_bases + 1)) != KMAT_OK) return rc;
    if ((rc = km_reserve_reads(c, n_reads)) != KMAT_OK) return rc;
    const int pgrid = km_probe_grid(n_reads);
    const uint32_t max_np = max_len;
    if (max_np > 256) {
        uint32_t slots = 1024; while (slots < 2 * max_np) slots <<= 1;
        const int warps = 148 * 6 * KM_PROBE_WARPS_HOST;
        if (slots > c->long_slots || warps > c->long_warps) {
            cudaFree(c->d_long_sets); c->d_long_sets = nullptr;
            KM_CUDA(cudaMalloc((void **)&c->d_long_sets, (size_t)warps * slots * 8));
            c->long_slots = slots; c->long_warps = warps;
        }
    }
    if (!c->d_cands) { if ((rc = km_grow(&c->d_cands, &c->cap_cands, (uint64_t)n_reads * 24 + 4096)) != KMAT_OK) return rc; }
    if (c->opt.want_lineage && !c->d_lin) { if ((rc = km_grow(&c->d_lin, &c->cap_lin, (uint64_t)n_reads * 24 + 4096)) != KMAT_OK) return rc; }
    KM_CUDA(cudaMemsetAsync(c->d_cursors, 0, 16, st));
    if (c->collect_stats) KM_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(KmStatsDev), st));
    KM_CUDA(cudaEventRecord(c->ev[0], st));
    rc = km_launch_encode_probe(c->db, d_bases, d_offs, n_reads, c->d_hit, c->d_hdr, nullptr, nullptr, c->d_long_sets, c->long_slots,
                                pgrid, c->collect_stats ? c->d_stats : nullptr, 1, st);
    if (rc != KMAT_OK) return rc;
    KM_CUDA(cudaEventRecord(c->ev[1], st));
    KmScoreParams P;
    P.C = km_ctx_dev(c);
    P.offs = d_offs; P.n_reads = n_reads; P.hit = c->d_hit; P.hdr = c->d_hdr; P.out = d_out;
    P.cands = c->d_cands; P.cand_cursor = c->d_cursors; P.cand_cap = c->cap_cands;
    P.lin = c->d_lin; P.lin_cursor = c->d_cursors + 1; P.lin_cap = c->cap_lin;
    P.big_scratch = c->d_big; P.stats = c->collect_stats ? c->d_stats : nullptr;
    const int grid = std::max(1, std::min<int>(c->score_grid, (int)((n_reads + KB_WARPS - 1) / KB_WARPS)));
    km_cand_kernel<<<grid, KB_WARPS * 32, c->score_smem, st>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    KM_CUDA(cudaEventRecord(c->ev[3], st));
    km_score_kernel<<<(n_reads + KS_THREADS - 1) / KS_THREADS, KS_THREADS, 0, st>>>(P);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    KM_CUDA(cudaEventRecord(c->ev[2], st));
    return KMAT_OK;
}

extern "C" int kmat_label_batch_device(kmat_ctx *c, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint64_t total_bases,
                                       uint32_t max_read_len, kmat_read_result *d_out, void *stream) {
    if (!c || !d_offs || (n_reads && !d_bases)) { kmat_set_error("kmat_label_batch_device: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    if (!n_reads) return KMAT_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = km_reserve_reads(c, n_reads);
    if (rc != KMAT_OK) return rc;
    if (!d_out) d_out = c->d_out;
    return km_run_device(c, d_bases, d_offs, n_reads, total_bases, max_read_len, d_out, st);
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
extern "C" int kmat_ctx_set_stats(kmat_ctx *c, int enable) {
    if (!c) return KMAT_ERR_ARG;
    c->collect_stats = enable ? 1 : 0;
    return KMAT_OK;
}
extern "C" int kmat_ctx_sync(kmat_ctx *c) {
    if (!c) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(c->device));
    KM_CUDA(cudaStreamSynchronize(c->stream));
    return KMAT_OK;
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

extern "C" int kmat_label_batch(kmat_ctx *c, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                                kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands, kmat_pair *lineage, uint64_t lineage_cap,
                                uint64_t *n_lineage) {
    if (!c || !offs || !out || (n_reads && !bases)) { kmat_set_error("kmat_label_batch: bad argument"); return KMAT_ERR_ARG; }
    if (n_cands) *n_cands = 0;
    if (n_lineage) *n_lineage = 0;
    if (!n_reads) return KMAT_OK;
    KM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint64_t total = offs[n_reads] - offs[0];
    uint32_t max_len = 0;
    for (uint32_t r = 0; r < n_reads; r++) max_len = std::max<uint32_t>(max_len, (uint32_t)(offs[r + 1] - offs[r]));
    int rc;
    // pinned staging (the caller's buffers are ordinary host memory)
    if (total + 1 > c->hcap_bases) { cudaFreeHost(c->h_bases); c->h_bases = nullptr; c->hcap_bases = total + total / 4 + 4096; KM_CUDA(cudaMallocHost((void **)&c->h_bases, c->hcap_bases)); }
    if (n_reads + 1 > c->hcap_reads) {
        cudaFreeHost(c->h_offs); cudaFreeHost(c->h_out); c->h_offs = nullptr; c->h_out = nullptr;
        c->hcap_reads = n_reads + n_reads / 4 + 64;
        KM_CUDA(cudaMallocHost((void **)&c->h_offs, (size_t)c->hcap_reads * 8));
        KM_CUDA(cudaMallocHost((void **)&c->h_out, (size_t)c->hcap_reads * sizeof(kmat_read_result)));
    }
    memcpy(c->h_bases, bases + offs[0], total);
    for (uint32_t r = 0; r <= n_reads; r++) c->h_offs[r] = offs[r] - offs[0];
    if ((rc = km_grow(&c->d_bases, &c->cap_bases, total + 1)) != KMAT_OK) return rc;
    if ((rc = km_reserve_reads(c, n_reads)) != KMAT_OK) return rc;
    KM_CUDA(cudaMemcpyAsync(c->d_bases, c->h_bases, total, cudaMemcpyHostToDevice, st));
    KM_CUDA(cudaMemcpyAsync(c->d_offs, c->h_offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    for (int attempt = 0; attempt < 3; attempt++) {
        rc = km_run_device(c, c->d_bases, c->d_offs, n_reads, total, max_len, c->d_out, st);
        if (rc != KMAT_OK) return rc;
        unsigned long long cur[2];
        KM_CUDA(cudaMemcpyAsync(cur, c->d_cursors, 16, cudaMemcpyDeviceToHost, st));
        KM_CUDA(cudaStreamSynchronize(st));
        bool again = false;
        if (cur[0] > c->cap_cands) { if ((rc = km_grow(&c->d_cands, &c->cap_cands, cur[0])) != KMAT_OK) return rc; again = true; }
        if (c->opt.want_lineage && cur[1] > c->cap_lin) { if ((rc = km_grow(&c->d_lin, &c->cap_lin, cur[1])) != KMAT_OK) return rc; again = true; }
        if (again) continue;                     // candidate buffer was too small: re-run with the exact size
        KM_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, (size_t)n_reads * sizeof(kmat_read_result), cudaMemcpyDeviceToHost, st));
        if (n_cands) *n_cands = cur[0];
        if (n_lineage) *n_lineage = c->opt.want_lineage ? cur[1] : 0;
        int ret = KMAT_OK;
        if (cands) { if (cur[0] <= cands_cap) { if (cur[0]) KM_CUDA(cudaMemcpyAsync(cands, c->d_cands, cur[0] * sizeof(kmat_pair), cudaMemcpyDeviceToHost, st)); } else ret = KMAT_ERR_OVERFLOW; }
        if (lineage && c->opt.want_lineage) { if (cur[1] <= lineage_cap) { if (cur[1]) KM_CUDA(cudaMemcpyAsync(lineage, c->d_lin, cur[1] * sizeof(kmat_pair), cudaMemcpyDeviceToHost, st)); } else ret = KMAT_ERR_OVERFLOW; }
        KM_CUDA(cudaStreamSynchronize(st));
        memcpy(out, c->h_out, (size_t)n_reads * sizeof(kmat_read_result));
        if (ret == KMAT_ERR_OVERFLOW) kmat_set_error("candidate buffer too small: need %llu pairs", cur[0]);
        return ret;
    }
    kmat_set_error("candidate buffer kept overflowing");
    return KMAT_ERR_CUDA;
}
