// kmat_shard.cuh -- DB-sharded mode (SURVEY.md 8(e) mode B); included at the end of kmat_label.cu.
//
// When the table does not fit one GPU it is partitioned by kmat_shard_of() (a hash of the canonical k-mer) and every
// rank holds one shard.  A batch of reads stays on its "home" rank; only k-mers travel:
//
//   home   kmat_shard_encode   K1 (encode + dedup, no probe) -> first-occurrence k-mers, grouped by owner shard
//   ------ exchange 1: all-to-all of the (mixed) k-mers, 8 bytes each -------------------------------------------
//   owner  kmat_shard_serve    probe the local shard; reply = one hit word per query (+ for list hits the resolved
//                              list record, packed per source rank)
//   ------ exchange 2: all-to-all of the hit words (4 bytes each) and of the list records ------------------------
//   home   kmat_shard_finish   hit words back to their read positions, then K3 / K4 exactly as in replicated mode,
//                              reading list records from the received payload instead of the local resolved pool
//
// The exchange itself is the caller's: NCCL all-to-all between one-process-per-GPU ranks (lmat_b200/sharded.py over
// torch.distributed) or peer copies inside one process.  Everything after the lookup stays on the home rank, so the
// results are bit-identical to the replicated table's.
#include <cub/device/device_scan.cuh>

struct KmShardSeg { unsigned long long start[KM_MAX_SHARDS + 1]; };

struct KmShardState {
    // home side
    uint64_t *d_xq = nullptr; uint64_t cap_xq = 0;          // mixed k-mer per base offset (first occurrences only)
    uint64_t *d_q = nullptr; uint64_t cap_q = 0;            // queries grouped by owner
    uint32_t *d_origin = nullptr; uint64_t cap_origin = 0;  // base offset of each query, relative to the pass
    unsigned long long *d_counts = nullptr;                 // [0..16) counts, [16..32) scatter cursors
    KmPass pass{}; uint32_t *hit = nullptr; int variant = 0; int n_shards = 0;
    uint64_t counts[KM_MAX_SHARDS] = {}; uint64_t n_q = 0;
    bool open = false;
    // owner side
    uint32_t *d_reply = nullptr; uint64_t cap_reply = 0;
    uint32_t *d_len = nullptr, *d_pos = nullptr; uint64_t cap_len = 0, cap_pos = 0;
    uint32_t *d_payload = nullptr; uint64_t cap_payload = 0;
    void *d_scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    unsigned long long *d_bounds = nullptr;                 // payload word offset at every source boundary
};

static void km_shard_free(KmShardState *s) {
    if (!s) return;
    cudaFree(s->d_xq); cudaFree(s->d_q); cudaFree(s->d_origin); cudaFree(s->d_counts); cudaFree(s->d_reply); cudaFree(s->d_len);
    cudaFree(s->d_pos); cudaFree(s->d_payload); cudaFree(s->d_scan_tmp); cudaFree(s->d_bounds);
    delete s;
}
static int km_shard_state(kmat_ctx *c, KmShardState **out) {
    if (!c->shard) {
        KmShardState *s = new KmShardState();
        c->shard = s;
        KM_CUDA(cudaMalloc((void **)&s->d_counts, 2 * KM_MAX_SHARDS * sizeof(unsigned long long)));
        KM_CUDA(cudaMalloc((void **)&s->d_bounds, (KM_MAX_SHARDS + 1) * sizeof(unsigned long long)));
    }
    *out = c->shard;
    return KMAT_OK;
}

__device__ __forceinline__ int km_seg_of(const KmShardSeg &g, int n, unsigned long long i) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < n && i >= g.start[s + 1]) s++;
    return s;
}

// ---- home: count and scatter the first-occurrence k-mers by owner -------------------------------------------------
__global__ void __launch_bounds__(256) km_shard_count_kernel(KmDbDev db, const uint32_t *__restrict__ hit, const uint64_t *__restrict__ xq, uint64_t n_pos,
                                                             uint32_t n_shards, unsigned long long *counts) {
    // per-lane private counter of owner `lane` (n_shards <= 16 < 32), fed by one ballot per owner: no contended atomics
    __shared__ unsigned int hist[KM_MAX_SHARDS];
    if (threadIdx.x < KM_MAX_SHARDS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned int mine = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31); i0 < n_pos; i0 += stride) {       // warp-uniform bounds
        const uint64_t i = i0 + lane;
        const bool m = i < n_pos && hit[i] == KM_HIT_MISS;
        const uint32_t owner = m ? km_owner_of_key(db, xq[i], n_shards) : 0xFFFFFFFFu;
        for (uint32_t o = 0; o < n_shards; o++) {
            const uint32_t b = __ballot_sync(KM_FULL, owner == o);
            if ((uint32_t)lane == o) mine += __popc(b);
        }
    }
    if ((uint32_t)lane < n_shards && mine) atomicAdd(&hist[lane], mine);
    __syncthreads();
    if (threadIdx.x < n_shards && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}
#define KM_SCATTER_TILE 4096
__global__ void __launch_bounds__(256) km_shard_scatter_kernel(KmDbDev db, const uint32_t *__restrict__ hit, const uint64_t *__restrict__ xq, uint64_t n_pos,
                                                               uint32_t n_shards, KmShardSeg base, unsigned long long *cursors,
                                                               uint64_t *q, uint32_t *origin) {
    // A CTA takes tiles of 4096 positions: it counts the tile's queries per owner in shared memory, reserves their places
    // with ONE global atomic per owner and tile (a warp-level atomic per 32 positions on 16 hot addresses cost 4 ms per
    // 2^20 reads), then walks the tile again and hands out the places from shared-memory counters.
    __shared__ unsigned int s_cnt[KM_MAX_SHARDS];
    __shared__ unsigned long long s_base[KM_MAX_SHARDS];
    const int lane = threadIdx.x & 31;
    const uint64_t n_tiles = (n_pos + KM_SCATTER_TILE - 1) / KM_SCATTER_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * KM_SCATTER_TILE;
        if (threadIdx.x < KM_MAX_SHARDS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        unsigned int mine = 0;
        for (int j = 0; j < KM_SCATTER_TILE / 256; j++) {
            const uint64_t i = t0 + (uint64_t)j * 256 + threadIdx.x;
            const bool m = i < n_pos && hit[i] == KM_HIT_MISS;
            const uint32_t owner = m ? km_owner_of_key(db, xq[i], n_shards) : 0xFFFFFFFFu;
            for (uint32_t o = 0; o < n_shards; o++) {
                const uint32_t bal = __ballot_sync(KM_FULL, owner == o);
                if ((uint32_t)lane == o) mine += __popc(bal);
            }
        }
        if ((uint32_t)lane < n_shards && mine) atomicAdd(&s_cnt[lane], mine);
        __syncthreads();
        if (threadIdx.x < n_shards) {
            const unsigned int cn = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = base.start[threadIdx.x] + (cn ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)cn) : 0ull);
            s_cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        for (int j = 0; j < KM_SCATTER_TILE / 256; j++) {
            const uint64_t i = t0 + (uint64_t)j * 256 + threadIdx.x;
            const bool m = i < n_pos && hit[i] == KM_HIT_MISS;
            const uint64_t x = m ? xq[i] : 0;
            const uint32_t owner = m ? km_owner_of_key(db, x, n_shards) : 0xFFFFFFFFu;
            const uint32_t act = __ballot_sync(KM_FULL, m);
            if (m) {
                const uint32_t grp = __match_any_sync(act, owner);
                const int leader = __ffs(grp) - 1;
                unsigned int off = 0;
                if (lane == leader) off = atomicAdd(&s_cnt[owner], (unsigned int)__popc(grp));
                off = __shfl_sync(grp, off, leader);
                const unsigned long long dst = s_base[owner] + off + __popc(grp & ((1u << lane) - 1));
                q[dst] = x; origin[dst] = (uint32_t)i;
            }
        }
        __syncthreads();
    }
}

// ---- owner: probe, measure, pack ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) km_shard_probe_kernel(KmDbDev db, const uint32_t *__restrict__ pool2, int mul, int permissive,
                                                             const uint64_t *__restrict__ q, uint64_t n, uint32_t *reply, uint32_t *len) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { len[i] = 0; return; }                    // sentinel for the exclusive scan
    uint32_t extra;
    const uint32_t hw = km_probe_x(db, q[i], extra);
    uint32_t l = 0;
    if (hw != KM_HIT_MISS && (hw & KM_HIT_LIST)) {
        const uint32_t *rec = pool2 + (size_t)(hw & 0x7FFFFFFFu) * mul;
        const uint32_t h = rec[0];
        if (h == KR_ERR_BAD) l = 1;                         // the home rank reports the bad stored id
        else l = permissive ? 2 + (h & 0xFFFFu) + rec[1] : 1 + (h & 0xFFFFu);
    }
    reply[i] = hw; len[i] = l;
}
__global__ void __launch_bounds__(256) km_shard_pack_kernel(const uint32_t *__restrict__ pool2, int mul, uint32_t *reply, const uint32_t *__restrict__ len,
                                                            const uint32_t *__restrict__ pos, uint64_t n, KmShardSeg src, int n_shards,
                                                            uint32_t *payload, unsigned long long *bounds) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i <= (uint64_t)n_shards) bounds[i] = pos[src.start[i]];             // payload word offset at every source boundary
    if (i >= n) return;
    const uint32_t l = len[i];
    if (!l) return;
    const uint32_t hw = reply[i];
    const uint32_t *rec = pool2 + (size_t)(hw & 0x7FFFFFFFu) * mul;
    const uint32_t at = pos[i];
    for (uint32_t w = 0; w < l; w++) payload[at + w] = rec[w];
    const int s = km_seg_of(src, n_shards, i);
    reply[i] = KM_HIT_LIST | (at - pos[src.start[s]]);                      // offset inside the source's own payload segment
}

// ---- home: replies back to their read positions --------------------------------------------------------------------
__global__ void __launch_bounds__(256) km_shard_fix_kernel(const uint32_t *__restrict__ reply, const uint32_t *__restrict__ origin, uint64_t n_q,
                                                           KmShardSeg qseg, KmShardSeg pbase, int n_shards, uint32_t *hit) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_q) return;
    uint32_t hw = reply[i];
    if (hw != KM_HIT_MISS && (hw & KM_HIT_LIST)) {
        const int s = km_seg_of(qseg, n_shards, i);
        hw = KM_HIT_LIST | (uint32_t)(pbase.start[s] + (hw & 0x7FFFFFFFu));
    }
    hit[origin[i]] = hw;
}

extern "C" int kmat_shard_encode(kmat_ctx *c, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint64_t total_bases,
                                 uint32_t max_read_len, int n_shards, const uint64_t **d_queries, uint64_t *counts, void *stream) {
    if (!c || !d_offs || !counts || !d_queries || n_shards < 1 || n_shards > KM_MAX_SHARDS || (n_reads && !d_bases)) { kmat_set_error("kmat_shard_encode: bad argument"); return KMAT_ERR_ARG; }
    if (total_bases >= (1ull << 32)) { kmat_set_error("kmat_shard_encode: a sharded pass is limited to 2^32 bases (got %llu); split the batch", (unsigned long long)total_bases); return KMAT_ERR_UNSUPPORTED; }
    if (c->db->shard_count != n_shards) { kmat_set_error("kmat_shard_encode: the ctx's table is shard %d of %d, not of %d", c->db->shard_index, c->db->shard_count, n_shards); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    KmShardState *s;
    int rc;
    if ((rc = km_shard_state(c, &s)) != KMAT_OK) return rc;
    for (int i = 0; i < n_shards; i++) counts[i] = 0;
    *d_queries = nullptr;
    s->open = false; s->n_q = 0; s->n_shards = n_shards;
    if (!n_reads) { s->pass = KmPass{d_bases, d_offs, 0, 0, 0, max_read_len, nullptr, true}; s->open = true; return KMAT_OK; }
    if ((rc = km_reserve_cands(c, n_reads, max_read_len)) != KMAT_OK) return rc;
    s->pass = KmPass{d_bases, d_offs, n_reads, 0, total_bases, max_read_len, nullptr, true};
    if ((rc = km_prepare_pass(c, s->pass, st, &s->hit, &s->variant)) != KMAT_OK) return rc;
    if ((rc = km_grow(&s->d_xq, &s->cap_xq, total_bases + 1)) != KMAT_OK) return rc;
    // positions no k-mer starts at (the last k-1 bases of a read) are never written by the encode kernel
    KM_CUDA(cudaMemsetAsync(c->d_hit, 0xFF, (size_t)total_bases * 4, st));
    rc = km_launch_encode_probe(c->db, d_bases, d_offs, n_reads, max_read_len, s->hit, c->d_hdr, nullptr, nullptr, c->d_long_sets, c->long_slots,
                                km_probe_grid(n_reads), nullptr, 0, st, 0, s->d_xq);
    if (rc != KMAT_OK) return rc;
    KM_CUDA(cudaMemsetAsync(s->d_counts, 0, 2 * KM_MAX_SHARDS * sizeof(unsigned long long), st));
    const int grid = (int)std::min<uint64_t>((total_bases + 255) / 256, 148ull * 8);
    km_shard_count_kernel<<<grid, 256, 0, st>>>(km_db_dev(c->db), c->d_hit, s->d_xq, total_bases, (uint32_t)n_shards, s->d_counts);
    g_km_launches++;
    unsigned long long h_counts[KM_MAX_SHARDS];
    KM_CUDA(cudaMemcpyAsync(h_counts, s->d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, st));
    KM_CUDA(cudaStreamSynchronize(st));
    KmShardSeg base;
    unsigned long long tot = 0;
    for (int i = 0; i <= KM_MAX_SHARDS; i++) { base.start[i] = tot; if (i < n_shards) { counts[i] = s->counts[i] = h_counts[i]; tot += h_counts[i]; } }
    s->n_q = tot;
    if ((rc = km_grow(&s->d_q, &s->cap_q, tot + 1)) != KMAT_OK) return rc;
    if ((rc = km_grow(&s->d_origin, &s->cap_origin, tot + 1)) != KMAT_OK) return rc;
    km_shard_scatter_kernel<<<(int)std::min<uint64_t>((total_bases + KM_SCATTER_TILE - 1) / KM_SCATTER_TILE, 148ull * 8), 256, 0, st>>>(km_db_dev(c->db), c->d_hit, s->d_xq, total_bases, (uint32_t)n_shards, base, s->d_counts + KM_MAX_SHARDS, s->d_q, s->d_origin);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    *d_queries = s->d_q;
    s->open = true;
    return KMAT_OK;
}

extern "C" int kmat_shard_serve(kmat_ctx *c, const uint64_t *d_queries, const uint64_t *counts, int n_shards, const uint32_t **d_reply,
                                const uint32_t **d_payload, uint64_t *payload_counts, void *stream) {
    if (!c || !counts || !d_reply || !d_payload || !payload_counts || n_shards < 1 || n_shards > KM_MAX_SHARDS) { kmat_set_error("kmat_shard_serve: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    KmShardState *s;
    int rc;
    if ((rc = km_shard_state(c, &s)) != KMAT_OK) return rc;
    KmShardSeg src;
    unsigned long long n = 0;
    for (int i = 0; i <= KM_MAX_SHARDS; i++) { src.start[i] = n; if (i < n_shards) n += counts[i]; }
    for (int i = 0; i < n_shards; i++) payload_counts[i] = 0;
    *d_reply = nullptr; *d_payload = nullptr;
    if (!n) return KMAT_OK;
    if (!d_queries) { kmat_set_error("kmat_shard_serve: bad argument"); return KMAT_ERR_ARG; }
    if (n >= (1ull << 32) - 1) { kmat_set_error("kmat_shard_serve: too many queries in one round"); return KMAT_ERR_UNSUPPORTED; }
    if ((rc = km_grow(&s->d_reply, &s->cap_reply, n + 1)) != KMAT_OK) return rc;
    if ((rc = km_grow(&s->d_len, &s->cap_len, n + 2)) != KMAT_OK) return rc;
    if ((rc = km_grow(&s->d_pos, &s->cap_pos, n + 2)) != KMAT_OK) return rc;
    const KmCtxDev X = km_ctx_dev(c);
    const int grid = (int)((n + 1 + 255) / 256);
    km_shard_probe_kernel<<<grid, 256, 0, st>>>(X.db, X.pool2, X.pool2_mul, X.opt.permissive != 0, d_queries, n, s->d_reply, s->d_len);
    g_km_launches++;
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, s->d_len, s->d_pos, (uint32_t)(n + 1), st);
    if (need > s->scan_tmp_bytes) {
        KM_CUDA(cudaStreamSynchronize(st));
        cudaFree(s->d_scan_tmp); s->d_scan_tmp = nullptr;
        KM_CUDA(cudaMalloc(&s->d_scan_tmp, need + 256));
        s->scan_tmp_bytes = need + 256;
    }
    size_t tmp_bytes = s->scan_tmp_bytes;
    cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, tmp_bytes, s->d_len, s->d_pos, (uint32_t)(n + 1), st);
    g_km_launches++;
    // total payload words = pos[n]; the payload buffer is sized before the pack kernel runs
    uint32_t total_words = 0;
    KM_CUDA(cudaMemcpyAsync(&total_words, s->d_pos + n, 4, cudaMemcpyDeviceToHost, st));
    KM_CUDA(cudaStreamSynchronize(st));
    if (total_words >= (1u << 31)) { kmat_set_error("kmat_shard_serve: list payload of %u words exceeds the 31-bit offset range; use smaller rounds", total_words); return KMAT_ERR_UNSUPPORTED; }
    if ((rc = km_grow(&s->d_payload, &s->cap_payload, (uint64_t)total_words + 8)) != KMAT_OK) return rc;
    km_shard_pack_kernel<<<grid, 256, 0, st>>>(X.pool2, X.pool2_mul, s->d_reply, s->d_len, s->d_pos, n, src, n_shards, s->d_payload, s->d_bounds);
    g_km_launches++;
    unsigned long long h_bounds[KM_MAX_SHARDS + 1];
    KM_CUDA(cudaMemcpyAsync(h_bounds, s->d_bounds, sizeof h_bounds, cudaMemcpyDeviceToHost, st));
    KM_CUDA(cudaStreamSynchronize(st));
    KM_CUDA(cudaGetLastError());
    for (int i = 0; i < n_shards; i++) payload_counts[i] = h_bounds[i + 1] - h_bounds[i];
    *d_reply = s->d_reply; *d_payload = s->d_payload;
    return KMAT_OK;
}

extern "C" int kmat_shard_finish(kmat_ctx *c, const uint32_t *d_reply, const uint32_t *d_payload, const uint64_t *payload_counts, int n_shards,
                                 kmat_read_result *d_out, void *stream) {
    if (!c || !payload_counts || n_shards < 1 || n_shards > KM_MAX_SHARDS) { kmat_set_error("kmat_shard_finish: bad argument"); return KMAT_ERR_ARG; }
    KmShardState *s = c->shard;
    if (!s || !s->open || s->n_shards != n_shards) { kmat_set_error("kmat_shard_finish: no open pass (call kmat_shard_encode first)"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    s->open = false;
    if (!s->pass.n_reads) return KMAT_OK;
    if (s->n_q && !d_reply) { kmat_set_error("kmat_shard_finish: bad argument"); return KMAT_ERR_ARG; }
    int rc;
    KmPass L = s->pass;
    if (!d_out) {
        uint64_t cap = c->cap_out_dev;
        if ((rc = km_grow(&c->d_out_dev, &cap, L.n_reads)) != KMAT_OK) return rc;
        c->cap_out_dev = (uint32_t)cap;
        d_out = c->d_out_dev;
    }
    L.d_out = d_out;
    KmShardSeg qseg, pbase;
    unsigned long long a = 0, b = 0;
    for (int i = 0; i <= KM_MAX_SHARDS; i++) { qseg.start[i] = a; pbase.start[i] = b; if (i < n_shards) { a += s->counts[i]; b += payload_counts[i]; } }
    if (b >= (1ull << 31)) { kmat_set_error("kmat_shard_finish: list payload exceeds the 31-bit offset range; use smaller rounds"); return KMAT_ERR_UNSUPPORTED; }
    if (s->n_q) {
        km_shard_fix_kernel<<<(int)((s->n_q + 255) / 256), 256, 0, st>>>(d_reply, s->d_origin, s->n_q, qseg, pbase, n_shards, c->d_hit);
        g_km_launches++;
        KM_CUDA(cudaGetLastError());
    }
    KM_CUDA(cudaEventRecord(c->ev[0], st)); KM_CUDA(cudaEventRecord(c->ev[1], st));
    // list records come from the received payload (mul 1); a batch without list hits never dereferences it
    rc = km_launch_cand_score(c, L, 0, L.n_reads, s->hit, s->variant, 0, st, d_payload ? d_payload : c->d_pool2, d_payload ? 1 : c->pool2_mul, c->ev[3]);
    if (rc != KMAT_OK) return rc;
    KM_CUDA(cudaEventRecord(c->ev[2], st));
    return KMAT_OK;
}

/* Results of the last finished sharded pass when kmat_shard_finish was called with d_out == NULL: device pointer of the
 * n_reads results, plus the device candidate pairs (rank_label after sort) they index through cand_off. */
extern "C" int kmat_ctx_device_results(kmat_ctx *c, const kmat_read_result **d_out, const kmat_pair **d_cands, uint64_t *n_cands) {
    if (!c) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(c->device));
    if (d_out) *d_out = c->d_out_dev;
    if (d_cands) *d_cands = c->d_cands;
    if (n_cands) {
        unsigned long long cur[2] = {0, 0};
        KM_CUDA(cudaMemcpy(cur, c->d_cursors, 16, cudaMemcpyDeviceToHost));
        *n_cands = cur[0];
    }
    return KMAT_OK;
}

// =====================================================================================================================
// Direct variant of the DB-sharded mode: no exchange rounds at all.  Every rank maps the bucket arrays, stashes and
// resolved list pools of ALL shards into its own address space (CUDA IPC between one-process-per-GPU ranks; plain peer
// access or the same device inside one process) and the unchanged K1+K2 / K3 kernels send each gather to the memory of
// the shard that owns the k-mer: over NVLink 5 / NVSwitch every peer is one hop away at full bandwidth, a 32-byte bucket
// read is one NVLink read request.  Nothing else moves; the labels are those of the replicated table by construction.
// =====================================================================================================================
struct KmPeerBlob {
    uint32_t magic; int32_t pid, device, shard_index, shard_count;
    int32_t bucket_bits, rem_bits, kmer_bits, tid_bytes, pool2_mul, rkmer, permissive, max_count;
    int32_t line_m, line_bits, pool_shared;
    uint32_t n_stash; uint64_t pool_words, line_first;
    uint64_t p_lines, p_slots, p_stash_x, p_stash_hit, p_pool2;                 // raw device pointers (valid inside the exporting process)
    cudaIpcMemHandle_t h_lines, h_slots, h_stash_x, h_stash_hit, h_pool2;       // the same allocations for other processes
};
static_assert(sizeof(KmPeerBlob) <= sizeof(kmat_peer_info), "kmat_peer_info too small");
#define KM_PEER_MAGIC 0x4B4D5045u

extern "C" int kmat_ctx_peer_export(kmat_ctx *c, kmat_peer_info *out) {
    if (!c || !out) { kmat_set_error("kmat_ctx_peer_export: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    KM_CUDA(cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    KmPeerBlob b;
    memset(&b, 0, sizeof b);
    const kmat_db *db = c->db;
    b.magic = KM_PEER_MAGIC; b.pid = (int32_t)getpid(); b.device = c->device; b.shard_index = db->shard_index; b.shard_count = db->shard_count;
    b.bucket_bits = db->geom.bucket_bits; b.rem_bits = db->geom.rem_bits; b.kmer_bits = db->geom.kmer_bits; b.tid_bytes = db->tid_bytes;
    b.pool2_mul = c->pool2_mul; b.rkmer = c->opt.rkmer_mode != 0; b.permissive = c->opt.permissive != 0; b.max_count = c->opt.max_count;
    b.n_stash = db->n_stash; b.pool_words = db->pool_words;
    b.pool_shared = db->pool_shared ? 1 : 0;
    b.line_m = db->geom.line_m; b.line_bits = db->geom.line_bits; b.line_first = db->line_first; b.p_lines = (uint64_t)db->d_lines;
    if (db->d_lines) KM_CUDA(cudaIpcGetMemHandle(&b.h_lines, db->d_lines));
    b.p_slots = (uint64_t)db->d_slots; b.p_stash_x = (uint64_t)db->d_stash_x; b.p_stash_hit = (uint64_t)db->d_stash_hit; b.p_pool2 = (uint64_t)c->d_pool2;
    if (db->d_slots) KM_CUDA(cudaIpcGetMemHandle(&b.h_slots, db->d_slots));
    if (db->n_stash) { KM_CUDA(cudaIpcGetMemHandle(&b.h_stash_x, db->d_stash_x)); KM_CUDA(cudaIpcGetMemHandle(&b.h_stash_hit, db->d_stash_hit)); }
    if (c->d_pool2) KM_CUDA(cudaIpcGetMemHandle(&b.h_pool2, c->d_pool2));
    memcpy(out, &b, sizeof b);
    return KMAT_OK;
}

// dst |= src over a resolved list pool (shared-pool shards: every word is non-zero in at most one shard's pool)
__global__ void __launch_bounds__(256) km_pool_merge_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint64_t words) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < words; i += stride) {
        const uint32_t v = src[i];
        if (v) dst[i] = v;
    }
}

extern "C" int kmat_ctx_peer_attach(kmat_ctx *c, int n_shards, const kmat_peer_info *all) {
    if (!c || !all || n_shards < 1 || n_shards > KM_MAX_SHARDS) { kmat_set_error("kmat_ctx_peer_attach: bad argument"); return KMAT_ERR_ARG; }
    if (c->d_peers) { kmat_set_error("kmat_ctx_peer_attach: peers already attached"); return KMAT_ERR_ARG; }
    const kmat_db *db = c->db;
    if (db->shard_count != n_shards) { kmat_set_error("kmat_ctx_peer_attach: the ctx's table is shard %d of %d, not of %d", db->shard_index, db->shard_count, n_shards); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    std::vector<KmPeer> peers((size_t)n_shards);
    const int me = (int)getpid();
    for (int s = 0; s < n_shards; s++) {
        KmPeerBlob b;
        memcpy(&b, &all[s], sizeof b);
        if (b.magic != KM_PEER_MAGIC || b.shard_index != s || b.shard_count != n_shards) { kmat_set_error("kmat_ctx_peer_attach: entry %d is not the export of shard %d of %d", s, s, n_shards); return KMAT_ERR_ARG; }
        if (b.line_m != db->geom.line_m || b.line_bits != db->geom.line_bits || b.kmer_bits != db->geom.kmer_bits || b.tid_bytes != db->tid_bytes) {
            kmat_set_error("kmat_ctx_peer_attach: shard %d has a different table geometry (2^%d lines, here 2^%d)", s, b.line_bits, db->geom.line_bits); return KMAT_ERR_UNSUPPORTED; }
        if (b.pool2_mul != c->pool2_mul || b.rkmer != (c->opt.rkmer_mode != 0) || b.permissive != (c->opt.permissive != 0) || b.max_count != c->opt.max_count) {
            kmat_set_error("kmat_ctx_peer_attach: shard %d's context was created with different options (-g / -s / rkmer)", s); return KMAT_ERR_ARG; }
        if (!b.pool_shared && b.pool_words > (1ull << KM_PEER_SHIFT)) { kmat_set_error("kmat_ctx_peer_attach: shard %d's list pool (%llu words) exceeds the 2^%d-word offset range of direct mode", s, (unsigned long long)b.pool_words, KM_PEER_SHIFT); return KMAT_ERR_UNSUPPORTED; }
        KmPeer &p = peers[(size_t)s];
        p.n_stash = b.n_stash; p.pool_base = KM_PEER_NO_BASE;
        p.line_first = b.line_first; p.lines = nullptr; p.slots = nullptr; p.rem_bits = b.rem_bits; p.bucket_mask = (1ull << b.bucket_bits) - 1;
        if (s == db->shard_index) {
            p.lines = db->d_lines; p.slots = db->d_slots; p.stash_x = db->d_stash_x; p.stash_hit = db->d_stash_hit; p.pool2 = c->d_pool2;
        } else if (b.pid == me) {
            // same process: the exporter's pointers are ours too; another device needs peer access switched on
            if (b.device != c->device) {
                int can = 0;
                KM_CUDA(cudaDeviceCanAccessPeer(&can, c->device, b.device));
                if (!can) { kmat_set_error("kmat_ctx_peer_attach: device %d cannot access device %d", c->device, b.device); return KMAT_ERR_UNSUPPORTED; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { kmat_set_error("cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e)); cudaGetLastError(); return KMAT_ERR_CUDA; }
                cudaGetLastError();
            }
            p.lines = (const uint64_t *)b.p_lines; p.slots = (const uint64_t *)b.p_slots; p.stash_x = (const uint64_t *)b.p_stash_x; p.stash_hit = (const uint32_t *)b.p_stash_hit; p.pool2 = (const uint32_t *)b.p_pool2;
        } else {
            auto open = [&](const cudaIpcMemHandle_t &h, const void **dst) -> int {
                void *q = nullptr;
                KM_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
                c->ipc_mapped.push_back(q);
                *dst = q;
                return KMAT_OK;
            };
            int rc;
            if (b.p_lines && (rc = open(b.h_lines, (const void **)&p.lines)) != KMAT_OK) return rc;
            if (b.p_slots && (rc = open(b.h_slots, (const void **)&p.slots)) != KMAT_OK) return rc;
            p.stash_x = nullptr; p.stash_hit = nullptr; p.pool2 = nullptr;
            if (b.n_stash) { if ((rc = open(b.h_stash_x, (const void **)&p.stash_x)) != KMAT_OK) return rc; if ((rc = open(b.h_stash_hit, (const void **)&p.stash_hit)) != KMAT_OK) return rc; }
            if (b.p_pool2) { if ((rc = open(b.h_pool2, (const void **)&p.pool2)) != KMAT_OK) return rc; }
        }
    }
    // List records.  Default: copy every shard's resolved pool into one local pool (a bulk NVLink copy, once) -- list
    // hits then never leave the GPU again; the hit words carry offsets into the concatenation.  When the pools together
    // are too large for that (KMAT_PEER_LISTS=fetch forces it), list hits are tagged with their owner instead and the
    // records of every pass are fetched from the owners (km_peer_fetch_kernel).
    bool all_shared = db->pool_shared;
    for (int s = 0; s < n_shards; s++) { KmPeerBlob b; memcpy(&b, &all[s], sizeof b); all_shared = all_shared && b.pool_shared && b.pool_words == db->pool_words; }
    if (all_shared && db->pool_words) {
        // every shard was built from the whole table's arrays and indexes the same list pool, but has resolved only the lists of
        // its own k-mers (km_resolve_lists leaves the rest zero): OR the peers' resolved pools into this rank's -- one bulk
        // read of every peer's pool over NVLink, once -- and every list hit is answered locally with its offset unchanged.
        // A peer that merges at the same time only ever turns a zero word into its final value, so the order does not matter.
        const uint64_t words = db->pool_words * (uint64_t)c->pool2_mul;
        for (int s = 0; s < n_shards; s++) {
            peers[(size_t)s].pool_base = 0;
            if (s == db->shard_index) continue;
            if (!peers[(size_t)s].pool2) { kmat_set_error("kmat_ctx_peer_attach: shard %d exports no resolved list pool", s); return KMAT_ERR_ARG; }
            km_pool_merge_kernel<<<148 * 8, 256, 0, c->stream>>>(c->d_pool2, peers[(size_t)s].pool2, words);
            g_km_launches++;
        }
        KM_CUDA(cudaStreamSynchronize(c->stream));
        c->d_pool2_all = c->d_pool2; c->pool2_all_alias = true;
    } else if (all_shared) {
        for (int s = 0; s < n_shards; s++) peers[(size_t)s].pool_base = 0;
    } else {
        uint64_t total_raw = 0;
        std::vector<uint64_t> raw(n_shards);
        for (int s = 0; s < n_shards; s++) { KmPeerBlob b; memcpy(&b, &all[s], sizeof b); raw[s] = b.pool_words; total_raw += b.pool_words; }
        size_t free_b = 0, total_b = 0;
        KM_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t bytes = (total_raw * (uint64_t)c->pool2_mul + 64) * 4;
        const char *mode = getenv("KMAT_PEER_LISTS");
        bool replicate = total_raw > 0 && total_raw < (1ull << 31) && bytes < free_b / 4;
        if (mode && strcmp(mode, "fetch") == 0) replicate = false;
        if (mode && strcmp(mode, "replicate") == 0 && total_raw > 0 && total_raw < (1ull << 31)) replicate = true;
        if (replicate) {
            KM_CUDA(cudaMalloc((void **)&c->d_pool2_all, bytes));
            uint64_t at = 0;
            for (int s = 0; s < n_shards; s++) {
                if (raw[s]) KM_CUDA(cudaMemcpy(c->d_pool2_all + at * c->pool2_mul, peers[(size_t)s].pool2, raw[s] * (uint64_t)c->pool2_mul * 4, cudaMemcpyDefault));
                peers[(size_t)s].pool_base = (uint32_t)at;
                at += raw[s];
            }
        }
    }
    KM_CUDA(cudaMalloc((void **)&c->d_peers, peers.size() * sizeof(KmPeer)));
    KM_CUDA(cudaMemcpy(c->d_peers, peers.data(), peers.size() * sizeof(KmPeer), cudaMemcpyHostToDevice));
    c->n_peers = (uint32_t)n_shards;
    return KMAT_OK;
}


// ---- direct mode: bring the list records of a pass home ---------------------------------------------------------------
// After K1+K2 a list hit word is LIST | owner << 27 | offset into the OWNER's resolved pool.  One thread per hit word
// reads the first two 32-byte sectors of the record from the owner's memory (two NVLink reads in flight per list hit,
// thousands per SM), appends the record to the pass's local buffer and rewrites the hit word to LIST | local offset
// (pool2_mul = 1, exactly what K3 gets in the exchange variant).
struct KmFetchParams {
    uint32_t *hit; uint64_t n_pos; const KmPeer *peers; int mul, permissive;
    uint32_t *recs; unsigned long long *cur; unsigned long long cap;
};
#define KM_FETCH_TILE 4096
__global__ void __launch_bounds__(256) km_peer_fetch_kernel(KmFetchParams F) {
    // A CTA takes a tile of 4096 hit words: the list hits among them (~3 %) are first queued in shared memory, then one
    // thread per queued hit issues its two remote sector reads -- so a tile waits ONCE for the NVLink round trip (tens of
    // microseconds while the peers' probe kernels keep the fabric saturated), not once per 32 hit words.
    __shared__ uint32_t s_pos[KM_FETCH_TILE];
    __shared__ uint32_t s_n, s_warp[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t n_tiles = (F.n_pos + KM_FETCH_TILE - 1) / KM_FETCH_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * KM_FETCH_TILE;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int j = 0; j < KM_FETCH_TILE / 256; j++) {
            const uint64_t i = t0 + (uint64_t)j * 256 + threadIdx.x;
            const uint32_t hw = i < F.n_pos ? F.hit[i] : KM_HIT_INVALID;
            const bool is_list = hw != KM_HIT_INVALID && hw != KM_HIT_MISS && (hw & KM_HIT_LIST);
            const uint32_t bal = __ballot_sync(KM_FULL, is_list);
            if (bal) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_n, (uint32_t)__popc(bal));
                base = __shfl_sync(KM_FULL, base, 0);
                if (is_list) s_pos[base + __popc(bal & ((1u << lane) - 1))] = (uint32_t)(i - t0);
            }
        }
        __syncthreads();
        const uint32_t n = s_n;
        for (uint32_t e0 = 0; e0 < n; e0 += 256) {
            const uint32_t e = e0 + threadIdx.x;
            const bool act = e < n;
            uint32_t S[16];
            const uint32_t *rec = nullptr;
            uint32_t w0 = 0, len = 0;
            uint64_t i = 0;
            if (act) {
                i = t0 + s_pos[e];
                const uint32_t hw = F.hit[i];
                const KmPeer pr = F.peers[(hw >> KM_PEER_SHIFT) & (KM_MAX_SHARDS - 1)];
                rec = pr.pool2 + (size_t)(hw & KM_PEER_OFFMASK) * F.mul;
                const uint64_t *a0 = (const uint64_t *)((uintptr_t)rec & ~(uintptr_t)31);
                w0 = (uint32_t)(((uintptr_t)rec & 31) >> 2);
                uint64_t q[8];
                km_load_bucket(a0, q[0], q[1], q[2], q[3]);
                km_load_bucket(a0 + 4, q[4], q[5], q[6], q[7]);
#pragma unroll
                for (int j = 0; j < 8; j++) { S[2 * j] = (uint32_t)q[j]; S[2 * j + 1] = (uint32_t)(q[j] >> 32); }
                const uint32_t h = S[w0];
                if (h == KR_ERR_BAD) len = 1;                               // the candidate kernel reports the bad stored id
                else len = F.permissive ? 2 + (h & 0xFFFFu) + S[w0 + 1] : 1 + (h & 0xFFFFu);
            }
            // CTA-wide allocation in the local record buffer: one global atomic per round
            uint32_t incl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(KM_FULL, incl, d); if (lane >= d) incl += t; }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t v = s_warp[w]; if (w < wid) before += v; total += v; }
            if (threadIdx.x == 0) s_base = atomicAdd(F.cur, (unsigned long long)total);
            __syncthreads();
            if (act) {
                const unsigned long long dst = s_base + before + (incl - len);
                if (dst + len > F.cap) { atomicAdd(F.cur + 1, 1ull); F.hit[i] = KM_HIT_MISS; }
                else {
                    for (uint32_t w = 0; w < len; w++) F.recs[dst + w] = w0 + w < 16 ? S[w0 + w] : rec[w];
                    F.hit[i] = KM_HIT_LIST | (uint32_t)dst;
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

static int km_peer_prepare(kmat_ctx *c, const KmPass &L, cudaStream_t st) {
    if (!c->d_peer_cur) { KM_CUDA(cudaMalloc((void **)&c->d_peer_cur, 16)); KM_CUDA(cudaMemsetAsync(c->d_peer_cur, 0, 16, st)); }
    // room for one short record at every other k-mer position (C2: ~5 list hits of ~6 words per 131 positions)
    uint64_t want = std::max<uint64_t>(1u << 20, L.total_bases / 2 * (uint64_t)c->peer_grow);
    if (const char *e = getenv("KMAT_TEST_PEER_RECS")) want = (uint64_t)atoll(e) * (uint64_t)c->peer_grow;      // tests: force the overflow path
    if (want >= (1ull << 31)) want = (1ull << 31) - 1;
    if (want > c->cap_peer_recs) { KM_CUDA(cudaStreamSynchronize(st)); int rc = km_grow(&c->d_peer_recs, &c->cap_peer_recs, want); if (rc != KMAT_OK) return rc; }
    KM_CUDA(cudaMemsetAsync(c->d_peer_cur, 0, 8, st));               // words used; the dropped counter is monotonic
    return KMAT_OK;
}
static int km_peer_fetch(kmat_ctx *c, const KmPass &L, cudaStream_t st) {
    KmFetchParams F;
    F.hit = c->d_hit; F.n_pos = L.total_bases; F.peers = c->d_peers; F.mul = c->pool2_mul; F.permissive = c->opt.permissive != 0;
    F.recs = c->d_peer_recs; F.cur = c->d_peer_cur; F.cap = std::min<uint64_t>(c->cap_peer_recs, (1ull << 31) - 1);
    if (const char *e = getenv("KMAT_TEST_PEER_RECS")) F.cap = std::min<uint64_t>(F.cap, (uint64_t)atoll(e) * (uint64_t)c->peer_grow);
    const int grid = (int)std::min<uint64_t>((L.total_bases + KM_FETCH_TILE - 1) / KM_FETCH_TILE, (uint64_t)c->sms * 8);
    km_peer_fetch_kernel<<<std::max(1, grid), 256, 0, st>>>(F);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}
