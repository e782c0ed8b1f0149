// kmat_gene.cuh -- gene_label's per-read work on the GPU; included at the end of kmat_label.cu.
//
// Restates retrieve_kmer_labels + the top-gene pick of proc_line in src/gene_label.cpp (:217-301): the unique canonical
// k-mers of the read (K1, shared with read_label: first occurrence wins) are looked up in a gene DB (K2; 32-bit ids,
// no pruning, no id map: TaxNodeStat::begin(kmer, NULL)), every id of every hit list counts once per k-mer, the ids
// keep their first-appearance order (geneid_lst, :249-258), std::sort by count descending (Cmp, :84-88; libstdc++'s
// order on ties, kmstd::sort) puts the call in front and its score is count / unique k-mers (:297-299).
struct KmGeneParams {
    KmDbDev db;
    const uint64_t *offs; uint32_t n_reads;
    const uint32_t *hit;
    kmat_gene_result *out;
    const uint32_t *stored;        // dense id -> stored id for 32-bit tables (NULL: the id is the stored id)
    uint32_t *big_q; unsigned int *big_cnt; uint32_t big_cap;     // reads with more than KB_CMAX genes: queued for km_gene_big_kernel
};
struct KgCountDesc {
    __device__ bool operator()(const uint2 &a, const uint2 &b) const { return a.y > b.y; }
};

__global__ void __launch_bounds__(KB_WARPS * 32) km_gene_kernel(KmGeneParams P) {
    __shared__ uint2 s_sort[KB_WARPS][KB_CMAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.x * KB_WARPS + wib, n_warps = gridDim.x * KB_WARPS;
    const int k = P.db.kmer_len;
    for (uint32_t r = warp_global; r < P.n_reads; r += n_warps) {
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        kmat_gene_result res;
        res.status = 0; res.valid_kmers = 0; res.n_genes = 0; res.gene = 0; res.count = 0; res.score = 0.0f;
        if (np <= 0) { if (lane == 0) P.out[r] = res; continue; }          // ri_len < k_size: nothing is printed (:278-282)
        KbCand K;
        K.nid[0] = K.nid[1] = KMAT_NONE; K.key[0] = K.key[1] = 0xFFFFFFFFu; K.leaf[0] = K.leaf[1] = 0;
        int C = 0, cnt = 0;
        bool overflow = false;
        const int nch = (np + 31) >> 5;
        for (int c = 0; c < nch && !overflow; c++) {
            const int p = (c << 5) + lane;
            const uint32_t hw = p < np ? __ldg(P.hit + off + p) : KM_HIT_INVALID;
            uint32_t a = 0, lo = 0;
            if (hw != KM_HIT_INVALID) {
                cnt++;                                                    // ++valid_cnt: a unique k-mer (:245)
                if (hw != KM_HIT_MISS) {
                    if (hw & KM_HIT_LIST) { lo = hw & 0x7FFFFFFFu; a = kb_list_count(P.db, lo); }
                    else a = 1;
                }
            }
            for (uint32_t seqno = 0;; seqno++) {
                uint32_t val = KMAT_NONE;
                if (seqno < a) val = (hw & KM_HIT_LIST) ? kb_list_id(P.db, lo, seqno) : hw;
                uint32_t pending = __ballot_sync(KM_FULL, val != KMAT_NONE);
                if (!pending) break;
                while (pending) {                                         // one round per distinct id among the lanes
                    const int leader = __ffs(pending) - 1;
                    const uint32_t v = __shfl_sync(KM_FULL, val, leader);
                    const uint32_t grp = __ballot_sync(KM_FULL, val == v);
                    const int idx = kb_find_or_add(K, C, v, lane);
                    if (idx < 0) { overflow = true; break; }
                    if (lane == (idx & 31)) {
                        const uint32_t key = ((uint32_t)min((c << 5) + leader, 0xFFFF) << 16) | min(seqno, 0xFFFFu);
                        if (idx < 32) { K.key[0] = min(K.key[0], key); K.leaf[0] += __popc(grp); }
                        else { K.key[1] = min(K.key[1], key); K.leaf[1] += __popc(grp); }
                    }
                    pending &= ~grp;
                }
                if (overflow) break;
            }
        }
        cnt = km_warp_sum(cnt);
        overflow = __any_sync(KM_FULL, overflow) || np > 0xFFFF;           // first-appearance keys hold 16-bit positions
        res.valid_kmers = (uint32_t)cnt; res.n_genes = (uint32_t)C;
        if (overflow) {                                                    // more than 64 genes (or a very long read): the big kernel takes it
            res.status = KMAT_ERR_UNSUPPORTED;
            if (lane == 0) {
                if (P.big_q) { const unsigned int q = atomicAdd(P.big_cnt, 1u); if (q < P.big_cap) P.big_q[q] = r; }
                P.out[r] = res;
            }
            continue;
        }
        if (C == 0) { if (lane == 0) P.out[r] = res; continue; }           // geneid_lst.empty(): nothing is printed (:309-312)
        // geneid_lst order = rank of the first-appearance key
        uint32_t ord[2] = {0, 0};
        for (int j = 0; j < C; j++) {
            const uint32_t kj = __shfl_sync(KM_FULL, j < 32 ? K.key[0] : K.key[1], j & 31);
#pragma unroll
            for (int s = 0; s < 2; s++) ord[s] += kj < K.key[s];
        }
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (lane + 32 * s < C) s_sort[wib][ord[s]] = make_uint2(P.stored ? P.stored[K.nid[s]] : K.nid[s], K.leaf[s]);
        __syncwarp();
        if (lane == 0) {
            kmstd::sort(s_sort[wib], C, KgCountDesc());                    // sort(gsort.begin(), gsort.end(), Cmp()) (:297)
            res.status = 1; res.gene = s_sort[wib][0].x; res.count = s_sort[wib][0].y;
            res.score = __fdiv_rn((float)res.count, (float)cnt);          // (float)gsort[0].second / (float)cnt (:298)
            P.out[r] = res;
        }
        __syncwarp();
    }
}

// The reads km_gene_kernel could not hold in registers: one warp per read, up to KG_CBIG distinct genes in shared memory
// (id, first-appearance key, count), same member walk, same final std::sort.
#define KG_CBIG 1024
#define KG_WARPS 8
struct KgBigW { uint32_t gid[KG_CBIG], cnt[KG_CBIG]; unsigned long long key[KG_CBIG]; uint2 sorted[KG_CBIG]; };     // 24 KB per warp
__global__ void __launch_bounds__(KG_WARPS * 32) km_gene_big_kernel(KmGeneParams P) {
    extern __shared__ __align__(16) unsigned char kg_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    KgBigW &W = *(KgBigW *)(kg_smem + (size_t)wib * sizeof(KgBigW));
    const uint32_t warp_global = blockIdx.x * KG_WARPS + wib, n_warps = gridDim.x * KG_WARPS;
    const int k = P.db.kmer_len;
    const uint32_t nq = min(*P.big_cnt, P.big_cap);
    for (uint32_t q = warp_global; q < nq; q += n_warps) {
        const uint32_t r = P.big_q[q];
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        kmat_gene_result res;
        res.status = 0; res.valid_kmers = 0; res.n_genes = 0; res.gene = 0; res.count = 0; res.score = 0.0f;
        int C = 0, cnt = 0;
        bool overflow = false;
        const int nch = (np + 31) >> 5;
        for (int c = 0; c < nch && !overflow; c++) {
            const int p = (c << 5) + lane;
            const uint32_t hw = p < np ? __ldg(P.hit + off + p) : KM_HIT_INVALID;
            uint32_t a = 0, lo = 0;
            if (hw != KM_HIT_INVALID) {
                cnt++;
                if (hw != KM_HIT_MISS) { if (hw & KM_HIT_LIST) { lo = hw & 0x7FFFFFFFu; a = kb_list_count(P.db, lo); } else a = 1; }
            }
            for (uint32_t seqno = 0;; seqno++) {
                uint32_t val = KMAT_NONE;
                if (seqno < a) val = (hw & KM_HIT_LIST) ? kb_list_id(P.db, lo, seqno) : hw;
                uint32_t pending = __ballot_sync(KM_FULL, val != KMAT_NONE);
                if (!pending) break;
                while (pending) {
                    const int leader = __ffs(pending) - 1;
                    const uint32_t v = __shfl_sync(KM_FULL, val, leader);
                    const uint32_t grp = __ballot_sync(KM_FULL, val == v);
                    int idx = -1;
                    for (int t = 0; t < C && idx < 0; t += 32) {
                        const uint32_t f = __ballot_sync(KM_FULL, t + lane < C && W.gid[t + lane] == v);
                        if (f) idx = t + __ffs(f) - 1;
                    }
                    if (idx < 0) {
                        if (C >= KG_CBIG) { overflow = true; break; }
                        idx = C++;
                        if (lane == 0) { W.gid[idx] = v; W.cnt[idx] = 0; W.key[idx] = ~0ull; }
                        __syncwarp();
                    }
                    if (lane == 0) {
                        const unsigned long long key = ((unsigned long long)((c << 5) + leader) << 32) | seqno;
                        if (key < W.key[idx]) W.key[idx] = key;
                        W.cnt[idx] += __popc(grp);
                    }
                    pending &= ~grp;
                }
                if (overflow) break;
            }
            __syncwarp();
        }
        cnt = km_warp_sum(cnt);
        overflow = __any_sync(KM_FULL, overflow);
        res.valid_kmers = (uint32_t)cnt; res.n_genes = (uint32_t)C;
        if (overflow) { res.status = KMAT_ERR_UNSUPPORTED; if (lane == 0) P.out[r] = res; continue; }
        if (C == 0) { if (lane == 0) P.out[r] = res; continue; }
        for (int i = lane; i < C; i += 32) {                                   // geneid_lst order = rank of the first-appearance key
            uint32_t ord = 0;
            const unsigned long long ki = W.key[i];
            for (int j = 0; j < C; j++) ord += W.key[j] < ki;
            W.sorted[ord] = make_uint2(P.stored ? P.stored[W.gid[i]] : W.gid[i], W.cnt[i]);
        }
        __syncwarp();
        if (lane == 0) {
            kmstd::sort(W.sorted, C, KgCountDesc());
            res.status = 1; res.gene = W.sorted[0].x; res.count = W.sorted[0].y;
            res.score = __fdiv_rn((float)res.count, (float)cnt);
            P.out[r] = res;
        }
        __syncwarp();
    }
}

extern "C" int kmat_gene_batch(const kmat_db *db, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_gene_result *out) {
    if (!db || !offs || !out || (n_reads && !bases)) { kmat_set_error("kmat_gene_batch: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= db->device) { kmat_set_error("CUDA device %d not available", db->device); return KMAT_ERR_NO_DEVICE; }
    if (!n_reads) return KMAT_OK;
    KM_CUDA(cudaSetDevice(db->device));
    const uint32_t chunk_reads = 1u << 20;
    const uint64_t chunk_bases = (uint64_t)256 << 20;
    char *d_b = nullptr; uint64_t *d_o = nullptr; uint32_t *d_hit = nullptr; int2 *d_hdr = nullptr; kmat_gene_result *d_out = nullptr;
    unsigned long long *d_long = nullptr; uint32_t long_slots = 0;
    uint64_t cap_b = 0; uint32_t cap_r = 0;
    int rc = KMAT_OK;
    uint32_t *d_bigq = nullptr; unsigned int *d_bigcnt = nullptr;
    const uint32_t bigq_cap = chunk_reads;
    if (cudaMalloc((void **)&d_bigq, (size_t)bigq_cap * 4) != cudaSuccess || cudaMalloc((void **)&d_bigcnt, 4) != cudaSuccess) {
        cudaFree(d_bigq); cudaGetLastError(); kmat_set_error("kmat_gene_batch: out of device memory"); return KMAT_ERR_NOMEM;
    }
    KM_CUDA(cudaFuncSetAttribute(km_gene_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KG_WARPS * sizeof(KgBigW))));   // per call: the attribute is per device
    auto cleanup = [&] { cudaFree(d_b); cudaFree(d_o); cudaFree(d_hit); cudaFree(d_hdr); cudaFree(d_out); cudaFree(d_long); cudaFree(d_bigq); cudaFree(d_bigcnt); };
    for (uint32_t r0 = 0; r0 < n_reads && rc == KMAT_OK;) {
        uint32_t r1 = (uint32_t)std::min<uint64_t>(n_reads, (uint64_t)r0 + chunk_reads);
        while (r1 > r0 + 1 && offs[r1] - offs[r0] > chunk_bases) r1 = r0 + (r1 - r0) / 2;
        const uint32_t n = r1 - r0;
        const uint64_t nb = offs[r1] - offs[r0];
        uint32_t max_len = 0;
        for (uint32_t r = r0; r < r1; r++) max_len = std::max<uint32_t>(max_len, (uint32_t)(offs[r + 1] - offs[r]));
        if (nb + 1 > cap_b) { cudaFree(d_b); cudaFree(d_hit); d_b = nullptr; d_hit = nullptr; cap_b = nb + nb / 4 + 1024;
                              if (cudaMalloc((void **)&d_b, cap_b) != cudaSuccess || cudaMalloc((void **)&d_hit, cap_b * 4) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; } }
        if (n > cap_r) { cudaFree(d_o); cudaFree(d_hdr); cudaFree(d_out); d_o = nullptr; d_hdr = nullptr; d_out = nullptr; cap_r = n + n / 4 + 64;
                         if (cudaMalloc((void **)&d_o, ((size_t)cap_r + 1) * 8) != cudaSuccess || cudaMalloc((void **)&d_hdr, (size_t)cap_r * sizeof(int2)) != cudaSuccess ||
                             cudaMalloc((void **)&d_out, (size_t)cap_r * sizeof(kmat_gene_result)) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; } }
        if (max_len > 256) {
            uint32_t slots = 1024; while (slots < 2 * max_len) slots <<= 1;
            if (slots > long_slots) { cudaFree(d_long); d_long = nullptr; long_slots = slots;
                                      if (cudaMalloc((void **)&d_long, (size_t)148 * 6 * KM_PROBE_WARPS_HOST * slots * 8) != cudaSuccess) { rc = KMAT_ERR_NOMEM; break; } }
        }
        std::vector<uint64_t> lo(n + 1);
        for (uint32_t i = 0; i <= n; i++) lo[i] = offs[r0 + i] - offs[r0];
        if (cudaMemcpy(d_b, bases + offs[r0], nb, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d_o, lo.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        rc = km_launch_encode_probe(db, d_b, d_o, n, max_len, d_hit, d_hdr, nullptr, nullptr, d_long, long_slots, km_probe_grid(n), nullptr, 1, 0, 0, nullptr);
        if (rc != KMAT_OK) break;
        KmGeneParams P;
        P.db = km_db_dev(db); P.offs = d_o; P.n_reads = n; P.hit = d_hit; P.out = d_out; P.stored = db->d_stored_tids;
        P.big_q = d_bigq; P.big_cnt = d_bigcnt; P.big_cap = bigq_cap;
        if (cudaMemsetAsync(d_bigcnt, 0, 4) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        km_gene_kernel<<<std::max(1u, std::min((n + KB_WARPS - 1) / KB_WARPS, 148u * 4)), KB_WARPS * 32>>>(P);
        g_km_launches++;
        km_gene_big_kernel<<<148, KG_WARPS * 32, KG_WARPS * sizeof(KgBigW)>>>(P);           // the reads with more than 64 genes (usually none)
        g_km_launches++;
        if (cudaMemcpy(out + r0, d_out, (size_t)n * sizeof(kmat_gene_result), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = KMAT_ERR_CUDA; break; }
        r0 = r1;
    }
    if (rc == KMAT_ERR_NOMEM) kmat_set_error("kmat_gene_batch: out of device memory");
    if (rc == KMAT_ERR_CUDA) { kmat_set_error("kmat_gene_batch: %s", cudaGetErrorString(cudaGetLastError())); }
    cleanup();
    return rc;
}
