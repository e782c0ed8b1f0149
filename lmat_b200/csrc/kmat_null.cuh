// kmat_null.cuh -- null-model generation: rand_read_label's per-read work on the GPU; included at the
// end of kmat_label.cu.
//
// Reference: src/rand_read_label.cpp.  Per read of the OMP loop (:687-701): genRandRead (:85-103) draws a read for GC
// bucket i % 10, proc_line (:367-397) runs rkmer.hpp's retrieve_kmer_labels (rkmer.hpp:76-294: read_label's candidate
// sets WITHOUT the human collapse; valid_kmers counts duplicate positions too), counts for every taxid the positions whose
// set holds it (:382-393) and construct_labels (:184-213) keeps per (taxid, bucket) the maximum of count / valid_kmers
// and the number of reads; the per-thread maps are merged by max / sum (:702-735) and written as <ofbase>.rand_lst
// (:736-755).
//
// Here: K1+K2 (kmat_db.cu) and K3 (km_cand_kernel, with opts.rkmer_mode) run unchanged -- K3's (node, hits) pairs ARE
// cnt_tids -- and km_nullacc_kernel folds them into two device arrays with atomicMax / atomicAdd (both order
// independent, so the result does not depend on scheduling).  Reads come from the caller (kmat_null_batch) or are drawn
// on the device by km_randgen_kernel from a counter-based generator (kmat_null_random).

// ---- read generator -------------------------------------------------------------------------------------------------
// The reference's generator is the process-wide glibc rand() seeded with time(0), shared by all OMP threads: not
// reproducible, and inherently serial.  Replacement: splitmix64 finaliser keyed by (seed, run index), one 64-bit draw per
// base.  Same distribution as genRandRead + std::random_shuffle: the number of g/c bases is the reference's float
// expression, WHICH positions hold them is a uniformly random subset (selection sampling: position i takes a g/c with
// probability remaining_gc / remaining_positions), and each base's letter within its class is a fair coin.
KM_HD uint64_t kn_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
KM_HD uint64_t kn_read_key(uint64_t seed, uint64_t index) { return kn_mix64(seed ^ kn_mix64(index + 0x9E3779B97F4A7C15ull)); }
KM_HD uint64_t kn_draw(uint64_t key, uint64_t j) { return kn_mix64(key + j * 0x9E3779B97F4A7C15ull); }

#define KN_GEN_THREADS 256
#define KN_GEN_SMEM (96 * 1024)

__global__ void __launch_bounds__(KN_GEN_THREADS) km_randgen_kernel(uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t rl,
                                                                   uint32_t reads_per_block, char *bases, uint64_t *offs) {
    extern __shared__ char s_reads[];
    const uint32_t r0 = blockIdx.x * reads_per_block;
    const uint32_t nb = min(reads_per_block, n_reads - r0);
    if (threadIdx.x < nb) {
        const uint32_t r = r0 + threadIdx.x;
        const uint64_t index = first_index + r;
        const uint64_t key = kn_read_key(seed, index);
        const int bucket = (int)(index % KMAT_NULL_BUCKETS);                                     // :693
        const int gc_draw = bucket * 10 + (int)((((kn_draw(key, 0) >> 32) * 10ull) >> 32));      // uniform in gc_range[bucket] (:88, :677-685)
        const float gc_pcnt = (float)((double)gc_draw / 100.0);                                  // :89
        uint32_t rem_gc = (uint32_t)__fmul_rn(gc_pcnt, (float)rl);                               // :90
        char *dst = s_reads + (size_t)threadIdx.x * rl;
        for (uint32_t i = 0; i < rl; i++) {
            const uint64_t d = kn_draw(key, 1 + i);
            const uint32_t rem_pos = rl - i;
            const bool gc = (uint32_t)(((d >> 32) * (uint64_t)rem_pos) >> 32) < rem_gc;
            const bool coin = d & 1;
            dst[i] = gc ? (coin ? 'g' : 'c') : (coin ? 'a' : 't');
            rem_gc -= gc;
        }
        offs[r] = (uint64_t)r * rl;
        if (r == n_reads - 1) offs[n_reads] = (uint64_t)n_reads * rl;
    }
    __syncthreads();
    char *g = bases + (size_t)r0 * rl;
    const size_t total = (size_t)nb * rl;
    for (size_t i = threadIdx.x; i < total; i += KN_GEN_THREADS) g[i] = s_reads[i];
}

static int km_launch_randgen(uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t read_len, char *d_bases, uint64_t *d_offs, cudaStream_t st) {
    if (read_len == 0 || read_len > KN_GEN_SMEM) { kmat_set_error("kmat_null: read_len %u out of range (1..%d)", read_len, KN_GEN_SMEM); return KMAT_ERR_ARG; }
    const uint32_t rpb = std::max(1u, std::min<uint32_t>(KN_GEN_THREADS, KN_GEN_SMEM / read_len));
    KM_CUDA(cudaFuncSetAttribute(km_randgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KN_GEN_SMEM));   // per launch: the attribute is per device
    km_randgen_kernel<<<(n_reads + rpb - 1) / rpb, KN_GEN_THREADS, (size_t)rpb * read_len, st>>>(seed, first_index, n_reads, read_len, rpb, d_bases, d_offs);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}

// ---- accumulation ---------------------------------------------------------------------------------------------------
struct KmNullParams {
    const kmat_read_result *out; const kmat_pair *cands; uint32_t n_reads; uint64_t first_index;
    uint32_t *max_bits, *cnt; unsigned long long *n_err;
};
// One thread per read: construct_labels of rand_read_label.cpp (:184-213) over K3's (node, hits) pairs.
__global__ void __launch_bounds__(256) km_nullacc_kernel(KmNullParams N) {
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    if (r >= N.n_reads) return;
    const kmat_read_result *o = N.out + r;
    const int status = o->status;
    if (status == KMAT_ST_ERROR) { atomicAdd(N.n_err, 1ull); return; }
    if (status != KMAT_ST_PENDING && status != KMAT_ST_PENDING_BIG && status != KMAT_ST_PENDING_HUGE) return;                   // shorter than k, no valid k-mer, or no taxid at all: nothing to count
    const uint32_t bucket = (uint32_t)((N.first_index + r) % KMAT_NULL_BUCKETS);
    const float valid = (float)o->valid_kmers;
    const kmat_pair *cp = N.cands + o->cand_off;
    const uint32_t n = o->n_cand;
    for (uint32_t i = 0; i < n; i++) {
        const kmat_pair p = cp[i];
        const float label_prob = __fdiv_rn((float)(int)__float_as_uint(p.score), valid);        // (float)found_genome_cnt / (float)cand_kmer_cnt (:195)
        const size_t cell = (size_t)p.tid * KMAT_NULL_BUCKETS + bucket;
        atomicMax(N.max_bits + cell, __float_as_uint(label_prob));                              // label_prob >= 0: float order == bit order (:203-209)
        atomicAdd(N.cnt + cell, 1u);                                                            // match_cnt[taxid][gcbucket] += 1 (:210)
    }
}

static int km_null_alloc(kmat_ctx *c) {
    if (c->d_null_max) return KMAT_OK;
    const size_t cells = (size_t)c->h.nodeA.size() * KMAT_NULL_BUCKETS;
    KM_CUDA(cudaMalloc((void **)&c->d_null_max, cells * 4));
    KM_CUDA(cudaMalloc((void **)&c->d_null_cnt, cells * 4));
    KM_CUDA(cudaMalloc((void **)&c->d_null_err, 8));
    KM_CUDA(cudaMemsetAsync(c->d_null_max, 0, cells * 4, c->stream));
    KM_CUDA(cudaMemsetAsync(c->d_null_cnt, 0, cells * 4, c->stream));
    KM_CUDA(cudaMemsetAsync(c->d_null_err, 0, 8, c->stream));
    return KMAT_OK;
}

static int km_launch_nullacc(kmat_ctx *c, const KmScoreParams &P, uint32_t r0, uint32_t n, cudaStream_t s2) {
    KmNullParams N;
    N.out = P.out; N.cands = P.cands; N.n_reads = n; N.first_index = c->null_first + r0;
    N.max_bits = c->d_null_max; N.cnt = c->d_null_cnt; N.n_err = c->d_null_err;
    km_nullacc_kernel<<<(n + 255) / 256, 256, 0, s2>>>(N);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}

static int km_null_check(kmat_ctx *c, const char *who) {
    if (!c) { kmat_set_error("%s: bad argument", who); return KMAT_ERR_ARG; }
    if (!c->opt.rkmer_mode) { kmat_set_error("%s: the ctx was not created with opts.rkmer_mode = 1", who); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    return km_null_alloc(c);
}

// K1+K2, K3 and the accumulation over reads resident on the device (offsets relative to d_bases)
static int km_null_pass(kmat_ctx *c, const char *d_bases, const uint64_t *d_offs, uint32_t n, uint64_t total_bases, uint32_t max_len, uint64_t first_index) {
    int rc;
    uint64_t cap = c->cap_out_dev;
    if ((rc = km_grow(&c->d_out_dev, &cap, n)) != KMAT_OK) return rc;
    c->cap_out_dev = (uint32_t)cap;
    if ((rc = km_grow_cands(c, (uint64_t)n * KB_CMAX + 4096)) != KMAT_OK) return rc;        // K3 emits <= KB_CMAX pairs per read: no overflow possible
    c->null_first = first_index;
    KmPass L{d_bases, d_offs, n, 0, total_bases, max_len, c->d_out_dev, true};
    return km_run_device(c, L, c->stream);
}

extern "C" int kmat_null_reset(kmat_ctx *c) {
    int rc = km_null_check(c, "kmat_null_reset");
    if (rc != KMAT_OK) return rc;
    const size_t cells = (size_t)c->h.nodeA.size() * KMAT_NULL_BUCKETS;
    KM_CUDA(cudaMemsetAsync(c->d_null_max, 0, cells * 4, c->stream));
    KM_CUDA(cudaMemsetAsync(c->d_null_cnt, 0, cells * 4, c->stream));
    KM_CUDA(cudaMemsetAsync(c->d_null_err, 0, 8, c->stream));
    KM_CUDA(cudaStreamSynchronize(c->stream));
    return KMAT_OK;
}

static const uint32_t KN_CHUNK_READS = 1u << 20;
static const uint64_t KN_CHUNK_BASES = (uint64_t)256 << 20;

extern "C" int kmat_null_batch(kmat_ctx *c, const char *bases, const uint64_t *offs, uint32_t n_reads, uint64_t first_index) {
    int rc = km_null_check(c, "kmat_null_batch");
    if (rc != KMAT_OK) return rc;
    if (!offs || (n_reads && !bases)) { kmat_set_error("kmat_null_batch: bad argument"); return KMAT_ERR_ARG; }
    std::vector<uint64_t> lo;
    for (uint32_t r0 = 0; r0 < n_reads;) {
        uint32_t r1 = (uint32_t)std::min<uint64_t>(n_reads, (uint64_t)r0 + KN_CHUNK_READS);
        while (r1 > r0 + 1 && offs[r1] - offs[r0] > KN_CHUNK_BASES) r1 = r0 + (r1 - r0) / 2;
        const uint32_t n = r1 - r0;
        const uint64_t nb = offs[r1] - offs[r0];
        uint32_t max_len = 0;
        lo.resize((size_t)n + 1);
        for (uint32_t i = 0; i <= n; i++) lo[i] = offs[r0 + i] - offs[r0];
        for (uint32_t i = 0; i < n; i++) max_len = std::max<uint32_t>(max_len, (uint32_t)(lo[i + 1] - lo[i]));
        KM_CUDA(cudaStreamSynchronize(c->stream));                      // the previous chunk still reads the buffers
        if ((rc = km_grow(&c->d_null_bases, &c->cap_null_bases, nb + 16)) != KMAT_OK) return rc;
        if ((rc = km_grow(&c->d_null_offs, &c->cap_null_offs, (uint64_t)n + 1)) != KMAT_OK) return rc;
        KM_CUDA(cudaMemcpyAsync(c->d_null_bases, bases + offs[r0], nb, cudaMemcpyHostToDevice, c->stream));
        KM_CUDA(cudaMemcpyAsync(c->d_null_offs, lo.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        KM_CUDA(cudaStreamSynchronize(c->stream));                      // lo is reused by the next chunk
        if ((rc = km_null_pass(c, c->d_null_bases, c->d_null_offs, n, nb, max_len, first_index + r0)) != KMAT_OK) return rc;
        r0 = r1;
    }
    KM_CUDA(cudaStreamSynchronize(c->stream));
    return KMAT_OK;
}

extern "C" int kmat_null_random(kmat_ctx *c, uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t read_len) {
    int rc = km_null_check(c, "kmat_null_random");
    if (rc != KMAT_OK) return rc;
    if (read_len == 0) { kmat_set_error("kmat_null_random: read_len must be > 0"); return KMAT_ERR_ARG; }
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(KN_CHUNK_READS, KN_CHUNK_BASES / read_len));
    for (uint32_t r0 = 0; r0 < n_reads; r0 += chunk) {
        const uint32_t n = std::min(chunk, n_reads - r0);
        const uint64_t nb = (uint64_t)n * read_len;
        // same stream throughout: the generator of chunk i+1 queues behind the kernels of chunk i
        if (nb + 16 > c->cap_null_bases || (uint64_t)n + 1 > c->cap_null_offs) KM_CUDA(cudaStreamSynchronize(c->stream));
        if ((rc = km_grow(&c->d_null_bases, &c->cap_null_bases, nb + 16)) != KMAT_OK) return rc;
        if ((rc = km_grow(&c->d_null_offs, &c->cap_null_offs, (uint64_t)n + 1)) != KMAT_OK) return rc;
        if ((rc = km_launch_randgen(seed, first_index + r0, n, read_len, c->d_null_bases, c->d_null_offs, c->stream)) != KMAT_OK) return rc;
        if ((rc = km_null_pass(c, c->d_null_bases, c->d_null_offs, n, nb, read_len, first_index + r0)) != KMAT_OK) return rc;
    }
    KM_CUDA(cudaStreamSynchronize(c->stream));
    return KMAT_OK;
}

extern "C" int kmat_null_draw_reads(int device, uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t read_len, char *bases) {
    if (!bases || !read_len) { kmat_set_error("kmat_null_draw_reads: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= device) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    if (!n_reads) return KMAT_OK;
    KM_CUDA(cudaSetDevice(device));
    char *d_b = nullptr; uint64_t *d_o = nullptr;
    const size_t nb = (size_t)n_reads * read_len;
    if (cudaMalloc((void **)&d_b, nb) != cudaSuccess || cudaMalloc((void **)&d_o, ((size_t)n_reads + 1) * 8) != cudaSuccess) {
        cudaFree(d_b); cudaFree(d_o); cudaGetLastError(); kmat_set_error("kmat_null_draw_reads: out of device memory"); return KMAT_ERR_NOMEM;
    }
    int rc = km_launch_randgen(seed, first_index, n_reads, read_len, d_b, d_o, 0);
    if (rc == KMAT_OK && cudaMemcpy(bases, d_b, nb, cudaMemcpyDeviceToHost) != cudaSuccess) {
        kmat_set_error("kmat_null_draw_reads: %s", cudaGetErrorString(cudaGetLastError())); rc = KMAT_ERR_CUDA;
    }
    cudaFree(d_b); cudaFree(d_o);
    return rc;
}

extern "C" int kmat_null_fetch(kmat_ctx *c, uint32_t *tids, float *max_frac, uint64_t *counts, uint32_t cap_rows, uint32_t *n_rows, uint64_t *reads_error) {
    int rc = km_null_check(c, "kmat_null_fetch");
    if (rc != KMAT_OK) return rc;
    const size_t nodes = c->h.nodeA.size(), cells = nodes * KMAT_NULL_BUCKETS;
    std::vector<uint32_t> mx(cells), cn(cells);
    unsigned long long nerr = 0;
    KM_CUDA(cudaStreamSynchronize(c->stream));
    KM_CUDA(cudaMemcpy(mx.data(), c->d_null_max, cells * 4, cudaMemcpyDeviceToHost));
    KM_CUDA(cudaMemcpy(cn.data(), c->d_null_cnt, cells * 4, cudaMemcpyDeviceToHost));
    KM_CUDA(cudaMemcpy(&nerr, c->d_null_err, 8, cudaMemcpyDeviceToHost));
    if (reads_error) *reads_error = nerr;
    std::vector<std::pair<uint32_t, uint32_t>> rows;            // (tid, node id): max_match holds a taxid once a read hit it (:198-206)
    for (size_t nid = 0; nid < nodes; nid++) {
        bool any = false;
        for (int b = 0; b < KMAT_NULL_BUCKETS; b++) any |= cn[nid * KMAT_NULL_BUCKETS + b] != 0;
        if (any) rows.emplace_back(c->h.nodeA[nid].tid, (uint32_t)nid);
    }
    std::sort(rows.begin(), rows.end());                        // std::map<TID_T, ...> iteration order (:745)
    if (n_rows) *n_rows = (uint32_t)rows.size();
    if (rows.size() > cap_rows) { kmat_set_error("kmat_null_fetch: %zu rows, capacity %u", rows.size(), cap_rows); return KMAT_ERR_OVERFLOW; }
    for (size_t i = 0; i < rows.size(); i++) {
        if (tids) tids[i] = rows[i].first;
        for (int b = 0; b < KMAT_NULL_BUCKETS; b++) {
            const uint32_t bits = mx[(size_t)rows[i].second * KMAT_NULL_BUCKETS + b];
            if (max_frac) memcpy(&max_frac[i * KMAT_NULL_BUCKETS + b], &bits, 4);
            if (counts) counts[i * KMAT_NULL_BUCKETS + b] = cn[(size_t)rows[i].second * KMAT_NULL_BUCKETS + b];
        }
    }
    return KMAT_OK;
}
