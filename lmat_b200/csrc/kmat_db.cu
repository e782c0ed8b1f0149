// kmat_db.cu -- the HBM-resident k-mer table: build (insertion kernel), the K1 (encode) and K2 (probe)
// kernels, their parity hooks, and the random-gather roofline probe.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "kmat_device.cuh"
#include "kmat_priv.h"

std::atomic<unsigned long long> g_km_launches{0};
extern "C" uint64_t kmat_launch_count(void) { return g_km_launches.load(); }

extern "C" int kmat_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int kmat_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes) {
    if (kmat_device_count() <= device || device < 0) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    KM_CUDA(cudaSetDevice(device));
    size_t f = 0, t = 0;
    KM_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return KMAT_OK;
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
static int km_choose_bucket_bits(uint64_t n, int kmer_bits) {
    int b = 4;
    while (((uint64_t)1 << b) < n) b++;               // <= 1 key per 4-slot bucket on average: ~1.5 % of the home buckets are full
    if (getenv("KMAT_TEST_TIGHT_TABLE")) b = b > 6 ? b - 2 : 4;   // tests: a nearly full table exercises displacement, the stash and doubling
    if (b < kmer_bits - KM_REM_BITS) b = kmer_bits - KM_REM_BITS;
    if (b > kmer_bits) b = kmer_bits;
    return b;
}
// First level: 2^b 128-byte lines for n k-mers (the GLOBAL count: every shard of a table uses the same b).  Default
// density: at most 2 k-mers per 16-slot line on average (profiles/r02_line_table.md: ~4 % of the k-mers then overflow into
// the second level); KMAT_LINE_DENSITY=<k-mers per line> trades memory against second-level probes.
static int km_choose_line_bits(uint64_t n, int k, int m) {
    if (!m || getenv("KMAT_NO_LINE_LEVEL")) return 0;
    double dens = 2.0;
    if (const char *e = getenv("KMAT_LINE_DENSITY")) { const double v = atof(e); if (v >= 0.25 && v <= 16.0) dens = v; }
    int b = km_line_min_bits(k);
    while (b < 2 * m && (double)((uint64_t)1 << b) * dens < (double)n) b++;
    if (getenv("KMAT_TEST_TIGHT_TABLE")) b = std::max(km_line_min_bits(k), b - 3);       // tests: many overflowing sectors
    return km_line_geometry_ok(k, m, b) ? b : 0;
}

extern "C" uint32_t kmat_shard_of(uint64_t kmer, int kmer_length, int shard_count) {
    if (shard_count <= 1) return 0;
    const int m = getenv("KMAT_NO_LINE_LEVEL") ? 0 : km_line_m(kmer_length);
    if (m) return km_line_owner_of_g(km_mzr_mix2(km_mzr_of(kmer, kmer_length, m).c, m), m, (uint32_t)shard_count);
    return km_owner_of_x(km_mix(kmer, 2 * kmer_length), (uint32_t)shard_count);
}

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------
#define KM_STASH_CAP 65536u
// second level (or the whole table when there is no first level).  `sel` (optional): indices into kmers / payload of the
// k-mers to insert (those that overflowed their first-level sector), n = their number.
__global__ void km_insert_kernel(const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ payload, const uint32_t *__restrict__ sel, uint64_t n,
                                 unsigned long long *slots, uint64_t bucket_mask, int kmer_bits, int rem_bits,
                                 unsigned int *stash_n, uint64_t *stash_x, uint32_t *stash_hit,
                                 uint32_t shard_index, uint32_t shard_count, unsigned long long *n_kept) {
    unsigned long long kept = 0;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = sel ? sel[i0] : i0;
        const uint64_t x = km_mix(kmers[i], kmer_bits);
        if (!sel && shard_count > 1 && km_owner_of_x(x, shard_count) != shard_index) continue;     // another shard's k-mer
        kept++;
        const uint64_t home = x >> rem_bits, rem = x & ((1ull << rem_bits) - 1);
        const uint32_t pl = payload[i];
        const uint64_t base = (1ull << 63) | ((uint64_t)((pl >> 31) & 1) << 62) | (rem << 32) | (pl & 0x7FFFFFFFu);
        bool done = false;
        for (int d = 0; d <= KM_MAX_DISP && !done; d++) {
            unsigned long long *b = slots + ((home + d) & bucket_mask) * KM_SLOTS_PER_BUCKET;
            const unsigned long long v = base | ((uint64_t)d << 60);
            for (int s = 0; s < KM_SLOTS_PER_BUCKET && !done; s++) {
                if (b[s] == 0ull && atomicCAS(b + s, 0ull, v) == 0ull) done = true;
            }
        }
        if (!done) {                                   // KM_MAX_DISP + 1 full buckets: the key goes to the stash
            const unsigned int q = atomicAdd(stash_n, 1u);
            if (q < KM_STASH_CAP) { stash_x[q] = x; stash_hit[q] = pl; }
        }
    }
    kept = __reduce_add_sync(0xffffffffu, (unsigned)kept);      // < 2^32 per warp pass: grid-stride, 32 lanes
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_kept, kept);
}
// first level: every k-mer of this shard goes to its sector (kmat_mzr.h) if one of the four slots is free; the others are
// listed in `ovf` for the second level
__global__ void km_line_insert_kernel(const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ payload, uint64_t n,
                                      unsigned long long *lines, uint64_t line_first, int k, int m, int line_bits,
                                      uint32_t shard_index, uint32_t shard_count, unsigned long long *n_kept,
                                      uint32_t *ovf, unsigned long long *n_ovf, uint64_t ovf_cap) {
    unsigned long long kept = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = km_line_x(kmers[i], k, m, line_bits);
        if (shard_count > 1 && km_line_owner_of_g(km_line_g_of_x(x, k, m, line_bits), m, shard_count) != shard_index) continue;
        kept++;
        const uint32_t pl = payload[i];
        const unsigned long long v = (1ull << 63) | ((uint64_t)((pl >> 31) & 1) << 62) | ((x & ((1ull << KM_MZR_KEY_BITS) - 1)) << 32) | (pl & 0x7FFFFFFFu);
        unsigned long long *b = lines + (((x >> KM_LINE_XSHIFT) - line_first) * 4 + ((x >> KM_MZR_KEY_BITS) & 3u)) * KM_SLOTS_PER_BUCKET;
        bool done = false;
        for (int s = 0; s < KM_SLOTS_PER_BUCKET && !done; s++)
            if (b[s] == 0ull && atomicCAS(b + s, 0ull, v) == 0ull) done = true;
        if (!done) {
            const unsigned long long q = atomicAdd(n_ovf, 1ull);
            if (q < ovf_cap) ovf[q] = (uint32_t)i;
        }
    }
    kept = __reduce_add_sync(0xffffffffu, (unsigned)kept);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_kept, kept);
}
// after the insertion has finished: the sectors of the overflowed k-mers get their flag (slot 0 of a full sector is occupied)
__global__ void km_line_flag_kernel(const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ ovf, uint64_t n_ovf,
                                    unsigned long long *lines, uint64_t line_first, int k, int m, int line_bits) {
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n_ovf; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = km_line_x(kmers[ovf[q]], k, m, line_bits);
        unsigned long long *b = lines + (((x >> KM_LINE_XSHIFT) - line_first) * 4 + ((x >> KM_MZR_KEY_BITS) & 3u)) * KM_SLOTS_PER_BUCKET;
        if (!(b[0] & KM_LINE_OVF)) atomicOr(b, KM_LINE_OVF);
    }
}
__global__ void km_prefix_bits_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint32_t *bits, int shift) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = kmers[i] >> shift;
        atomicOr(bits + (p >> 5), 1u << (p & 31));
    }
}

// device temporaries of a table build: freed on every path out
struct KmBuildTmp {
    unsigned long long *d_cnt = nullptr;      // [0] k-mers kept, [1] first-level overflows
    uint32_t *d_ovf = nullptr;
    unsigned int *d_sn = nullptr; uint64_t *d_sx = nullptr; uint32_t *d_sh = nullptr;
    ~KmBuildTmp() { cudaFree(d_cnt); cudaFree(d_ovf); cudaFree(d_sn); cudaFree(d_sx); cudaFree(d_sh); }
};

// second level over the k-mers `sel` (or all n when sel == NULL, filtered by shard): sized for `expect` keys, doubled while
// the stash overflows
static int km_build_buckets(kmat_db *db, const uint64_t *d_kmers, const uint32_t *d_payload, const uint32_t *d_sel, uint64_t n, uint64_t expect,
                            int shard_index, int shard_count, KmBuildTmp &T, unsigned long long *kept_out) {
    const int kmer_bits = 2 * db->kmer_len;
    if (!T.d_sn) {
        KM_CUDA(cudaMalloc((void **)&T.d_sn, sizeof(unsigned int)));
        KM_CUDA(cudaMalloc((void **)&T.d_sx, (size_t)KM_STASH_CAP * 8)); KM_CUDA(cudaMalloc((void **)&T.d_sh, (size_t)KM_STASH_CAP * 4));
    }
    for (int b = km_choose_bucket_bits(expect, kmer_bits);; b++) {
        if (b > kmer_bits) { kmat_set_error("hash table build failed: displacement limit at maximum size"); return KMAT_ERR_UNSUPPORTED; }
        db->geom.bucket_bits = b; db->geom.rem_bits = kmer_bits - b;
        db->n_buckets = 1ull << b;
        const size_t bytes = db->n_buckets * KM_SLOTS_PER_BUCKET * sizeof(uint64_t);
        KM_CUDA(cudaMalloc((void **)&db->d_slots, bytes));
        KM_CUDA(cudaMemset(db->d_slots, 0, bytes));
        KM_CUDA(cudaMemset(T.d_sn, 0, sizeof(unsigned int)));
        KM_CUDA(cudaMemset(T.d_cnt, 0, 8));
        if (n) {
            const int threads = 256;
            const int blocks = (int)std::min<uint64_t>((n + threads - 1) / threads, 148ull * 16);
            km_insert_kernel<<<blocks, threads>>>(d_kmers, d_payload, d_sel, n, (unsigned long long *)db->d_slots, db->n_buckets - 1,
                                                  kmer_bits, db->geom.rem_bits, T.d_sn, T.d_sx, T.d_sh, (uint32_t)shard_index, (uint32_t)shard_count, T.d_cnt);
            g_km_launches++;
            KM_CUDA(cudaGetLastError());
        }
        unsigned int sn = 0;
        KM_CUDA(cudaMemcpy(&sn, T.d_sn, sizeof sn, cudaMemcpyDeviceToHost));
        KM_CUDA(cudaMemcpy(kept_out, T.d_cnt, 8, cudaMemcpyDeviceToHost));
        if (sn <= KM_STASH_CAP) {
            // the stash: sorted by mixed key on the host (a few hundred entries at most), searched by km_probe_buckets
            if (sn) {
                std::vector<uint64_t> sx(sn); std::vector<uint32_t> sh(sn), ord(sn);
                KM_CUDA(cudaMemcpy(sx.data(), T.d_sx, (size_t)sn * 8, cudaMemcpyDeviceToHost));
                KM_CUDA(cudaMemcpy(sh.data(), T.d_sh, (size_t)sn * 4, cudaMemcpyDeviceToHost));
                for (unsigned int i = 0; i < sn; i++) ord[i] = i;
                std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b2) { return sx[a] < sx[b2]; });
                std::vector<uint64_t> sx2(sn); std::vector<uint32_t> sh2(sn);
                for (unsigned int i = 0; i < sn; i++) { sx2[i] = sx[ord[i]]; sh2[i] = sh[ord[i]]; }
                KM_CUDA(cudaMalloc((void **)&db->d_stash_x, (size_t)sn * 8)); KM_CUDA(cudaMalloc((void **)&db->d_stash_hit, (size_t)sn * 4));
                KM_CUDA(cudaMemcpy(db->d_stash_x, sx2.data(), (size_t)sn * 8, cudaMemcpyHostToDevice));
                KM_CUDA(cudaMemcpy(db->d_stash_hit, sh2.data(), (size_t)sn * 4, cudaMemcpyHostToDevice));
            }
            db->n_stash = sn;
            return KMAT_OK;
        }
        cudaFree(db->d_slots); db->d_slots = nullptr;       // too many displaced keys for the stash: double the table
    }
}

// n k-mers in the arrays; n_global = size of the whole table (decides the first level's geometry, which every shard must
// share); prefiltered: the arrays hold this shard's k-mers only (the host upload path), else the kernels filter by owner.
static int km_db_alloc_and_insert(kmat_db *db, const uint64_t *d_kmers, const uint32_t *d_payload, uint64_t n, uint64_t n_global, bool prefiltered,
                                  int shard_index, int shard_count) {
    const int kmer_bits = 2 * db->kmer_len;
    KmBuildTmp T;
    KM_CUDA(cudaMalloc((void **)&T.d_cnt, 16));
    KM_CUDA(cudaMemset(T.d_cnt, 0, 16));
    unsigned long long kept = 0;
    db->geom.kmer_bits = kmer_bits;
    db->geom.line_m = km_line_m(db->kmer_len);
    db->geom.line_bits = km_choose_line_bits(n_global, db->kmer_len, db->geom.line_m);
    if (!db->geom.line_bits || n >= (1ull << 32)) { db->geom.line_m = 0; db->geom.line_bits = 0; }      // the overflow list holds 32-bit indices
    if (db->geom.line_m) {
        const int m = db->geom.line_m, lb = db->geom.line_bits;
        db->line_first = km_line_shard_first((uint32_t)shard_index, (uint32_t)shard_count, m, lb);
        db->n_lines = km_line_shard_count((uint32_t)shard_index, (uint32_t)shard_count, m, lb);
        const size_t bytes = db->n_lines * 128;
        KM_CUDA(cudaMalloc((void **)&db->d_lines, bytes));
        KM_CUDA(cudaMemset(db->d_lines, 0, bytes));
        // a shard keeps ~n / shard_count of the k-mers; the overflow list is sized for all of them
        const uint64_t ovf_cap = prefiltered ? n + 16 : n / shard_count + n / (4 * (uint64_t)shard_count) + 4096;
        KM_CUDA(cudaMalloc((void **)&T.d_ovf, ovf_cap * 4));
        unsigned long long n_ovf = 0;
        if (n) {
            const int blocks = (int)std::min<uint64_t>((n + 255) / 256, 148ull * 16);
            km_line_insert_kernel<<<blocks, 256>>>(d_kmers, d_payload, n, (unsigned long long *)db->d_lines, db->line_first, db->kmer_len, m, lb,
                                                   (uint32_t)shard_index, (uint32_t)shard_count, T.d_cnt, T.d_ovf, T.d_cnt + 1, ovf_cap);
            g_km_launches++;
            KM_CUDA(cudaGetLastError());
            unsigned long long h[2];
            KM_CUDA(cudaMemcpy(h, T.d_cnt, 16, cudaMemcpyDeviceToHost));
            kept = h[0]; n_ovf = h[1];
            if (n_ovf > ovf_cap) { kmat_set_error("table build: %llu first-level overflows exceed the list of %llu (uneven shards?)", n_ovf, (unsigned long long)ovf_cap); return KMAT_ERR_UNSUPPORTED; }
            if (n_ovf) {
                km_line_flag_kernel<<<(int)std::min<uint64_t>((n_ovf + 255) / 256, 148ull * 16), 256>>>(d_kmers, T.d_ovf, n_ovf, (unsigned long long *)db->d_lines, db->line_first, db->kmer_len, m, lb);
                g_km_launches++;
                KM_CUDA(cudaGetLastError());
            }
        }
        db->n_overflow = n_ovf;
        if (n_ovf) {
            unsigned long long k2 = 0;
            int rc = km_build_buckets(db, d_kmers, d_payload, T.d_ovf, n_ovf, n_ovf, 0, 1, T, &k2);
            if (rc != KMAT_OK) return rc;
        }
    } else {
        // a shard keeps ~n / shard_count of the k-mers (the owner hash is uniform); the table is sized for that plus 5 %
        const uint64_t expect = (shard_count > 1 && !prefiltered) ? n / shard_count + n / (20 * (uint64_t)shard_count) + 1024 : n;
        int rc = km_build_buckets(db, d_kmers, d_payload, nullptr, n, expect, shard_index, prefiltered ? 1 : shard_count, T, &kept);
        if (rc != KMAT_OK) return rc;
    }
    // bitmap of the reference's non-empty top-tier prefixes: only used to count "prefix miss" lookups exactly
    // for the algorithmic-bytes statistic; 2^27 bits = 16 MiB
    db->prefix_shift = db->kmer_len == 18 ? 9 : 13;
    const uint64_t nprefix = (kmer_bits - db->prefix_shift) >= 40 ? 0 : (1ull << (kmer_bits - db->prefix_shift));
    if (nprefix && nprefix <= (1ull << 32)) {
        const size_t words = (size_t)((nprefix + 31) / 32);
        KM_CUDA(cudaMalloc((void **)&db->d_prefix_bits, words * 4));
        KM_CUDA(cudaMemset(db->d_prefix_bits, 0, words * 4));
        if (n) {
            km_prefix_bits_kernel<<<(int)std::min<uint64_t>((n + 255) / 256, 148ull * 16), 256>>>(d_kmers, n, db->d_prefix_bits, db->prefix_shift);
            g_km_launches++;
            KM_CUDA(cudaGetLastError());
        }
        db->prefix_bytes = words * 4;
    }
    KM_CUDA(cudaDeviceSynchronize());
    db->n_kmers = kept;
    db->shard_index = shard_index; db->shard_count = shard_count;
    return KMAT_OK;
}

KmDbDev km_db_dev(const kmat_db *db) {
    KmDbDev d;
    d.peers = nullptr; d.n_peers = 0;
    d.lines = db->d_lines; d.line_first = db->line_first; d.n_lines = db->n_lines; d.line_m = db->geom.line_m; d.line_bits = db->geom.line_bits;
    d.slots = db->d_slots; d.bucket_mask = db->n_buckets - 1; d.kmer_bits = db->geom.kmer_bits; d.rem_bits = db->geom.rem_bits;
    d.kmer_len = db->kmer_len; d.tid_bytes = db->tid_bytes; d.pool = db->d_pool; d.prefix_bits = db->d_prefix_bits;
    d.prefix_shift = db->prefix_shift;
    d.stash_x = db->d_stash_x; d.stash_hit = db->d_stash_hit; d.n_stash = db->n_stash;
    return d;
}

static int km_db_build(int device, int kmer_len, int tid_bytes, uint64_t n, uint64_t n_global, bool prefiltered, const uint64_t *d_kmers,
                       const uint32_t *d_payload, const uint32_t *d_pool, uint64_t pool_words, uint32_t n_stored_ids,
                       int shard_index, int shard_count, kmat_db **out) {
    kmat_db *db = new kmat_db();
    db->device = device; db->kmer_len = kmer_len; db->tid_bytes = tid_bytes; db->n_sid = tid_bytes == 2 ? 65536u : n_stored_ids;
    db->pool_words = pool_words;
    auto body = [&]() -> int {
        if (pool_words) {
            KM_CUDA(cudaMalloc((void **)&db->d_pool, pool_words * 4));
            KM_CUDA(cudaMemcpy(db->d_pool, d_pool, pool_words * 4, cudaMemcpyDeviceToDevice));
        }
        return km_db_alloc_and_insert(db, d_kmers, d_payload, n, n_global, prefiltered, shard_index, shard_count);
    };
    const int rc = body();
    if (rc != KMAT_OK) { kmat_db_free(db); return rc; }      // every device buffer of a failed build goes back
    *out = db;
    return KMAT_OK;
}

extern "C" int kmat_db_build_device(int device, int kmer_len, int tid_bytes, uint64_t n, const uint64_t *d_kmers,
                                    const uint32_t *d_payload, const uint32_t *d_pool, uint64_t pool_words, uint32_t n_stored_ids,
                                    int shard_index, int shard_count, kmat_db **out) {
    if (!out || (tid_bytes != 2 && tid_bytes != 4) || kmer_len < 8 || kmer_len > 28 || shard_count < 1 || shard_count > KM_MAX_SHARDS || shard_index < 0 || shard_index >= shard_count) { kmat_set_error("kmat_db_build_device: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= device) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    if (pool_words >= (1ull << 31)) { kmat_set_error("list pool of %llu words exceeds the 31-bit offset range", (unsigned long long)pool_words); return KMAT_ERR_UNSUPPORTED; }
    KM_CUDA(cudaSetDevice(device));
    const int rc = km_db_build(device, kmer_len, tid_bytes, n, n, false, d_kmers, d_payload, d_pool, pool_words, n_stored_ids, shard_index, shard_count, out);
    if (rc == KMAT_OK && shard_count > 1) (*out)->pool_shared = true;        // the payloads of every shard index this one pool
    return rc;
}

// device copies of the host arrays of an upload: freed on every path out
struct KmUploadTmp {
    uint64_t *d_k = nullptr; uint32_t *d_p = nullptr, *d_pool = nullptr;
    ~KmUploadTmp() { cudaFree(d_k); cudaFree(d_p); cudaFree(d_pool); }
};

// Host table -> device.  List pool record: 16-bit ids: [u16 count][u16 id]*count ; 32-bit ids: [u32 count][u32 id]*count,
// padded to 4 bytes; a record of <= 32 bytes never straddles a 32-byte sector, so a list fetch is one sector.
extern "C" int kmat_db_upload(const kmat_table *t, int device, int shard_index, int shard_count, kmat_db **out) {
    if (!t || !out || shard_count < 1 || shard_count > KM_MAX_SHARDS || shard_index < 0 || shard_index >= shard_count) { kmat_set_error("kmat_db_upload: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= device) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    if (t->kmer_len < 8 || t->kmer_len > 28) { kmat_set_error("k-mer length %d unsupported", t->kmer_len); return KMAT_ERR_UNSUPPORTED; }
    std::vector<uint64_t> kmers; std::vector<uint32_t> payload, pool;
    std::vector<uint32_t> stored;      // 32-bit DBs: distinct stored tids, ascending; lists hold indices into it
    if (t->tid_bytes == 4) {
        stored.assign(t->ids, t->ids + t->n_ids);
        std::sort(stored.begin(), stored.end());
        stored.erase(std::unique(stored.begin(), stored.end()), stored.end());
    }
    auto sid_of = [&](uint32_t id) -> uint32_t {
        if (t->tid_bytes == 2) return id & 0xFFFF;
        return (uint32_t)(std::lower_bound(stored.begin(), stored.end(), id) - stored.begin());
    };
    kmers.reserve(t->n_kmers / shard_count + 16); payload.reserve(t->n_kmers / shard_count + 16);
    for (uint64_t i = 0; i < t->n_kmers; i++) {
        if (i && t->kmers[i] <= t->kmers[i - 1]) { kmat_set_error("table: k-mers are not strictly ascending at index %llu", (unsigned long long)i); return KMAT_ERR_FORMAT; }
        if (shard_count > 1 && (int)kmat_shard_of(t->kmers[i], t->kmer_len, shard_count) != shard_index) continue;
        const uint64_t a = t->offs[i], c = t->offs[i + 1] - a;
        if (t->offs[i + 1] < a || t->offs[i + 1] > t->n_ids) { kmat_set_error("table: list offsets of k-mer %llu are not ascending / exceed the id array", (unsigned long long)i); return KMAT_ERR_FORMAT; }
        if (c == 0) continue;                         // cannot occur in a SortedDb (every record has >= 1 tid)
        if (c >= 32768) { kmat_set_error("taxid list of %llu entries: counts >= 32768 turn label_vec[pos].first negative in the reference (int16_t, read_label.cpp:49); unsupported", (unsigned long long)c); return KMAT_ERR_UNSUPPORTED; }
        kmers.push_back(t->kmers[i]);
        if (c == 1) { payload.push_back(sid_of(t->ids[a])); continue; }
        size_t words = t->tid_bytes == 2 ? (2 + 2 * c + 3) / 4 : 1 + c;
        size_t at = pool.size();
        if (words <= 8 && (at % 8) + words > 8) at = (at + 7) & ~(size_t)7;      // keep short records inside one sector
        if (at + words >= (1ull << 31)) { kmat_set_error("list pool exceeds the 31-bit offset range"); return KMAT_ERR_UNSUPPORTED; }
        pool.resize(at + words, 0);
        if (t->tid_bytes == 2) {
            uint16_t *p = (uint16_t *)(pool.data() + at);
            p[0] = (uint16_t)c;
            for (uint64_t j = 0; j < c; j++) p[1 + j] = (uint16_t)t->ids[a + j];
        } else {
            pool[at] = (uint32_t)c;
            for (uint64_t j = 0; j < c; j++) pool[at + 1 + j] = sid_of(t->ids[a + j]);
        }
        payload.push_back(0x80000000u | (uint32_t)at);
    }
    KM_CUDA(cudaSetDevice(device));
    KmUploadTmp U;
    const uint64_t n = kmers.size();
    if (n) {
        KM_CUDA(cudaMalloc((void **)&U.d_k, n * 8)); KM_CUDA(cudaMalloc((void **)&U.d_p, n * 4));
        KM_CUDA(cudaMemcpy(U.d_k, kmers.data(), n * 8, cudaMemcpyHostToDevice));
        KM_CUDA(cudaMemcpy(U.d_p, payload.data(), n * 4, cudaMemcpyHostToDevice));
    }
    if (!pool.empty()) {
        KM_CUDA(cudaMalloc((void **)&U.d_pool, pool.size() * 4));
        KM_CUDA(cudaMemcpy(U.d_pool, pool.data(), pool.size() * 4, cudaMemcpyHostToDevice));
    }
    // the host loop above already kept this shard's k-mers only (and only their lists); the first level's geometry follows
    // the size of the WHOLE table so that every shard computes the same keys
    int rc = km_db_build(device, t->kmer_len, t->tid_bytes, n, t->n_kmers, true, U.d_k, U.d_p, U.d_pool, pool.size(), (uint32_t)stored.size(), shard_index, shard_count, out);
    if (rc == KMAT_OK && !stored.empty()) {
        if (cudaMalloc((void **)&(*out)->d_stored_tids, stored.size() * 4) != cudaSuccess ||
            cudaMemcpy((*out)->d_stored_tids, stored.data(), stored.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); kmat_db_free(*out); *out = nullptr; kmat_set_error("kmat_db_upload: out of device memory"); return KMAT_ERR_NOMEM;
        }
    }
    if (rc == KMAT_OK) (*out)->stored_tids = stored;
    return rc;
}

// HBM one device needs to hold (one shard of) this table and to build it: both levels from the real geometry, the list pools
// (stored + resolved), the upload temporaries that are live next to them, and room for the per-batch buffers.
extern "C" uint64_t kmat_table_device_bytes(const kmat_table *t, int shard_count) {
    if (!t || shard_count < 1) return 0;
    const uint64_t n = t->n_kmers, n_s = n / (uint64_t)shard_count + 1;
    const int m = km_line_m(t->kmer_len), lb = (n < (1ull << 32)) ? km_choose_line_bits(n, t->kmer_len, m) : 0;
    uint64_t bytes = 0;
    if (lb) {
        bytes += (((uint64_t)1 << lb) / (uint64_t)shard_count + 2) * 128;                         // first level
        bytes += ((uint64_t)1 << km_choose_bucket_bits(n_s / 12 + 16, 2 * t->kmer_len)) * 32;      // second level: ~5 % of the k-mers, doubled once
        bytes += n_s * 4;                                                                           // overflow list during the build
    } else bytes += ((uint64_t)1 << km_choose_bucket_bits(n_s + n_s / 20, 2 * t->kmer_len)) * 32;
    const uint64_t list_ids = t->n_ids > n ? t->n_ids - n : 0;                                      // ids beyond one per k-mer sit in lists
    const uint64_t pool = (list_ids * 2 * (uint64_t)t->tid_bytes + list_ids) / (uint64_t)shard_count;  // records incl. counts and padding (upper bound)
    bytes += pool * (t->tid_bytes == 2 ? 4 : 3);                                                    // stored pool + its upload copy + the resolved pool
    bytes += n_s * 12 + (16ull << 20);                                                              // k-mers + payloads of the upload, prefix bitmap
    bytes += 3ull << 30;                                                                            // batch buffers of a 131072-read batch pipeline, candidates, scratch
    return bytes;
}

extern "C" uint64_t kmat_db_size(const kmat_db *db) { return db ? db->n_kmers : 0; }
extern "C" uint64_t kmat_db_bytes(const kmat_db *db) { return db ? db->n_lines * 128 + (db->d_slots ? db->n_buckets * 32 : 0) + db->pool_words * 4 + db->prefix_bytes + (uint64_t)db->n_stash * 12 : 0; }
extern "C" uint64_t kmat_db_overflow(const kmat_db *db) { return db ? db->n_overflow : 0; }      /* k-mers in the second level of a two-level table */
extern "C" int kmat_db_kmer_length(const kmat_db *db) { return db ? db->kmer_len : 0; }
extern "C" int kmat_db_device(const kmat_db *db) { return db ? db->device : -1; }
extern "C" void kmat_db_free(kmat_db *db) {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaFree(db->d_lines); cudaFree(db->d_slots); cudaFree(db->d_pool); cudaFree(db->d_prefix_bits); cudaFree(db->d_stash_x); cudaFree(db->d_stash_hit); cudaFree(db->d_stored_tids);
    delete db;
}

// ---------------------------------------------------------------------------------------------
// K2 parity hook: thread per k-mer
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t km_list_count(const KmDbDev &db, uint32_t off) {
    return db.tid_bytes == 2 ? (uint32_t)(*(const uint16_t *)(db.pool + off)) : db.pool[off];
}
__device__ __forceinline__ uint32_t km_list_id(const KmDbDev &db, uint32_t off, uint32_t j) {
    return db.tid_bytes == 2 ? (uint32_t)((const uint16_t *)(db.pool + off))[1 + j] : db.pool[off + 1 + j];
}
__global__ void km_lookup_count_kernel(KmDbDev db, const uint64_t *__restrict__ kmers, uint32_t n, uint32_t *hit, uint64_t *cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t extra;
    const uint32_t h = km_probe(db, kmers[i], extra);
    hit[i] = h;
    cnt[i] = h == KM_HIT_MISS ? 0 : (h & KM_HIT_LIST) ? km_list_count(db, h & 0x7FFFFFFFu) : 1;
}
__global__ void km_lookup_fill_kernel(KmDbDev db, const uint32_t *__restrict__ hit, const uint64_t *__restrict__ off, uint32_t n,
                                      uint32_t *ids, uint64_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = hit[i];
    if (h == KM_HIT_MISS) return;
    const uint64_t o = off[i];
    if (h & KM_HIT_LIST) {
        const uint32_t lo = h & 0x7FFFFFFFu, c = km_list_count(db, lo);
        for (uint32_t j = 0; j < c; j++) if (o + j < cap) ids[o + j] = km_list_id(db, lo, j);
    } else if (o < cap) ids[o] = h;
}

extern "C" int kmat_lookup_batch(const kmat_db *db, const uint64_t *kmers, uint32_t n, uint64_t *hit_off, uint32_t *ids,
                                 uint64_t ids_cap, uint64_t *n_ids) {
    if (!db || (n && (!kmers || !hit_off))) { kmat_set_error("kmat_lookup_batch: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(db->device));
    hit_off[0] = 0;
    if (n_ids) *n_ids = 0;
    if (!n) return KMAT_OK;
    uint64_t *d_k, *d_cnt, *d_off; uint32_t *d_hit, *d_ids = nullptr;
    KM_CUDA(cudaMalloc((void **)&d_k, (size_t)n * 8)); KM_CUDA(cudaMalloc((void **)&d_cnt, (size_t)(n + 1) * 8));
    KM_CUDA(cudaMalloc((void **)&d_off, (size_t)(n + 1) * 8)); KM_CUDA(cudaMalloc((void **)&d_hit, (size_t)n * 4));
    KM_CUDA(cudaMemcpy(d_k, kmers, (size_t)n * 8, cudaMemcpyHostToDevice));
    KM_CUDA(cudaMemset(d_cnt, 0, (size_t)(n + 1) * 8));
    KmDbDev dd = km_db_dev(db);
    km_lookup_count_kernel<<<(n + 255) / 256, 256>>>(dd, d_k, n, d_hit, d_cnt);
    g_km_launches++;
    void *tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_off, n + 1);
    KM_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_cnt, d_off, n + 1);
    g_km_launches++;
    KM_CUDA(cudaMemcpy(hit_off, d_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
    const uint64_t total = hit_off[n];
    if (n_ids) *n_ids = total;
    int rc = KMAT_OK;
    if (total > ids_cap || (total && !ids)) rc = KMAT_ERR_OVERFLOW;
    else if (total) {
        KM_CUDA(cudaMalloc((void **)&d_ids, total * 4));
        km_lookup_fill_kernel<<<(n + 255) / 256, 256>>>(dd, d_hit, d_off, n, d_ids, total);
        g_km_launches++;
        KM_CUDA(cudaMemcpy(ids, d_ids, total * 4, cudaMemcpyDeviceToHost));
        if (db->tid_bytes == 4) for (uint64_t i = 0; i < total; i++) ids[i] = db->stored_tids[ids[i]];   // dense id -> stored tid
    }
    cudaFree(d_k); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_hit); cudaFree(d_ids); cudaFree(tmp);
    KM_CUDA(cudaGetLastError());
    return rc;
}

// ---------------------------------------------------------------------------------------------
// K1 + K2: encode, dedup, probe.  One warp per read, streaming over 32-base chunks.
//
// Restates the rolling encoder of retrieve_kmer_labels (read_label.cpp:943-950, 978-1017): 2-bit MSB-first
// packing, canonical = min(forward, reverse complement), any non-ACGT byte resets the run, the first
// occurrence of a canonical k-mer in the read wins (no_dups, :1010,1017), and the GC bookkeeping of
// :994-1008 / :1205-1206 (bases of every run of >= k valid bases are counted once).
// ---------------------------------------------------------------------------------------------
#define KM_DEDUP_SLOTS 512          // per-warp shared-memory set; reads with more k-mer positions use a global one
#define KM_PROBE_WARPS 8

struct KmProbeParams {
    KmDbDev db;
    const char *bases; const uint64_t *offs; uint32_t n_reads;
    uint32_t *hit;                 // per base offset (k-mer start position p of read r -> hit[offs[r] + p])
    int2 *hdr;                     // per read: {valid_kmers, bin_sel}
    uint64_t *out_kmers; uint8_t *out_flags;    // optional (encode parity hook)
    unsigned long long *long_sets; uint32_t long_slots;   // global dedup sets for long reads: one per warp in the grid
    KmStatsDev *stats;             // optional
    int do_probe;
    uint64_t *xq;                  // DB-sharded mode (with do_probe == 0): mixed k-mer of every first occurrence, per base offset
    int skip_mid;                  // 1: reads of KM_LONG_MIN..KM_LONG_MAX bases are left to km_encode_probe_long_kernel
};
// Reads of this many bases get a whole CTA (km_encode_probe_long_kernel): their dedup set fits shared memory
#define KM_LONG_MIN 257
#define KM_LONG_MAX 12000
#define KM_LONG_SLOTS 16384
#define KM_LONG_THREADS 512
#define KM_LONG_POS_BITS 24

__device__ __forceinline__ int km_code(unsigned char ch) {
    // ENCODE macro, read_label.cpp:943-950: a/A 0, c/C 1, g/G 2, t/T 3, anything else resets (-1).  Branch-free:
    // letters live in 0x40..0x7F and (ch & 31) is 1, 3, 7, 20 for A, C, G, T in either case; bits 2:1 of the ASCII
    // code are 00, 01, 11, 10 for A, C, G, T, which x ^ (x >> 1) turns into 0, 1, 2, 3.
    const uint32_t c = ch;
    const bool ok = (c & 0xC0u) == 0x40u && ((0x0010008Au >> (c & 31u)) & 1u);
    const uint32_t x = (c >> 1) & 3u;
    return ok ? (int)(x ^ (x >> 1)) : -1;
}
__device__ __forceinline__ uint64_t kb_shfl_u64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(KM_FULL, (uint32_t)v, src), hi = __shfl_sync(KM_FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t km_revcomp(uint64_t fwd, int kmer_bits) {
    uint64_t x = ~fwd << (64 - kmer_bits);            // complement; k-mer now left-aligned
    x = __brevll(x);                                   // reverses base order and the two bits inside each base
    return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
}

__global__ void __launch_bounds__(KM_PROBE_WARPS * 32) km_encode_probe_kernel(KmProbeParams P) {
    __shared__ unsigned long long s_set[KM_PROBE_WARPS][KM_DEDUP_SLOTS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp_global = blockIdx.x * KM_PROBE_WARPS + wib, n_warps = gridDim.x * KM_PROBE_WARPS;
    const int k = P.db.kmer_len, kmer_bits = P.db.kmer_bits;
    const uint64_t kmask = (1ull << kmer_bits) - 1;
    const int ebits = 64 - kmer_bits;                   // epoch tag above the k-mer: stale entries read as empty
    const uint32_t emax = (ebits >= 31 ? 0x7FFFFFFFu : ((1u << ebits) - 1));
    unsigned long long *sset = s_set[wib];
    for (int i = lane; i < KM_DEDUP_SLOTS; i += 32) sset[i] = 0;
    unsigned long long *lset = P.long_sets ? P.long_sets + (size_t)warp_global * P.long_slots : nullptr;
    uint32_t epoch_s = 0, epoch_l = emax;               // the long set is cleared on first use
    unsigned long long st_lookups = 0, st_hits = 0, st_lists = 0, st_extra = 0, st_pmiss = 0;
    __syncwarp();

    for (uint32_t r = warp_global; r < P.n_reads; r += n_warps) {
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        const int np = len - k + 1;
        if (P.skip_mid && len >= KM_LONG_MIN && len <= KM_LONG_MAX) continue;
        // pick the dedup set
        unsigned long long *set; uint32_t smask; uint32_t epoch;
        if (np <= KM_DEDUP_SLOTS / 2 || !lset) {
            if (++epoch_s > emax) { for (int i = lane; i < KM_DEDUP_SLOTS; i += 32) sset[i] = 0; epoch_s = 1; __syncwarp(); }
            set = sset; smask = KM_DEDUP_SLOTS - 1; epoch = epoch_s;
        } else {
            if (++epoch_l > emax) { for (uint32_t i = lane; i < P.long_slots; i += 32) lset[i] = 0; epoch_l = 1; __syncwarp(); }
            set = lset; smask = P.long_slots - 1; epoch = epoch_l;
        }
        const unsigned long long etag = ebits >= 64 ? 0ull : ((unsigned long long)epoch << kmer_bits);
        uint64_t prev = 0;             // packed previous 32 bases
        uint32_t pinv = 0xFFFFFFFFu;   // invalid-base mask of the previous chunk (before the read: all invalid)
        uint32_t pgc = 0;
        int valid = 0, vgc = 0, vtot = 0;
        const int nchunks = (len + 31) >> 5;
        for (int c = 0; c < nchunks; c++) {
            const int j = (c << 5) + lane;                                   // base index handled by this lane
            const int code = j < len ? km_code((unsigned char)P.bases[off + j]) : -1;
            const uint32_t cinv = __ballot_sync(KM_FULL, code < 0);
            const uint32_t cgc = __ballot_sync(KM_FULL, code == 1 || code == 2);
            const uint32_t cc = code < 0 ? 0u : (uint32_t)code;
            const uint32_t hi = __reduce_or_sync(KM_FULL, lane < 16 ? cc << (30 - 2 * lane) : 0u);
            const uint32_t lo = __reduce_or_sync(KM_FULL, lane >= 16 ? cc << (62 - 2 * lane) : 0u);
            const uint64_t cur = ((uint64_t)hi << 32) | lo;
            // k-mer ENDING at base j (start position p = j - k + 1)
            const int s = 62 - 2 * lane;
            const uint64_t fwd = ((cur >> s) | (s ? (prev << (64 - s)) : 0ull)) & kmask;
            const uint64_t inv64 = ((uint64_t)cinv << 32) | pinv, gc64 = ((uint64_t)cgc << 32) | pgc;
            const uint64_t wmask = (1ull << k) - 1;
            const int wsh = 32 + lane - k + 1;                               // >= 0 because k <= 32
            const bool ok = ((inv64 >> wsh) & wmask) == 0;                   // no invalid base in [p, j]
            const bool ok_prev = wsh > 0 && ((inv64 >> (wsh - 1)) & wmask) == 0;   // the window ending at j-1
            const int p = j - k + 1;
            uint64_t canon = 0; bool first = false;
            if (ok) {
                const uint64_t rc = km_revcomp(fwd, kmer_bits);
                canon = fwd < rc ? fwd : rc;                                   // read_label.cpp:1009
            }
            // GC bookkeeping (:994-1008): a run's first k-mer adds its k bases, each further k-mer adds one
            const int add_tot = ok ? (ok_prev ? 1 : k) : 0;
            const int add_gc = ok ? (ok_prev ? (int)((cgc >> lane) & 1) : __popcll((gc64 >> wsh) & wmask)) : 0;
            valid += ok; vtot += add_tot; vgc += add_gc;
            // dedup: lower position wins.  Chunks are visited in order; inside a chunk the lowest lane of each
            // equal-k-mer group is the leader and inserts.
            const uint32_t okmask = __ballot_sync(KM_FULL, ok);
            if (ok) {
                const uint32_t grp = __match_any_sync(okmask, canon);
                const int leader = __ffs(grp) - 1;
                bool fresh = false;
                if (lane == leader) {
                    const unsigned long long key = etag | canon;
                    uint32_t h = (uint32_t)((canon * 0x9E3779B97F4A7C15ull) >> 40) & smask;
                    for (;;) {
                        const unsigned long long cur_e = ((volatile unsigned long long *)set)[h];   // L1-bypassing: lanes CAS the same set
                        if (cur_e == key) break;                                               // seen earlier in this read
                        const bool stale = ebits < 64 ? ((cur_e >> kmer_bits) != epoch) : (cur_e == 0);
                        if (stale) {
                            const unsigned long long old = atomicCAS(set + h, cur_e, key);
                            if (old == cur_e) { fresh = true; break; }
                            if (old == key) break;
                            continue;                                                          // someone else took it: re-read
                        }
                        h = (h + 1) & smask;
                    }
                }
                fresh = __shfl_sync(grp, fresh, leader);
                first = fresh && lane == leader;
            }
            uint32_t hw = KM_HIT_INVALID;
            if (first) {
                hw = KM_HIT_MISS;
                if (P.do_probe) {
                    uint32_t extra;
                    hw = km_probe(P.db, canon, extra);
                    if (P.stats) {
                        st_lookups++; st_extra += extra;
                        if (hw != KM_HIT_MISS) { st_hits++; if (hw & KM_HIT_LIST) st_lists++; }
                        else if (P.db.prefix_bits) {
                            const uint64_t pf = canon >> P.db.prefix_shift;
                            if (!((P.db.prefix_bits[pf >> 5] >> (pf & 31)) & 1)) st_pmiss++;
                        }
                    }
                }
            }
            if (p >= 0 && j < len) {
                P.hit[off + p] = hw;
                if (P.xq && first) P.xq[off + p] = km_key(P.db, canon);
                if (P.out_kmers) { P.out_kmers[off + p] = ok ? canon : 0; P.out_flags[off + p] = ok ? (first ? 1 : 2) : 0; }
            }
            prev = cur; pinv = cinv; pgc = cgc;
        }
        valid = km_warp_sum(valid); vgc = km_warp_sum(vgc); vtot = km_warp_sum(vtot);
        if (lane == 0) {
            // gc_pcnt = ((float)vgc / (float)vtot) * 100.0 [double] -> float; bin = gc_pcnt / 10   (:1205-1206)
            const float frac = __fdiv_rn((float)vgc, (float)vtot);
            const float gc_pcnt = __double2float_rn(__dmul_rn((double)frac, 100.0));
            const float q = __fdiv_rn(gc_pcnt, 10.0f);
            P.hdr[r] = make_int2(valid, vtot > 0 ? (int)q : 0);
        }
    }
    if (P.stats) {
        st_lookups = km_warp_sum((int)st_lookups); st_hits = km_warp_sum((int)st_hits); st_lists = km_warp_sum((int)st_lists);
        st_extra = km_warp_sum((int)st_extra); st_pmiss = km_warp_sum((int)st_pmiss);
        if (lane == 0) {
            atomicAdd(&P.stats->lookups, st_lookups); atomicAdd(&P.stats->hits, st_hits); atomicAdd(&P.stats->list_hits, st_lists);
            atomicAdd(&P.stats->extra_buckets, st_extra); atomicAdd(&P.stats->prefix_miss, st_pmiss);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 + K2 for short reads (the Illumina case): variant of the kernel above for batches whose reads have at most
// 32 * NCH bases (NCH <= 8) and k <= 24.  Same results; what changes is how many table requests are in flight.
// Measured on B200 (profiles/r01_probe_kernel_notes.md): uniform random table reads are limited to ~37 G requests/s
// chip-wide whatever their width (and a second request to the same sector costs as much as the first), so the
// kernel must (a) issue exactly one 32-byte request per lookup and (b) keep enough of them in flight.
//   * the bases of the warp's NEXT read (and the offsets of the one after) are loaded while the current one is
//     processed, so no iteration starts with an exposed global load;
//   * all NCH chunks are encoded and deduplicated first; then every lane issues the home-bucket gather (one
//     LDG.E.256) of ALL its first-occurrence k-mers back to back and only then looks at the first one: NCH
//     independent gathers per lane, 32 * NCH per warp, instead of one per chunk.  That costs 8 registers per
//     gather (2 CTAs of 8 warps per SM at ~100 registers) and is what lifts the kernel from 25 to 34.5 G lookups/s;
//   * dedup without a hash set: duplicates inside a read are rare, so every valid k-mer sets one bit of a per-warp
//     bitmap (atomicOr on shared memory, SETN bits hashed from the mixed k-mer).  Whoever finds its bit clear is the
//     first k-mer with that hash; the few that find it set (a real duplicate or a hash collision) are "suspects"
//     and are settled exactly, one at a time and warp-wide: the suspect's k-mer is broadcast, every lane compares
//     it with its own NCH k-mers, an equal k-mer at a lower position makes the suspect a duplicate, and equal
//     k-mers at higher positions are marked duplicates themselves.  The outcome is "the lowest position of every
//     distinct k-mer survives" (read_label.cpp:1010-1017) whatever order the atomics were served in.  Keys are the
//     mixed k-mers (km_mix is a bijection), which the probe needs anyway;
//   * the home bucket answers ~98.5 % of the lookups; a full home bucket continues with km_probe_x(d = 1).
// Dynamic shared memory per warp: SETN / 8 B dedup bitmap (16 Ki bits for 150 bp reads: ~0.5 suspects per read; the warp-wide
// settlement of a suspect costs ~50 instructions, 2 per read at 4 Ki bits were 5 % of the kernel).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// K1 + K2 for long reads (257 .. 12000 bases, k <= 20): one CTA per read.  The any-length kernel above keeps the dedup set
// of such a read in global memory, which costs two more random DRAM requests per k-mer than the table gather itself
// (the path is request-rate bound: 70 ms per 10^9 bases).  Here the set lives in shared memory (16 K slots of
// [canonical k-mer : position], 128 KB): phase 1 inserts every k-mer window with "lowest position wins" (atomicMin on the
// packed word), phase 2 re-encodes, asks the set whether a window is the first occurrence and gathers the buckets of the
// first occurrences, four chunks in flight per warp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t km_fetch_left(uint32_t cur, uint32_t prv, int lane, int sh);     // below, with the short-read kernel
// KEY: also the table key of the lane's k-mer when the table has a first level (the stateless sliding minimum: this chunk's
// level-0 keys and those of the previous chunk's tail, w - 1 linear steps; host model in tests/mzr_check.cpp, 3d) -- the k-mer
// need not be valid, every lane takes part in the shuffles
template <bool KEY>
__device__ __forceinline__ void kl_encode_chunk(const KmProbeParams &P, uint64_t off, int len, int c, int lane, int k, uint64_t kmask, int kmer_bits,
                                                uint64_t &canon, bool &ok, bool &ok_prev, uint32_t &cgc_bit, int &gc_win, uint64_t *key = nullptr) {
    // the 32 bases of chunk c and of the chunk before it (re-read: they are in L1/L2), packed as in the any-length kernel
    const int j = (c << 5) + lane;
    const int code = j < len ? km_code((unsigned char)P.bases[off + j]) : -1;
    const int jp = j - 32;
    const int codep = jp >= 0 ? km_code((unsigned char)P.bases[off + jp]) : -1;
    const uint32_t cinv = __ballot_sync(KM_FULL, code < 0), pinv = __ballot_sync(KM_FULL, codep < 0);
    const uint32_t cgc = __ballot_sync(KM_FULL, code == 1 || code == 2), pgc = __ballot_sync(KM_FULL, codep == 1 || codep == 2);
    const uint32_t cc = code < 0 ? 0u : (uint32_t)code, cp = codep < 0 ? 0u : (uint32_t)codep;
    const uint64_t cur = ((uint64_t)__reduce_or_sync(KM_FULL, lane < 16 ? cc << (30 - 2 * lane) : 0u) << 32) | __reduce_or_sync(KM_FULL, lane >= 16 ? cc << (62 - 2 * lane) : 0u);
    const uint64_t prev = ((uint64_t)__reduce_or_sync(KM_FULL, lane < 16 ? cp << (30 - 2 * lane) : 0u) << 32) | __reduce_or_sync(KM_FULL, lane >= 16 ? cp << (62 - 2 * lane) : 0u);
    const int s = 62 - 2 * lane;
    const uint64_t fwd = ((cur >> s) | (s ? (prev << (64 - s)) : 0ull)) & kmask;
    const uint64_t inv64 = ((uint64_t)cinv << 32) | pinv, gc64 = ((uint64_t)cgc << 32) | pgc;
    const uint64_t wmask = (1ull << k) - 1;
    const int wsh = 32 + lane - k + 1;
    ok = ((inv64 >> wsh) & wmask) == 0;
    ok_prev = wsh > 0 && ((inv64 >> (wsh - 1)) & wmask) == 0;
    canon = 0;
    bool fwd_is_canon = true;
    if (ok) { const uint64_t rc = km_revcomp(fwd, kmer_bits); fwd_is_canon = fwd < rc; canon = fwd_is_canon ? fwd : rc; }
    cgc_bit = (cgc >> lane) & 1;
    gc_win = __popcll((gc64 >> wsh) & wmask);
    if (KEY) {
        const int lm = P.db.line_m;
        if (lm) {
            const int w = k - lm + 1;
            const uint32_t hh = km_slide_hash(fwd, lm), hp = km_slide_hash(prev >> s, lm);       // the m-mer ending at this lane's base, here and one chunk back
            const uint32_t r0 = km_slide_r0(hh), l0 = km_slide_l0(hh), r0p = km_slide_r0(hp), l0p = km_slide_l0(hp);
            uint32_t kr = r0, kl = l0;
            for (int d = 1; d < w; d++) { kr = km_slide_r(kr, km_fetch_left(r0, r0p, lane, d), d); kl = km_slide_l(kl, km_fetch_left(l0, l0p, lane, d), d); }
            *key = ok ? km_line_x_of(canon, km_slide_finish(kr, kl, fwd, fwd_is_canon, k, lm), k, lm, P.db.line_bits) : 0ull;
        } else *key = ok ? km_mix(canon, kmer_bits) : 0ull;
    }
}
__device__ __forceinline__ uint32_t kl_slot(uint64_t canon) { return (uint32_t)((canon * 0x9E3779B97F4A7C15ull) >> 40) & (KM_LONG_SLOTS - 1); }

__global__ void __launch_bounds__(KM_LONG_THREADS, 1) km_encode_probe_long_kernel(KmProbeParams P) {
    extern __shared__ __align__(16) unsigned long long kl_set[];          // KM_LONG_SLOTS entries, ~0 = empty
    __shared__ int s_valid, s_vgc, s_vtot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = KM_LONG_THREADS / 32;
    const int k = P.db.kmer_len, kmer_bits = P.db.kmer_bits;
    const uint64_t kmask = (1ull << kmer_bits) - 1;
    unsigned long long st_lookups = 0, st_hits = 0, st_lists = 0, st_extra = 0, st_pmiss = 0;
    for (int i = threadIdx.x; i < KM_LONG_SLOTS; i += KM_LONG_THREADS) kl_set[i] = ~0ull;
    if (threadIdx.x == 0) { s_valid = 0; s_vgc = 0; s_vtot = 0; }
    __syncthreads();
    for (uint32_t r = blockIdx.x; r < P.n_reads; r += gridDim.x) {
        const uint64_t off = P.offs[r];
        const int len = (int)(P.offs[r + 1] - off);
        if (len < KM_LONG_MIN || len > KM_LONG_MAX) continue;                 // block-uniform
        const int nchunks = (len + 31) >> 5;
        // ---- phase 1: every k-mer window into the set, lowest position wins; GC bookkeeping (:994-1008)
        int valid = 0, vgc = 0, vtot = 0;
        for (int c = wid; c < nchunks; c += NW) {
            uint64_t canon; bool ok, ok_prev; uint32_t gcb; int gcw;
            kl_encode_chunk<false>(P, off, len, c, lane, k, kmask, kmer_bits, canon, ok, ok_prev, gcb, gcw);
            valid += ok; vtot += ok ? (ok_prev ? 1 : k) : 0; vgc += ok ? (ok_prev ? (int)gcb : gcw) : 0;
            if (ok) {
                const int p = (c << 5) + lane - k + 1;
                const unsigned long long e = (canon << KM_LONG_POS_BITS) | (unsigned long long)p;
                uint32_t h = kl_slot(canon);
                for (;;) {
                    unsigned long long cur_e = kl_set[h];
                    if (cur_e == ~0ull) { cur_e = atomicCAS(&kl_set[h], ~0ull, e); if (cur_e == ~0ull) break; }
                    if ((cur_e >> KM_LONG_POS_BITS) == canon) { atomicMin(&kl_set[h], e); break; }
                    h = (h + 1) & (KM_LONG_SLOTS - 1);
                }
            }
        }
        valid = km_warp_sum(valid); vgc = km_warp_sum(vgc); vtot = km_warp_sum(vtot);
        if (lane == 0) { atomicAdd(&s_valid, valid); atomicAdd(&s_vgc, vgc); atomicAdd(&s_vtot, vtot); }
        __syncthreads();
        // ---- phase 2: first occurrences probe the table, four chunks of gathers in flight per warp
        for (int c0 = wid * 4; c0 < nchunks; c0 += NW * 4) {
            uint64_t xk[4], cn[4], bk[4][4]; bool first[4]; uint32_t owner[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                first[u] = false; xk[u] = 0; cn[u] = 0; owner[u] = 0;
                bk[u][0] = bk[u][1] = bk[u][2] = bk[u][3] = 0;
                const int c = c0 + u;
                if (c >= nchunks) continue;                                   // warp-uniform
                uint64_t canon, key; bool ok, ok_prev; uint32_t gcb; int gcw;
                kl_encode_chunk<true>(P, off, len, c, lane, k, kmask, kmer_bits, canon, ok, ok_prev, gcb, gcw, &key);
                if (ok) {
                    const int p = (c << 5) + lane - k + 1;
                    uint32_t h = kl_slot(canon);
                    for (;;) {
                        const unsigned long long cur_e = kl_set[h];
                        if ((cur_e >> KM_LONG_POS_BITS) == canon) { first[u] = (int)(cur_e & ((1ull << KM_LONG_POS_BITS) - 1)) == p; break; }
                        h = (h + 1) & (KM_LONG_SLOTS - 1);
                    }
                    if (first[u]) {
                        cn[u] = canon;
                        xk[u] = key;                                          // = km_key(P.db, canon), from the sliding minimum instead of the definition
                        if (P.do_probe) km_first_load(P.db, xk[u], owner[u], bk[u][0], bk[u][1], bk[u][2], bk[u][3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = c0 + u;
                if (c >= nchunks) continue;
                const int j = (c << 5) + lane, p = j - k + 1;
                uint32_t hw = KM_HIT_INVALID;
                if (first[u]) {
                    hw = KM_HIT_MISS;
                    if (P.do_probe) {
                        uint32_t extra = 0;
                        hw = km_first_finish(P.db, xk[u], owner[u], bk[u][0], bk[u][1], bk[u][2], bk[u][3], extra);
                        if (P.stats) {
                            st_lookups++; st_extra += extra;
                            if (hw != KM_HIT_MISS) { st_hits++; if (hw & KM_HIT_LIST) st_lists++; }
                            else if (P.db.prefix_bits) { const uint64_t pf = cn[u] >> P.db.prefix_shift; if (!((P.db.prefix_bits[pf >> 5] >> (pf & 31)) & 1)) st_pmiss++; }
                        }
                    }
                }
                if (p >= 0 && j < len) P.hit[off + p] = hw;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const float frac = __fdiv_rn((float)s_vgc, (float)s_vtot);                     // :1205-1206
            const float gc_pcnt = __double2float_rn(__dmul_rn((double)frac, 100.0));
            const float q = __fdiv_rn(gc_pcnt, 10.0f);
            P.hdr[r] = make_int2(s_valid, s_vtot > 0 ? (int)q : 0);
            s_valid = 0; s_vgc = 0; s_vtot = 0;
        }
        for (int i = threadIdx.x; i < KM_LONG_SLOTS; i += KM_LONG_THREADS) kl_set[i] = ~0ull;
        __syncthreads();
    }
    if (P.stats) {
        st_lookups = km_warp_sum((int)st_lookups); st_hits = km_warp_sum((int)st_hits); st_lists = km_warp_sum((int)st_lists); st_extra = km_warp_sum((int)st_extra);
        st_pmiss = km_warp_sum((int)st_pmiss);
        if (lane == 0) { atomicAdd(&P.stats->lookups, st_lookups); atomicAdd(&P.stats->hits, st_hits); atomicAdd(&P.stats->list_hits, st_lists); atomicAdd(&P.stats->extra_buckets, st_extra); atomicAdd(&P.stats->prefix_miss, st_pmiss); }
    }
}

#ifndef KM_FAST_CTAS
#define KM_FAST_CTAS 2
#endif
// value of the lane `sh` bases to the left: this chunk's lanes, or the tail of the previous chunk (kmat_mzr.h, sliding minimum).
// Which of the two a SOURCE lane hands out depends on its own index only (the last sh lanes serve the next chunk's first
// lanes), so one shuffle does it.
__device__ __forceinline__ uint32_t km_fetch_left(uint32_t cur, uint32_t prv, int lane, int sh) {
    return __shfl_sync(KM_FULL, lane < 32 - sh ? cur : prv, (lane - sh) & 31);
}
// LINE: the table has a minimizer-ordered first level (kmat_mzr.h).  Every lane hashes the m-mer that ends at its base, a
// sliding minimum over the lanes gives each k-mer its minimizer, and the lanes that hold the k-mers of one super-k-mer then
// read different sectors of ONE 128-byte line in the same instruction: ~48 requests per 150 bp read instead of 131.
template <int NCH, int SETN, bool STATS, bool PEERS, bool LINE>
__global__ void __launch_bounds__(KM_PROBE_WARPS * 32, NCH <= 5 ? KM_FAST_CTAS : 1) km_encode_probe_fast_kernel(KmProbeParams P) {
    if (!PEERS) P.db.n_peers = 0;                 // the replicated table's instantiation carries no owner logic at all
    extern __shared__ __align__(16) unsigned char km_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *bitmap = (uint32_t *)(km_smem + (size_t)wib * (SETN / 8));
    const uint64_t warp_global = blockIdx.x * KM_PROBE_WARPS + wib, n_warps = gridDim.x * KM_PROBE_WARPS;
    const int k = P.db.kmer_len, kmer_bits = P.db.kmer_bits;
    const uint64_t kmask = (1ull << kmer_bits) - 1;
    const int lm = P.db.line_m, lb = P.db.line_bits;
    int sl1 = 0, sl2 = 0, sl3 = 0;
    if (LINE) km_slide_shifts(k - lm + 1, sl1, sl2, sl3);
    for (int i = lane; i < SETN / 32; i += 32) bitmap[i] = 0;
    unsigned long long st_lookups = 0, st_hits = 0, st_lists = 0, st_extra = 0, st_pmiss = 0;
    __syncwarp();

    const uint64_t n = P.n_reads;
    uint64_t r = warp_global;
    uint64_t offA = 0, offB = 0; int lenA = 0, lenB = 0;          // read r (A) and read r + n_warps (B)
    if (r < n) { offA = P.offs[r]; lenA = (int)(P.offs[r + 1] - offA); }
    if (r + n_warps < n) { offB = P.offs[r + n_warps]; lenB = (int)(P.offs[r + n_warps + 1] - offB); }
    unsigned char raw[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) { const int j = (c << 5) + lane; raw[c] = j < lenA ? (unsigned char)P.bases[offA + j] : (unsigned char)0; }

    for (; r < n; r += n_warps) {
        const uint64_t off = offA; const int len = lenA;
        int code[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) code[c] = km_code(raw[c]);
        // ---- next read's bases, and the offsets of the one after it: in flight during this iteration
        offA = offB; lenA = lenB;
#pragma unroll
        for (int c = 0; c < NCH; c++) { const int j = (c << 5) + lane; raw[c] = j < lenA ? (unsigned char)P.bases[offA + j] : (unsigned char)0; }
        { const uint64_t r2 = r + 2 * n_warps; lenB = 0; if (r2 < n) { offB = P.offs[r2]; lenB = (int)(P.offs[r2 + 1] - offB); } }

        // ---- encode every chunk (read_label.cpp:943-950, 978-1009, GC bookkeeping :994-1008).  ONE copy of the chunk body in
        //      the instruction stream (it is the bulk of the kernel's code and the fully unrolled form stalled on instruction
        //      fetch, profiles/r02e): the chunk's code is selected into a scalar and its key selected back, so code[] / xk[]
        //      stay in registers
        uint64_t xk[NCH];                  // table key of the canonical k-mer ending at base j = 32 c + lane
        uint64_t canon_s[STATS ? NCH : 1];
#pragma unroll
        for (int c = 0; c < NCH; c++) { xk[c] = 0; if (STATS) canon_s[c] = 0; }
        uint32_t okbits = 0;
        uint64_t prev = 0; uint32_t pinv = 0xFFFFFFFFu, pgc = 0;
        int valid = 0, vgc = 0, vtot = 0;
        uint32_t sp_r[3] = {KM_SLIDE_NONE, KM_SLIDE_NONE, KM_SLIDE_NONE}, sp_l[3] = {KM_SLIDE_NONE, KM_SLIDE_NONE, KM_SLIDE_NONE};
        const int nch = (len + 31) >> 5;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
            int code1 = code[0];
#pragma unroll
            for (int q = 1; q < NCH; q++) if (c == q) code1 = code[q];
            const uint32_t cinv = __ballot_sync(KM_FULL, code1 < 0);
            const uint32_t cgc = __ballot_sync(KM_FULL, code1 == 1 || code1 == 2);
            const uint32_t cc = code1 < 0 ? 0u : (uint32_t)code1;
            const uint32_t hi = __reduce_or_sync(KM_FULL, lane < 16 ? cc << (30 - 2 * lane) : 0u);
            const uint32_t lo = __reduce_or_sync(KM_FULL, lane >= 16 ? cc << (62 - 2 * lane) : 0u);
            const uint64_t cur = ((uint64_t)hi << 32) | lo;
            const int s = 62 - 2 * lane;
            const uint64_t fwd = ((cur >> s) | (s ? (prev << (64 - s)) : 0ull)) & kmask;
            const uint64_t inv64 = ((uint64_t)cinv << 32) | pinv, gc64 = ((uint64_t)cgc << 32) | pgc;
            const uint64_t wmask = (1ull << k) - 1;
            const int wsh = 32 + lane - k + 1;
            const bool ok = ((inv64 >> wsh) & wmask) == 0;
            const bool ok_prev = wsh > 0 && ((inv64 >> (wsh - 1)) & wmask) == 0;
            uint32_t sl_kr = 0, sl_kl = 0;
            if (LINE) {                                          // every lane takes part: one hash per base, three doubling steps
                const uint32_t hh = km_slide_hash(fwd, lm);
                const uint32_t r0 = km_slide_r0(hh), l0 = km_slide_l0(hh);
                const uint32_t r1 = km_slide_r(r0, km_fetch_left(r0, sp_r[0], lane, sl1), sl1), l1 = km_slide_l(l0, km_fetch_left(l0, sp_l[0], lane, sl1), sl1);
                const uint32_t r2 = km_slide_r(r1, km_fetch_left(r1, sp_r[1], lane, sl2), sl2), l2 = km_slide_l(l1, km_fetch_left(l1, sp_l[1], lane, sl2), sl2);
                sl_kr = km_slide_r(r2, km_fetch_left(r2, sp_r[2], lane, sl3), sl3);
                sl_kl = km_slide_l(l2, km_fetch_left(l2, sp_l[2], lane, sl3), sl3);
                sp_r[0] = r0; sp_r[1] = r1; sp_r[2] = r2; sp_l[0] = l0; sp_l[1] = l1; sp_l[2] = l2;
            }
            uint64_t x1 = 0, canon1 = 0;
            if (ok) {
                const uint64_t rc = km_revcomp(fwd, kmer_bits);
                canon1 = fwd < rc ? fwd : rc;                                  // read_label.cpp:1009
                if (LINE) x1 = km_line_x_of(canon1, km_slide_finish(sl_kr, sl_kl, fwd, fwd < rc, k, lm), k, lm, lb);
                else x1 = km_mix(canon1, kmer_bits);
                okbits |= 1u << c;
            }
#pragma unroll
            for (int q = 0; q < NCH; q++) if (c == q) { xk[q] = x1; if (STATS) canon_s[q] = canon1; }
            const int add_tot = ok ? (ok_prev ? 1 : k) : 0;
            const int add_gc = ok ? (ok_prev ? (int)((cgc >> lane) & 1) : __popcll((gc64 >> wsh) & wmask)) : 0;
            valid += ok; vtot += add_tot; vgc += add_gc;
            prev = cur; pinv = cinv; pgc = cgc;
        }
        // ---- dedup: the lowest position of every distinct k-mer wins
        uint32_t first = 0, suspect = 0;       // per chunk bits of this lane
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if ((okbits >> c) & 1) {
                const uint32_t h = KM_SET_HASH(xk[c]) & (SETN - 1);
                const uint32_t old = atomicOr(bitmap + (h >> 5), 1u << (h & 31));
                if ((old >> (h & 31)) & 1) suspect |= 1u << c; else first |= 1u << c;
            }
        }
        if (__any_sync(KM_FULL, suspect != 0)) {
#pragma unroll 1
            for (int c = 0; c < nch; c++) {
                uint32_t pend = __ballot_sync(KM_FULL, (suspect >> c) & 1);
                if (!pend) continue;
                uint64_t xc = xk[0];
#pragma unroll
                for (int q = 1; q < NCH; q++) if (c == q) xc = xk[q];
                while (pend) {                                        // rare: ~ (k-mers per read)^2 / (2 SETN) suspects per read
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const uint64_t key = kb_shfl_u64(xc, src);
                    const int ps = (c << 5) + src;                    // the suspect's position index (base index of its last base)
                    bool lower = false;
#pragma unroll
                    for (int c2 = 0; c2 < NCH; c2++) {
                        if (((okbits >> c2) & 1) && xk[c2] == key) {
                            const int pq = (c2 << 5) + lane;
                            if (pq < ps) lower = true;
                            else if (pq > ps) first &= ~(1u << c2);    // a later copy of the suspect's k-mer is never the first
                        }
                    }
                    const bool dup = __any_sync(KM_FULL, lower);
                    if (lane == src && !dup) first |= 1u << c;
                }
            }
        }
        // the words this lane touched go back to zero for the next read (every set bit lies in such a word)
#pragma unroll
        for (int c = 0; c < NCH; c++) if ((okbits >> c) & 1) bitmap[(KM_SET_HASH(xk[c]) & (SETN - 1)) >> 5] = 0;
        // ---- probe the first occurrences: the first request of every lookup of the read is issued before the first one is
        //      looked at (NCH independent LDG.256 per lane in flight), then one hit word per k-mer start position
        uint64_t bk[NCH][4];
        uint32_t owner[PEERS ? NCH : 1];
        uint32_t hwv[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) hwv[c] = ((first >> c) & 1) ? KM_HIT_MISS : KM_HIT_INVALID;
        if (P.do_probe) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                bk[c][0] = bk[c][1] = bk[c][2] = bk[c][3] = 0;
                if (PEERS) owner[c] = 0;
                if ((first >> c) & 1) {
                    if (PEERS) km_first_load(P.db, xk[c], owner[c], bk[c][0], bk[c][1], bk[c][2], bk[c][3]);      // direct sharded mode: the gather goes to the owner's memory
                    else if (LINE) km_load_sector(P.db.lines + (((xk[c] >> KM_LINE_XSHIFT) - P.db.line_first) * 4 + ((xk[c] >> KM_MZR_KEY_BITS) & 3u)) * KM_SLOTS_PER_BUCKET, bk[c][0], bk[c][1], bk[c][2], bk[c][3]);
                    else km_load_bucket(P.db.slots + ((xk[c] >> P.db.rem_bits) & P.db.bucket_mask) * KM_SLOTS_PER_BUCKET, bk[c][0], bk[c][1], bk[c][2], bk[c][3]);
                }
            }
            uint32_t more = 0;                 // lookups that have to go on: a flagged sector (second level) / a full home bucket
            uint32_t st_x = 0;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                if ((first >> c) & 1) {
                    uint32_t hw = KM_HIT_MISS, extra = 0;
                    if (PEERS) hw = km_first_finish(P.db, xk[c], owner[c], bk[c][0], bk[c][1], bk[c][2], bk[c][3], extra);
                    else if (LINE) { if (km_sector_match(bk[c][0], bk[c][1], bk[c][2], bk[c][3], xk[c] & ((1ull << KM_MZR_KEY_BITS) - 1), hw) == 2) more |= 1u << c; }
                    else if (km_bucket_match(bk[c][0], bk[c][1], bk[c][2], bk[c][3], xk[c] & ((1ull << P.db.rem_bits) - 1), 0, hw) == 2) more |= 1u << c;
                    hwv[c] = hw; st_x += extra;
                }
            }
            // the rare continuations of the whole read together: their loads are issued back to back by the lanes that need
            // them instead of one dependent chain per chunk with one or two lanes active (9 % of the stall samples before)
            if (!PEERS && __any_sync(KM_FULL, more != 0)) {
                if (LINE) {
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        if ((more >> c) & 1) {
                            xk[c] = km_mix(km_line_kmer_of(xk[c], k, lm, lb), kmer_bits);         // the second level is addressed by the mixed k-mer
                            if (P.db.slots) km_load_bucket(P.db.slots + ((xk[c] >> P.db.rem_bits) & P.db.bucket_mask) * KM_SLOTS_PER_BUCKET, bk[c][0], bk[c][1], bk[c][2], bk[c][3]);
                            else bk[c][0] = bk[c][1] = bk[c][2] = bk[c][3] = 0;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if ((more >> c) & 1) {
                        uint32_t hw = KM_HIT_MISS, extra = 1;
                        if (LINE) { if (km_bucket_match(bk[c][0], bk[c][1], bk[c][2], bk[c][3], xk[c] & ((1ull << P.db.rem_bits) - 1), 0, hw) == 2) { uint32_t e2 = 0; hw = km_probe_buckets(P.db, xk[c], 0u, e2, 1); extra += 1 + e2; } }
                        else { uint32_t e2 = 0; hw = km_probe_buckets(P.db, xk[c], 0u, e2, 1); extra += e2; }
                        hwv[c] = hw; st_x += extra;
                    }
                }
            }
            if (STATS) {
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if ((first >> c) & 1) {
                        st_lookups++;
                        if (hwv[c] != KM_HIT_MISS) { st_hits++; if (hwv[c] & KM_HIT_LIST) st_lists++; }
                        else if (P.db.prefix_bits) {
                            const uint64_t pf = canon_s[c] >> P.db.prefix_shift;
                            if (!((P.db.prefix_bits[pf >> 5] >> (pf & 31)) & 1)) st_pmiss++;
                        }
                    }
                }
                st_extra += st_x;
            }
        }
        {
            uint32_t *hp = P.hit + off + lane - (k - 1);                    // hit word of the k-mer that ENDS at this lane's base of chunk 0
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int j = (c << 5) + lane;
                if (j >= k - 1 && j < len) hp[c << 5] = hwv[c];
            }
            if (P.xq) {                                                     // DB-sharded encode only: the table keys travel to their owners
                uint64_t *xp = P.xq + off + lane - (k - 1);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int j = (c << 5) + lane;
                    if (j >= k - 1 && j < len && ((first >> c) & 1)) xp[c << 5] = xk[c];
                }
            }
        }
        valid = km_warp_sum(valid); vgc = km_warp_sum(vgc); vtot = km_warp_sum(vtot);
        if (lane == 0) {
            const float frac = __fdiv_rn((float)vgc, (float)vtot);                         // :1205-1206
            const float gc_pcnt = __double2float_rn(__dmul_rn((double)frac, 100.0));
            const float q = __fdiv_rn(gc_pcnt, 10.0f);
            P.hdr[r] = make_int2(valid, vtot > 0 ? (int)q : 0);
        }
        __syncwarp();                      // the bitmap is clean again before the next read's atomics
    }
    if (STATS) {
        st_lookups = km_warp_sum((int)st_lookups); st_hits = km_warp_sum((int)st_hits); st_lists = km_warp_sum((int)st_lists);
        st_extra = km_warp_sum((int)st_extra); st_pmiss = km_warp_sum((int)st_pmiss);
        if (lane == 0) {
            atomicAdd(&P.stats->lookups, st_lookups); atomicAdd(&P.stats->hits, st_hits); atomicAdd(&P.stats->list_hits, st_lists);
            atomicAdd(&P.stats->extra_buckets, st_extra); atomicAdd(&P.stats->prefix_miss, st_pmiss);
        }
    }
}

template <int NCH, int SETN, bool STATS, bool PEERS, bool LINE>
static int km_launch_fast(const KmProbeParams &P, int ctas_per_sm, cudaStream_t stream) {
    const int smem = KM_PROBE_WARPS * (SETN / 8);
    // per instantiation AND per device (function attributes belong to a device's context; a process may drive several
    // GPUs, one thread each): CTAs that fit the device at once (persistent grid)
    static std::atomic<int> resident_of[KM_MAX_DEVICES], sms_of[KM_MAX_DEVICES];
    int dev = 0;
    KM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= KM_MAX_DEVICES) { kmat_set_error("device index %d out of range", dev); return KMAT_ERR_ARG; }
    int resident = resident_of[dev].load(), sms = sms_of[dev].load();
    if (!resident) {
        KM_CUDA(cudaFuncSetAttribute(km_encode_probe_fast_kernel<NCH, SETN, STATS, PEERS, LINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        KM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, km_encode_probe_fast_kernel<NCH, SETN, STATS, PEERS, LINE>, KM_PROBE_WARPS * 32, smem));
        KM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (const char *e = getenv("KMAT_PROBE_CTAS")) { const int v = atoi(e); if (v > 0 && v < per_sm) per_sm = v; }
        resident = std::max(1, per_sm) * sms;
        sms_of[dev].store(sms); resident_of[dev].store(resident);       // racing threads compute the same values
    }
    const uint32_t want = (P.n_reads + KM_PROBE_WARPS - 1) / KM_PROBE_WARPS;
    const uint32_t cap = ctas_per_sm > 0 ? std::min<uint32_t>((uint32_t)resident, (uint32_t)(ctas_per_sm * sms)) : (uint32_t)resident;
    const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(want, cap));
    km_encode_probe_fast_kernel<NCH, SETN, STATS, PEERS, LINE><<<grid, KM_PROBE_WARPS * 32, smem, stream>>>(P);
    return KMAT_OK;
}
template <int NCH, int SETN>
static int km_launch_fast_pick(const KmProbeParams &P, bool stats, bool peers, bool line, int ctas_per_sm, cudaStream_t stream) {
    if (line) {
        if (peers) return stats ? km_launch_fast<NCH, SETN, true, true, true>(P, ctas_per_sm, stream) : km_launch_fast<NCH, SETN, false, true, true>(P, ctas_per_sm, stream);
        return stats ? km_launch_fast<NCH, SETN, true, false, true>(P, ctas_per_sm, stream) : km_launch_fast<NCH, SETN, false, false, true>(P, ctas_per_sm, stream);
    }
    if (peers) return stats ? km_launch_fast<NCH, SETN, true, true, false>(P, ctas_per_sm, stream) : km_launch_fast<NCH, SETN, false, true, false>(P, ctas_per_sm, stream);
    return stats ? km_launch_fast<NCH, SETN, true, false, false>(P, ctas_per_sm, stream) : km_launch_fast<NCH, SETN, false, false, false>(P, ctas_per_sm, stream);
}

int km_launch_encode_probe(const kmat_db *db, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint32_t max_len,
                           uint32_t *d_hit, int2 *d_hdr, uint64_t *d_kmers, uint8_t *d_flags, unsigned long long *d_long_sets,
                           uint32_t long_slots, int grid, KmStatsDev *d_stats, int do_probe, cudaStream_t stream, int ctas_per_sm, uint64_t *d_xq,
                           const KmPeer *d_peers, uint32_t n_peers) {
    KmProbeParams P;
    P.db = km_db_dev(db); P.db.peers = d_peers; P.db.n_peers = d_peers ? n_peers : 0; P.bases = d_bases; P.offs = d_offs; P.n_reads = n_reads; P.hit = d_hit; P.hdr = d_hdr;
    P.out_kmers = d_kmers; P.out_flags = d_flags; P.long_sets = d_long_sets; P.long_slots = long_slots; P.stats = d_stats;
    P.do_probe = do_probe; P.xq = d_xq; P.skip_mid = 0;
    const bool fast = !d_kmers && !d_flags && max_len <= 256 && db->kmer_len <= 24 && !getenv("KMAT_NO_FAST_PROBE");
    int rc = KMAT_OK;
    const bool peers = P.db.n_peers != 0, line = P.db.line_m != 0;
    if (fast && max_len <= 160) rc = km_launch_fast_pick<5, 16384>(P, d_stats != nullptr, peers, line, ctas_per_sm, stream);
    else if (fast) rc = km_launch_fast_pick<8, 32768>(P, d_stats != nullptr, peers, line, ctas_per_sm, stream);
    else {
        // mixed / long batches: reads of 257 .. 12000 bases get a CTA each (shared-memory dedup set), the any-length kernel
        // takes the rest
        const bool use_long = !d_kmers && !d_flags && !d_xq && max_len >= KM_LONG_MIN && db->geom.kmer_bits + KM_LONG_POS_BITS <= 64 && db->kmer_len <= 32 && !getenv("KMAT_NO_LONG_PROBE");
        if (use_long) {
            // per launch: the attribute belongs to the current device's context (several GPUs per process), and setting it is cheap
            KM_CUDA(cudaFuncSetAttribute(km_encode_probe_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_LONG_SLOTS * 8));
            P.skip_mid = 1;
            km_encode_probe_long_kernel<<<(int)std::max<uint32_t>(1, std::min<uint32_t>(n_reads, 148u)), KM_LONG_THREADS, KM_LONG_SLOTS * 8, stream>>>(P);
            g_km_launches++;
            KM_CUDA(cudaGetLastError());
        }
        km_encode_probe_kernel<<<grid, KM_PROBE_WARPS * 32, 0, stream>>>(P);
    }
    if (rc != KMAT_OK) return rc;
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}

int km_probe_grid(uint32_t n_reads) {
    // persistent-style grid: 148 SMs x resident CTAs (8 warps, 32 KB smem -> 6 CTAs/SM by smem, 8 by threads)
    const uint32_t want = (n_reads + KM_PROBE_WARPS - 1) / KM_PROBE_WARPS;
    return (int)std::max<uint32_t>(1, std::min<uint32_t>(want, 148u * 6));
}

extern "C" int kmat_encode_batch(const kmat_db *db, const char *bases, const uint64_t *offs, uint32_t n_reads, uint64_t *kmers,
                                 uint8_t *flags, int32_t *valid_kmers, int32_t *bin_sel) {
    if (!db || !offs || (n_reads && !bases)) { kmat_set_error("kmat_encode_batch: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(db->device));
    if (!n_reads) return KMAT_OK;
    const uint64_t total = offs[n_reads];
    uint32_t max_np = 0;
    for (uint32_t r = 0; r < n_reads; r++) max_np = std::max<uint32_t>(max_np, (uint32_t)(offs[r + 1] - offs[r]));
    char *d_b; uint64_t *d_o, *d_k; uint32_t *d_hit; int2 *d_hdr; uint8_t *d_f; unsigned long long *d_long = nullptr;
    KM_CUDA(cudaMalloc((void **)&d_b, total + 1)); KM_CUDA(cudaMalloc((void **)&d_o, (size_t)(n_reads + 1) * 8));
    KM_CUDA(cudaMalloc((void **)&d_k, (total + 1) * 8)); KM_CUDA(cudaMalloc((void **)&d_hit, (total + 1) * 4));
    KM_CUDA(cudaMalloc((void **)&d_hdr, (size_t)n_reads * sizeof(int2))); KM_CUDA(cudaMalloc((void **)&d_f, total + 1));
    KM_CUDA(cudaMemcpy(d_b, bases, total, cudaMemcpyHostToDevice));
    KM_CUDA(cudaMemcpy(d_o, offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice));
    KM_CUDA(cudaMemset(d_k, 0, (total + 1) * 8)); KM_CUDA(cudaMemset(d_f, 0, total + 1));
    const int grid = km_probe_grid(n_reads);
    uint32_t long_slots = 0;
    if (max_np > KM_DEDUP_SLOTS / 2) {
        long_slots = 1024; while (long_slots < 2 * max_np) long_slots <<= 1;
        KM_CUDA(cudaMalloc((void **)&d_long, (size_t)grid * KM_PROBE_WARPS * long_slots * 8));
    }
    int rc = km_launch_encode_probe(db, d_b, d_o, n_reads, max_np, d_hit, d_hdr, d_k, d_f, d_long, long_slots, grid, nullptr, 0, 0, 0, nullptr);
    if (rc == KMAT_OK) {
        std::vector<int2> hdr(n_reads);
        KM_CUDA(cudaMemcpy(hdr.data(), d_hdr, (size_t)n_reads * sizeof(int2), cudaMemcpyDeviceToHost));
        if (kmers) KM_CUDA(cudaMemcpy(kmers, d_k, total * 8, cudaMemcpyDeviceToHost));
        if (flags) KM_CUDA(cudaMemcpy(flags, d_f, total, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < n_reads; r++) { if (valid_kmers) valid_kmers[r] = hdr[r].x; if (bin_sel) bin_sel[r] = hdr[r].y; }
    }
    cudaFree(d_b); cudaFree(d_o); cudaFree(d_k); cudaFree(d_hit); cudaFree(d_hdr); cudaFree(d_f); cudaFree(d_long);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// random-gather roofline probe: uniform random aligned loads over a large span
// ---------------------------------------------------------------------------------------------
// access modes of the probe: 8 / 16 / 32 = plain loads of that width (32 = the table's LDG.256); the 1xx modes are
// 8-byte loads with different cache operators, used to find out what a random access costs in DRAM traffic
template <int MODE>
__global__ void km_gather_kernel(const uint8_t *__restrict__ base, uint64_t n_units, uint64_t n_gathers, uint64_t seed, unsigned long long *sink) {
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_gathers; i += (uint64_t)gridDim.x * blockDim.x) {
        // 2xx modes: G consecutive lanes share one random 128-byte line (what a table ordered by minimizer would look like
        // to the probe kernel: neighbouring k-mers of a read in one line); n_gathers still counts lanes
        const int G = MODE == 204 ? 8 : (MODE == 201 || MODE == 202 || MODE == 205) ? 4 : 1;
        uint64_t x = ((MODE >= 200 ? i / G : i) + seed) * 0x9E3779B97F4A7C15ull;
        x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
        const uint64_t u = MODE >= 200 ? (uint64_t)(((unsigned __int128)x * (n_units / 4)) >> 64) * 4 : (uint64_t)(((unsigned __int128)x * n_units) >> 64);
        const uint8_t *p = base + u * 32;                     // one access per 32-byte sector (2xx: the line's first sector)
        unsigned long long v = 0;
        if (MODE == 201) { uint64_t a, b, c, d; km_load_bucket((const uint64_t *)(p + (i & 3) * 32), a, b, c, d); v = a ^ b ^ c ^ d; }     // 4 lanes, one sector each
        else if (MODE == 202) { uint64_t a, b, c, d; km_load_bucket((const uint64_t *)p, a, b, c, d); v = a ^ b ^ c ^ d; }                 // 4 lanes, the same sector
        else if (MODE == 203 || MODE == 205) {                                                                                             // every lane reads the whole line
#pragma unroll
            for (int q = 0; q < 4; q++) { uint64_t a, b, c, d; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + q * 32)); v ^= a ^ b ^ c ^ d; }
        }
        else if (MODE == 204) { uint64_t a, b; asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p + (i & 7) * 16)); v = a ^ b; }   // 8 lanes, 16 bytes each
        else
        if (MODE == 8) v = *(const unsigned long long *)p;
        else if (MODE == 16) { const ulonglong2 w = *(const ulonglong2 *)p; v = w.x ^ w.y; }
        else if (MODE == 32) { uint64_t a, b, c, d; km_load_bucket((const uint64_t *)p, a, b, c, d); v = a ^ b ^ c ^ d; }
        else if (MODE == 101) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 102) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 103) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 104) asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 105) asm volatile("ld.global.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 106) asm volatile("ld.global.lu.u64 %0, [%1];" : "=l"(v) : "l"(p));
        else if (MODE == 107) { uint64_t a, b, c, d; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); v = a ^ b ^ c ^ d; }
        else if (MODE == 108) { uint64_t a, b; asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); v = a ^ b; }
        else if (MODE == 109) { uint64_t a, b; asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); v = a ^ b; }
        else if (MODE == 110) {
            // one 32-byte bulk copy (TMA, non-tensor) per lane into shared memory, one mbarrier per warp
            __shared__ __align__(32) unsigned char s_buf[256 * 32];
            __shared__ __align__(8) unsigned long long s_bar[8];
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[w]);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_buf + threadIdx.x * 32);
            const uint32_t iter = (uint32_t)((i - (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x)) / ((uint64_t)gridDim.x * blockDim.x));
            if (iter == 0) { if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar)); __syncwarp(); }
            const uint32_t act = __activemask();
            if (lane == __ffs(act) - 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(32u * __popc(act)));
            __syncwarp(act);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" :: "r"(dst), "l"(p), "r"(bar) : "memory");
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(bar), "r"(iter & 1) : "memory");
            v = *(const unsigned long long *)(s_buf + threadIdx.x * 32);
            __syncwarp(act);
        }
        acc += v;
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}
static int km_gather_bench_impl(int device, int mem_device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                                double *gathers_per_s, double *sector_gbps);
extern "C" int kmat_gather_bench(int device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                                 double *gathers_per_s, double *sector_gbps) {
    return km_gather_bench_impl(device, device, span_bytes, access_bytes, n_gathers, iters, gathers_per_s, sector_gbps);
}
extern "C" int kmat_gather_bench_peer(int device, int mem_device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                                      double *gathers_per_s, double *sector_gbps) {
    return km_gather_bench_impl(device, mem_device, span_bytes, access_bytes, n_gathers, iters, gathers_per_s, sector_gbps);
}
static int km_gather_bench_impl(int device, int mem_device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                                double *gathers_per_s, double *sector_gbps) {
    if (kmat_device_count() <= device || kmat_device_count() <= mem_device) { kmat_set_error("CUDA device %d / %d not available", device, mem_device); return KMAT_ERR_NO_DEVICE; }
    const int mode = access_bytes;
    if (mode != 8 && mode != 16 && mode != 32 && !(mode >= 101 && mode <= 110) && !(mode >= 201 && mode <= 205)) return KMAT_ERR_ARG;
    uint8_t *buf; unsigned long long *sink;
    KM_CUDA(cudaSetDevice(mem_device));
    KM_CUDA(cudaMalloc((void **)&buf, span_bytes));
    KM_CUDA(cudaMemset(buf, 1, span_bytes));
    KM_CUDA(cudaDeviceSynchronize());
    KM_CUDA(cudaSetDevice(device));
    if (mem_device != device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(mem_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaFree(buf); kmat_set_error("cudaDeviceEnablePeerAccess(%d): %s", mem_device, cudaGetErrorString(e)); cudaGetLastError(); return KMAT_ERR_UNSUPPORTED; }
        cudaGetLastError();
    }
    KM_CUDA(cudaMalloc((void **)&sink, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint64_t n_units = span_bytes / 32;
    float best = 1e30f;
    for (int it = 0; it < iters + 1; it++) {
        cudaEventRecord(e0);
        const int blocks = 148 * 16, threads = 256;
#define KM_G(M) case M: km_gather_kernel<M><<<blocks, threads>>>(buf, n_units, n_gathers, 977 * it, sink); break;
        switch (mode) { KM_G(8) KM_G(16) KM_G(32) KM_G(101) KM_G(102) KM_G(103) KM_G(104) KM_G(105) KM_G(106) KM_G(107) KM_G(108) KM_G(109) KM_G(110) KM_G(201) KM_G(202) KM_G(203) KM_G(204) KM_G(205) }
#undef KM_G
        g_km_launches++;
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
    KM_CUDA(cudaGetLastError());
    if (gathers_per_s) *gathers_per_s = (double)n_gathers / (best * 1e-3);
    if (sector_gbps) *sector_gbps = (double)n_gathers * 32.0 / (best * 1e-3) / 1e9;
    return KMAT_OK;
}

// cudaLimitMaxL2FetchGranularity hint (32, 64 or 128 bytes) for the current device: random 32-byte probes do not
// benefit from wider DRAM fetches.  Returns the value in effect afterwards (or a negative error).
extern "C" int kmat_set_l2_fetch_granularity(int device, int bytes) {
    if (kmat_device_count() <= device) return KMAT_ERR_NO_DEVICE;
    KM_CUDA(cudaSetDevice(device));
    if (bytes > 0) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes); if (e != cudaSuccess) cudaGetLastError(); }
    size_t v = 0;
    KM_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
    return (int)v;
}

#include "kmat_kcov.cuh"
