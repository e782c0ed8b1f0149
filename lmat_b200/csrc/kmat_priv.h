// kmat_priv.h -- private host-side declarations of the CUDA translation units.
#ifndef KMAT_PRIV_H
#define KMAT_PRIV_H
#include <cuda_runtime.h>

#include <atomic>
#include <vector>

#include "kmat_internal.h"

#define KM_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            kmat_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_));  \
            cudaGetLastError();                                                                           \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? KMAT_ERR_NO_DEVICE    \
                   : (e_ == cudaErrorMemoryAllocation) ? KMAT_ERR_NOMEM : KMAT_ERR_CUDA;                  \
        }                                                                                                 \
    } while (0)

extern std::atomic<unsigned long long> g_km_launches;
#define KM_MAX_DEVICES 64     // per-device caches of launch geometry (function attributes are per device)

struct kmat_db {
    int device = 0, kmer_len = 0, tid_bytes = 2;
    KmTableGeom geom{};
    uint64_t n_buckets = 0, n_kmers = 0, pool_words = 0, prefix_bytes = 0;
    uint64_t *d_lines = nullptr; uint64_t n_lines = 0, line_first = 0, n_overflow = 0;    // first level (kmat_mzr.h), this shard's line range
    uint64_t *d_slots = nullptr;                                                           // second level
    uint32_t *d_pool = nullptr;
    uint32_t *d_prefix_bits = nullptr;
    uint64_t *d_stash_x = nullptr; uint32_t *d_stash_hit = nullptr; uint32_t n_stash = 0;   // overflow stash (see km_probe_x)
    int prefix_shift = 13;
    uint32_t n_sid = 65536;
    bool pool_shared = false;              // a shard built from the whole table's arrays (kmat_db_build_device): every shard holds the SAME list pool
    int shard_index = 0, shard_count = 1;  // DB-sharded mode: this table holds the k-mers with kmat_shard_of() == shard_index
    std::vector<uint32_t> stored_tids;     // 32-bit tables: dense stored id -> tid
    uint32_t *d_stored_tids = nullptr;     //   the same on the device (gene_label path)
};

struct KmDbDev;
struct KmPeer;
struct KmStatsDev;
KmDbDev km_db_dev(const kmat_db *db);
int km_probe_grid(uint32_t n_reads);
int km_launch_encode_probe(const kmat_db *db, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint32_t max_len,
                           uint32_t *d_hit, int2 *d_hdr, uint64_t *d_kmers, uint8_t *d_flags, unsigned long long *d_long_sets,
                           uint32_t long_slots, int grid, KmStatsDev *d_stats, int do_probe, cudaStream_t stream, int ctas_per_sm /* 0 = all that fit */,
                           uint64_t *d_xq /* DB-sharded mode: mixed first-occurrence k-mers per base offset, else NULL */,
                           const KmPeer *d_peers = nullptr, uint32_t n_peers = 0 /* direct sharded mode: probes go to the owner's memory */);
#define KM_PROBE_WARPS_HOST 8
#endif
