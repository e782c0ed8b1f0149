// kmat_device.cuh -- device-side structures and small helpers shared by the kernels.
#ifndef KMAT_DEVICE_CUH
#define KMAT_DEVICE_CUH
#include <cuda_runtime.h>

#include <cstdint>

#include "kmat_internal.h"

#define KM_FULL 0xffffffffu

// hit word written by the probe kernel for every k-mer start position of a read
#define KM_HIT_INVALID 0xFFFFFFFFu   // no valid k-mer starts here, or it duplicates an earlier one (label_vec[pos].first = -1)
#define KM_HIT_MISS 0xFFFFFFFEu      // valid first occurrence, not in the table (first = 0, empty set)
#define KM_HIT_LIST 0x80000000u      // bit 31: payload is a list-pool offset (4-byte words); else a stored id

// DB-sharded mode, direct variant (kmat_ctx_peer_attach): where shard `owner` lives.  The pointers are device addresses
// valid on THIS GPU: the shard's own memory, another GPU's memory mapped through CUDA IPC / peer access (the loads then
// travel over NVLink), or simply another allocation of the same device (virtual ranks in the tests).
struct KmPeer {
    const uint64_t *lines;        // the owner's first level (its line range starts at global line `line_first`), may be null
    uint64_t line_first;
    const uint64_t *slots;        // the owner's bucket array (second level; every shard sizes its own)
    uint64_t bucket_mask; int rem_bits;
    const uint64_t *stash_x; const uint32_t *stash_hit;
    const uint32_t *pool2;        // the owner ctx's resolved list pool
    uint32_t n_stash;
    uint32_t pool_base;           // KM_PEER_NO_BASE: list hits are tagged with the owner (records fetched from its pool);
                                  // else the owner's resolved pool has been copied into this rank's concatenated pool at this
                                  // (raw) word offset and list hit words carry offsets into that local copy
};
#define KM_PEER_NO_BASE 0xFFFFFFFFu
#define KM_PEER_SHIFT 27          // direct mode: a list hit word is LIST | owner << 27 | pool word offset (< 2^27)
#define KM_PEER_OFFMASK 0x07FFFFFFu

struct KmDbDev {
    const KmPeer *peers;          // direct sharded mode: n_peers entries indexed by km_owner_of_key; else null
    uint32_t n_peers;
    const uint64_t *lines;        // first level: 128-byte lines ordered by minimizer (kmat_mzr.h); null when line_m == 0
    uint64_t line_first, n_lines; // global index of lines[0] and the number of lines here (a shard holds the line range of its owner)
    int line_m, line_bits;        // minimizer length (0: no first level), log2 of the GLOBAL line count
    const uint64_t *slots;        // second level (or the whole table when line_m == 0): n_buckets * 4
    uint64_t bucket_mask;
    int kmer_bits, rem_bits, kmer_len, tid_bytes;
    const uint32_t *pool;         // list pool, 4-byte words
    const uint32_t *prefix_bits;  // bitmap over the reference's top-tier prefixes (stats only), may be null
    int prefix_shift;             // BITS_PER_2ND of the reference layout (13 for k=20, 9 for k=18)
    const uint64_t *stash_x;      // overflow stash: mixed keys (km_mix) of the few k-mers whose KM_MAX_DISP+1 buckets were
    const uint32_t *stash_hit;    //   all full at build time, ascending, and their hit words; searched only after 4 full buckets
    uint32_t n_stash;
};

struct KmStatsDev {
    unsigned long long lookups, hits, list_hits, list_ids, extra_buckets, prefix_miss, list_sectors;
    unsigned long long reads_fast, reads_slow, reads_error;
};

// 256-bit load of one bucket (one 32-byte sector): LDG.E.NA.LTC64B.256 on sm_100a.  L1 is bypassed (no reuse) and the
// L2 fill is limited to 64 bytes: by default a miss brings the whole 128-byte line in from DRAM (measured: 127 B of
// DRAM traffic per random gather, 64 B with this hint, same request rate - profiles/r01_gather_modes.md).
__device__ __forceinline__ void km_load_bucket(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// One sector of a first-level line.  The lanes of a warp that hold the k-mers of one super-k-mer read different sectors of
// the same line in the same instruction: one request (profiles/r02a_line_gather.jsonl).
__device__ __forceinline__ void km_load_sector(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// One bucket of the probe sequence against (rem, displacement d): 0 = found (hw set), 1 = absent for good (a free
// slot: the key cannot have been displaced further), 2 = bucket full, look at the next one.
__device__ __forceinline__ int km_bucket_match(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, uint64_t rem, int d, uint32_t &hw) {
    const uint64_t want = (1ull << 63) | ((uint64_t)d << 60) | (rem << 32);
    const uint64_t keymask = ~((1ull << 62) | 0xFFFFFFFFull);
    uint64_t hit = 0;
    if ((s0 & keymask) == want) hit = s0;
    if ((s1 & keymask) == want) hit = s1;
    if ((s2 & keymask) == want) hit = s2;
    if ((s3 & keymask) == want) hit = s3;
    if (hit) { hw = (uint32_t)hit | (((hit >> 62) & 1) ? KM_HIT_LIST : 0u); return 0; }
    hw = KM_HIT_MISS;
    return (s0 && s1 && s2 && s3) ? 2 : 1;
}
// One sector of a first-level line against the 28-bit key: 0 = found (hw set), 1 = absent for good (the sector never
// overflowed), 2 = the sector overflowed at build time: the key may be in the second level.
__device__ __forceinline__ int km_sector_match(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, uint64_t key, uint32_t &hw) {
    const uint64_t want = (1ull << 63) | (key << 32);
    const uint64_t keymask = ~((1ull << 62) | KM_LINE_OVF | 0xFFFFFFFFull);
    uint64_t hit = 0;
    if ((s0 & keymask) == want) hit = s0;
    if ((s1 & keymask) == want) hit = s1;
    if ((s2 & keymask) == want) hit = s2;
    if ((s3 & keymask) == want) hit = s3;
    if (hit) { hw = (uint32_t)hit | (((hit >> 62) & 1) ? KM_HIT_LIST : 0u); return 0; }
    hw = KM_HIT_MISS;
    return (s0 & KM_LINE_OVF) ? 2 : 1;
}

// The table key of a canonical k-mer: [line][sector][key] when the table has a first level, else the mixed k-mer.
__device__ __forceinline__ uint64_t km_key(const KmDbDev &db, uint64_t canon) {
    return db.line_m ? km_line_x(canon, db.kmer_len, db.line_m, db.line_bits) : km_mix(canon, db.kmer_bits);
}
// Owner shard of a table key (DB-sharded modes): by minimizer line, so that the k-mers of a super-k-mer stay together
// and a k-mer's two levels live on the same shard; by a second hash of the mixed k-mer when there is no first level.
__device__ __forceinline__ uint32_t km_owner_of_key(const KmDbDev &db, uint64_t x, uint32_t n_shards) {
    if (n_shards <= 1) return 0;
    if (db.line_m) return km_line_owner_of_g(km_line_g_of_x(x, db.kmer_len, db.line_m, db.line_bits), db.line_m, n_shards);
    return km_owner_of_x(x, n_shards);
}
// address of the sector table key x lives in: this table's lines, or (direct sharded mode) the owner's
__device__ __forceinline__ const uint64_t *km_sector_of(const KmDbDev &db, uint64_t x, uint32_t owner) {
    const uint64_t line = x >> KM_LINE_XSHIFT, sector = (x >> KM_MZR_KEY_BITS) & 3u;
    if (!db.n_peers) return db.lines + ((line - db.line_first) * 4 + sector) * KM_SLOTS_PER_BUCKET;
    const uint64_t *pl = (const uint64_t *)__ldg((const unsigned long long *)&db.peers[owner].lines);
    const uint64_t first = __ldg((const unsigned long long *)&db.peers[owner].line_first);
    return pl + ((line - first) * 4 + sector) * KM_SLOTS_PER_BUCKET;
}
// A hit word found in shard `owner`: list offsets are local to the owner's pool, so the owner rides along (direct mode)
__device__ __forceinline__ uint32_t km_tag_owner(const KmDbDev &db, uint32_t hw, uint32_t owner) {
    if (!db.n_peers || hw == KM_HIT_MISS || !(hw & KM_HIT_LIST)) return hw;
    const uint32_t base = __ldg(&db.peers[owner].pool_base);
    return base == KM_PEER_NO_BASE ? hw | (owner << KM_PEER_SHIFT) : KM_HIT_LIST | (base + (hw & 0x7FFFFFFFu));
}
// Second level (or the whole table when there is no first level): open addressing over 4-slot buckets with the mixed
// k-mer xb, starting at displacement d0; the overflow stash after KM_MAX_DISP + 1 full buckets.
__device__ __forceinline__ uint32_t km_probe_buckets(const KmDbDev &db, uint64_t xb, uint32_t owner, uint32_t &extra, int d0) {
    const uint64_t *slots = db.slots; uint64_t bucket_mask = db.bucket_mask; int rem_bits = db.rem_bits;
    const uint64_t *stash_x = db.stash_x; const uint32_t *stash_hit = db.stash_hit;
    uint32_t lo = 0, hi = db.n_stash;
    if (db.n_peers) { const KmPeer pr = db.peers[owner]; slots = pr.slots; bucket_mask = pr.bucket_mask; rem_bits = pr.rem_bits; stash_x = pr.stash_x; stash_hit = pr.stash_hit; hi = pr.n_stash; }
    if (!slots) return KM_HIT_MISS;                        // an empty second level
    const uint64_t home = xb >> rem_bits;
    const uint64_t rem = xb & ((1ull << rem_bits) - 1);
#pragma unroll 1
    for (int d = d0; d <= KM_MAX_DISP; d++) {
        uint64_t s0, s1, s2, s3;
        km_load_bucket(slots + ((home + d) & bucket_mask) * KM_SLOTS_PER_BUCKET, s0, s1, s2, s3);
        uint32_t hw;
        if (km_bucket_match(s0, s1, s2, s3, rem, d, hw) != 2) return km_tag_owner(db, hw, owner);
        extra++;
    }
    // every bucket of the probe window is full: the key, if present, sits in the stash
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t v = stash_x[mid];
        if (v == xb) return km_tag_owner(db, stash_hit[mid], owner);
        if (v < xb) lo = mid + 1; else hi = mid;
    }
    return KM_HIT_MISS;
}
// Probe with the table key x = km_key(kmer): returns the hit word.  extra = number of additional requests (second level,
// buckets past a full one).  level0_done: the caller has already looked at the first-level sector (table with a first
// level) or at the home bucket (table without) and has to go on.
__device__ __forceinline__ uint32_t km_probe_x(const KmDbDev &db, uint64_t x, uint32_t &extra, int level0_done = 0) {
    extra = 0;
    const uint32_t owner = db.n_peers ? km_owner_of_key(db, x, db.n_peers) : 0u;
    if (!db.line_m) return km_probe_buckets(db, x, owner, extra, level0_done);
    if (!level0_done) {
        // a key outside this shard's line range (a lookup against one shard of a split table): not here
        if (!db.n_peers && (x >> KM_LINE_XSHIFT) - db.line_first >= db.n_lines) return KM_HIT_MISS;
        uint64_t s0, s1, s2, s3;
        km_load_sector(km_sector_of(db, x, owner), s0, s1, s2, s3);
        uint32_t hw;
        if (km_sector_match(s0, s1, s2, s3, x & ((1ull << KM_MZR_KEY_BITS) - 1), hw) != 2) return km_tag_owner(db, hw, owner);
    }
    extra = 1;
    return km_probe_buckets(db, km_mix(km_line_kmer_of(x, db.kmer_len, db.line_m, db.line_bits), db.kmer_bits), owner, extra, 0);
}
// The same in two halves, for kernels that issue the first request of many lookups before they look at any of them:
// km_first_load starts the first-level sector load (or the home-bucket load), km_first_finish matches it and goes on to the
// second level / the next bucket when it has to.
__device__ __forceinline__ void km_first_load(const KmDbDev &db, uint64_t x, uint32_t &owner, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    owner = db.n_peers ? km_owner_of_key(db, x, db.n_peers) : 0u;
    if (db.line_m) {
        if (!db.n_peers && (x >> KM_LINE_XSHIFT) - db.line_first >= db.n_lines) { a = b = c = d = 0; return; }     // not this shard's line: reads as an empty sector
        km_load_sector(km_sector_of(db, x, owner), a, b, c, d);
        return;
    }
    const uint64_t *slots = db.slots; uint64_t bucket_mask = db.bucket_mask; int rem_bits = db.rem_bits;
    if (db.n_peers) { const KmPeer pr = db.peers[owner]; slots = pr.slots; bucket_mask = pr.bucket_mask; rem_bits = pr.rem_bits; }
    km_load_bucket(slots + ((x >> rem_bits) & bucket_mask) * KM_SLOTS_PER_BUCKET, a, b, c, d);
}
__device__ __forceinline__ uint32_t km_first_finish(const KmDbDev &db, uint64_t x, uint32_t owner, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint32_t &extra) {
    uint32_t hw;
    extra = 0;
    int r;
    if (db.line_m) r = km_sector_match(a, b, c, d, x & ((1ull << KM_MZR_KEY_BITS) - 1), hw);
    else {
        int rem_bits = db.rem_bits;
        if (db.n_peers) rem_bits = db.peers[owner].rem_bits;
        r = km_bucket_match(a, b, c, d, x & ((1ull << rem_bits) - 1), 0, hw);
    }
    if (r != 2) return km_tag_owner(db, hw, owner);
    hw = km_probe_x(db, x, extra, 1);
    if (!db.line_m) extra++;
    return hw;
}
__device__ __forceinline__ uint32_t km_probe(const KmDbDev &db, uint64_t kmer, uint32_t &extra) {
    return km_probe_x(db, km_key(db, kmer), extra);
}

__device__ __forceinline__ int km_warp_sum(int v) { return __reduce_add_sync(KM_FULL, v); }
__device__ __forceinline__ unsigned long long km_warp_or64(unsigned long long v) {
    unsigned lo = __reduce_or_sync(KM_FULL, (unsigned)v), hi = __reduce_or_sync(KM_FULL, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

// glibc >= 2.27 logf (sysdeps/ieee754/flt-32/e_logf.c), evaluated in double exactly like the host libm;
// the reference scores with std::log(float) (read_label.cpp:688).  Exhaustively identical to the host
// logf for every positive finite float (tests/test_logf_stdsort.py checks the C twin of this routine).
__device__ __forceinline__ float km_logf(float x) {
    const double T[16][2] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
        {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
        {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
        {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
        {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const double Ln2 = 0x1.62e42fefa39efp-1;
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return __uint_as_float(0xff800000u);            // -inf
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __uint_as_float(0x7fc00000u);
        ix = __float_as_uint(__fmul_rn(x, 8388608.0f));
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15);
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = T[i][0], logc = T[i][1];
    const double z = (double)__uint_as_float(iz);
    const double r = __dadd_rn(__dmul_rn(z, invc), -1.0);
    const double y0 = __dadd_rn(logc, __dmul_rn((double)k, Ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

#endif
