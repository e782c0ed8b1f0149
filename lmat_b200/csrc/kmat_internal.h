// kmat_internal.h -- structures shared by the host side (kmat_host.cpp) and the CUDA side (kmat_device.cu).
#ifndef KMAT_INTERNAL_H
#define KMAT_INTERNAL_H
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/kmat.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void kmat_set_error(const char *fmt, ...);

// ---------------------------------------------------------------------------------------------
// Host logical table: what SortedDb::begin_/next expose (SortedDb.hpp:188-385), flattened.
// ---------------------------------------------------------------------------------------------
struct kmat_table {
    int kmer_len = 0;
    int tid_bytes = 2;
    uint64_t n_kmers = 0, n_ids = 0;
    const uint64_t *kmers = nullptr;   // ascending
    const uint64_t *offs = nullptr;    // n_kmers + 1
    const uint32_t *ids = nullptr;     // stored ids, list order
    // ownership
    std::vector<uint64_t> own_kmers, own_offs;
    std::vector<uint32_t> own_ids;
    void *map_base = nullptr;
    size_t map_len = 0;
};

// ---------------------------------------------------------------------------------------------
// Parsed run-time inputs
// ---------------------------------------------------------------------------------------------
struct kmat_null_model {
    int kmer_cnt = 0;                 // key of _rand_hits (read_label.cpp:571)
    bool loaded = false;              // the file existed (maps were created)
    std::vector<uint32_t> tid;        // rows, file order with duplicates collapsed (last wins)
    std::vector<uint16_t> cls;        // class id per row
    std::vector<float> cut;           // rows x nbins
};
struct kmat_inputs {
    bool has_tree = false, has_conv = false, has_prune = false, models_requested = false;
    std::vector<uint32_t> node_tid, node_parent;          // -c, file order (later duplicates win)
    std::vector<uint32_t> depth_tid, depth_val;           // -e
    std::vector<uint32_t> rank_tid;                       // -w
    std::vector<uint8_t> rank_code;                       //   1 = "strain", 2 = "species", 0 = other
    std::vector<uint32_t> conv_stored, conv_tid;          // -f
    std::vector<uint32_t> prune_tid, prune_rank;          // -m
    std::vector<uint32_t> plasmid_tid;                    // -r
    // -n
    int nbins = 0;
    std::vector<kmat_null_model> models;                  // one per distinct key, list-file order
    std::vector<int> read_len_vec;                        // starts as {0} (read_label.cpp:60), sorted after load
    std::vector<int> read_len_avgs;                       // starts as {0} (:61)
    std::vector<std::string> class_names;                 // ids 0..9 = gNum2rank keys (read_label.cpp:534-547)
    std::vector<int32_t> class_ranknum;                   // gRank2num[class] (:519-532), 0 for unknown strings
};

// ---------------------------------------------------------------------------------------------
// Device-side node records (SoA of two 16-byte records per node; nid = dense node index)
// ---------------------------------------------------------------------------------------------
#define KMAT_NONE 0xFFFFFFFFu
enum : uint32_t {
    KM_META_DEPTH_MASK = 0xFFFFu,      // -e depth (missing -> 0, see DESIGN.md "depth of unknown tid")
    KM_META_RANK_SHIFT = 16,           // 2 bits: 1 strain, 2 species
    KM_META_HUMAN = 1u << 18,          // isHuman (tid_checks.hpp:15-28)
    KM_META_DROP = 1u << 19,           // tid == 20999999 || badGenomes (read_label.cpp:82-104,1038)
    KM_META_PHIX = 1u << 20,           // isPhiX (tid_checks.hpp:13)
    KM_META_PLASMID = 1u << 21,        // isPlasmid (read_label.cpp:69)
    KM_META_INTREE = 1u << 22
};
struct KmNodeA { uint32_t tid, parent, meta, species_anc; };
struct KmNodeB { uint32_t tin, tout, path_off, path_len; };

struct KmHostCtx {                     // everything kmat_ctx uploads, built by kmat_host.cpp
    std::vector<KmNodeA> nodeA;
    std::vector<KmNodeB> nodeB;
    std::vector<uint32_t> paths;       // concatenated strict-ancestor lists, nearest first
    std::vector<uint32_t> prune_rank;  // per nid (tid_rank_map, missing -> 0); empty when -m absent
    std::vector<uint32_t> sid2nid;     // stored id -> nid, KMAT_NONE when the -f map lacks it
    uint32_t nid_human = KMAT_NONE, nid_one = KMAT_NONE;
    // null models
    int nbins = 0, n_models = 0, n_classes = 0;
    std::vector<int16_t> model_of_cand; // 65536 entries: closest()/getReadLen() folded, -1 = no model
    std::vector<int32_t> mrow;          // n_models x n_nodes -> row or -1
    std::vector<float> cut;             // rows x nbins
    std::vector<uint8_t> cls;           // rows
    std::vector<int32_t> class_ranknum;
};

// Build the node universe + model tables.  stored_tids: for 32-bit DBs the distinct stored tids
// (sid -> tid); empty for 16-bit DBs, where sid is the raw 16-bit value resolved through -f.
int kmat_build_host_ctx(const kmat_inputs &in, int tid_bytes, const std::vector<uint32_t> &stored_tids, KmHostCtx &out);

// ---------------------------------------------------------------------------------------------
// Hash-table geometry (see DESIGN.md "HBM table layout")
// ---------------------------------------------------------------------------------------------
// slot (u64):  [63] occupied  [62] is_list  [61:60] displacement  [59:32] remainder (28 bits)  [31:0] payload
// bucket = 4 slots = 32 bytes = one DRAM sector, fetched with one LDG.256.
#define KM_SLOTS_PER_BUCKET 4
#define KM_REM_BITS 28
#define KM_MAX_DISP 3
struct KmTableGeom {
    int kmer_bits;      // 2k
    int bucket_bits;    // number of buckets of the second level = 2^bucket_bits
    int rem_bits;       // 2k - bucket_bits  (<= KM_REM_BITS)
    int line_m;         // minimizer length of the first level (0: no first level)
    int line_bits;      // GLOBAL number of lines = 2^line_bits (a shard holds the lines of its owner range)
};

#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#else
#define KM_HD inline
#endif
// Bijective mix of the 2k-bit k-mer space, so that (bucket, remainder) identifies the k-mer.
KM_HD uint64_t km_mix(uint64_t x, int kmer_bits) {
    const uint64_t mask = kmer_bits >= 64 ? ~0ull : ((1ull << kmer_bits) - 1);
    const int s = kmer_bits / 2;
    x = (x * 0x9E3779B97F4A7C15ull) & mask;
    x ^= x >> s;
    x = (x * 0xD6E8FEB86659FD93ull) & mask;
    x ^= x >> s;
    x = (x * 0xCA5A826395121157ull) & mask;
    x ^= x >> s;
    return x;
}
// The first level of the table is ordered by minimizer (kmat_mzr.h): the key of a canonical k-mer is then km_line_x(),
// [line][sector][28-bit key]; tables whose k has no line geometry (km_line_m(k) == 0) use the mixed k-mer above for
// everything.  KM_SET_HASH: the key bits that index the probe kernel's per-read dedup bitmap.
#include "kmat_mzr.h"
#define KM_SET_HASH(x) ((((uint32_t)(x) ^ (uint32_t)((x) >> 30)) * 0x9E3779B1u) >> 16)
// Owner shard of a mixed key x = km_mix(kmer) in DB-sharded mode (SURVEY.md 8(e) mode B: hash prefix of the canonical
// k-mer; a raw k-mer prefix would be skewed).  A second multiply/xor-shift round so that owner and bucket index (the top
// bits of x) are independent; multiply-shift range reduction instead of a modulo.
#define KM_MAX_SHARDS 16
KM_HD uint32_t km_owner_of_x(uint64_t x, uint32_t n_shards) {
    uint64_t y = x * 0xA24BAED4963EE407ull;
    y ^= y >> 29;
    y *= 0x9FB21C651E98DF25ull;
    y ^= y >> 32;
    return (uint32_t)(((y >> 32) * (uint64_t)n_shards) >> 32);
}
#endif
