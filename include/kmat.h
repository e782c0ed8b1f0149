/* kmat.h -- C ABI of libkmat, a B200-native (sm_100a) implementation of LMAT's read_label hot path.
 *
 * Drop-in boundary.  The reference (LivGen/LMAT v1.2.4_2018a) has no FFI layer; the seam this ABI
 * replaces is the in-process call
 *     proc_line(tax_tree, len, read, k, INDEXDB<DBTID_T>* table, ofs, threshold, sopt, max_count, ...)
 *                                                                       (src/read_label.cpp:1211-1212, call site :1746)
 * and, beneath it, the duck-typed table contract
 *     bool begin_(u64 kmer, u16& count, u32& offset, u8& page);  void next(u32&, u8&, tid_T&);
 *     char get_kmer_length();  size_t size();                (src/kmerdb/SortedDb.hpp:188,366,433,438)
 * Each entry point below names the reference interface it stands in for.  Plain C types only: the
 * caller owns every host buffer, the library owns all device memory, every call returns 0 or a
 * negative KMAT_ERR_* (no exit(), no exceptions).  There is NO CPU fallback: calls that need the GPU
 * fail with KMAT_ERR_NO_DEVICE / KMAT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef KMAT_H
#define KMAT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KMAT_ABI_VERSION 1

enum {
    KMAT_OK = 0,
    KMAT_ERR_ARG = -1,         /* bad argument                                                      */
    KMAT_ERR_NO_DEVICE = -2,   /* no CUDA device / driver                                           */
    KMAT_ERR_CUDA = -3,        /* a CUDA call failed (see kmat_last_error)                          */
    KMAT_ERR_NOMEM = -4,
    KMAT_ERR_IO = -5,          /* file missing / unreadable (reference: cerr + return -1)           */
    KMAT_ERR_FORMAT = -6,      /* malformed input file / DB image                                   */
    KMAT_ERR_UNSUPPORTED = -7, /* valid for the reference but outside what this build handles       */
    KMAT_ERR_BAD_TAXID = -8,   /* stored id missing from the -f map: reference asserts (TaxNodeStat.hpp:140-144,235-238) */
    KMAT_ERR_TREE = -9,        /* taxonomy is not a forest (missing parent / cycle): reference exits or hangs (TaxTree.hpp:73-77) */
    KMAT_ERR_OVERFLOW = -10    /* caller-provided output buffer too small; required size is reported */
};

/* ---- opaque handles ------------------------------------------------------------------------- */
typedef struct kmat_table kmat_table;   /* host-side logical table: ascending k-mers + CSR lists of stored ids     */
typedef struct kmat_db kmat_db;         /* device-resident bucketised hash table (one per device / shard)          */
typedef struct kmat_inputs kmat_inputs; /* parsed run-time text inputs (-c -e -w -f -m -r -n)                       */
typedef struct kmat_ctx kmat_ctx;       /* device-resident taxonomy + null models + options, bound to one kmat_db  */

const char *kmat_strerror(int code);
const char *kmat_last_error(void);      /* thread-local detail of the last failure */
int kmat_abi_version(void);
int kmat_device_count(void);            /* 0 when no usable GPU (never an error) */
int kmat_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);   /* HBM of `device` right now */

/* ---- table ingest (host) ---------------------------------------------------------------------
 * Replaces: `perm(&taxtable,..); mopen(db,"r",0)` + the SortedDb members reached through begin_/next
 * (read_label.cpp:1481-1489; layout SortedDb.hpp:143-148,453-481).  A reference-side binding passes
 * the three arrays of its mapped SortedDb object. */
int kmat_table_from_sorteddb(const uint64_t *top_tier_block, uint64_t tt_block_count, int bits_per_2nd,
                             const void *kmer_table /* 8-byte kmer_record[] */, uint64_t n_records,
                             const char *storage_space, uint64_t storage_bytes,
                             int kmer_length, int tid_bytes /* sizeof(DBTID_T): 2 or 4 */, kmat_table **out);
/* From a logical dump: kmers strictly ascending, offs[n+1], ids = stored ids (as next() would yield them). */
int kmat_table_from_arrays(const uint64_t *kmers, const uint64_t *offs, const uint32_t *ids, uint64_t n_kmers,
                           int kmer_length, int tid_bytes, kmat_table **out);
/* Open a DB file: a flat ".kmat" image written by kmat_table_save, or the KMPERM01 heap image that
 * oracle/_ref/make_db_table writes (real perm-je heaps: parity unpinned, KMAT_ERR_FORMAT). */
int kmat_table_open(const char *path, int tid_bytes, kmat_table **out);
int kmat_table_save(const kmat_table *, const char *path);
/* Build the logical table straight from tax_histo files.  Replaces make_db_table main() +
 * SortedDb<tid_T>::add_data (src/make_db_table.cpp:105-433, src/kmerdb/SortedDb.cpp:84-751): files in ascending k-mer
 * order, optional 32->16-bit id map (-f), run of the rank-priority pruning (-g N -m ranks), sorted human k-mer stream
 * (-j) and adaptor k-mer set (-u).  The result holds exactly the lists a reader of the reference-built DB sees through
 * begin_/next, in the same order; save it with kmat_table_save or upload it with kmat_db_upload. */
typedef struct {
    int32_t kmer_length;        /* -k */
    int32_t tax_histo_format;   /* 1 (default); 0 = -h: kmerPrefixCounter records (u32 counts, sanity word every 1000) */
    int32_t tid_cutoff;         /* -g, 0 = no pruning */
    int32_t reserved;
    uint64_t stopper;           /* -q, 0 = none: records 0..stopper of every file */
    const char *map16;          /* -f, NULL: 32-bit ids are stored */
    const char *numrank;        /* -m, only read when tid_cutoff > 0 (make_db_table.cpp:303) */
    const char *human_kmers;    /* -j */
    const char *adaptor_kmers;  /* -u */
} kmat_build_opts;
void kmat_build_opts_default(kmat_build_opts *);
int kmat_table_build(const char *const *files, int n_files, const kmat_build_opts *, kmat_table **out);
uint64_t kmat_table_size(const kmat_table *);        /* SortedDb::size()            (SortedDb.hpp:438) */
int kmat_table_kmer_length(const kmat_table *);      /* SortedDb::get_kmer_length() (SortedDb.hpp:433) */
int kmat_table_tid_bytes(const kmat_table *);
/* Borrowed views for inspection (valid until kmat_table_free). */
int kmat_table_view(const kmat_table *, const uint64_t **kmers, const uint64_t **offs, const uint32_t **ids,
                    uint64_t *n_ids);
void kmat_table_free(kmat_table *);

/* ---- device table ---------------------------------------------------------------------------- */
/* Build the HBM hash table on `device`.  shard_count > 1 keeps only k-mers whose
 * kmat_shard_of(kmer) == shard_index (DB-sharded mode, SURVEY.md 8(e) mode B). */
int kmat_db_upload(const kmat_table *, int device, int shard_index, int shard_count, kmat_db **out);
/* Same, from DEVICE arrays already resident on `device` (used to build very large synthetic tables
 * without a host round trip): payload[i] = stored id for singletons, or (1u<<31 | pool offset in
 * 4-byte words) for lists; pool = the list pool ([count][ids...] records). */
int kmat_db_build_device(int device, int kmer_length, int tid_bytes, uint64_t n_kmers,
                         const uint64_t *d_kmers, const uint32_t *d_payload,
                         const uint32_t *d_pool, uint64_t pool_words, uint32_t n_stored_ids,
                         int shard_index, int shard_count /* keep kmat_shard_of() == shard_index only; 0, 1 = all */,
                         kmat_db **out);
uint32_t kmat_shard_of(uint64_t kmer, int kmer_length, int shard_count);
/* HBM (bytes) one device needs to build and hold one of `shard_count` shards of the table, batch buffers included: lets a
 * host decide between a replicated and a sharded table before it uploads anything. */
uint64_t kmat_table_device_bytes(const kmat_table *, int shard_count);
uint64_t kmat_db_size(const kmat_db *);
uint64_t kmat_db_bytes(const kmat_db *);             /* device bytes held */
uint64_t kmat_db_overflow(const kmat_db *);          /* k-mers of a two-level table that live in its second level */
int kmat_db_kmer_length(const kmat_db *);
int kmat_db_device(const kmat_db *);
void kmat_db_free(kmat_db *);

/* K2 parity hook.  Replaces TaxNodeStat::begin(kmer) + next() loop with no pruning
 * (TaxNodeStat.hpp:41-58,208-256).  Host buffers; hit_off has n+1 entries; ids receives the stored
 * ids of every hit in list order.  On KMAT_ERR_OVERFLOW *n_ids is the capacity needed. */
int kmat_lookup_batch(const kmat_db *, const uint64_t *kmers, uint32_t n, uint64_t *hit_off, uint32_t *ids,
                      uint64_t ids_cap, uint64_t *n_ids);

/* K1 parity hook.  Replaces the rolling encoder of retrieve_kmer_labels (read_label.cpp:978-1017,
 * 1205-1206) for a batch of reads.  kmers/flags are indexed by base offset (offs[r] + p for k-mer
 * start position p; the k-1 tail slots of each read are unused): flags 0 = no valid k-mer, 1 = valid
 * first occurrence, 2 = valid duplicate.  valid_kmers[r], bin_sel[r] per read. */
int kmat_encode_batch(const kmat_db *, const char *bases, const uint64_t *offs, uint32_t n_reads,
                      uint64_t *kmers, uint8_t *flags, int32_t *valid_kmers, int32_t *bin_sel);

/* ---- run-time inputs --------------------------------------------------------------------------
 * Replaces the loaders in read_label main(): TaxTree ctor (-c, TaxTree.hpp:24-57), depth map (-e,
 * :1573-1582), gRank_table (-w, :1560-1567), conv_map (-f, :1585-1602), tid_rank_map (-m, :1543-1559),
 * loadLowNumPlasmids (-r, :499-510), loadRandHits (-n, :512-678; lmat_dir = $LMAT_DIR).  Any path may
 * be NULL (option absent). */
int kmat_inputs_load(const char *tree, const char *depth, const char *rank, const char *conv16, const char *numrank,
                     const char *plasmids, const char *null_list, const char *lmat_dir, kmat_inputs **out);
void kmat_inputs_free(kmat_inputs *);

typedef struct {
    int32_t min_kmer;      /* -j, default 35 (run_rl.sh passes 30)        read_label.cpp:1337,1364 */
    int32_t min_fnd_kmer;  /* -z, default 1                                :1337,1367 */
    float sdiff;           /* -b, ScoreOptions::_diff_thresh, default 1.0  :488,1388  */
    float hbias;           /* -l, ScoreOptions::_diff_thresh2, default 3.0 :488,1391  */
    float min_score;       /* -x, default 0 (host tallies only)            :1336,1373 */
    int32_t max_count;     /* -g, default 65535 = uint16_t(~0)             :1346,1422 */
    int32_t permissive;    /* -s                                           :1382      */
    int32_t phix_screen;   /* default 1; -h clears                         :41,1355   */
    int32_t want_lineage;  /* also return the valid_cand list (printed on MultiMatch without -p, :917-927) */
    int32_t rkmer_mode;    /* 1: the ctx serves kmat_null_* (rand_read_label): src/rkmer.hpp's retrieve_kmer_labels, i.e. NO
                              human collapse (absent at rkmer.hpp:119-121) and no -j / -z thresholds (min_kmer and
                              min_fnd_kmer are forced to 0); kmat_label_batch* then return KMAT_ERR_ARG.  default 0 */
} kmat_opts;
void kmat_opts_default(kmat_opts *);

int kmat_ctx_create(const kmat_db *, const kmat_inputs *, const kmat_opts *, kmat_ctx **out);
int kmat_ctx_set_opts(kmat_ctx *, const kmat_opts *);
void kmat_ctx_destroy(kmat_ctx *);

/* ---- per-read results ------------------------------------------------------------------------- */
enum { KMAT_DIRECT = 0, KMAT_MULTI = 1, KMAT_PARTIAL = 2, KMAT_NOMATCH = 3, KMAT_LCA_ERROR = 4 }; /* match_t, read_label.cpp:202 */
enum {
    KMAT_ST_SHORT_LEN = 0,   /* len < k: "-1 -1 -1\t-1 -1\t<len> <k> ReadTooShort"              :1217-1218 */
    KMAT_ST_SHORT_VALID = 1, /* valid_kmers < -j: "... <valid> <min_kmer> ReadTooShort"          :1232-1233 */
    KMAT_ST_NODBHITS = 2,    /* no taxid: "-1 -1 <valid>\t-1 -1\t<len> <k> NoDbHits"             :1270-1271 */
    KMAT_ST_SILENT = 3,      /* construct_labels early NoMatch: NOTHING is written, tallied NoDbHits :727-733,1248-1253 */
    KMAT_ST_PHIX = 4,        /* PhiX / artificial-sequence bypass                                 :841-848  */
    KMAT_ST_LABELED = 5,     /* normal line                                                        :894-937  */
    KMAT_ST_ERROR = 6        /* this read could not be processed; see err */
};
typedef struct { uint32_t tid; float score; } kmat_pair;
typedef struct {
    int32_t status;          /* KMAT_ST_*                                                    */
    int32_t n1, n2;          /* the two integers of the ReadTooShort / NoDbHits lines        */
    int32_t valid_kmers;     /* retrieve_kmer_labels().first                                  */
    int32_t cand_kmer_cnt;   /* construct_labels: positions with label_vec[pos].first >= 0    */
    int32_t match;           /* KMAT_DIRECT ...                                               */
    uint32_t tid;            /* best_guess.first                                              */
    float score;             /* best_guess.second                                             */
    float log_avg, stdev;    /* first two numbers of the normal line                          */
    uint32_t n_cand;         /* rank_label after sort(TCmp), ascending (printed descending)   */
    uint32_t n_lin;          /* valid_cand list (only when opts.want_lineage)                 */
    uint64_t cand_off;       /* offsets into the cands / lineage output arrays                */
    uint64_t lin_off;
    int32_t bin_sel;         /* GC bin                                                         */
    int32_t err;             /* KMAT_ERR_* for status == KMAT_ST_ERROR, else 0                */
} kmat_read_result;

/* THE hot path.  Replaces one proc_line() call per read (read_label.cpp:1211-1279) for a batch:
 * bases = concatenated reads (any bytes; non-ACGT resets the k-mer run), offs[n+1] byte offsets.
 * Host buffers in, host buffers out; device work runs on an internal stream and the call returns
 * when the results are in `out`.  cands/lineage may be NULL (then n_cand/n_lin are still reported);
 * on KMAT_ERR_OVERFLOW *n_cands / *n_lineage hold the capacities needed and `out` is valid. */
int kmat_label_batch(kmat_ctx *, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                     kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands,
                     kmat_pair *lineage, uint64_t lineage_cap, uint64_t *n_lineage);

/* ---- compact interface: 2-bit packed reads in, 32-byte results out ------------------------------------------------
 * Same labels as kmat_label_batch with a third of the bytes on the PCIe / host-memory path (46 B instead of 158 B in and
 * 32 B + one pair list instead of 64 B + lists out per 150-base read): what the parallel reader hands to a GPU worker and
 * what bench.py's end-to-end leg measures.
 *   codes    : 2 bits per base (a/A 0, c/C 1, g/G 2, t/T 3: the ENCODE macro, read_label.cpp:943-950) indexed by the
 *              base's GLOBAL offset in the batch: word w holds bases 16 w .. 16 w + 15, base i in bits 2 (i % 16)..+1
 *   inv_pos  : ascending offsets of the bases that are none of those eight characters (they reset the k-mer run)
 *   offs     : as for kmat_label_batch (n_reads + 1 base offsets)
 * kmat_pack_reads produces codes / inv_pos from ASCII on `threads` host threads (codes: kmat_pack_words(total) words;
 * KMAT_ERR_OVERFLOW with *n_inv = the count needed when inv_cap is too small). */
typedef struct {
    uint32_t tid; float score, log_avg, stdev;
    uint32_t list_off;       /* first pair of this read's list in `list`                                             */
    uint16_t n_list;         /* rank_label (ctx with want_lineage == 0: the -p line) or valid_cand (want_lineage == 1)  */
    uint16_t valid_kmers, cand_kmer_cnt;
    uint16_t flags;          /* status | match << 3 | bin_sel << 6 | (-err) << 10                                     */
    uint32_t n_cand;
} kmat_read_result32;
uint64_t kmat_pack_words(uint64_t total_bases);
int kmat_pack_reads(const char *bases, uint64_t total_bases, int threads, uint32_t *codes, uint64_t *inv_pos, uint64_t inv_cap, uint64_t *n_inv);
int kmat_label_batch_packed(kmat_ctx *, const uint32_t *codes, const uint64_t *inv_pos, uint64_t n_inv, const uint64_t *offs, uint32_t n_reads,
                            kmat_read_result32 *out, kmat_pair *list, uint64_t list_cap, uint64_t *n_list);
/* The same with the pair list in RUN-LENGTH form: equal scores are adjacent in rank_label / valid_cand, so a taxid word with bit 31
 * set is followed by the score (float bits) that holds for it and for the unflagged taxid words after it -- ~5 instead of 8 bytes
 * per pair on the way out, which is what bounds several GPUs behind one host.  list_off of a record is the offset of the read's
 * words in `words`, n_list its number of pairs; kmat_list_decode rebuilds them (returns the words consumed).  Taxonomies with a
 * taxid >= 2^31: KMAT_ERR_UNSUPPORTED.  KMAT_ERR_OVERFLOW: *n_words = the capacity needed. */
int kmat_label_batch_packed_rl(kmat_ctx *, const uint32_t *codes, const uint64_t *inv_pos, uint64_t n_inv, const uint64_t *offs, uint32_t n_reads,
                               kmat_read_result32 *out, uint32_t *words, uint64_t words_cap, uint64_t *n_words);
uint32_t kmat_list_decode(const uint32_t *words, uint32_t n_list, kmat_pair *out);
/* 32-byte record -> the 64-byte one kmat_format_tail / kmat_tally_class take (the two integers of the ReadTooShort /
 * NoDbHits lines are the read's length, k and -j: read_label.cpp:1217-1218, 1232-1233, 1270-1271) */
void kmat_result_expand(const kmat_read_result32 *in, uint32_t read_len, int kmer_length, int min_kmer, int want_lineage, kmat_read_result *out);

/* ---- DB-sharded mode (SURVEY.md 8(e) mode B: table partitioned by kmat_shard_of over n_shards ranks) ---------------
 * One pass over a batch = three device phases around two exchange steps that the CALLER performs (NCCL all-to-all
 * between one-process-per-GPU ranks -- lmat_b200/sharded.py over torch.distributed -- or peer copies in one process).
 * Every rank calls all three phases every round, with n_reads == 0 when it has no reads left.  Results are identical
 * to the replicated table's.  All pointers named d_* are device pointers on the ctx's device.
 *
 * 1. home: K1 (encode, dedup) and grouping of the first-occurrence k-mers by owner.  *d_queries (library-owned, valid
 *    until the next kmat_shard_encode) holds the mixed k-mers, owner 0's first; counts[o] = how many go to owner o. */
int kmat_shard_encode(kmat_ctx *, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads, uint64_t total_bases,
                      uint32_t max_read_len, int n_shards, const uint64_t **d_queries, uint64_t *counts /* host [n_shards] */,
                      void *stream);
/* 2. owner: probe this rank's shard.  d_queries = the received k-mers, source rank 0's first, counts[s] from source s.
 *    *d_reply: one 32-bit hit word per query, same order (goes back to the sources with the same split);
 *    *d_payload: the resolved list records of the list hits, packed per source; payload_counts[s] = 32-bit words for
 *    source s.  Both buffers are library-owned and valid until the next kmat_shard_serve. */
int kmat_shard_serve(kmat_ctx *, const uint64_t *d_queries, const uint64_t *counts /* host [n_shards] */, int n_shards,
                     const uint32_t **d_reply, const uint32_t **d_payload, uint64_t *payload_counts /* host [n_shards] */,
                     void *stream);
/* 3. home: d_reply = the hit words received for this rank's queries (owner 0's first, i.e. the order of *d_queries),
 *    d_payload = the received list records, owner 0's first, payload_counts[o] words from owner o.  Writes the hit
 *    words back to their read positions and runs K3 / K4; results to d_out (or, if NULL, to a library buffer
 *    readable through kmat_ctx_device_results).  Does not synchronise. */
int kmat_shard_finish(kmat_ctx *, const uint32_t *d_reply, const uint32_t *d_payload, const uint64_t *payload_counts /* host [n_shards] */,
                      int n_shards, kmat_read_result *d_out, void *stream);
/* Device-resident results of the last pass that was given d_out == NULL, and the device candidate pairs that
 * cand_off / n_cand index (rank_label after sort(TCmp), ascending).  Synchronises. */
int kmat_ctx_device_results(kmat_ctx *, const kmat_read_result **d_out, const kmat_pair **d_cands, uint64_t *n_cands);

/* The exchange itself, inside the library: NCCL send/recv groups (libnccl.so.2 is loaded on first use; KMAT_ERR_UNSUPPORTED
 * when it is absent).  One kmat_comm per rank; rank r's ctx must sit on shard r of `world`.  The 128-byte unique id is made
 * by one rank and handed to the others by the caller (a broadcast of the launcher in use, or shared memory between the
 * threads of one process); kmat_comm_init is collective. */
typedef struct kmat_comm kmat_comm;
int kmat_comm_unique_id(unsigned char *id128);
int kmat_comm_init(int device, int rank, int world, const unsigned char *id128, kmat_comm **out);
void kmat_comm_free(kmat_comm *);
/* COLLECTIVE pass over this rank's device-resident reads (possibly none): rounds of `round_reads` reads (0 = 2^20) through
 * kmat_shard_encode -> send/recv -> kmat_shard_serve -> send/recv -> kmat_shard_finish until every rank has finished.
 * h_offs / d_offs = the n_reads + 1 absolute base offsets on the host and on the device; results to d_out[0 .. n_reads).
 * stats (optional, host, 4 entries): unique lookups sent, queries served, list-record words received, rounds. */
int kmat_shard_label_device(kmat_ctx *, kmat_comm *, const char *d_bases, const uint64_t *h_offs, const uint64_t *d_offs, uint32_t n_reads,
                            uint32_t round_reads, kmat_read_result *d_out, uint64_t *stats, void *stream);
/* The same with host buffers in and out: kmat_label_batch for a sharded table.  COLLECTIVE (a rank without reads passes
 * n_reads = 0).  KMAT_ERR_OVERFLOW: the pass has run to its end (the ranks stay in step); *n_cands / *n_lineage hold the
 * capacities needed. */
int kmat_shard_label_batch(kmat_ctx *, kmat_comm *, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                           kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands, kmat_pair *lineage, uint64_t lineage_cap, uint64_t *n_lineage);

/* ---- gene_label -------------------------------------------------------------------------------
 * Replaces retrieve_kmer_labels + the top-gene pick of proc_line in src/gene_label.cpp (:217-301) for a batch of reads
 * against a gene DB (a table of 32-bit gene ids; no id map, no pruning).  Host buffers in and out. */
typedef struct {
    int32_t status;        /* 1 = a line is printed; 0 = no gene hit / read shorter than k (the reference prints nothing);
                              < 0 = KMAT_ERR_* (more than 64 distinct genes in one read: KMAT_ERR_UNSUPPORTED)            */
    uint32_t valid_kmers;  /* cnt: unique canonical k-mers of the read                                       (:245)      */
    uint32_t n_genes;      /* geneid_lst.size()                                                                            */
    uint32_t gene;         /* gsort[0].first after sort(Cmp)                                                 (:297-299)  */
    uint32_t count;        /* gsort[0].second: k-mers of the read that carry this gene                                     */
    float score;           /* (float)count / (float)cnt                                                      (:298)      */
} kmat_gene_result;
int kmat_gene_batch(const kmat_db *, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_gene_result *out);

/* ---- null-model generation: rand_read_label -------------------------------
 * Replaces the body of rand_read_label's OMP loop and its merge phase (src/rand_read_label.cpp:687-735): genRandRead
 * (:85-103), proc_line (:367-397) over rkmer.hpp's retrieve_kmer_labels (rkmer.hpp:76-294), construct_labels (:184-213).
 * The ctx must have been created with opts.rkmer_mode = 1.  Read i of the run (0-based, all "threads" concatenated)
 * belongs to GC bucket i % KMAT_NULL_BUCKETS (:693); per (taxid, bucket) the ctx accumulates on the device the maximum
 * over reads of hits / valid_kmers (float, :195) and the number of reads that hit the taxid. */
#define KMAT_NULL_BUCKETS 10
int kmat_null_reset(kmat_ctx *);
/* Caller-provided reads (host buffers), read r of the batch having run index first_index + r. */
int kmat_null_batch(kmat_ctx *, const char *bases, const uint64_t *offs, uint32_t n_reads, uint64_t first_index);
/* Reads drawn ON THE DEVICE, n_reads of read_len bases with run indices first_index.., from a counter-based generator
 * keyed by (seed, run index): the same (seed, index) gives the same read whatever the batching or the number of GPUs.
 * Distribution as genRandRead: gc_draw uniform in [10 b, 10 b + 9], num_gc = (unsigned)((float)(gc_draw / 100.0) * len)
 * positions chosen uniformly at random hold g/c (fair coin), the others a/t (fair coin). */
int kmat_null_random(kmat_ctx *, uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t read_len);
/* The same reads written to a host buffer (n_reads * read_len bytes) instead of being labeled (test hook). */
int kmat_null_draw_reads(int device, uint64_t seed, uint64_t first_index, uint32_t n_reads, uint32_t read_len, char *bases);
/* Accumulated rows in ascending taxid order = the lines of <ofbase>.rand_lst (:736-755): every taxid hit by at least one
 * read.  tids[cap_rows], max_frac / counts[cap_rows * KMAT_NULL_BUCKETS]; *n_rows receives the number of rows (also
 * when it exceeds cap_rows: KMAT_ERR_OVERFLOW, nothing written).  *reads_error: reads the kernels could not process
 * (more than 64 candidate taxids); they contribute nothing. */
int kmat_null_fetch(kmat_ctx *, uint32_t *tids, float *max_frac, uint64_t *counts, uint32_t cap_rows, uint32_t *n_rows,
                    uint64_t *reads_error);
/* Merge rows of several contexts / GPUs (max of maxima, sum of counts; what :702-735 does across threads) and write the
 * .rand_lst text.  sets = n_sets pointers to (tids, max_frac, counts) triples with n_rows[i] rows each, each ascending. */
int kmat_null_write(const char *path, int n_sets, const uint32_t *const *tids, const float *const *max_frac,
                    const uint64_t *const *counts, const uint32_t *n_rows);

/* ---- DB-sharded mode, direct variant (SURVEY.md 8(e) mode B without exchange rounds) ----------
 * Every rank holds one shard (kmat_db_upload / kmat_db_build_device with shard_index / shard_count) and a ctx over it.
 * kmat_ctx_peer_export describes the rank's shard (CUDA IPC handles of its bucket array, stash and resolved list pool);
 * the ranks exchange these blobs by any means (MPI / torch.distributed all-gather; plain memory inside one process)
 * and each calls kmat_ctx_peer_attach with all of them, indexed by shard.  From then on kmat_label_batch and
 * kmat_label_batch_device of that ctx label reads against the WHOLE table: the probe kernel sends every bucket gather
 * to the owner's memory (NVLink peer reads), list records are read from the owner's pool the same way.  A rank must keep
 * its table and ctx alive until every peer has finished (barrier before kmat_ctx_destroy / kmat_db_free).
 * All shards must have the same geometry and the contexts the same -g / -s options (checked). */
typedef struct { unsigned char opaque[512]; } kmat_peer_info;
int kmat_ctx_peer_export(kmat_ctx *, kmat_peer_info *out);
int kmat_ctx_peer_attach(kmat_ctx *, int n_shards, const kmat_peer_info *all /* [n_shards], entry s = export of shard s */);

/* ---- content_summ: k-mer coverage of the called taxids ---------------------
 * Replaces retrieve_kmer_labels / storeKmers (src/content_summ.cpp:114-160: per read and per k of the -k list, every
 * DISTINCT canonical k-mer of the read counts once for the read's taxid) and the merge + histogram of compKmerCov
 * (:527-571).  `groups[r]` is the caller's dense index (< 2^21) of the taxid read r counts for, KMAT_GROUP_SKIP = the
 * read is left out.  k <= 20, at most 8 values. */
#define KMAT_GROUP_SKIP 0xFFFFFFFFu
typedef struct kmat_kcov kmat_kcov;
int kmat_kcov_create(int device, const int32_t *k_sizes, int n_k, kmat_kcov **out);
int kmat_kcov_add(kmat_kcov *, const char *bases, const uint64_t *offs, const uint32_t *groups, uint32_t n_reads);
int kmat_kcov_finish(kmat_kcov *);     /* merge everything added so far; required before kmat_kcov_query */
/* For (k_sizes[k_index], group): number of distinct k-mers, sum of their counts, and the histogram of the counts --
 * hist_count[i] ascending, hist_n[i] k-mers seen in exactly hist_count[i] reads (what :556-569 prints).  *n_hist receives
 * the number of histogram entries (KMAT_ERR_OVERFLOW if cap > 0 is too small; cap == 0 only queries). */
int kmat_kcov_query(kmat_kcov *, int k_index, uint32_t group, uint64_t *distinct, uint64_t *total, uint32_t *hist_count, uint64_t *hist_n,
                    uint32_t cap, uint32_t *n_hist);
void kmat_kcov_free(kmat_kcov *);

/* Page-locked host memory for the buffers of kmat_label_batch (optional; NULL when no device / out of memory). */
void *kmat_host_alloc(size_t bytes);
void kmat_host_free(void *);

/* Device-resident variant used for kernel-only timing and multi-batch pipelines: inputs already in
 * HBM (d_bases, d_offs on the ctx's device); results stay on the device (d_out) unless NULL.
 * max_read_len bounds the longest read of the batch (sizes the per-warp dedup sets).
 * stream = a cudaStream_t cast to void* (NULL = the ctx's own stream); does not synchronise. */
int kmat_label_batch_device(kmat_ctx *, const char *d_bases, const uint64_t *d_offs, uint32_t n_reads,
                            uint64_t total_bases, uint32_t max_read_len, kmat_read_result *d_out, void *stream);
/* Waits for the ctx's own stream.  KMAT_ERR_OVERFLOW: the candidate buffer of the last device-resident pass was too small
 * (reads in KMAT_ST_ERROR / KMAT_ERR_OVERFLOW); it is enlarged for the next call, run the batch again. */
int kmat_ctx_sync(kmat_ctx *);
/* Statistics of the last batch (device counters read back): unique k-mer lookups issued, hits, list
 * hits, total list ids, and the algorithmic table bytes of SURVEY.md 8(d). */
typedef struct {
    uint64_t lookups, hits, list_hits, list_ids, probe_extra_buckets, algorithmic_bytes;
    uint64_t reads_fast, reads_slow, reads_error;
} kmat_batch_stats;
int kmat_ctx_last_stats(kmat_ctx *, kmat_batch_stats *);
/* Turn the statistics counters on (default) or off for subsequent batches. */
int kmat_ctx_set_stats(kmat_ctx *, int enable);
/* Intra-pass pipeline.  sub_batches > 1: a pass is cut into that many sub-batches (<= 16) and the encode+probe kernel
 * of sub-batch i+1 runs next to the candidate / scoring kernels of sub-batch i on a second stream; 0 or 1: the three
 * kernels run one after the other (default; needed for kmat_ctx_last_kernel_ms' per-kernel split); -1: automatic
 * (8 sub-batches for passes of >= 2^19 short reads).  Results do not depend on the setting.  On B200 the overlap
 * gains nothing for the 150 bp workload (profiles/r01_probe_kernel_notes.md), hence serial by default. */
int kmat_ctx_set_pipeline(kmat_ctx *, int sub_batches);
/* Device time of the three kernels of the last batch (CUDA events on the launching stream): the
 * encode+probe kernel (K1+K2), the candidate-set kernel (K3) and the scoring/LCA kernel (K4).
 * Synchronises on the batch.  In pipelined passes the kernels overlap: probe_ms is then the whole pass, the others 0. */
int kmat_ctx_last_kernel_ms(kmat_ctx *, float *probe_ms, float *cand_ms, float *score_ms);
/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t kmat_launch_count(void);

/* Text after "hdr\tread\t" exactly as the reference writes it (read_label.cpp:1218,1233,1271,
 * 844-848,894-937; floats via ostream<<float == "%g").  prn_all = -p (2: -p together with -y, which also prints the
 * candidates with a negative score, :901).  Returns bytes written
 * (0 for KMAT_ST_SILENT: the reference writes nothing, not even '\n') or <0 if cap is too small. */
int kmat_format_tail(const kmat_read_result *, const kmat_pair *cands, const kmat_pair *lineage, int prn_all,
                     char *buf, size_t cap);

/* kmat_label_batch with those tails formatted ON THE DEVICE (K5): replaces the per-read output formatting of
 * proc_line / construct_labels (read_label.cpp:894-937, 1218, 1233, 1271, 844-848) for a batch; the caller pastes
 * "hdr\tread\t" and the tail together.  text_ref[i] = offset << KMAT_TEXT_LEN_BITS | length of read i's tail inside `text`
 * (the same bytes kmat_format_tail writes), or KMAT_TEXT_ON_HOST for a read the device formatter leaves to
 * kmat_format_tail: a number that prints in exponent notation, a tail longer than 768 bytes, no room left in `text`, a read
 * in KMAT_ST_ERROR.  *n_text = bytes of `text` in use.  text_cap may be anything (0: every read is left to the host). */
#define KMAT_TEXT_LEN_BITS 24
#define KMAT_TEXT_ON_HOST (~0ull)
int kmat_label_batch_text(kmat_ctx *, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                          kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands,
                          kmat_pair *lineage, uint64_t lineage_cap, uint64_t *n_lineage,
                          int prn_all, char *text, uint64_t text_cap, uint64_t *n_text, uint64_t *text_ref);

/* ---- read ingest (host) -------------------------------------------------------------------------
 * Replaces the single-producer FASTA/FASTQ parser of read_label main() (read_label.cpp:1651-1713) and the
 * header substitution of :1728-1732, quirks included: FASTA lines of length <= 1 are ignored and wrapped lines
 * are joined; a read is emitted when the next '>' line (or EOF) arrives; with fastq != 0 the '+'/'-' line
 * emits the read paired with the PREVIOUS record's header and the quality line is skipped; an empty header
 * becomes "unknown_hdr:<n>", n = 1-based ordinal of the read.  path "-" = stdin. */
typedef struct kmat_reader kmat_reader;
typedef struct kmat_read_batch kmat_read_batch;   /* one batch of reads: owned buffers, reusable */
int kmat_reader_open(const char *path, int fastq, kmat_reader **out);
/* Same with `threads` parser threads.  FASTA in a regular file is mapped and cut at header lines into ~8 MB segments
 * that are parsed in parallel and handed out in file order, one segment per kmat_reader_next call (max_reads /
 * max_bases then do not apply); the (header, read) sequence is identical to the sequential reader's.  FASTQ files are
 * cut at '@' lines that verifiably start a record (followed by sequence lines, a '+' / '-' line, one quality line and
 * another '@' line or the end of the file); the reference's pairing of a FASTQ read with the previous record's header
 * is carried across the cuts; a segment of a malformed file that does not end between two records is detected after
 * the fact and the rest of the file is then parsed sequentially from that segment on, so the sequence is identical to
 * the sequential reader's for any input.  stdin and threads <= 1 fall back to the sequential reader. */
int kmat_reader_open_mt(const char *path, int fastq, int threads, kmat_reader **out);
void kmat_reader_close(kmat_reader *);
kmat_read_batch *kmat_read_batch_new(void);
/* The same, with the bases handed out from page-locked memory (kmat_host_alloc): kmat_label_batch then copies them to the
 * device by DMA instead of through the driver's staging buffer.  Falls back to pageable memory when pinning fails. */
kmat_read_batch *kmat_read_batch_new_pinned(void);
void kmat_read_batch_free(kmat_read_batch *);
/* Fill `b` with up to max_reads reads / about max_bases bases (at least one read).  Returns the number of
 * reads (0 = end of input) or a negative KMAT_ERR_*. */
int64_t kmat_reader_next(kmat_reader *, uint32_t max_reads, uint64_t max_bases, kmat_read_batch *b);
/* Borrowed views, valid until the batch is refilled or freed: bases/offs[n+1] as kmat_label_batch takes
 * them, hdrs/hdr_offs[n+1] the headers, first_ordinal the 1-based ordinal of read 0. */
int kmat_read_batch_view(const kmat_read_batch *, const char **bases, const uint64_t **offs, const char **hdrs,
                         const uint64_t **hdr_offs, uint32_t *n_reads, uint64_t *first_ordinal);

/* Per-read tally of proc_line (read_label.cpp:1241-1277): which counter a finished read increments.
 * Returns 0 = counted for (tid, score) [track_taxids / track_tscores], 1 = ReadTooShort, 2 = NoDbHits,
 * 3 = LowScore, -1 = nothing (NaN score). */
int kmat_tally_class(const kmat_read_result *, float min_score, int32_t min_kmer);

/* Test hook for the device formatter (K5): printf("%g") text of n floats as the device writes it, 16 bytes each (NUL padded;
 * first byte 0xFF = a value the device leaves to the host formatter). */
int kmat_test_format_floats(int device, const float *vals, uint32_t n, char *out16);

/* Random-access HBM roofline probe: uniform random `access_bytes`-wide loads
 * (8, 16 or 32) over a `span_bytes` device allocation; returns achieved gathers/s. */
int kmat_gather_bench(int device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                      double *gathers_per_s, double *sector_gbps);

/* The same with the gathered allocation on another GPU (`mem_device`, peer access over NVLink): the ceiling of the direct
 * sharded mode's remote bucket reads.  access_bytes also selects load flavours (see km_gather_kernel). */
int kmat_gather_bench_peer(int device, int mem_device, uint64_t span_bytes, int access_bytes, uint64_t n_gathers, int iters,
                           double *gathers_per_s, double *sector_gbps);

/* cudaLimitMaxL2FetchGranularity hint (32 / 64 / 128) on `device`; bytes <= 0 only queries.  Returns the
 * granularity in effect, or a negative error. */
int kmat_set_l2_fetch_granularity(int device, int bytes);

#ifdef __cplusplus
}
#endif
#endif
