/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * kmat_oracle: a plain-C, single-threaded CPU restatement of the reference's read_label hot path
 * (LivGen/LMAT v1.2.4_2018a, src/read_label.cpp + src/kmerdb/{SortedDb,TaxNodeStat,TaxTree}.hpp).
 * It exists only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can CHECK
 * the CUDA path.  Nothing under lmat_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/golden/ holds the outputs of the unmodified reference binaries (oracle/_ref/*, built by
 * oracle/Makefile; generating scripts tests/golden/make_*golden.py) which this restatement must reproduce byte for byte
 * (tests/test_oracle_golden.py, test_gene_label.py, test_null_model.py).
 *
 * The DB is consumed in the reference's own in-memory layout (SortedDb.hpp:143-148,453-481), so the
 * lookup restated here is the reference's two-level search, not the product's hash table.
 */
#ifndef KMAT_ORACLE_H
#define KMAT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t tid; float score; } kmo_pair;

/* read_label.cpp:202 match_t */
enum { KMO_DIRECT = 0, KMO_MULTI = 1, KMO_PARTIAL = 2, KMO_NOMATCH = 3, KMO_LCA_ERROR = 4 };
/* which of proc_line's output branches a read took (read_label.cpp:1211-1279) */
enum {
    KMO_ST_SHORT_LEN = 0,   /* len < k                          :1217-1218 */
    KMO_ST_SHORT_VALID = 1, /* valid_kmers < -j                 :1232-1233 */
    KMO_ST_NODBHITS = 2,    /* taxid_lst empty                  :1270-1271 */
    KMO_ST_SILENT = 3,      /* construct_labels early return    :727-733 (nothing printed) */
    KMO_ST_PHIX = 4,        /* PhiX bypass                      :841-848 */
    KMO_ST_LABELED = 5      /* normal line                      :894-937 */
};

typedef struct {
    int32_t status;
    int32_t n1, n2;         /* the two integers of the ReadTooShort / NoDbHits lines */
    int32_t valid_kmers;
    int32_t cand_kmer_cnt;
    int32_t match;          /* match_t */
    uint32_t tid;
    float score;
    float log_avg, stdev;
    uint32_t n_cand;        /* sorted rank_label (ascending TCmp order) */
    uint32_t n_lin;         /* valid_cand lineage list (printed without -p on MultiMatch) */
    uint64_t cand_off, lin_off;
    int32_t bin_sel;
    int32_t err;            /* nonzero: the reference would have hit UB/assert on this read */
} kmo_result;

typedef struct {
    int min_kmer;       /* -j, default 35 */
    int min_fnd_kmer;   /* -z, default 1  */
    float sdiff;        /* -b, default 1.0 */
    float hbias;        /* -l, default 3.0 */
    float min_score;    /* -x, default 0 */
    int max_count;      /* -g, default 65535 (uint16_t ~0) */
    int permissive;     /* -s */
    int phix_screen;    /* default 1, -h clears */
    int prn_all;        /* -p */
    int prn_read;       /* default 1, -a clears */
    int rkmer;          /* 1: src/rkmer.hpp's retrieve_kmer_labels (rand_read_label): no human collapse (rkmer.hpp:119-121);
                           set by kmo_null_batch, never by read_label */
} kmo_opts;

typedef struct {
    const uint64_t *top_tier;  /* SortedDb::top_tier_block                     */
    const uint8_t *kmer_table; /* kmer_record[], 8 bytes each                  */
    const uint8_t *storage;    /* SortedDb::m_storage_space                    */
    int kmer_len;              /* 18 or 20 (SortedDb.hpp:190-198)              */
    int tid_bytes;             /* sizeof(DBTID_T): 2 or 4                      */
} kmo_db;

typedef struct kmo_ctx kmo_ctx;

kmo_ctx *kmo_ctx_new(void);
void kmo_ctx_free(kmo_ctx *);
void kmo_default_opts(kmo_opts *);
void kmo_set_opts(kmo_ctx *, const kmo_opts *);
void kmo_set_db(kmo_ctx *, const kmo_db *);
/* taxonomy tree as parsed from the -c file: node ids and their parents (TaxTree.hpp:24-57) */
int kmo_set_tree(kmo_ctx *, uint32_t n, const uint32_t *tid, const uint32_t *parent);
int kmo_set_depth(kmo_ctx *, uint32_t n, const uint32_t *tid, const uint32_t *depth);       /* -e */
/* -w: code 1 = "strain", 2 = "species", 0 = any other rank string */
int kmo_set_ranks(kmo_ctx *, uint32_t n, const uint32_t *tid, const uint8_t *code);
int kmo_set_conv(kmo_ctx *, uint32_t n, const uint32_t *tid16, const uint32_t *tid32);      /* -f */
int kmo_set_prune_ranks(kmo_ctx *, uint32_t n, const uint32_t *tid, const uint32_t *rank);  /* -m */
int kmo_set_plasmids(kmo_ctx *, uint32_t n, const uint32_t *tid);                           /* -r */
/* -n: restates loadRandHits (read_label.cpp:512-678); returns number of model files loaded or <0 */
int kmo_load_null_models(kmo_ctx *, const char *list_path, const char *lmat_dir);

/* Label n reads.  bases/offs: concatenated reads, offs has n+1 entries.  results[n].  Candidate and
 * lineage pairs are appended to growable arrays owned by the ctx; fetch them with kmo_pairs(). */
int kmo_label_batch(kmo_ctx *, const char *bases, const uint64_t *offs, uint32_t n, kmo_result *results);
const kmo_pair *kmo_cands(const kmo_ctx *);
const kmo_pair *kmo_lineage(const kmo_ctx *);

/* K2 hook: look kmers up; for each, hit_off[i+1]-hit_off[i] stored ids (raw, before -f conversion)
 * are written to ids (capacity cap).  Returns total ids or <0 on overflow. */
int64_t kmo_lookup_batch(const kmo_db *, const uint64_t *kmers, uint32_t n, uint64_t *hit_off, uint32_t *ids,
                         uint64_t cap);

/* gene_label.cpp:217-301 for one read against a gene DB: number of unique canonical k-mers (cnt), the gene id std::sort
 * puts first and its k-mer count; returns the number of distinct gene ids hit (0 = the reference prints nothing). */
int kmo_gene_label_read(const kmo_db *, const char *seq, int len, uint32_t *valid_cnt, uint32_t *gene, uint32_t *count);

/* ---- rand_read_label: null-model generation ---------------------------------------------
 * glibc's srand()/rand() (stdlib/random_r.c, TYPE_3: x^31 + x^3 + 1 additive feedback, 310 outputs discarded after
 * seeding) restated, so that the reads the UNMODIFIED reference draws under a fixed time(0) are known here. */
typedef struct { int32_t st[31]; int f, r; } kmo_glibc_rand;
void kmo_srand(kmo_glibc_rand *, unsigned seed);
int kmo_rand(kmo_glibc_rand *);
/* genRandRead (rand_read_label.cpp:85-103) + libstdc++'s std::random_shuffle(first, last) (bits/stl_algo.h: for i in
 * 1..n-1 swap(i, rand() % (i+1))), for reads first..first+n-1 of the OMP loop (:692-699) on ONE thread: read i is drawn
 * for GC bucket i % 10 with range [10*b, 10*b+9] (:673-685).  bases receives n * read_len lower-case letters. */
void kmo_gen_rand_reads(kmo_glibc_rand *, uint64_t first, uint64_t n, int read_len, char *bases);
/* proc_line + construct_labels of rand_read_label.cpp (:367-397, :184-213) for n reads; read i belongs to GC bucket
 * (first_index + i) % 10.  Accumulates per (tid, bucket) the maximum hit fraction and the number of reads; the
 * accumulator lives in the ctx until kmo_null_reset. */
int kmo_null_batch(kmo_ctx *, const char *bases, const uint64_t *offs, uint32_t n, uint64_t first_index);
uint32_t kmo_null_rows(const kmo_ctx *);
/* rows in ascending tid order (the std::map order the .rand_lst file is written in, :745-754): tids[rows],
 * max_frac[rows*10], cnt[rows*10] */
void kmo_null_get(const kmo_ctx *, uint32_t *tids, float *max_frac, int32_t *cnt);
void kmo_null_reset(kmo_ctx *);

/* K1 hook: canonical k-mers of one read exactly as retrieve_kmer_labels walks it
 * (read_label.cpp:978-1017).  out_kmer[p]: canonical k-mer at position p; out_flag[p]: 0 = no valid
 * k-mer ends here, 1 = valid first occurrence, 2 = valid duplicate.  Returns valid_kmers; bin_sel
 * receives the GC bin (read_label.cpp:1205-1206). */
int kmo_encode_read(const char *seq, int len, int k, uint64_t *out_kmer, uint8_t *out_flag, int *bin_sel);

/* Format the text the reference writes after "hdr\tread\t" for one read (may be empty: the silent
 * NoMatch quirk writes nothing, not even a newline).  Returns bytes written (excluding NUL). */
int kmo_format_tail(const kmo_ctx *, const kmo_result *, char *buf, size_t cap);

/* libm-free float log used by the scoring (bit-identical to glibc >= 2.27 logf; see kmat_oracle.c) */
float kmo_logf(float x);

/* libstdc++ std::sort / heap emulation hooks, exported so tests can pin them against g++ */
typedef int (*kmo_less_fn)(const kmo_pair *a, const kmo_pair *b, void *ctx);
void kmo_std_sort(kmo_pair *first, size_t n, kmo_less_fn less, void *ctx);
void kmo_heap_push(kmo_pair *heap, size_t *n, kmo_pair v); /* priority_queue<MyPair>::push, key = tid field */
kmo_pair kmo_heap_pop(kmo_pair *heap, size_t *n);          /* top() + pop() */

#ifdef __cplusplus
}
#endif
#endif
