/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See kmat_oracle.h.
 *
 * Plain-C restatement of the reference read_label hot path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Compiled with -ffp-contract=off: the
 * reference is built for baseline x86-64 (no -march, CMakeLists.txt:139-148), so its float math has
 * no FMA contraction.
 */
#include "kmat_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------------------------------- */
/* small containers                                                                             */
/* ------------------------------------------------------------------------------------------- */
typedef struct { uint32_t *key, *val; uint8_t *used; uint32_t cap, n; } u32map;

static void u32map_init(u32map *m, uint32_t want) {
    uint32_t cap = 16;
    while (cap < want * 2u + 2u) cap <<= 1;
    m->cap = cap; m->n = 0;
    m->key = (uint32_t *)calloc(cap, 4); m->val = (uint32_t *)calloc(cap, 4); m->used = (uint8_t *)calloc(cap, 1);
}
static void u32map_free(u32map *m) { free(m->key); free(m->val); free(m->used); memset(m, 0, sizeof *m); }
static uint32_t u32hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
static int u32map_find(const u32map *m, uint32_t k, uint32_t *out) {
    if (!m->cap) return 0;
    uint32_t i = u32hash(k) & (m->cap - 1);
    while (m->used[i]) { if (m->key[i] == k) { if (out) *out = m->val[i]; return 1; } i = (i + 1) & (m->cap - 1); }
    return 0;
}
static void u32map_put(u32map *m, uint32_t k, uint32_t v) {
    if ((m->n + 1) * 2 > m->cap) {
        u32map b; u32map_init(&b, m->cap);
        for (uint32_t i = 0; i < m->cap; i++) if (m->used[i]) u32map_put(&b, m->key[i], m->val[i]);
        u32map_free(m); *m = b;
    }
    uint32_t i = u32hash(k) & (m->cap - 1);
    while (m->used[i]) { if (m->key[i] == k) { m->val[i] = v; return; } i = (i + 1) & (m->cap - 1); }
    m->used[i] = 1; m->key[i] = k; m->val[i] = v; m->n++;
}

typedef struct { uint32_t *v; size_t n, cap; } u32vec;
static void u32vec_push(u32vec *a, uint32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 16; a->v = (uint32_t *)realloc(a->v, a->cap * 4); }
    a->v[a->n++] = x;
}
typedef struct { kmo_pair *v; size_t n, cap; } pairvec;
static void pairvec_push(pairvec *a, kmo_pair x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (kmo_pair *)realloc(a->v, a->cap * sizeof(kmo_pair)); }
    a->v[a->n++] = x;
}

/* ------------------------------------------------------------------------------------------- */
/* glibc logf, restated.  The reference calls std::log(float) == logf (read_label.cpp:688; the  */
/* oracle binary imports exactly logf@GLIBC_2.27).  glibc >= 2.27 ships the */
/* ARM "optimized routines" logf (sysdeps/ieee754/flt-32/e_logf.c, e_logf_data.c): 16-entry      */
/* table, degree-3 polynomial, evaluated in double.  tests/test_logf.py pins this against the    */
/* host libm (exhaustively when asked).                                                         */
/* ------------------------------------------------------------------------------------------- */
static const double kLogfT[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
static const double kLogfLn2 = 0x1.62e42fefa39efp-1;
static const double kLogfA[3] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2};

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float kmo_logf(float x) {
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
        ix = f2u(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) % 16);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = kLogfT[i][0], logc = kLogfT[i][1];
    double z = (double)u2f(iz);
    double r = z * invc - 1;
    double y0 = logc + (double)k * kLogfLn2;
    double r2 = r * r;
    double y = kLogfA[1] * r + kLogfA[2];
    y = kLogfA[0] * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

/* ------------------------------------------------------------------------------------------- */
/* libstdc++ std::sort (bits/stl_algo.h: __introsort_loop, __final_insertion_sort, threshold 16) */
/* and heap primitives (bits/stl_heap.h), restated so tie order is the reference's.             */
/* ------------------------------------------------------------------------------------------- */
typedef struct { kmo_less_fn less; void *ctx; } cmp_t;
#define LESS(c, a, b) ((c)->less((a), (b), (c)->ctx))

static void ss_swap(kmo_pair *a, kmo_pair *b) { kmo_pair t = *a; *a = *b; *b = t; }

static void ss_push_heap(kmo_pair *first, ptrdiff_t hole, ptrdiff_t top, kmo_pair value, const cmp_t *c) {
    ptrdiff_t parent = (hole - 1) / 2;
    while (hole > top && LESS(c, first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void ss_adjust_heap(kmo_pair *first, ptrdiff_t hole, ptrdiff_t len, kmo_pair value, const cmp_t *c) {
    const ptrdiff_t top = hole;
    ptrdiff_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(c, first + child, first + (child - 1))) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    ss_push_heap(first, hole, top, value, c);
}
static void ss_pop_heap(kmo_pair *first, kmo_pair *last, kmo_pair *result, const cmp_t *c) {
    kmo_pair value = *result;
    *result = *first;
    ss_adjust_heap(first, 0, last - first, value, c);
}
static void ss_make_heap(kmo_pair *first, kmo_pair *last, const cmp_t *c) {
    ptrdiff_t len = last - first;
    if (len < 2) return;
    ptrdiff_t parent = (len - 2) / 2;
    for (;;) {
        kmo_pair value = first[parent];
        ss_adjust_heap(first, parent, len, value, c);
        if (parent == 0) return;
        parent--;
    }
}
static void ss_partial_sort_all(kmo_pair *first, kmo_pair *last, const cmp_t *c) {
    /* std::__partial_sort(first, last, last): heap_select over an empty tail, then sort_heap */
    ss_make_heap(first, last, c);
    while (last - first > 1) { --last; ss_pop_heap(first, last, last, c); }
}
static void ss_move_median_to_first(kmo_pair *result, kmo_pair *a, kmo_pair *b, kmo_pair *cc, const cmp_t *c) {
    if (LESS(c, a, b)) {
        if (LESS(c, b, cc)) ss_swap(result, b);
        else if (LESS(c, a, cc)) ss_swap(result, cc);
        else ss_swap(result, a);
    } else if (LESS(c, a, cc)) ss_swap(result, a);
    else if (LESS(c, b, cc)) ss_swap(result, cc);
    else ss_swap(result, b);
}
static kmo_pair *ss_unguarded_partition(kmo_pair *first, kmo_pair *last, kmo_pair *pivot, const cmp_t *c) {
    for (;;) {
        while (LESS(c, first, pivot)) ++first;
        --last;
        while (LESS(c, pivot, last)) --last;
        if (!(first < last)) return first;
        ss_swap(first, last);
        ++first;
    }
}
static void ss_introsort_loop(kmo_pair *first, kmo_pair *last, long depth_limit, const cmp_t *c) {
    while (last - first > 16) {
        if (depth_limit == 0) { ss_partial_sort_all(first, last, c); return; }
        --depth_limit;
        kmo_pair *mid = first + (last - first) / 2;
        ss_move_median_to_first(first, first + 1, mid, last - 1, c);
        kmo_pair *cut = ss_unguarded_partition(first + 1, last, first, c);
        ss_introsort_loop(cut, last, depth_limit, c);
        last = cut;
    }
}
static void ss_unguarded_linear_insert(kmo_pair *last, const cmp_t *c) {
    kmo_pair val = *last;
    kmo_pair *next = last - 1;
    while (LESS(c, &val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void ss_insertion_sort(kmo_pair *first, kmo_pair *last, const cmp_t *c) {
    if (first == last) return;
    for (kmo_pair *i = first + 1; i != last; ++i) {
        if (LESS(c, i, first)) {
            kmo_pair val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(kmo_pair));
            *first = val;
        } else ss_unguarded_linear_insert(i, c);
    }
}
void kmo_std_sort(kmo_pair *first, size_t n, kmo_less_fn less, void *ctx) {
    cmp_t c = {less, ctx};
    kmo_pair *last = first + n;
    if (first == last) return;
    long lg = 0;
    for (size_t t = n; t > 1; t >>= 1) lg++;
    ss_introsort_loop(first, last, lg * 2, &c);
    if (last - first > 16) {
        ss_insertion_sort(first, first + 16, &c);
        for (kmo_pair *i = first + 16; i != last; ++i) ss_unguarded_linear_insert(i, &c);
    } else ss_insertion_sort(first, last, &c);
}

/* priority_queue<MyPair> (SortedDb.hpp:128-139: operator< compares .first == the rank number).
 * Stored here as kmo_pair{tid = MyPair.second, score = (float)rank is NOT used}; ranks are kept in a
 * parallel encoding: we pack rank into .score's bits to stay within one element type. */
static int rank_less(const kmo_pair *a, const kmo_pair *b, void *ctx) {
    (void)ctx;
    return f2u(a->score) < f2u(b->score);
}
void kmo_heap_push(kmo_pair *heap, size_t *n, kmo_pair v) {
    cmp_t c = {rank_less, NULL};
    heap[*n] = v; (*n)++;
    ss_push_heap(heap, (ptrdiff_t)*n - 1, 0, heap[*n - 1], &c);
}
kmo_pair kmo_heap_pop(kmo_pair *heap, size_t *n) {
    cmp_t c = {rank_less, NULL};
    kmo_pair top = heap[0];
    if (*n > 1) ss_pop_heap(heap, heap + *n - 1, heap + *n - 1, &c);
    (*n)--;
    return top;
}

/* ------------------------------------------------------------------------------------------- */
/* context                                                                                      */
/* ------------------------------------------------------------------------------------------- */
#define KMO_MAX_CLASSES 256
typedef struct {
    int kmer_cnt;                  /* key in _rand_hits (read_label.cpp:571) */
    int nbins;
    int loaded;                    /* file existed: maps were created */
    u32map row;                    /* tid -> row index */
    float *cut;                    /* rows x nbins */
    uint16_t *cls;                 /* rows, index into ctx->class_names */
    uint32_t nrows, caprows;
} null_model;

struct kmo_ctx {
    kmo_db db;
    kmo_opts opt;
    u32map tree;                   /* tid -> parent tid */
    int has_tree;
    u32map depth;                  /* sopt._imap */
    u32map rank;                   /* gRank_table: code 1 strain, 2 species, 0 other */
    u32map conv; int has_conv;     /* conv_map[tid16] = tid32 */
    u32map prune; int has_prune;   /* tid_rank_map */
    u32map plasmid;                /* gLowNumPlasmid */
    /* null models */
    null_model *models; int n_models;
    int *read_len_vec; int n_len;  /* starts as {0} (read_label.cpp:60) */
    int *read_len_avgs; int n_avg; /* starts as {0} (:61) */
    int models_loaded;             /* loadRandHits ran (gRank2num populated) */
    char *class_names[KMO_MAX_CLASSES]; int n_classes;
    pairvec cands, lineage;
    /* rand_read_label accumulators (max_match / match_cnt, rand_read_label.cpp:182-183) */
    u32map null_row; u32vec null_tids; float *null_max; int32_t *null_cnt; uint32_t null_cap;
};

static int class_id(kmo_ctx *c, const char *s) {
    for (int i = 0; i < c->n_classes; i++) if (strcmp(c->class_names[i], s) == 0) return i;
    if (c->n_classes >= KMO_MAX_CLASSES) return -1;
    c->class_names[c->n_classes] = strdup(s);
    return c->n_classes++;
}
/* gRank2num (read_label.cpp:519-532); unknown strings get 0 through operator[] (:786) */
static int rank2num(const char *s) {
    static const char *names[] = {"no_rank", "ethnic", "region", "species", "genus", "family", "order", "class",
                                  "phylum", "kingdom", "depth=0"};
    static const int nums[] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
    for (int i = 0; i < 11; i++) if (strcmp(s, names[i]) == 0) return nums[i];
    return 0;
}
/* gNum2rank (read_label.cpp:534-547): insert() does not overwrite, so 0 stays "no_rank" */
static const char *num2rank(int n) {
    static const char *names[] = {"no_rank", "region", "species", "genus", "family", "order", "class", "phylum",
                                  "kingdom", "depth=0"};
    return (n >= 0 && n < 10) ? names[n] : "";
}

kmo_ctx *kmo_ctx_new(void) {
    kmo_ctx *c = (kmo_ctx *)calloc(1, sizeof *c);
    kmo_default_opts(&c->opt);
    c->read_len_vec = (int *)calloc(1, sizeof(int)); c->n_len = 1;
    c->read_len_avgs = (int *)calloc(1, sizeof(int)); c->n_avg = 1;
    for (int i = 0; i < 10; i++) class_id(c, num2rank(i));   /* the gNum2rank keys track[] can default-insert */
    return c;
}
void kmo_ctx_free(kmo_ctx *c) {
    if (!c) return;
    u32map_free(&c->tree); u32map_free(&c->depth); u32map_free(&c->rank); u32map_free(&c->conv);
    u32map_free(&c->prune); u32map_free(&c->plasmid);
    for (int i = 0; i < c->n_models; i++) { u32map_free(&c->models[i].row); free(c->models[i].cut); free(c->models[i].cls); }
    free(c->models); free(c->read_len_vec); free(c->read_len_avgs);
    for (int i = 0; i < c->n_classes; i++) free(c->class_names[i]);
    free(c->cands.v); free(c->lineage.v);
    u32map_free(&c->null_row); free(c->null_tids.v); free(c->null_max); free(c->null_cnt);
    free(c);
}
void kmo_default_opts(kmo_opts *o) {
    /* read_label.cpp:1336-1347 and ScoreOptions ctor :488 */
    o->min_kmer = 35; o->min_fnd_kmer = 1; o->sdiff = 1.0f; o->hbias = 3.0f; o->min_score = 0.0f;
    o->max_count = 65535; o->permissive = 0; o->phix_screen = 1; o->prn_all = 0; o->prn_read = 1; o->rkmer = 0;
}
void kmo_set_opts(kmo_ctx *c, const kmo_opts *o) { c->opt = *o; }
void kmo_set_db(kmo_ctx *c, const kmo_db *d) { c->db = *d; }
static int fill_map(u32map *m, uint32_t n, const uint32_t *k, const uint32_t *v) {
    u32map_free(m); u32map_init(m, n);
    for (uint32_t i = 0; i < n; i++) u32map_put(m, k[i], v[i]);   /* later entries overwrite, like operator[]= */
    return 0;
}
int kmo_set_tree(kmo_ctx *c, uint32_t n, const uint32_t *tid, const uint32_t *parent) { c->has_tree = 1; return fill_map(&c->tree, n, tid, parent); }
int kmo_set_depth(kmo_ctx *c, uint32_t n, const uint32_t *tid, const uint32_t *d) { return fill_map(&c->depth, n, tid, d); }
int kmo_set_ranks(kmo_ctx *c, uint32_t n, const uint32_t *tid, const uint8_t *code) {
    /* gRank_table.insert(make_pair(tid,rank)) -- insert keeps the FIRST entry for a tid (read_label.cpp:1565) */
    u32map_free(&c->rank); u32map_init(&c->rank, n);
    for (uint32_t i = 0; i < n; i++) if (!u32map_find(&c->rank, tid[i], NULL)) u32map_put(&c->rank, tid[i], code[i]);
    return 0;
}
int kmo_set_conv(kmo_ctx *c, uint32_t n, const uint32_t *t16, const uint32_t *t32) { c->has_conv = 1; return fill_map(&c->conv, n, t16, t32); }
int kmo_set_prune_ranks(kmo_ctx *c, uint32_t n, const uint32_t *tid, const uint32_t *r) { c->has_prune = n > 0; return fill_map(&c->prune, n, tid, r); }
int kmo_set_plasmids(kmo_ctx *c, uint32_t n, const uint32_t *tid) {
    u32map_free(&c->plasmid); u32map_init(&c->plasmid, n);
    for (uint32_t i = 0; i < n; i++) u32map_put(&c->plasmid, tid[i], 1);
    return 0;
}
const kmo_pair *kmo_cands(const kmo_ctx *c) { return c->cands.v; }
const kmo_pair *kmo_lineage(const kmo_ctx *c) { return c->lineage.v; }

/* ------------------------------------------------------------------------------------------- */
/* loadRandHits -- read_label.cpp:512-678                                                        */
/* ------------------------------------------------------------------------------------------- */
static int int_cmp(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

int kmo_load_null_models(kmo_ctx *c, const char *list_path, const char *lmat_dir) {
    FILE *fl = fopen(list_path, "r");
    if (!fl) return -1;                               /* :514-517 (reference just returns) */
    c->models_loaded = 1;
    int read_len; char fname[4096];
    int loaded = 0;
    while (fscanf(fl, "%d %4095s", &read_len, fname) == 2) {           /* :553 */
        char path[8192];
        if (lmat_dir) snprintf(path, sizeof path, "%s/%s", lmat_dir, fname);   /* :555-558 */
        else snprintf(path, sizeof path, "%s", fname);
        c->read_len_vec = (int *)realloc(c->read_len_vec, (size_t)(c->n_len + 1) * sizeof(int));
        c->read_len_vec[c->n_len++] = read_len;                        /* :562 (before the existence check) */
        FILE *pre = fopen(path, "r");
        if (!pre) continue;                                            /* :564-568 */
        fclose(pre);
        gzFile gz = gzopen(path, "rb");
        if (!gz) continue;
        /* rand_hits_all[read_len] : operator[] -> one map per distinct key; a repeated key appends to it */
        null_model *m = NULL;
        for (int i = 0; i < c->n_models; i++) if (c->models[i].kmer_cnt == read_len) m = &c->models[i];
        if (!m) {
            c->models = (null_model *)realloc(c->models, (size_t)(c->n_models + 1) * sizeof(null_model));
            m = &c->models[c->n_models++];
            memset(m, 0, sizeof *m);
            m->kmer_cnt = read_len;
            u32map_init(&m->row, 1024);
        }
        m->loaded = 1;
        static char buff[20004];                                       /* :577-578 */
        if (!gzgets(gz, buff, sizeof buff)) { gzclose(gz); continue; }
        int num_bins = atoi(buff);                                     /* :579-583 */
        if (num_bins <= 0) { gzclose(gz); return -2; }
        if (m->nbins && m->nbins != num_bins) { gzclose(gz); return -3; }
        m->nbins = num_bins;
        float *save_ecoli = (float *)malloc((size_t)num_bins * sizeof(float));
        for (int b = 0; b < num_bins; b++) save_ecoli[b] = 0.5f;       /* :584 */
        float *cutoff = (float *)malloc((size_t)num_bins * sizeof(float));
        unsigned *revisit = (unsigned *)malloc((size_t)num_bins * sizeof(unsigned));
        while (gzgets(gz, buff, sizeof buff)) {                        /* :585 */
            char *save = NULL;
            char *tok = strtok_r(buff, " \t\r\n", &save);
            if (!tok) continue;
            uint32_t taxid = (uint32_t)strtoul(tok, NULL, 10);
            tok = strtok_r(NULL, " \t\r\n", &save);
            if (!tok) continue;
            char val[256];
            const char *dash = strchr(tok, '-');                       /* :591-593 */
            if (!dash) { free(save_ecoli); free(cutoff); free(revisit); gzclose(gz); return -4; }
            size_t vl = (size_t)(dash - tok); if (vl > 255) vl = 255;
            memcpy(val, tok, vl); val[vl] = 0;
            if (vl >= 3 && val[0] == 'n' && val[1] == 'o' && val[2] == '_') strcpy(val, "genus");   /* :594-601 */
            int nrev = 0;
            float max_val = 0;
            for (int b = 0; b < num_bins; b++) cutoff[b] = 0;           /* :603 */
            for (int bin = 0; bin < num_bins; ++bin) {                 /* :604-630 */
                int num_obs = 0, kmer_cnt = 0;
                char *t1 = strtok_r(NULL, " \t\r\n", &save), *t2 = strtok_r(NULL, " \t\r\n", &save),
                     *t3 = strtok_r(NULL, " \t\r\n", &save);
                if (t1) num_obs = atoi(t1);
                if (t2) max_val = strtof(t2, NULL);
                if (t3) kmer_cnt = atoi(t3);
                if (num_obs == 0 && kmer_cnt >= 100000) { max_val = 0.5f; cutoff[bin] = max_val; }
                else if (num_obs == 0 && kmer_cnt < 100000) revisit[nrev++] = (unsigned)bin;
                if (num_obs > 0) {
                    cutoff[bin] = max_val;
                    if (taxid == 562) save_ecoli[bin] = cutoff[bin];
                }
                if (taxid == 28384) { strcpy(val, "genus"); memcpy(cutoff, save_ecoli, (size_t)num_bins * sizeof(float)); }
            }
            for (int r = 0; r < nrev; r++) {                           /* :631-665 */
                unsigned it = revisit[r];
                int j = (int)it - 1;
                unsigned i = it + 1;
                while (j >= 0 || i < (unsigned)num_bins) {
                    float a_val = 0.0f, b_val = 0.0f;
                    if (j >= 0) a_val = cutoff[j];
                    if (i < (unsigned)num_bins) b_val = cutoff[i];
                    if (a_val > 0 && b_val > 0) cutoff[it] = a_val > b_val ? a_val : b_val;   /* std::max(a,b) */
                    else if (a_val > 0) cutoff[it] = a_val;
                    else if (b_val > 0) cutoff[it] = b_val;
                    if (cutoff[it] > 0) break;
                    --j; ++i;
                }
                if (cutoff[it] <= 0) cutoff[it] = 0.5f;
            }
            int cid = class_id(c, val);
            if (cid < 0) { free(save_ecoli); free(cutoff); free(revisit); gzclose(gz); return -5; }
            uint32_t row;
            if (!u32map_find(&m->row, taxid, &row)) {                  /* :666-667 operator[]= : last line wins */
                if (m->nrows == m->caprows) {
                    m->caprows = m->caprows ? m->caprows * 2 : 1024;
                    m->cut = (float *)realloc(m->cut, (size_t)m->caprows * (size_t)num_bins * sizeof(float));
                    m->cls = (uint16_t *)realloc(m->cls, (size_t)m->caprows * sizeof(uint16_t));
                }
                row = m->nrows++;
                u32map_put(&m->row, taxid, row);
            }
            memcpy(m->cut + (size_t)row * (size_t)num_bins, cutoff, (size_t)num_bins * sizeof(float));
            m->cls[row] = (uint16_t)cid;
        }
        free(save_ecoli); free(cutoff); free(revisit);
        gzclose(gz);
        loaded++;
    }
    fclose(fl);
    qsort(c->read_len_vec, (size_t)c->n_len, sizeof(int), int_cmp);    /* :672 */
    free(c->read_len_avgs);
    c->n_avg = c->n_len - 1;                                           /* :674-677 */
    c->read_len_avgs = (int *)calloc((size_t)(c->n_avg > 0 ? c->n_avg : 1), sizeof(int));
    for (int i = 1; i < c->n_len; i++) c->read_len_avgs[i - 1] = (c->read_len_vec[i - 1] + c->read_len_vec[i]) / 2;
    return loaded;
}

/* closest / getReadLen -- read_label.cpp:107-133.  When the loop falls through, the reference reads
 * read_len_vec[avgs.size()]: the last element once models are loaded; one PAST the end (UB) when no
 * -n was given.  In that case _rand_hits is empty and the value cannot matter; we return -1. */
static int closest(const kmo_ctx *c, int value) {
    int i;
    for (i = 0; i < c->n_avg; i++) if (value <= c->read_len_avgs[i]) return c->read_len_vec[i];
    return i < c->n_len ? c->read_len_vec[i] : -1;
}
static int getReadLen(const kmo_ctx *c, int rl) { int len = closest(c, rl); return len > 0 ? len : 80; }

/* ------------------------------------------------------------------------------------------- */
/* tree helpers                                                                                 */
/* ------------------------------------------------------------------------------------------- */
/* TaxTree::getPathToRoot -- TaxTree.hpp:60-91.  Returns length; *err set where the reference would
 * exit(-1) (missing parent on the first hop) or dereference end() (later hops). */
static size_t path_to_root(const kmo_ctx *c, uint32_t tid, u32vec *out, int *err) {
    out->n = 0;
    uint32_t parent;
    if (!u32map_find(&c->tree, tid, &parent)) return 0;          /* registerFailure: empty path */
    uint32_t cur = tid;
    size_t guard = 0;
    while (parent != cur) {
        uint32_t pp;
        if (!u32map_find(&c->tree, parent, &pp)) { if (err) *err = 1; return out->n; }
        cur = parent; parent = pp;
        u32vec_push(out, cur);
        if (++guard > 100000) { if (err) *err = 2; return out->n; }   /* cycle: the reference would not terminate */
    }
    return out->n;
}
/* isAncestor -- read_label.cpp:138-150 */
static int is_ancestor(const kmo_ctx *c, uint32_t anc, uint32_t desc, u32vec *tmp, int *err) {
    path_to_root(c, desc, tmp, err);
    for (size_t p = 0; p < tmp->n; p++) if (tmp->v[p] == anc) return 1;
    return 0;
}
/* (*dmap.find(tid)).second with libstdc++: a missing key dereferences end(), whose "value" overlays
 * the tree header's node count -> .second reads 0 (verified against g++ 13 in this image). */
static uint32_t depth_of(const kmo_ctx *c, uint32_t tid) { uint32_t d = 0; u32map_find(&c->depth, tid, &d); return d; }
static int is_human(uint32_t t) { return t == 9606 || t == 63221 || t == 741158; }   /* tid_checks.hpp:15-28 */
static int is_phix(uint32_t t) { return t == 374840 || t == 10847 || t == 32630; }   /* tid_checks.hpp:13 */
static int bad_genome(uint32_t t) { return t == 12721 || t == 693660; }              /* read_label.cpp:82-104 */
static int is_plasmid(const kmo_ctx *c, uint32_t t) {                                /* read_label.cpp:69 */
    return (t >= 10000000u && t < 11000000u) || u32map_find(&c->plasmid, t, NULL);
}
static int rank_code(const kmo_ctx *c, uint32_t t) { uint32_t r = 0; return u32map_find(&c->rank, t, &r) ? (int)r : 0; }

/* ------------------------------------------------------------------------------------------- */
/* SortedDb::begin_18 / begin_20 / next -- SortedDb.hpp:202-385                                  */
/* ------------------------------------------------------------------------------------------- */
#define KMO_PAGE_SIZE 4294701056ull
#define KMO_MAX_PAGE 255
typedef struct { uint16_t count; uint32_t offset; uint8_t page; int found; int err; } dbcur;

static dbcur db_begin(const kmo_db *db, uint64_t kmer) {
    dbcur cur; memset(&cur, 0, sizeof cur);
    int bits; uint64_t mask;
    if (db->kmer_len == 20) { bits = 13; mask = 0x1fff; }          /* :41-44 */
    else if (db->kmer_len == 18) { bits = 9; mask = 0x1ff; }       /* :36-39 */
    else { cur.err = 1; return cur; }                              /* :195-197 assert(0) */
    uint64_t tt = db->top_tier[kmer >> bits];
    if (tt == 0) return cur;
    uint16_t k_count = (uint16_t)(tt >> 48);
    uint64_t koff = tt & 0x0000ffffffffffffull;
    const uint8_t *recs = db->kmer_table + koff * 8;
    uint16_t want = (uint16_t)(kmer & mask);
    const uint8_t *hit = NULL;
    if (k_count == 1) {
        uint16_t lsb; memcpy(&lsb, recs, 2);
        if (lsb != want) return cur;
        hit = recs;
    } else {
        /* bsearch with kmer_rec_comp = lsb difference (SortedDb.cpp:18-25); lsb values are unique per bucket */
        size_t lo = 0, hi = k_count;
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            uint16_t lsb; memcpy(&lsb, recs + mid * 8, 2);
            int d = (int)want - (int)lsb;
            if (d == 0) { hit = recs + mid * 8; break; }
            if (d < 0) hi = mid; else lo = mid + 1;
        }
        if (!hit) return cur;
    }
    uint16_t page_id; uint32_t page_off;
    memcpy(&page_id, hit + 2, 2); memcpy(&page_off, hit + 4, 4);
    cur.page = (uint8_t)page_id;                                   /* narrowed to uint8_t page_out (:305) */
    cur.offset = page_off;
    cur.found = 1;
    if (cur.page == KMO_MAX_PAGE) cur.count = 1;                   /* :326-327 */
    else {
        if (kmer % 4096 == 0) {                                    /* :331-337 self-check echo */
            uint64_t km; memcpy(&km, db->storage + KMO_PAGE_SIZE * cur.page + cur.offset, 8);
            if (km != kmer) cur.err = 2;
            cur.offset += 8;
        }
        memcpy(&cur.count, db->storage + KMO_PAGE_SIZE * cur.page + cur.offset, 2);
        cur.offset += 2;
    }
    return cur;
}
static uint32_t db_next(const kmo_db *db, dbcur *cur) {            /* :366-385 */
    if (cur->page == KMO_MAX_PAGE) {
        /* taxid_out = offset_out_in, narrowed to tid_T */
        return db->tid_bytes == 2 ? (uint32_t)(uint16_t)cur->offset : cur->offset;
    }
    uint32_t v = 0;
    memcpy(&v, db->storage + KMO_PAGE_SIZE * cur->page + cur->offset, (size_t)db->tid_bytes);
    cur->offset += (uint32_t)db->tid_bytes;
    return v;
}

int64_t kmo_lookup_batch(const kmo_db *db, const uint64_t *kmers, uint32_t n, uint64_t *hit_off, uint32_t *ids,
                         uint64_t cap) {
    uint64_t w = 0;
    hit_off[0] = 0;
    for (uint32_t i = 0; i < n; i++) {
        dbcur cur = db_begin(db, kmers[i]);
        if (cur.err) return -2;
        if (cur.found) {
            for (unsigned j = 0; j < cur.count; j++) {
                if (w >= cap) return -1;
                ids[w++] = db_next(db, &cur);
            }
        }
        hit_off[i + 1] = w;
    }
    return (int64_t)w;
}

/* ------------------------------------------------------------------------------------------- */
/* gene_label: retrieve_kmer_labels + the top-gene pick of proc_line -- gene_label.cpp:217-301     */
/* Unique canonical k-mers of the read (first occurrence wins, :242-245); every id of every hit     */
/* list counts once per k-mer, ids in first-appearance order (:249-258); std::sort by count         */
/* descending (Cmp, :84-88; libstdc++ order on ties) and the front element is the call (:297-299). */
/* Returns the number of distinct gene ids (0: the reference prints nothing for the read).        */
/* ------------------------------------------------------------------------------------------- */
static int gene_cmp_less(const kmo_pair *a, const kmo_pair *b, void *ctx) { (void)ctx; return a->score > b->score; }
int kmo_gene_label_read(const kmo_db *db, const char *seq, int len, uint32_t *valid_cnt, uint32_t *gene, uint32_t *count) {
    const int k = db->kmer_len;
    *valid_cnt = 0; *gene = 0; *count = 0;
    if (len < k) return 0;
    const int np = len - k + 1;
    uint64_t *km = (uint64_t *)malloc((size_t)np * 8);
    uint8_t *fl = (uint8_t *)malloc((size_t)np);
    int bin;
    kmo_encode_read(seq, len, k, km, fl, &bin);
    kmo_pair *g = NULL; size_t ng = 0, cap = 0;
    uint32_t ids[65536];
    for (int p = 0; p < np; p++) {
        if (fl[p] != 1) continue;
        (*valid_cnt)++;
        uint64_t ho[2];
        const int64_t n = kmo_lookup_batch(db, &km[p], 1, ho, ids, 65536);
        for (int64_t j = 0; j < n; j++) {
            size_t q = 0;
            while (q < ng && g[q].tid != ids[j]) q++;
            if (q == ng) {
                if (ng == cap) { cap = cap ? cap * 2 : 16; g = (kmo_pair *)realloc(g, cap * sizeof *g); }
                g[ng].tid = ids[j]; g[ng].score = 1.0f; ng++;
            } else g[q].score += 1.0f;
        }
    }
    if (ng) { kmo_std_sort(g, ng, gene_cmp_less, NULL); *gene = g[0].tid; *count = (uint32_t)g[0].score; }
    free(g); free(km); free(fl);
    return (int)ng;
}

/* ------------------------------------------------------------------------------------------- */
/* TaxNodeStat::begin / next -- TaxNodeStat.hpp:60-256                                           */
/* Produces the tid sequence next() would hand out (32-bit ids) and taxidCount().              */
/* ------------------------------------------------------------------------------------------- */
typedef struct { u32vec tids; uint16_t count; int err; int printed_nl; } tns_out;

static uint32_t conv_tid(const kmo_ctx *c, uint32_t stored, int *err) {
    if (!c->has_conv) return stored;
    uint32_t t = 0;
    if (!u32map_find(&c->conv, stored, &t) || t == 0) { *err = 3; return 0; }   /* "bad taxid" assert (:140-144,235-238) */
    return t;
}

static void tns_run(const kmo_ctx *c, uint64_t kmer, tns_out *o, pairvec *heap) {
    o->tids.n = 0; o->count = 0; o->err = 0; o->printed_nl = 0;
    dbcur cur = db_begin(&c->db, kmer);
    if (cur.err) { o->err = cur.err; return; }
    if (!cur.found) return;                                        /* :70-73 */
    int tid_cut = c->opt.max_count;
    uint16_t m_taxid_count = cur.count;
    size_t hn = 0;
    if (tid_cut > 0 && m_taxid_count > tid_cut) {                  /* :76 */
        if (!c->has_prune) {                                       /* :78-81 p_map.size()==0 */
            m_taxid_count = 1;
        } else {                                                   /* :118-201 (strainspecies is never set, read_label.cpp:80) */
            heap->n = 0;
            for (int i = 0; i < m_taxid_count; i++) {
                uint32_t stored = db_next(&c->db, &cur);
                uint32_t tid = conv_tid(c, stored, &o->err);
                if (o->err) return;
                uint32_t r = 0;
                u32map_find(&c->prune, tid, &r);                   /* p_map[m_taxid]: missing -> 0 */
                kmo_pair pp; pp.tid = tid; pp.score = u2f(r);
                if (heap->n + 1 > heap->cap) { heap->cap = heap->cap ? heap->cap * 2 : 64; heap->v = (kmo_pair *)realloc(heap->v, heap->cap * sizeof(kmo_pair)); }
                size_t n = heap->n; kmo_heap_push(heap->v, &n, pp); heap->n = n;
            }
            hn = heap->n;
            while (hn > 0) {                                       /* :161-190 */
                uint32_t cur_priority = f2u(heap->v[0].score);
                while (f2u(heap->v[0].score) == cur_priority) {
                    kmo_heap_pop(heap->v, &hn);
                    if (hn == 0) break;
                }
                if (hn <= (size_t)tid_cut) { m_taxid_count = (uint16_t)hn; break; }
            }
            o->printed_nl = 1;                                     /* cout << "\n" (:192) */
            if (hn == 0) {                                         /* :193-199 */
                m_taxid_count = 1;
                kmo_pair pp; pp.tid = 1; pp.score = u2f(1);
                kmo_heap_push(heap->v, &hn, pp);
            }
        }
    }
    o->count = m_taxid_count;
    for (unsigned calls = 0; calls < m_taxid_count; calls++) {     /* next(): :208-256 */
        uint32_t tid;
        if (hn > 0) tid = kmo_heap_pop(heap->v, &hn).tid;
        else {
            uint32_t stored = db_next(&c->db, &cur);
            tid = conv_tid(c, stored, &o->err);
            if (o->err) return;
        }
        u32vec_push(&o->tids, tid);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* K1: the rolling encoder of retrieve_kmer_labels -- read_label.cpp:943-950, 978-1017, 1205-1206 */
/* ------------------------------------------------------------------------------------------- */
static int encode_base(char ch) {
    switch (ch) {
        case 'a': case 'A': return 0;
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 2;
        case 't': case 'T': return 3;
        default: return -1;
    }
}
/* small open-addressing u64 set for the per-read no_dups set */
typedef struct { uint64_t *k; uint8_t *u; size_t cap, n; } u64set;
static void u64set_reset(u64set *s, size_t want) {
    size_t cap = 64; while (cap < want * 2 + 2) cap <<= 1;
    if (cap > s->cap) { free(s->k); free(s->u); s->k = (uint64_t *)malloc(cap * 8); s->u = (uint8_t *)malloc(cap); s->cap = cap; }
    memset(s->u, 0, s->cap); s->n = 0;
}
static int u64set_insert(u64set *s, uint64_t k) {   /* returns 1 if newly inserted */
    uint64_t h = k * 0x9E3779B97F4A7C15ull; size_t i = (size_t)(h >> 20) & (s->cap - 1);
    while (s->u[i]) { if (s->k[i] == k) return 0; i = (i + 1) & (s->cap - 1); }
    s->u[i] = 1; s->k[i] = k; s->n++; return 1;
}

int kmo_encode_read(const char *seq, int len, int k, uint64_t *out_kmer, uint8_t *out_flag, int *bin_sel) {
    int np = len - k + 1;
    for (int p = 0; p < np; p++) { out_kmer[p] = 0; out_flag[p] = 0; }
    int kk = 0, highbits = (k - 1) * 2;
    uint64_t mask = k * 2 >= 64 ? ~0ull : (((uint64_t)1 << (k * 2)) - 1);
    uint64_t forward = 0, reverse = 0;
    int valid = 0, gc = 0, tot = 0, vgc = 0, vtot = 0;
    u64set seen; memset(&seen, 0, sizeof seen); u64set_reset(&seen, (size_t)(np > 0 ? np : 1));
    for (int j = 0; j < len; j++) {
        int t = encode_base(seq[j]);
        if (t < 0) { kk = 0; gc = 0; tot = 0; continue; }
        forward = ((forward << 2) | (uint64_t)t) & mask;
        reverse = ((uint64_t)(t ^ 3) << highbits) | (reverse >> 2);
        if (t == 1 || t == 2) { ++gc; ++tot; } else ++tot;
        if (++kk >= k) {
            valid++; vgc += gc; vtot += tot; gc = 0; tot = 0;
            uint64_t km = forward < reverse ? forward : reverse;
            int pos = j - k + 1;
            out_kmer[pos] = km;
            out_flag[pos] = u64set_insert(&seen, km) ? 1 : 2;
        }
    }
    free(seen.k); free(seen.u);
    if (bin_sel) {
        float gc_pcnt = (float)(((double)((float)vgc / (float)vtot)) * 100.0);   /* :1205 float*double -> float */
        *bin_sel = (int)(gc_pcnt / 10);                                           /* :1206 */
    }
    return valid;
}

/* ------------------------------------------------------------------------------------------- */
/* per-read state                                                                               */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int16_t first; u32vec set; } label_info;   /* label_info_t; set kept sorted ascending (std::set) */

static int set_insert(u32vec *s, uint32_t t) {   /* std::set<pair<tid,1>>::insert; returns 1 if new */
    size_t lo = 0, hi = s->n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (s->v[mid] < t) lo = mid + 1; else hi = mid; }
    if (lo < s->n && s->v[lo] == t) return 0;
    u32vec_push(s, 0);
    memmove(s->v + lo + 1, s->v + lo, (s->n - 1 - lo) * 4);
    s->v[lo] = t;
    return 1;
}
static int set_has(const u32vec *s, uint32_t t) {
    size_t lo = 0, hi = s->n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (s->v[mid] < t) lo = mid + 1; else hi = mid; }
    return lo < s->n && s->v[lo] == t;
}

typedef struct {
    label_info *label; int n_label, cap_label;
    u32vec taxid_lst;            /* first-appearance order */
    u32map tax2idx;
    u32map leaf_track;
    u32vec leaf_keys;
    u32vec path, path2, obs;
    pairvec sortbuf, heap;
    tns_out tns;
    u64set seen;
    int err;
} read_state;

static void add_taxid(read_state *rs, uint32_t tid) {     /* read_label.cpp:1117-1122 and :1192-1198 */
    if (!u32map_find(&rs->tax2idx, tid, NULL)) {
        u32map_put(&rs->tax2idx, tid, (uint32_t)rs->taxid_lst.n);
        u32vec_push(&rs->taxid_lst, tid);
    }
}
static int depth_desc_less(const kmo_pair *a, const kmo_pair *b, void *ctx) {   /* CmpDepth / CmpDepth1 :159-177 */
    const kmo_ctx *c = (const kmo_ctx *)ctx;
    return (int)depth_of(c, a->tid) > (int)depth_of(c, b->tid);
}
static int u32_cmp(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return (x > y) - (x < y); }

/* retrieve_kmer_labels -- read_label.cpp:974-1209 */
static void retrieve_kmer_labels(kmo_ctx *c, read_state *rs, const char *str, int slen, int klen, int *valid_out, int *bin_out) {
    int k = 0, highbits = (klen - 1) * 2;
    uint64_t mask = (((uint64_t)1 << (klen * 2)) - 1);
    uint64_t forward = 0, reverse = 0;
    int valid_kmers = 0, gc_cnt = 0, valid_gc_cnt = 0, valid_tot_cnt = 0, tot_cnt = 0;
    u64set_reset(&rs->seen, (size_t)rs->n_label);
    u32map_free(&rs->leaf_track); u32map_init(&rs->leaf_track, 64); rs->leaf_keys.n = 0;
    for (int j = 0; j < slen; j++) {
        int t = encode_base(str[j]);
        if (t < 0) { k = 0; gc_cnt = 0; tot_cnt = 0; continue; }                 /* ENCODE :943-950 */
        forward = ((forward << 2) | (uint64_t)t) & mask;                         /* :992 */
        reverse = ((uint64_t)(t ^ 3) << highbits) | (reverse >> 2);               /* :993 */
        if (t == 1 || t == 2) { ++gc_cnt; ++tot_cnt; } else ++tot_cnt;             /* :994-999 */
        if (++k >= klen) {                                                        /* :1002 */
            valid_kmers++; valid_gc_cnt += gc_cnt; valid_tot_cnt += tot_cnt; gc_cnt = 0; tot_cnt = 0;
            uint64_t kmer_id = forward < reverse ? forward : reverse;            /* :1009 */
            const int pos = j - klen + 1;
            if (c->opt.rkmer) rs->label[pos].first = 0;                           /* rkmer.hpp:104, before the dup check :106 */
            if (!u64set_insert(&rs->seen, kmer_id)) continue;                     /* :1010,1017 */
            rs->label[pos].first = 0;                                             /* :1015 */
            tns_run(c, kmer_id, &rs->tns, &rs->heap);                             /* :1019-1026 */
            if (rs->tns.err) { rs->err = rs->tns.err; return; }
            unsigned dcnt = 0;
            int seenHuman = 0;
            rs->obs.n = 0;
            for (size_t q = 0; q < rs->tns.tids.n; q++) {                         /* while(h->next()) :1031-1066 */
                uint32_t tid = rs->tns.tids.v[q];
                if (!c->opt.rkmer) {                                              /* absent from rkmer.hpp:119-121 */
                    if (is_human(tid) && seenHuman) continue;
                    else if (is_human(tid) && !seenHuman) { tid = 9606; seenHuman = 1; }
                }
                if (tid == 20999999u || bad_genome(tid)) continue;
                uint16_t ng = rs->tns.count;
                if (dcnt == 0) { if (ng <= 0) ng = 1; rs->label[pos].first = (int16_t)ng; }   /* :1040-1046 */
                u32vec_push(&rs->obs, tid);
                if (c->opt.permissive) { set_insert(&rs->label[pos].set, tid); add_taxid(rs, tid); }   /* :1050-1058 */
                dcnt++;
            }
            rs->sortbuf.n = 0;
            for (size_t q = 0; q < rs->obs.n; q++) { kmo_pair p; p.tid = rs->obs.v[q]; p.score = 0; pairvec_push(&rs->sortbuf, p); }
            kmo_std_sort(rs->sortbuf.v, rs->sortbuf.n, depth_desc_less, c);       /* :1073-1074 */
            if (c->opt.permissive) {                                              /* :1075-1102 */
                int last_depth = -1;
                for (size_t i = 0; i < rs->sortbuf.n; i++) {
                    uint32_t tid = rs->sortbuf.v[i].tid;
                    int depth = (int)depth_of(c, tid);
                    if (depth == 0) break;
                    if (last_depth == depth || last_depth == -1) {
                        path_to_root(c, tid, &rs->path, &rs->err);
                        for (size_t p = 0; p < rs->path.n; p++) { set_insert(&rs->label[pos].set, rs->path.v[p]); add_taxid(rs, rs->path.v[p]); }
                    } else break;
                    /* NB: the reference never updates last_depth (it stays -1), restated as is */
                }
            } else {                                                              /* :1103-1134 */
                /* non_leaf: unordered_set of every ancestor of a kept tid */
                rs->path2.n = 0;   /* flat non_leaf list */
                for (size_t i = 0; i < rs->sortbuf.n; i++) {
                    uint32_t tid = rs->sortbuf.v[i].tid;
                    int in_nl = 0;
                    for (size_t q = 0; q < rs->path2.n; q++) if (rs->path2.v[q] == tid) { in_nl = 1; break; }
                    if (in_nl) continue;
                    set_insert(&rs->label[pos].set, tid);                         /* :1111 */
                    uint32_t cnt;
                    if (u32map_find(&rs->leaf_track, tid, &cnt)) u32map_put(&rs->leaf_track, tid, cnt + 1);   /* :1112-1116 */
                    else { u32map_put(&rs->leaf_track, tid, 1); u32vec_push(&rs->leaf_keys, tid); }
                    add_taxid(rs, tid);                                           /* :1117-1122 */
                    path_to_root(c, tid, &rs->path, &rs->err);                    /* :1123-1129 */
                    for (size_t p = 0; p < rs->path.n; p++) u32vec_push(&rs->path2, rs->path.v[p]);
                }
            }
        }
    }
    if (!c->opt.permissive) {                                                     /* :1143-1204 */
        /* save_spec_rep: species -> (strain, count); leaf_track iterated in ascending tid order (std::map) */
        qsort(rs->leaf_keys.v, rs->leaf_keys.n, 4, u32_cmp);
        u32vec sp_keys = {0}, sp_strain = {0}, sp_cnt = {0};
        for (size_t q = 0; q < rs->leaf_keys.n; q++) {
            uint32_t stid = rs->leaf_keys.v[q], stid_cnt = 0;
            u32map_find(&rs->leaf_track, stid, &stid_cnt);
            if (rank_code(c, stid) == 1) {                                        /* "strain" :1150 */
                path_to_root(c, stid, &rs->path, &rs->err);
                for (size_t p = 0; p < rs->path.n; p++) {
                    uint32_t ptid = rs->path.v[p];
                    if (rank_code(c, ptid) == 2) {                                /* "species" :1156 */
                        size_t f = sp_keys.n;
                        for (size_t z = 0; z < sp_keys.n; z++) if (sp_keys.v[z] == ptid) { f = z; break; }
                        if (f == sp_keys.n) { u32vec_push(&sp_keys, ptid); u32vec_push(&sp_strain, stid); u32vec_push(&sp_cnt, stid_cnt); }
                        else if (stid_cnt > sp_cnt.v[f]) { sp_strain.v[f] = stid; sp_cnt.v[f] = stid_cnt; }
                        break;
                    }
                }
            }
        }
        for (int pos = 0; pos < rs->n_label; pos++) {                             /* :1178-1203 */
            if (rs->label[pos].first < 0) continue;
            /* iterate the std::set in ascending order while inserting into it: an element inserted after the
             * cursor is visited later; one inserted before it is not.  Iterate by value to restate that. */
            u32vec *s = &rs->label[pos].set;
            size_t idx = 0;
            while (idx < s->n) {
                uint32_t tid = s->v[idx];
                int is_rep = 0;
                for (size_t z = 0; z < sp_strain.n; z++) if (sp_strain.v[z] == tid) { is_rep = 1; break; }
                if (is_rep || rank_code(c, tid) != 1) {                           /* :1184 */
                    path_to_root(c, tid, &rs->path, &rs->err);
                    for (size_t p = 0; p < rs->path.n; p++) { set_insert(s, rs->path.v[p]); add_taxid(rs, rs->path.v[p]); }
                }
                /* advance to the successor of tid in the (possibly grown) set */
                size_t lo = 0, hi = s->n;
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (s->v[mid] <= tid) lo = mid + 1; else hi = mid; }
                idx = lo;
            }
        }
        free(sp_keys.v); free(sp_strain.v); free(sp_cnt.v);
    }
    float gc_pcnt = (float)(((double)((float)valid_gc_cnt / (float)valid_tot_cnt)) * 100.0);   /* :1205 */
    int bin_sel = (int)(gc_pcnt / 10);                                                           /* :1206 */
    *valid_out = valid_kmers; *bin_out = bin_sel;
}

/* ------------------------------------------------------------------------------------------- */
/* findReadLabelVer2 and helpers -- read_label.cpp:225-419                                       */
/* ------------------------------------------------------------------------------------------- */
static int map_depth_checked(const kmo_ctx *c, uint32_t tid) { uint32_t d = 0; u32map_find(&c->depth, tid, &d); return (int)d; }

static int addToCandLineage(const kmo_ctx *c, kmo_pair cand, pairvec *lineage, u32vec *tmp, int *err) {   /* :225-262 */
    int addNode = 0;
    if (lineage->n == 0) addNode = 1;
    else {
        unsigned cand_depth = (unsigned)map_depth_checked(c, cand.tid);
        addNode = 1;
        for (size_t i = 0; i < lineage->n; i++) {
            uint32_t taxid = lineage->v[i].tid;
            unsigned chk_depth = (unsigned)map_depth_checked(c, taxid);
            if (chk_depth > cand_depth && !is_ancestor(c, cand.tid, taxid, tmp, err)) { addNode = 0; break; }
            else if (chk_depth < cand_depth && !is_ancestor(c, taxid, cand.tid, tmp, err)) { addNode = 0; break; }
            else if (chk_depth == cand_depth) { addNode = 0; break; }
        }
    }
    if (addNode) pairvec_push(lineage, cand);
    return addNode;
}
static int cmpCompLineage(const kmo_ctx *c, kmo_pair cand, const pairvec *lineage, u32vec *no_good, float diff_thresh,
                          u32vec *tmp, int *err) {                                                          /* :264-282 */
    const float undef = -10000;
    int keep_going = 1;
    for (size_t i = 0; i < lineage->n; ++i) {
        if (is_ancestor(c, lineage->v[i].tid, cand.tid, tmp, err)) break;
        if (lineage->v[i].score != undef && (lineage->v[i].score - cand.score) > diff_thresh) { keep_going = 0; break; }
        if ((lineage->v[i].score - cand.score) <= diff_thresh) set_insert(no_good, lineage->v[i].tid);
    }
    return keep_going;
}

typedef struct { kmo_pair call; int match; } label_res;

static label_res findReadLabelVer2(kmo_ctx *c, const pairvec *rank_label, float diff_thresh, pairvec *cand_lin,
                                   const pairvec *all_cand /* unsorted (tid,score) in taxid_lst order */, float topScore,
                                   read_state *rs) {
    label_res out; out.match = KMO_NOMATCH; out.call.tid = 0; out.call.score = 0;
    uint32_t savePlasmidId = 0; int plasmidTopHit = 0;
    unsigned lowest_depth = 0, highest_depth = 0;
    kmo_pair lowest = {0, 0}, highest = {0, 0};
    int lidx = -1, linDone = 0;
    int n = (int)rank_label->n;
    cand_lin->n = 0;
    for (int i = n - 1; i >= 0; --i) {                                              /* :295-325 */
        kmo_pair rl = rank_label->v[i];
        if (rl.score >= topScore && is_plasmid(c, rl.tid)) { plasmidTopHit = 1; savePlasmidId = rl.tid; }
        if (!linDone && !addToCandLineage(c, rl, cand_lin, &rs->path, &rs->err)) { lidx = i; linDone = 1; }
        else if (!linDone) {
            unsigned d = depth_of(c, rl.tid);
            if (d > lowest_depth || i == n - 1) { lowest = rl; lowest_depth = d; }
            if (d < highest_depth || i == n - 1) { highest = rl; highest_depth = d; }
        }
        if (linDone && rl.score < topScore) break;
    }
    u32vec add_set = {0};
    if (highest_depth != 0) {                                                       /* :327-343 */
        u32vec path = {0};
        path_to_root(c, highest.tid, &path, &rs->err);
        for (size_t i = 0; i < path.n; ++i) {
            set_insert(&add_set, path.v[i]);
            kmo_pair val; val.tid = path.v[i]; val.score = -10000;
            for (size_t q = 0; q < all_cand->n; q++) if (all_cand->v[q].tid == path.v[i]) { val.score = all_cand->v[q].score; break; }
            pairvec_push(cand_lin, val);
        }
        free(path.v);
    }
    pairvec cand_lin_vec = {0};
    for (size_t i = 0; i < cand_lin->n; i++) pairvec_push(&cand_lin_vec, cand_lin->v[i]);
    kmo_std_sort(cand_lin_vec.v, cand_lin_vec.n, depth_desc_less, c);                /* :350-351 */
    u32vec no_good = {0};
    for (int i = lidx; i >= 0; --i) {                                               /* :355-362 */
        if (!set_has(&add_set, rank_label->v[i].tid)) {
            if (!cmpCompLineage(c, rank_label->v[i], &cand_lin_vec, &no_good, diff_thresh, &rs->path, &rs->err)) break;
        }
    }
    if (cand_lin->n == 0 && no_good.n == 0) out.match = KMO_NOMATCH;                 /* :364-365 */
    else if (cand_lin->n != 0 && no_good.n == 0) { out.call = lowest; out.match = KMO_DIRECT; }   /* :366-368 */
    else {                                                                          /* :369-409 */
        /* cand_vec = cand_lin sorted by depth desc: the same input and comparator as cand_lin_vec */
        float max_val = -10000;
        int found = 0; int root_idx = -1; uint32_t lca_tid = 0;
        for (size_t i = 0; i < cand_lin_vec.n; ++i) {
            max_val = cand_lin_vec.v[i].score > max_val ? cand_lin_vec.v[i].score : max_val;   /* std::max(a,b): b<a?a... see note */
            if (!set_has(&no_good, cand_lin_vec.v[i].tid)) { lca_tid = cand_lin_vec.v[i].tid; found = 1; root_idx = (int)i; break; }
        }
        if (!found) { out.call.tid = 0; out.call.score = -1; out.match = KMO_LCA_ERROR; }
        else {
            out.match = KMO_MULTI;
            int in_all = 0;
            for (size_t q = 0; q < all_cand->n; q++) if (all_cand->v[q].tid == lca_tid) { in_all = 1; break; }
            if (in_all && max_val < cand_lin_vec.v[root_idx].score) { out.match = KMO_PARTIAL; max_val = cand_lin_vec.v[root_idx].score; }
            out.call.tid = lca_tid; out.call.score = max_val;
        }
    }
    if (plasmidTopHit) {                                                            /* :410-416 */
        if (is_ancestor(c, out.call.tid, savePlasmidId, &rs->path, &rs->err)) out.call.tid = savePlasmidId;
    }
    free(add_set.v); free(cand_lin_vec.v); free(no_good.v);
    return out;
}

/* TCmp -- read_label.cpp:475-485 */
static int tcmp_less(const kmo_pair *a, const kmo_pair *b, void *ctx) {
    const kmo_ctx *c = (const kmo_ctx *)ctx;
    if ((double)fabsf(a->score - b->score) < 0.001) return (int)depth_of(c, a->tid) < (int)depth_of(c, b->tid);
    return a->score < b->score;
}

/* construct_labels -- read_label.cpp:692-941.  Returns status; fills res. */
static void construct_labels(kmo_ctx *c, read_state *rs, int bin_sel, kmo_result *res) {
    const unsigned num_tax_ids = (unsigned)rs->taxid_lst.n;
    unsigned cnt_fnd_kmers = 0;
    uint16_t cand_kmer_cnt = 0;
    for (int pos = 0; pos < rs->n_label; ++pos) {                                   /* :701-726 */
        if (rs->label[pos].first >= 0) ++cand_kmer_cnt;
        if (rs->label[pos].set.n) ++cnt_fnd_kmers;
    }
    res->cand_kmer_cnt = cand_kmer_cnt;
    if ((int)cnt_fnd_kmers < c->opt.min_fnd_kmer || cand_kmer_cnt < c->opt.min_kmer) {   /* :727-733 */
        res->status = KMO_ST_SILENT; res->match = KMO_NOMATCH; res->tid = 0; res->score = -1;
        return;
    }
    const int cand_kmer_cnt_match = getReadLen(c, cand_kmer_cnt);                   /* :736 */
    const null_model *mod = NULL;
    for (int i = 0; i < c->n_models; i++) if (c->models[i].loaded && c->models[i].kmer_cnt == cand_kmer_cnt_match) mod = &c->models[i];
    const int useRandMod = mod != NULL;                                             /* :738-739 */
    float *rank_first = (float *)calloc(num_tax_ids ? num_tax_ids : 1, sizeof(float));
    /* track: unordered_map<string,float>, keyed by class string */
    float track_val[KMO_MAX_CLASSES + 16]; uint8_t track_has[KMO_MAX_CLASSES + 16];
    memset(track_has, 0, sizeof track_has);
    int hasHuman = 0;
    for (unsigned tax_idx = 0; tax_idx < num_tax_ids; ++tax_idx) {                  /* :748-802 */
        float found_genome_cnt = 0;
        const uint32_t taxid = rs->taxid_lst.v[tax_idx];
        if (is_human(taxid)) hasHuman = 1;
        for (int pos = 0; pos < rs->n_label; ++pos) if (set_has(&rs->label[pos].set, taxid)) found_genome_cnt += 1;
        rank_first[tax_idx] = (float)found_genome_cnt / (float)cand_kmer_cnt;       /* :761 */
        float random_prob = 0.5f;
        uint32_t row = 0; int has_row = 0;
        if (!useRandMod) random_prob = 0.1f;
        else if ((has_row = u32map_find(&mod->row, taxid, &row))) {
            /* val_vec[bin_sel]; bin_sel == nbins (GC = 100 %) reads one past the vector (:770, UB).  With
             * glibc malloc that word is the next chunk's size field, a denormal float; we use 0. */
            const float val = (bin_sel >= 0 && bin_sel < mod->nbins) ? mod->cut[(size_t)row * (size_t)mod->nbins + (size_t)bin_sel] : 0.0f;
            if (bin_sel < 0 || bin_sel > mod->nbins) res->err = 4;
            random_prob = (float)((double)val + 0.0001);                           /* :771 */
        } else random_prob = 1.0f;                                                  /* :773-774 */
        if (useRandMod) {
            if (!has_row) { res->err = 5; free(rank_first); return; }               /* assert(chk != equiv_class.end()) :778 */
            const char *cval = c->class_names[mod->cls[row]];
            /* class ids for the gNum2rank strings may not be interned yet */
            int cid = class_id(c, cval);
            if (!track_has[cid]) { track_has[cid] = 1; track_val[cid] = random_prob; }   /* :780-782 */
            else track_val[cid] = random_prob > track_val[cid] ? random_prob : track_val[cid];   /* :793 std::max(random_prob,track) */
            const int cval_rank = rank2num(cval);                                   /* :786 / :794 */
            for (int ti = cval_rank - 1; ti >= 0; ti--) {                           /* :787-790 / :795-798 */
                int lid = class_id(c, num2rank(ti));
                if (!track_has[lid]) { track_has[lid] = 1; track_val[lid] = 0; }    /* operator[] default-inserts 0 */
                track_val[cid] = track_val[cid] < track_val[lid] ? track_val[lid] : track_val[cid];   /* std::max(track[cval],track[lower]) */
            }
        }
    }
    pairvec rank_label = {0}, all_cand = {0};
    int fndPhiX = 0;
    float log_sum = 0.0f, pos_log_sum = 0.0f, top_score = 0.0f, phiXscore = 0.0f;
    unsigned sig_hits = 0, pos_sig_hits = 0;
    for (unsigned tax_idx = 0; tax_idx < num_tax_ids; ++tax_idx) {                  /* :807-837 */
        const uint32_t taxid = rs->taxid_lst.v[tax_idx];
        const float label_prob = rank_first[tax_idx];
        float log_odds = label_prob;
        if (useRandMod) {
            uint32_t row = 0; u32map_find(&mod->row, taxid, &row);
            float random_prob = track_val[class_id(c, c->class_names[mod->cls[row]])];
            const float denom = random_prob <= 0 ? (float)0.00001 : random_prob;    /* :687 */
            log_odds = kmo_logf(label_prob / denom);                               /* :688 */
        }
        kmo_pair p; p.tid = taxid; p.score = log_odds;
        pairvec_push(&rank_label, p); pairvec_push(&all_cand, p);
        log_sum += log_odds; sig_hits++;
        if (log_odds > 0) { pos_sig_hits++; pos_log_sum += log_odds; }
        if (c->opt.phix_screen && is_phix(taxid)) { phiXscore = log_odds; fndPhiX = 1; }
        if (tax_idx == 0 || log_odds > top_score) top_score = log_odds;
    }
    free(rank_first);
    res->cand_off = c->cands.n; res->n_cand = 0; res->lin_off = c->lineage.n; res->n_lin = 0;
    if (c->opt.phix_screen && phiXscore >= top_score && fndPhiX) {                  /* :841-848 */
        res->status = KMO_ST_PHIX; res->match = KMO_DIRECT; res->tid = 32630; res->score = phiXscore;
    } else {
        float log_avg;
        unsigned use_sig_hits;
        const unsigned min_pos_examples = 3;
        if (pos_sig_hits > min_pos_examples) { use_sig_hits = pos_sig_hits; log_avg = pos_log_sum / (float)pos_sig_hits; }
        else { use_sig_hits = sig_hits; log_avg = sig_hits > 0 ? log_sum / (float)sig_hits : 0; }
        float log_std = 0;
        for (unsigned t = 0; t < num_tax_ids; ++t) {                                /* :865-880 */
            if (rank_label.v[t].score > 0) {
                if (pos_sig_hits > min_pos_examples) { const float val = log_avg - rank_label.v[t].score; log_std += (val * val); }
            }
            if (pos_sig_hits <= min_pos_examples) { const float val = log_avg - rank_label.v[t].score; log_std += (val * val); }
        }
        float stdev1 = use_sig_hits > 1 ? sqrtf(log_std / (float)(use_sig_hits - 1)) : 0;   /* :881 */
        res->status = KMO_ST_LABELED;
        res->log_avg = log_avg; res->stdev = stdev1;
        label_res lr; lr.match = KMO_NOMATCH; lr.call.tid = 0; lr.call.score = 0;
        pairvec valid_cand = {0};
        if (use_sig_hits > 0) {                                                     /* :882-912 */
            if (hasHuman) for (unsigned t = 0; t < num_tax_ids; ++t) if (is_human(rank_label.v[t].tid)) rank_label.v[t].score += (c->opt.hbias * stdev1);
            kmo_std_sort(rank_label.v, rank_label.n, tcmp_less, c);                 /* :893 */
            float thresh = stdev1 * c->opt.sdiff;                                   /* :895 */
            lr = findReadLabelVer2(c, &rank_label, thresh, &valid_cand, &all_cand, top_score, rs);
            for (size_t i = 0; i < rank_label.n; i++) pairvec_push(&c->cands, rank_label.v[i]);
            res->n_cand = (uint32_t)rank_label.n;
        }
        for (size_t i = 0; i < valid_cand.n; i++) pairvec_push(&c->lineage, valid_cand.v[i]);
        res->n_lin = (uint32_t)valid_cand.n;
        free(valid_cand.v);
        res->match = lr.match;
        if (lr.match == KMO_DIRECT || lr.match == KMO_MULTI || lr.match == KMO_PARTIAL) { res->tid = lr.call.tid; res->score = lr.call.score; }
        else { res->tid = 0; res->score = 0; }                                       /* best_guess stays (0,0) :839 */
    }
    free(rank_label.v); free(all_cand.v);
}

/* proc_line -- read_label.cpp:1211-1279 */
static void rs_free(read_state *rs) {
    for (int p = 0; p < rs->cap_label; p++) free(rs->label[p].set.v);
    free(rs->label); free(rs->taxid_lst.v); u32map_free(&rs->tax2idx); u32map_free(&rs->leaf_track); free(rs->leaf_keys.v);
    free(rs->path.v); free(rs->path2.v); free(rs->obs.v); free(rs->sortbuf.v); free(rs->heap.v); free(rs->tns.tids.v);
    free(rs->seen.k); free(rs->seen.u);
}
static void proc_read(kmo_ctx *c, read_state *rs, const char *seq, int len, kmo_result *res) {
    memset(res, 0, sizeof *res);
    const int k = c->db.kmer_len;
    res->cand_off = c->cands.n; res->lin_off = c->lineage.n;
    if (len < k) { res->status = KMO_ST_SHORT_LEN; res->n1 = len; res->n2 = k; res->match = KMO_NOMATCH; return; }
    int np = len - k + 1;
    if (np > rs->cap_label) {
        rs->label = (label_info *)realloc(rs->label, (size_t)np * sizeof(label_info));
        memset(rs->label + rs->cap_label, 0, (size_t)(np - rs->cap_label) * sizeof(label_info));
        rs->cap_label = np;
    }
    rs->n_label = np;
    for (int p = 0; p < np; p++) { rs->label[p].first = -1; rs->label[p].set.n = 0; }
    rs->taxid_lst.n = 0;
    u32map_free(&rs->tax2idx); u32map_init(&rs->tax2idx, 64);
    rs->err = 0;
    int valid = 0, bin_sel = 0;
    retrieve_kmer_labels(c, rs, seq, len, k, &valid, &bin_sel);
    res->valid_kmers = valid; res->bin_sel = bin_sel;
    if (rs->err) { res->err = rs->err; return; }
    if (valid < c->opt.min_kmer) { res->status = KMO_ST_SHORT_VALID; res->n1 = valid; res->n2 = c->opt.min_kmer; res->match = KMO_NOMATCH; return; }
    if (rs->taxid_lst.n == 0) { res->status = KMO_ST_NODBHITS; res->n1 = len; res->n2 = k; res->match = KMO_NOMATCH; return; }
    construct_labels(c, rs, bin_sel, res);
    if (rs->err && !res->err) res->err = rs->err;
}

int kmo_label_batch(kmo_ctx *c, const char *bases, const uint64_t *offs, uint32_t n, kmo_result *results) {
    read_state rs; memset(&rs, 0, sizeof rs);
    c->cands.n = 0; c->lineage.n = 0;
    for (uint32_t i = 0; i < n; i++) proc_read(c, &rs, bases + offs[i], (int)(offs[i + 1] - offs[i]), &results[i]);
    rs_free(&rs);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* rand_read_label (null-model generation) -- src/rand_read_label.cpp, src/rkmer.hpp               */
/* ------------------------------------------------------------------------------------------- */
/* glibc stdlib/random_r.c: __srandom_r / __random_r for the default TYPE_3 state (degree 31, separation 3) */
void kmo_srand(kmo_glibc_rand *g, unsigned seed) {
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    g->st[0] = word;
    for (int i = 1; i < 31; i++) {                 /* 16807 * word % 2147483647 without overflow */
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        g->st[i] = word;
    }
    g->f = 3; g->r = 0;
    for (int i = 0; i < 310; i++) (void)kmo_rand(g);
}
int kmo_rand(kmo_glibc_rand *g) {
    uint32_t v = (uint32_t)g->st[g->f] + (uint32_t)g->st[g->r];
    g->st[g->f] = (int32_t)v;
    if (++g->f >= 31) { g->f = 0; ++g->r; }
    else if (++g->r >= 31) g->r = 0;
    return (int)(v >> 1);
}
void kmo_gen_rand_reads(kmo_glibc_rand *g, uint64_t first, uint64_t n, int read_len, char *bases) {
    const unsigned rl = (unsigned)read_len;
    for (uint64_t i = 0; i < n; i++) {
        char *rb = bases + i * rl;
        const int gc_bucket = (int)((first + i) % 10);                 /* :693 */
        const int beg = gc_bucket * 10, end = gc_bucket * 10 + 9;      /* gc_range, :677-685: (int)lval, (int)(lval+width-1) */
        const int range = (end - beg) + 1;
        const int gc_draw = (kmo_rand(g) % range) + beg;               /* :88 */
        const float gc_pcnt = (float)(gc_draw / 100.0);                /* :89 */
        const unsigned num_gc = (unsigned)(gc_pcnt * (float)rl);       /* :90 (float * unsigned -> float) */
        for (unsigned q = 0; q < num_gc; q++) rb[q] = (kmo_rand(g) % 100) < 50 ? 'g' : 'c';      /* :91-95 */
        for (unsigned q = num_gc; q < rl; q++) rb[q] = (kmo_rand(g) % 100) < 50 ? 'a' : 't';     /* :96-100 */
        for (unsigned q = 1; q < rl; q++) {                            /* std::random_shuffle, libstdc++ stl_algo.h */
            const unsigned j = (unsigned)kmo_rand(g) % (q + 1);
            if (j != q) { char t = rb[q]; rb[q] = rb[j]; rb[j] = t; }
        }
    }
}

static uint32_t null_row_of(kmo_ctx *c, uint32_t tid) {
    uint32_t row;
    if (u32map_find(&c->null_row, tid, &row)) return row;
    if (!c->null_row.cap) u32map_init(&c->null_row, 1024);
    row = (uint32_t)c->null_tids.n;
    if (row >= c->null_cap) {
        uint32_t ncap = c->null_cap ? c->null_cap * 2 : 1024;
        c->null_max = (float *)realloc(c->null_max, (size_t)ncap * 10 * sizeof(float));
        c->null_cnt = (int32_t *)realloc(c->null_cnt, (size_t)ncap * 10 * sizeof(int32_t));
        memset(c->null_max + (size_t)c->null_cap * 10, 0, (size_t)(ncap - c->null_cap) * 10 * sizeof(float));
        memset(c->null_cnt + (size_t)c->null_cap * 10, 0, (size_t)(ncap - c->null_cap) * 10 * sizeof(int32_t));
        c->null_cap = ncap;
    }
    u32map_put(&c->null_row, tid, row);
    u32vec_push(&c->null_tids, tid);
    return row;
}

int kmo_null_batch(kmo_ctx *c, const char *bases, const uint64_t *offs, uint32_t n, uint64_t first_index) {
    read_state rs; memset(&rs, 0, sizeof rs);
    const int k = c->db.kmer_len;
    const int save_rkmer = c->opt.rkmer;
    c->opt.rkmer = 1;
    int rc = 0;
    u32map cnt_tids; memset(&cnt_tids, 0, sizeof cnt_tids);
    u32vec cnt_keys = {0};
    for (uint32_t i = 0; i < n && rc == 0; i++) {
        const char *seq = bases + offs[i];
        const int len = (int)(offs[i + 1] - offs[i]);
        const unsigned gcbucket = (unsigned)((first_index + i) % 10);
        if (len < k) continue;                                                       /* :372-374 */
        const int np = len - k + 1;
        if (np > rs.cap_label) {
            rs.label = (label_info *)realloc(rs.label, (size_t)np * sizeof(label_info));
            memset(rs.label + rs.cap_label, 0, (size_t)(np - rs.cap_label) * sizeof(label_info));
            rs.cap_label = np;
        }
        rs.n_label = np;
        for (int p = 0; p < np; p++) { rs.label[p].first = -1; rs.label[p].set.n = 0; }  /* :377 */
        rs.taxid_lst.n = 0;
        u32map_free(&rs.tax2idx); u32map_init(&rs.tax2idx, 64);
        rs.err = 0;
        int valid = 0, bin_sel = 0;
        retrieve_kmer_labels(c, &rs, seq, len, k, &valid, &bin_sel);                 /* rkmer.hpp:76-294 */
        if (rs.err) { rc = rs.err; break; }
        if (valid <= 0) continue;                                                    /* :381 */
        u32map_free(&cnt_tids); u32map_init(&cnt_tids, 64); cnt_keys.n = 0;
        for (int p = 0; p < np; p++)                                                 /* :382-393 */
            for (size_t q = 0; q < rs.label[p].set.n; q++) {
                const uint32_t tid = rs.label[p].set.v[q];
                uint32_t v;
                if (u32map_find(&cnt_tids, tid, &v)) u32map_put(&cnt_tids, tid, v + 1);
                else { u32map_put(&cnt_tids, tid, 1); u32vec_push(&cnt_keys, tid); }
            }
        for (size_t q = 0; q < cnt_keys.n; q++) {                                    /* construct_labels :184-213 */
            const uint32_t tid = cnt_keys.v[q];
            uint32_t found = 0; u32map_find(&cnt_tids, tid, &found);
            const float label_prob = (float)(int)found / (float)valid;               /* :195 */
            const uint32_t row = null_row_of(c, tid);
            float *mx = c->null_max + (size_t)row * 10 + gcbucket;
            if (*mx < label_prob) *mx = label_prob;                                   /* :203-210 (first insert: 0 -> prob) */
            c->null_cnt[(size_t)row * 10 + gcbucket] += 1;
        }
    }
    c->opt.rkmer = save_rkmer;
    u32map_free(&cnt_tids); free(cnt_keys.v);
    rs_free(&rs);
    return rc;
}
uint32_t kmo_null_rows(const kmo_ctx *c) { return (uint32_t)c->null_tids.n; }
void kmo_null_get(const kmo_ctx *c, uint32_t *tids, float *max_frac, int32_t *cnt) {
    const uint32_t n = (uint32_t)c->null_tids.n;
    uint32_t *ord = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    memcpy(ord, c->null_tids.v, (size_t)n * sizeof(uint32_t));
    qsort(ord, n, 4, u32_cmp);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t row = 0; u32map_find(&c->null_row, ord[i], &row);
        tids[i] = ord[i];
        memcpy(max_frac + (size_t)i * 10, c->null_max + (size_t)row * 10, 10 * sizeof(float));
        memcpy(cnt + (size_t)i * 10, c->null_cnt + (size_t)row * 10, 10 * sizeof(int32_t));
    }
    free(ord);
}
void kmo_null_reset(kmo_ctx *c) {
    u32map_free(&c->null_row); c->null_tids.n = 0;
    free(c->null_max); free(c->null_cnt); c->null_max = NULL; c->null_cnt = NULL; c->null_cap = 0;
}

/* Text after "hdr\tread\t" -- read_label.cpp:1218,1233,1271,844-848,894-937.  Floats go through
 * ostream operator<<(float): "%g" with precision 6. */
static const char *match_str(int m) {
    switch (m) { case KMO_DIRECT: return "DirectMatch"; case KMO_MULTI: return "MultiMatch"; case KMO_PARTIAL: return "PartialMultiMatch";
                 case KMO_NOMATCH: return "NoMatch"; default: return "LCA_ERROR"; }
}
int kmo_format_tail(const kmo_ctx *c, const kmo_result *r, char *buf, size_t cap) {
    size_t w = 0;
#define EMIT(...) do { int _n = snprintf(buf + w, w < cap ? cap - w : 0, __VA_ARGS__); if (_n < 0) return -1; w += (size_t)_n; } while (0)
    switch (r->status) {
        case KMO_ST_SHORT_LEN: case KMO_ST_SHORT_VALID: EMIT("-1 -1 -1\t-1 -1\t%d %d ReadTooShort\n", r->n1, r->n2); break;
        case KMO_ST_NODBHITS: EMIT("-1 -1 %d\t-1 -1\t%d %d NoDbHits\n", r->valid_kmers, r->n1, r->n2); break;
        case KMO_ST_SILENT: break;
        case KMO_ST_PHIX:
            EMIT("-1 -1 %d\t%u %g\t%u %g %s\n", r->cand_kmer_cnt, r->tid, (double)r->score, r->tid, (double)r->score, match_str(KMO_DIRECT));
            break;
        default: {
            EMIT("%g %g %d\t", (double)r->log_avg, (double)r->stdev, r->cand_kmer_cnt);
            if (c->opt.prn_all && r->n_cand) {     /* n_cand == 0 only when use_sig_hits == 0: block skipped */
                int prn = 0;
                for (int i = (int)r->n_cand - 1; i >= 0; --i) {
                    const kmo_pair *p = &c->cands.v[r->cand_off + (uint64_t)i];
                    if (p->score >= 0 || c->opt.prn_all > 1) { EMIT(" %u %g", p->tid, (double)p->score); prn = 1; }   /* prn_all = 2: -p with -y (read_label.cpp:901) */
                }
                if (!prn) EMIT("-1 -1");
                EMIT("\t");
            }
            if (r->match == KMO_DIRECT) EMIT("%u %g %s", r->tid, (double)r->score, match_str(r->match));
            else if (r->match == KMO_MULTI || r->match == KMO_PARTIAL) {
                if (!c->opt.prn_all) {
                    for (uint32_t i = 0; i < r->n_lin; i++) { const kmo_pair *p = &c->lineage.v[r->lin_off + i]; EMIT(" %u %g", p->tid, (double)p->score); }
                    if (!r->n_lin) EMIT("-1 -1");
                    EMIT("\t");
                }
                EMIT("%u %g %s", r->tid, (double)r->score, match_str(r->match));
            } else if (r->match == KMO_NOMATCH) EMIT("-1 -1 %s", match_str(r->match));
            else EMIT("-1 -1 Unmatched");
            EMIT("\n");
        }
    }
#undef EMIT
    if (w < cap) buf[w] = 0;
    return (int)w;
}
