// kmat_host.cpp -- host side of libkmat: table ingest, run-time input parsers, node-table construction,
// output formatting.  No CUDA here; the device side is kmat_device.cu.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdarg>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "kmat_internal.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024];
void kmat_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *kmat_last_error(void) { return g_err; }
extern "C" const char *kmat_strerror(int code) {
    switch (code) {
        case KMAT_OK: return "ok";
        case KMAT_ERR_ARG: return "bad argument";
        case KMAT_ERR_NO_DEVICE: return "no usable CUDA device (libkmat has no CPU fallback)";
        case KMAT_ERR_CUDA: return "CUDA call failed";
        case KMAT_ERR_NOMEM: return "out of memory";
        case KMAT_ERR_IO: return "I/O error";
        case KMAT_ERR_FORMAT: return "malformed input";
        case KMAT_ERR_UNSUPPORTED: return "unsupported input";
        case KMAT_ERR_BAD_TAXID: return "stored taxid missing from the 16-bit map";
        case KMAT_ERR_TREE: return "taxonomy is not a forest";
        case KMAT_ERR_OVERFLOW: return "output buffer too small";
        default: return "unknown error";
    }
}
extern "C" int kmat_abi_version(void) { return KMAT_ABI_VERSION; }
extern "C" void kmat_opts_default(kmat_opts *o) {
    // read_label.cpp:1336-1347 and the ScoreOptions ctor (:488)
    o->min_kmer = 35; o->min_fnd_kmer = 1; o->sdiff = 1.0f; o->hbias = 3.0f; o->min_score = 0.0f;
    o->max_count = 65535; o->permissive = 0; o->phix_screen = 1; o->want_lineage = 0; o->rkmer_mode = 0;
}

// ---------------------------------------------------------------------------------------------
// table ingest
// ---------------------------------------------------------------------------------------------
static const uint64_t kPageSize = 4294701056ull;   // SortedDb.hpp:27
static const unsigned kMaxPage = 255;               // SortedDb.hpp:28

// Walks the reference layout the way begin_20/begin_18 + next do (SortedDb.hpp:202-385): top-tier entry
// = count<<48 | first record; records {u16 lsb; u16 page; u32 offset}; page 255 = inline singleton;
// otherwise [u64 kmer echo iff kmer % 4096 == 0][u16 count][count x tid_T] at page*PAGE_SIZE+offset.
extern "C" int kmat_table_from_sorteddb(const uint64_t *tt, uint64_t tt_count, int bits2, const void *kmer_table,
                                        uint64_t n_records, const char *storage, uint64_t storage_bytes, int kmer_len,
                                        int tid_bytes, kmat_table **out) {
    if (!tt || !kmer_table || !out || (tid_bytes != 2 && tid_bytes != 4) || bits2 <= 0 || bits2 > 16) {
        kmat_set_error("kmat_table_from_sorteddb: bad argument");
        return KMAT_ERR_ARG;
    }
    kmat_table *t = new kmat_table();
    t->kmer_len = kmer_len; t->tid_bytes = tid_bytes;
    t->own_kmers.reserve(n_records); t->own_offs.reserve(n_records + 1);
    t->own_offs.push_back(0);
    const uint8_t *recs = (const uint8_t *)kmer_table;
    uint64_t last = 0; bool have_last = false;
    for (uint64_t p = 0; p < tt_count; p++) {
        uint64_t e = tt[p];
        if (!e) continue;
        uint64_t cnt = e >> 48, first = e & 0x0000ffffffffffffull;
        if (first + cnt > n_records) { delete t; kmat_set_error("top tier entry %llu points past the record table", (unsigned long long)p); return KMAT_ERR_FORMAT; }
        for (uint64_t i = 0; i < cnt; i++) {
            const uint8_t *r = recs + (first + i) * 8;
            uint16_t lsb, page16; uint32_t off;
            memcpy(&lsb, r, 2); memcpy(&page16, r + 2, 2); memcpy(&off, r + 4, 4);
            uint64_t kmer = (p << bits2) | lsb;
            if (have_last && kmer <= last) { delete t; kmat_set_error("records not ascending at k-mer %llu", (unsigned long long)kmer); return KMAT_ERR_FORMAT; }
            last = kmer; have_last = true;
            unsigned page = page16 & 0xff;                      // narrowed to uint8_t page_out (SortedDb.hpp:305)
            t->own_kmers.push_back(kmer);
            if (page == kMaxPage) {
                t->own_ids.push_back(tid_bytes == 2 ? (uint32_t)(uint16_t)off : off);   // next(): taxid_out = offset (:370-371)
            } else {
                uint64_t a = kPageSize * page + off;
                if (kmer % 4096 == 0) a += 8;                   // self-check echo (:331-337)
                if (!storage || a + 2 > storage_bytes) { delete t; kmat_set_error("list offset out of range"); return KMAT_ERR_FORMAT; }
                uint16_t c; memcpy(&c, storage + a, 2);
                a += 2;
                if (a + (uint64_t)c * tid_bytes > storage_bytes) { delete t; kmat_set_error("list runs past the storage space"); return KMAT_ERR_FORMAT; }
                for (unsigned j = 0; j < c; j++) {
                    uint32_t v = 0; memcpy(&v, storage + a + (uint64_t)j * tid_bytes, tid_bytes);
                    t->own_ids.push_back(v);
                }
            }
            t->own_offs.push_back(t->own_ids.size());
        }
    }
    t->n_kmers = t->own_kmers.size(); t->n_ids = t->own_ids.size();
    t->kmers = t->own_kmers.data(); t->offs = t->own_offs.data(); t->ids = t->own_ids.data();
    *out = t;
    return KMAT_OK;
}

extern "C" int kmat_table_from_arrays(const uint64_t *kmers, const uint64_t *offs, const uint32_t *ids, uint64_t n,
                                      int kmer_len, int tid_bytes, kmat_table **out) {
    if ((n && (!kmers || !offs)) || !out || (tid_bytes != 2 && tid_bytes != 4)) { kmat_set_error("kmat_table_from_arrays: bad argument"); return KMAT_ERR_ARG; }
    for (uint64_t i = 1; i < n; i++) if (kmers[i] <= kmers[i - 1]) { kmat_set_error("k-mers must be strictly ascending (index %llu)", (unsigned long long)i); return KMAT_ERR_FORMAT; }
    kmat_table *t = new kmat_table();
    t->kmer_len = kmer_len; t->tid_bytes = tid_bytes; t->n_kmers = n; t->n_ids = n ? offs[n] : 0;
    t->own_kmers.assign(kmers, kmers + n);
    if (n) t->own_offs.assign(offs, offs + n + 1); else t->own_offs.assign(1, 0);
    t->own_ids.assign(ids, ids + t->n_ids);
    t->kmers = t->own_kmers.data(); t->offs = t->own_offs.data(); t->ids = t->own_ids.data();
    *out = t;
    return KMAT_OK;
}

struct KmatFileHeader { char magic[8]; uint32_t version, kmer_len, tid_bytes, pad; uint64_t n_kmers, n_ids; };
static const char kFlatMagic[8] = {'K', 'M', 'A', 'T', 'T', 'B', 'L', '1'};

extern "C" int kmat_table_save(const kmat_table *t, const char *path) {
    if (!t || !path) return KMAT_ERR_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) { kmat_set_error("cannot write %s", path); return KMAT_ERR_IO; }
    KmatFileHeader h; memset(&h, 0, sizeof h);
    memcpy(h.magic, kFlatMagic, 8); h.version = 1; h.kmer_len = t->kmer_len; h.tid_bytes = t->tid_bytes; h.n_kmers = t->n_kmers; h.n_ids = t->n_ids;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    ok = ok && (t->n_kmers == 0 || fwrite(t->kmers, 8, t->n_kmers, f) == t->n_kmers);
    ok = ok && fwrite(t->offs, 8, t->n_kmers + 1, f) == t->n_kmers + 1;
    ok = ok && (t->n_ids == 0 || fwrite(t->ids, 4, t->n_ids, f) == t->n_ids);
    fclose(f);
    if (!ok) { kmat_set_error("short write to %s", path); return KMAT_ERR_IO; }
    return KMAT_OK;
}

// SortedDb object layout on x86-64 (SortedDb.hpp:453-481): int idx_config @0; size_t m_n_kmers @8;
// uint8_t m_kmer_length @16; char* m_storage_space @24; kmer_record* kmer_table @32; uint64_t*
// top_tier_block @40; size_t m_list_offset @48 (== size()); ...
extern "C" int kmat_table_open(const char *path, int tid_bytes, kmat_table **out) {
    if (!path || !out) return KMAT_ERR_ARG;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { kmat_set_error("Error: unable to open kmer db [%s]", path); return KMAT_ERR_IO; }   // read_label.cpp:1483
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 64) { close(fd); kmat_set_error("%s: too small to be a DB", path); return KMAT_ERR_FORMAT; }
    void *m = mmap(nullptr, sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { kmat_set_error("mmap(%s) failed", path); return KMAT_ERR_IO; }
    const uint8_t *b = (const uint8_t *)m;
    int rc;
    if (memcmp(b, kFlatMagic, 8) == 0) {
        KmatFileHeader h; memcpy(&h, b, sizeof h);
        // counts are checked against the file size one by one (a damaged header must not wrap the sum around)
        const uint64_t fsz = (uint64_t)sb.st_size;
        const bool counts_ok = h.n_kmers <= fsz / 16 && h.n_ids <= fsz / 4 && sizeof h + 8 * h.n_kmers + 8 * (h.n_kmers + 1) + 4 * h.n_ids <= fsz;
        if (h.version != 1 || !counts_ok) { munmap(m, sb.st_size); kmat_set_error("%s: truncated .kmat image", path); return KMAT_ERR_FORMAT; }
        const uint64_t *offs = (const uint64_t *)(b + sizeof h) + h.n_kmers;
        if (h.kmer_len < 1 || h.kmer_len > 32 || (h.tid_bytes != 2 && h.tid_bytes != 4) || offs[0] != 0 || offs[h.n_kmers] != h.n_ids) {
            munmap(m, sb.st_size); kmat_set_error("%s: inconsistent .kmat header", path); return KMAT_ERR_FORMAT;
        }
        kmat_table *t = new kmat_table();
        t->kmer_len = h.kmer_len; t->tid_bytes = h.tid_bytes; t->n_kmers = h.n_kmers; t->n_ids = h.n_ids;
        t->kmers = (const uint64_t *)(b + sizeof h);
        t->offs = t->kmers + h.n_kmers;
        t->ids = (const uint32_t *)(t->offs + h.n_kmers + 1);
        t->map_base = m; t->map_len = sb.st_size;
        *out = t;
        return KMAT_OK;
    } else if (memcmp(b, "KMPERM01", 8) == 0) {
        // stand-in PERM heap written by oracle/_ref/make_db_table (oracle/standins/jemalloc/pallocator.h)
        uint64_t base, fsize, brk, nreg, n0, obj;
        memcpy(&base, b + 8, 8); memcpy(&fsize, b + 16, 8); memcpy(&brk, b + 24, 8); memcpy(&nreg, b + 32, 8);
        memcpy(&n0, b + 40, 8); memcpy(&obj, b + 48, 8);
        if (nreg < 1 || n0 != 8 || obj < base || obj - base + 88 > (uint64_t)sb.st_size) { munmap(m, sb.st_size); kmat_set_error("%s: bad KMPERM01 header", path); return KMAT_ERR_FORMAT; }
        const uint8_t *o = b + (obj - base);
        int32_t idx_config; uint8_t klen; uint64_t p_storage, p_table, p_tt, n_rec;
        memcpy(&idx_config, o, 4); memcpy(&klen, o + 16, 1);
        memcpy(&p_storage, o + 24, 8); memcpy(&p_table, o + 32, 8); memcpy(&p_tt, o + 40, 8); memcpy(&n_rec, o + 48, 8);
        int bits2; uint64_t tt_count = 134217728ull;                                    // TT_BLOCK_COUNT_{18,20} (SortedDb.hpp:36,41)
        if (klen == 20) bits2 = 13; else if (klen == 18) bits2 = 9;
        else { munmap(m, sb.st_size); kmat_set_error("K size %d not supported by this application version!", (int)klen); return KMAT_ERR_UNSUPPORTED; }   // SortedDb.hpp:195-197
        if (p_tt < base || p_table < base || p_storage < base || p_tt - base + tt_count * 8 > (uint64_t)sb.st_size ||
            n_rec > (uint64_t)sb.st_size / 8 || p_table - base + n_rec * 8 > (uint64_t)sb.st_size || p_storage - base > (uint64_t)sb.st_size) {
            munmap(m, sb.st_size); kmat_set_error("%s: SortedDb pointers outside the image", path); return KMAT_ERR_FORMAT;
        }
        rc = kmat_table_from_sorteddb((const uint64_t *)(b + (p_tt - base)), tt_count, bits2, b + (p_table - base), n_rec,
                                      (const char *)(b + (p_storage - base)), (uint64_t)sb.st_size - (p_storage - base), klen, tid_bytes, out);
        munmap(m, sb.st_size);
        return rc;
    }
    munmap(m, sb.st_size);
    kmat_set_error("%s: neither a .kmat image nor a KMPERM01 heap (real perm-je heaps are not supported: layout unpinned)", path);
    return KMAT_ERR_FORMAT;
}
extern "C" uint64_t kmat_table_size(const kmat_table *t) { return t ? t->n_kmers : 0; }
extern "C" int kmat_table_kmer_length(const kmat_table *t) { return t ? t->kmer_len : 0; }
extern "C" int kmat_table_tid_bytes(const kmat_table *t) { return t ? t->tid_bytes : 0; }
extern "C" int kmat_table_view(const kmat_table *t, const uint64_t **kmers, const uint64_t **offs, const uint32_t **ids, uint64_t *n_ids) {
    if (!t) return KMAT_ERR_ARG;
    if (kmers) *kmers = t->kmers;
    if (offs) *offs = t->offs;
    if (ids) *ids = t->ids;
    if (n_ids) *n_ids = t->n_ids;
    return KMAT_OK;
}
extern "C" void kmat_table_free(kmat_table *t) {
    if (!t) return;
    if (t->map_base) munmap(t->map_base, t->map_len);
    delete t;
}

// ---------------------------------------------------------------------------------------------
// run-time input parsers
// ---------------------------------------------------------------------------------------------
static bool slurp(const char *path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss; ss << f.rdbuf(); out = ss.str();
    return true;
}
// "tid value" per line, read like `while (ifs >> a >> b)` / fscanf("%d%d") loops
static bool parse_u32_pairs(const char *path, std::vector<uint32_t> &a, std::vector<uint32_t> &b) {
    std::string s;
    if (!slurp(path, s)) return false;
    const char *p = s.c_str();
    char *e;
    for (;;) {
        long long x = strtoll(p, &e, 10); if (e == p) break; p = e;
        long long y = strtoll(p, &e, 10); if (e == p) break; p = e;
        a.push_back((uint32_t)x); b.push_back((uint32_t)y);
    }
    return true;
}

static int class_id(kmat_inputs *in, const std::string &s) {
    for (size_t i = 0; i < in->class_names.size(); i++) if (in->class_names[i] == s) return (int)i;
    static const char *names[] = {"no_rank", "ethnic", "region", "species", "genus", "family", "order", "class", "phylum", "kingdom", "depth=0"};
    static const int nums[] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9};           // gRank2num, read_label.cpp:519-532
    int rn = 0;                                                             // operator[] on an unknown string -> 0 (:786)
    for (int i = 0; i < 11; i++) if (s == names[i]) rn = nums[i];
    in->class_names.push_back(s); in->class_ranknum.push_back(rn);
    return (int)in->class_names.size() - 1;
}

// loadRandHits, read_label.cpp:512-678
static int load_null_models(kmat_inputs *in, const char *list_path, const char *lmat_dir) {
    std::ifstream lst(list_path);
    if (!lst) { fprintf(stderr, "Unexpected reading error (RandHits file list): %s\n", list_path); return KMAT_OK; }   // :514-517 returns silently
    in->models_requested = true;
    int read_len; std::string file;
    while (lst >> read_len >> file) {                                                    // :553
        if (lmat_dir) file = std::string(lmat_dir) + "/" + file;                         // :555-558
        else fprintf(stderr, "WARNING! Missing LMAT_DIR environment variable!\n");
        in->read_len_vec.push_back(read_len);                                            // :562
        FILE *pre = fopen(file.c_str(), "rb");
        if (!pre) { fprintf(stderr, "Unexpected reading error (RandHits file), skipping... %s\n", file.c_str()); continue; }   // :564-568
        fclose(pre);
        gzFile gz = gzopen(file.c_str(), "rb");
        if (!gz) continue;
        kmat_null_model *m = nullptr;
        for (auto &mm : in->models) if (mm.kmer_cnt == read_len) m = &mm;               // rand_hits_all[read_len] (:571)
        if (!m) { in->models.emplace_back(); m = &in->models.back(); m->kmer_cnt = read_len; }
        m->loaded = true;
        std::unordered_map<uint32_t, uint32_t> rowof;
        for (uint32_t i = 0; i < m->tid.size(); i++) rowof[m->tid[i]] = i;
        static const int buff_size = 20004;                                              // :577
        std::vector<char> buff(buff_size);
        if (!gzgets(gz, buff.data(), buff_size)) { gzclose(gz); continue; }
        int num_bins = atoi(buff.data());                                                // :579-583
        if (num_bins <= 0) { gzclose(gz); kmat_set_error("%s: num_bins must be > 0", file.c_str()); return KMAT_ERR_FORMAT; }
        if (in->nbins && in->nbins != num_bins) { gzclose(gz); kmat_set_error("%s: null models disagree on the bin count", file.c_str()); return KMAT_ERR_UNSUPPORTED; }
        in->nbins = num_bins;
        std::vector<float> save_ecoli(num_bins, 0.5f), cutoff(num_bins);                 // :584
        std::vector<unsigned> revisit;
        while (gzgets(gz, buff.data(), buff_size)) {                                     // :585
            std::istringstream is(buff.data());
            uint32_t taxid = 0; std::string class_str; float max_val = 0;
            is >> taxid >> class_str;
            if (class_str.empty()) continue;
            size_t pos = class_str.find('-');                                            // :591-593
            if (pos == std::string::npos) { gzclose(gz); kmat_set_error("%s: class string without '-'", file.c_str()); return KMAT_ERR_FORMAT; }
            std::string val = class_str.substr(0, pos);
            if (val.size() >= 3 && val[0] == 'n' && val[1] == 'o' && val[2] == '_') val = "genus";   // :594-601
            revisit.clear();
            std::fill(cutoff.begin(), cutoff.end(), 0.0f);                               // :603
            for (int bin = 0; bin < num_bins; ++bin) {                                   // :604-630
                int num_obs = 0, kmer_cnt = 0;
                is >> num_obs >> max_val >> kmer_cnt;
                if (num_obs == 0 && kmer_cnt >= 100000) { max_val = 0.5f; cutoff[bin] = max_val; }
                else if (num_obs == 0 && kmer_cnt < 100000) revisit.push_back(bin);
                if (num_obs > 0) { cutoff[bin] = max_val; if (taxid == 562) save_ecoli[bin] = cutoff[bin]; }
                if (taxid == 28384) { val = "genus"; cutoff = save_ecoli; }
            }
            for (unsigned it : revisit) {                                                // :631-665
                int j = (int)it - 1; unsigned i = it + 1;
                while (j >= 0 || i < cutoff.size()) {
                    float a_val = 0.0f, b_val = 0.0f;
                    if (j >= 0) a_val = cutoff[j];
                    if (i < cutoff.size()) b_val = cutoff[i];
                    if (a_val > 0 && b_val > 0) cutoff[it] = std::max(a_val, b_val);
                    else if (a_val > 0) cutoff[it] = a_val;
                    else if (b_val > 0) cutoff[it] = b_val;
                    if (cutoff[it] > 0) break;
                    --j; ++i;
                }
                if (cutoff[it] <= 0) cutoff[it] = 0.5f;
            }
            int cid = class_id(in, val);
            auto f = rowof.find(taxid);                                                  // :666-667 operator[]= : last line wins
            uint32_t row;
            if (f == rowof.end()) { row = (uint32_t)m->tid.size(); rowof[taxid] = row; m->tid.push_back(taxid); m->cls.push_back(0); m->cut.resize((size_t)(row + 1) * num_bins); }
            else row = f->second;
            m->cls[row] = (uint16_t)cid;
            std::copy(cutoff.begin(), cutoff.end(), m->cut.begin() + (size_t)row * num_bins);
        }
        gzclose(gz);
    }
    std::sort(in->read_len_vec.begin(), in->read_len_vec.end());                         // :672
    in->read_len_avgs.clear();                                                           // :674-677
    for (size_t i = 1; i < in->read_len_vec.size(); i++) in->read_len_avgs.push_back((in->read_len_vec[i - 1] + in->read_len_vec[i]) / 2);
    return KMAT_OK;
}

extern "C" int kmat_inputs_load(const char *tree, const char *depth, const char *rank, const char *conv16, const char *numrank,
                                const char *plasmids, const char *null_list, const char *lmat_dir, kmat_inputs **out) {
    if (!out) return KMAT_ERR_ARG;
    kmat_inputs *in = new kmat_inputs();
    in->read_len_vec.assign(1, 0); in->read_len_avgs.assign(1, 0);                       // read_label.cpp:60-61
    static const char *n2r[] = {"no_rank", "region", "species", "genus", "family", "order", "class", "phylum", "kingdom", "depth=0"};
    for (int i = 0; i < 10; i++) class_id(in, n2r[i]);                                   // ids 0..9 = gNum2rank keys (:534-547)
    if (tree) {
        // TaxTree ctor + TaxNode::read (TaxTree.hpp:24-57, TaxNode.hpp:131-147): 2 comment lines, a count line,
        // then per node the token stream "id nchild child*nchild parent", the rest of that line, a name line.
        std::ifstream f(tree);
        if (!f.is_open()) { delete in; kmat_set_error("failed to open %s for reading", tree); return KMAT_ERR_IO; }
        std::string line;
        std::getline(f, line); std::getline(f, line);
        int count; f >> count; std::getline(f, line);
        for (;;) {
            std::streampos p = f.tellg();
            if (f.eof() || !f.good() || (int)p == -1) break;
            uint32_t id, ct, child, parent;
            if (!(f >> id)) break;       // a trailing newline: the reference reads a phantom node here (UB)
            f >> ct;
            for (uint32_t j = 0; j < ct; j++) f >> child;
            f >> parent;
            std::getline(f, line); std::getline(f, line);
            in->node_tid.push_back(id); in->node_parent.push_back(parent);
        }
        in->has_tree = true;
    }
    if (depth && !parse_u32_pairs(depth, in->depth_tid, in->depth_val)) { delete in; kmat_set_error("ERROR! Unable to open: %s", depth); return KMAT_ERR_IO; }   // :1575-1577
    if (rank) {                                                                          // :1560-1567
        std::ifstream f(rank);
        uint32_t tid; std::string r;
        while (f >> tid >> r) { in->rank_tid.push_back(tid); in->rank_code.push_back(r == "strain" ? 1 : r == "species" ? 2 : 0); }
    }
    if (conv16) {                                                                        // :1585-1602 conv_map[dest] = src
        std::vector<uint32_t> src, dst;
        if (!parse_u32_pairs(conv16, src, dst)) { delete in; kmat_set_error("ERROR! Unable to read 16-bit map file:%s", conv16); return KMAT_ERR_IO; }
        for (size_t i = 0; i < src.size(); i++) { in->conv_stored.push_back((uint16_t)dst[i]); in->conv_tid.push_back(src[i]); }
        in->has_conv = true;
    }
    if (numrank) {                                                                       // :1543-1559 (always loads)
        if (parse_u32_pairs(numrank, in->prune_tid, in->prune_rank)) in->has_prune = !in->prune_tid.empty();
    }
    if (plasmids) {                                                                      // :499-510
        std::ifstream f(plasmids);
        if (!f) fprintf(stderr, "Unexpected reading error (plasmids): %s\n", plasmids);
        uint32_t pid;
        while (f >> pid) in->plasmid_tid.push_back(pid);
    }
    if (null_list) {
        int rc = load_null_models(in, null_list, lmat_dir);
        if (rc != KMAT_OK) { delete in; return rc; }
    }
    *out = in;
    return KMAT_OK;
}
extern "C" void kmat_inputs_free(kmat_inputs *in) { delete in; }

// ---------------------------------------------------------------------------------------------
// node universe, Euler intervals, root paths, model tables
// ---------------------------------------------------------------------------------------------
static bool is_human(uint32_t t) { return t == 9606 || t == 63221 || t == 741158; }      // tid_checks.hpp:15-28
static bool is_phix(uint32_t t) { return t == 374840 || t == 10847 || t == 32630; }      // tid_checks.hpp:13
static bool is_drop(uint32_t t) { return t == 20999999u || t == 12721 || t == 693660; }  // read_label.cpp:82-104,1038

int kmat_build_host_ctx(const kmat_inputs &in, int tid_bytes, const std::vector<uint32_t> &stored_tids, KmHostCtx &out) {
    // ---- universe of tids that can ever be touched: tree nodes, -f targets (or raw stored tids), 9606, 1
    std::vector<uint32_t> uni(in.node_tid);
    uni.insert(uni.end(), in.conv_tid.begin(), in.conv_tid.end());
    uni.insert(uni.end(), stored_tids.begin(), stored_tids.end());
    uni.push_back(9606); uni.push_back(1);
    std::sort(uni.begin(), uni.end());
    uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
    const uint32_t N = (uint32_t)uni.size();
    auto nid_of = [&](uint32_t tid) -> uint32_t {
        auto it = std::lower_bound(uni.begin(), uni.end(), tid);
        return (it != uni.end() && *it == tid) ? (uint32_t)(it - uni.begin()) : KMAT_NONE;
    };
    out.nodeA.assign(N, KmNodeA{0, 0, 0, KMAT_NONE});
    out.nodeB.assign(N, KmNodeB{0, 0, 0, 0});
    for (uint32_t i = 0; i < N; i++) { out.nodeA[i].tid = uni[i]; out.nodeA[i].parent = i; }
    // tree: later duplicates win, like (*this)[t->id()] = t (TaxTree.hpp:49)
    std::vector<uint32_t> parent_tid(N, 0); std::vector<uint8_t> in_tree(N, 0);
    for (size_t i = 0; i < in.node_tid.size(); i++) { uint32_t n = nid_of(in.node_tid[i]); parent_tid[n] = in.node_parent[i]; in_tree[n] = 1; }
    for (uint32_t n = 0; n < N; n++) {
        if (!in_tree[n]) continue;
        out.nodeA[n].meta |= KM_META_INTREE;
        if (parent_tid[n] == uni[n]) continue;                                   // root: its own parent
        uint32_t p = nid_of(parent_tid[n]);
        if (p == KMAT_NONE || !in_tree[p]) {                                     // TaxTree.hpp:73-77 "fatal error!" exit(-1)
            kmat_set_error("failed to find parent TaxNode for taxid %u whose parent is %u", uni[n], parent_tid[n]);
            return KMAT_ERR_TREE;
        }
        out.nodeA[n].parent = p;
    }
    // depth (-e): later lines win (operator[]=); missing -> 0 (what (*dmap.find(tid)).second reads with libstdc++)
    for (size_t i = 0; i < in.depth_tid.size(); i++) {
        uint32_t n = nid_of(in.depth_tid[i]);
        if (n == KMAT_NONE) continue;
        if (in.depth_val[i] > 0xFFFF) { kmat_set_error("depth %u of taxid %u exceeds 65535", in.depth_val[i], in.depth_tid[i]); return KMAT_ERR_UNSUPPORTED; }
        out.nodeA[n].meta = (out.nodeA[n].meta & ~KM_META_DEPTH_MASK) | in.depth_val[i];
    }
    // rank (-w): first line wins (map::insert)
    { std::vector<uint8_t> seen(N, 0);
      for (size_t i = 0; i < in.rank_tid.size(); i++) {
          uint32_t n = nid_of(in.rank_tid[i]);
          if (n == KMAT_NONE || seen[n]) continue;
          seen[n] = 1; out.nodeA[n].meta |= (uint32_t)in.rank_code[i] << KM_META_RANK_SHIFT;
      } }
    std::vector<uint32_t> plas(in.plasmid_tid); std::sort(plas.begin(), plas.end());
    for (uint32_t n = 0; n < N; n++) {
        uint32_t t = uni[n];
        if (is_human(t)) out.nodeA[n].meta |= KM_META_HUMAN;
        if (is_drop(t)) out.nodeA[n].meta |= KM_META_DROP;
        if (is_phix(t)) out.nodeA[n].meta |= KM_META_PHIX;
        if ((t >= 10000000u && t < 11000000u) || std::binary_search(plas.begin(), plas.end(), t)) out.nodeA[n].meta |= KM_META_PLASMID;
    }
    out.nid_human = nid_of(9606); out.nid_one = nid_of(1);
    // ---- children lists, cycle check, Euler intervals (preorder index / last index in subtree)
    std::vector<uint32_t> child_cnt(N + 1, 0);
    for (uint32_t n = 0; n < N; n++) if (out.nodeA[n].parent != n) child_cnt[out.nodeA[n].parent + 1]++;
    for (uint32_t n = 0; n < N; n++) child_cnt[n + 1] += child_cnt[n];
    std::vector<uint32_t> child(child_cnt[N]), fill(child_cnt.begin(), child_cnt.end() - 1);
    for (uint32_t n = 0; n < N; n++) if (out.nodeA[n].parent != n) child[fill[out.nodeA[n].parent]++] = n;
    uint32_t clock = 0; std::vector<uint8_t> visited(N, 0);
    std::vector<std::pair<uint32_t, uint32_t>> stack;
    for (uint32_t r = 0; r < N; r++) {
        if (out.nodeA[r].parent != r) continue;
        stack.push_back({r, child_cnt[r]}); out.nodeB[r].tin = clock++; visited[r] = 1;
        while (!stack.empty()) {
            auto &top = stack.back();
            uint32_t n = top.first;
            if (top.second < child_cnt[n + 1]) {
                uint32_t c = child[top.second++];
                out.nodeB[c].tin = clock++; visited[c] = 1;
                stack.push_back({c, child_cnt[c]});
            } else { out.nodeB[n].tout = clock - 1; stack.pop_back(); }
        }
    }
    for (uint32_t n = 0; n < N; n++) if (!visited[n]) { kmat_set_error("taxonomy has a parent cycle through taxid %u", uni[n]); return KMAT_ERR_TREE; }
    // ---- root paths (strict ancestors, nearest first: TaxTree::getPathToRoot) and first species ancestor
    uint64_t total = 0;
    for (uint32_t n = 0; n < N; n++) { uint32_t c = n, len = 0; while (out.nodeA[c].parent != c) { c = out.nodeA[c].parent; len++; } out.nodeB[n].path_len = len; total += len; }
    if (total >= 0xFFFFFFFFull) { kmat_set_error("taxonomy too deep: %llu path entries", (unsigned long long)total); return KMAT_ERR_UNSUPPORTED; }
    out.paths.resize(total ? total : 1);
    uint32_t w = 0;
    for (uint32_t n = 0; n < N; n++) {
        out.nodeB[n].path_off = w;
        uint32_t c = n;
        while (out.nodeA[c].parent != c) {
            c = out.nodeA[c].parent; out.paths[w++] = c;
            if (out.nodeA[n].species_anc == KMAT_NONE && ((out.nodeA[c].meta >> KM_META_RANK_SHIFT) & 3) == 2) out.nodeA[n].species_anc = c;
        }
    }
    // ---- stored id -> nid
    if (tid_bytes == 2) {
        out.sid2nid.assign(65536, KMAT_NONE);
        if (in.has_conv) { for (size_t i = 0; i < in.conv_stored.size(); i++) out.sid2nid[in.conv_stored[i] & 0xFFFF] = in.conv_tid[i] ? nid_of(in.conv_tid[i]) : KMAT_NONE; }
        else {
            // 16-bit table read without -f: the reference uses the stored value as the taxid (TaxNodeStat.hpp:240-250)
            for (uint32_t s = 0; s < 65536; s++) out.sid2nid[s] = nid_of(s);
        }
    } else {
        out.sid2nid.resize(stored_tids.size());
        for (size_t i = 0; i < stored_tids.size(); i++) out.sid2nid[i] = nid_of(stored_tids[i]);
    }
    // fold isHuman / dropped-tid flags into the stored-id table (bit 31 / bit 30; nid in the low 30 bits)
    if (N >= (1u << 30)) { kmat_set_error("taxonomy too large"); return KMAT_ERR_UNSUPPORTED; }
    for (auto &e : out.sid2nid) {
        if (e == KMAT_NONE) continue;
        const uint32_t meta = out.nodeA[e].meta;
        e |= (meta & KM_META_HUMAN) ? 0x80000000u : 0u;
        e |= (meta & KM_META_DROP) ? 0x40000000u : 0u;
    }
    // ---- pruning ranks (-m)
    if (in.has_prune) {
        out.prune_rank.assign(N, 0);
        for (size_t i = 0; i < in.prune_tid.size(); i++) { uint32_t n = nid_of(in.prune_tid[i]); if (n != KMAT_NONE) out.prune_rank[n] = in.prune_rank[i]; }
    }
    // ---- null models: closest()/getReadLen() (read_label.cpp:107-133) folded into a 65536-entry table
    out.nbins = in.nbins; out.n_classes = (int)in.class_names.size(); out.class_ranknum = in.class_ranknum;
    if (out.n_classes > 64) { kmat_set_error("more than 64 distinct null-model classes"); return KMAT_ERR_UNSUPPORTED; }
    std::vector<int> loaded_idx;
    for (size_t i = 0; i < in.models.size(); i++) if (in.models[i].loaded) loaded_idx.push_back((int)i);
    out.n_models = (int)loaded_idx.size();
    out.model_of_cand.assign(65536, -1);
    if (out.n_models) {
        for (int v = 0; v < 65536; v++) {
            size_t i; int len = -1;
            for (i = 0; i < in.read_len_avgs.size(); i++) if (v <= in.read_len_avgs[i]) { len = in.read_len_vec[i]; break; }
            if (i == in.read_len_avgs.size()) len = i < in.read_len_vec.size() ? in.read_len_vec[i] : -1;
            if (len <= 0) len = 80;                                                       // getReadLen fallback
            for (int m = 0; m < out.n_models; m++) if (in.models[loaded_idx[m]].kmer_cnt == len) out.model_of_cand[v] = (int16_t)m;
        }
        out.mrow.assign((size_t)out.n_models * N, -1);
        uint32_t rows = 0;
        for (int m = 0; m < out.n_models; m++) {
            const kmat_null_model &mm = in.models[loaded_idx[m]];
            for (size_t r = 0; r < mm.tid.size(); r++) {
                uint32_t n = nid_of(mm.tid[r]);
                if (n == KMAT_NONE) continue;
                out.mrow[(size_t)m * N + n] = (int32_t)rows;
                out.cls.push_back((uint8_t)mm.cls[r]);
                out.cut.insert(out.cut.end(), mm.cut.begin() + r * in.nbins, mm.cut.begin() + (r + 1) * in.nbins);
                rows++;
            }
        }
    }
    return KMAT_OK;
}

// ---------------------------------------------------------------------------------------------
// output formatting: read_label.cpp:1218,1233,1271,844-848,894-937
// ---------------------------------------------------------------------------------------------
static const char *match_str(int m) {
    switch (m) {
        case KMAT_DIRECT: return "DirectMatch";
        case KMAT_MULTI: return "MultiMatch";
        case KMAT_PARTIAL: return "PartialMultiMatch";
        case KMAT_NOMATCH: return "NoMatch";
        default: return "LCA_ERROR";
    }
}
// ---- number formatting.  The reference prints floats with ostream << float, i.e. printf("%g") of the promoted
// double (6 significant digits, trailing zeros stripped, exponent form outside [1e-4, 1e6)).  At tens of millions of
// reads per second the host writes ~20 such numbers per read, so snprintf (~300 ns) is the bottleneck of the writers.
// km_fmt_g is exact by construction: a float in [1e-4, 999999) is scaled to its six significant digits in integer
// arithmetic (see the function), rounded to nearest-even on the exact remainder; values outside that range take
// std::to_chars(general, 6), specified to equal printf("%.6g").  tests/fmt_exhaustive.cpp compares it with printf on
// every float of the range (279 M values), tests/test_abi_cpu.py on random bit patterns through kmat_format_tail.
static const char KM_DIGIT_PAIRS[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869"
    "707172737475767778798081828384858687888990919293949596979899";
static inline char *km_fmt_u32(char *p, uint32_t v) {
    // digit count first, then two digits per step from the back (taxids are 1-8 digits: at most 4 steps)
    const int n = v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6 : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
    char *q = p + n;
    while (v >= 100) { const uint32_t r = v % 100; v /= 100; q -= 2; memcpy(q, KM_DIGIT_PAIRS + 2 * r, 2); }
    if (v >= 10) memcpy(q - 2, KM_DIGIT_PAIRS + 2 * v, 2); else q[-1] = (char)('0' + v);
    return p + n;
}
static inline char *km_fmt_i32(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; return km_fmt_u32(p, (uint32_t)(-(int64_t)v)); }
    return km_fmt_u32(p, (uint32_t)v);
}
static char *km_fmt_g_slow(char *p, double x) {
    auto r = std::to_chars(p, p + 32, x, std::chars_format::general, 6);
    return r.ptr;
}
static inline char *km_fmt_g(char *p, float f) {
    uint32_t bits; memcpy(&bits, &f, 4);
    const uint32_t ab = bits & 0x7FFFFFFFu;
    // fixed notation with a 6-digit significand covers [1e-4, 999999): 0x38D1B718 is the smallest float >= 1e-4,
    // 0x497423F0 is 999999.0f; everything else (zero, tiny, huge, inf, nan) takes the library path
    if (ab - 0x38D1B718u >= 0x497423F0u - 0x38D1B718u) {
        if (ab == 0) { if (bits >> 31) *p++ = '-'; *p++ = '0'; return p; }
        return km_fmt_g_slow(p, (double)f);
    }
    // The value is m * 2^(b2 - 23) exactly (m = 24-bit significand).  e = floor(log10(value)): a binade holds at most one
    // power of ten, so e is the binade's lower bound E_LOW[b2] plus one when m reaches THR[b2], the smallest significand of
    // that binade whose value is >= 10^(E_LOW + 1) (2^24 = none).  Tables generated with exact rational arithmetic for
    // b2 = -14 .. 19.  The scaled significand m * 10^(5 - e) < 2^54 is then an exact integer, and shifting it right by
    // 23 - b2 leaves the six digits and the exact remainder: round to nearest, ties to even -- what printf("%g") does with
    // the exact binary value.  No floating-point operation, no fallback near ties.
    static const int8_t E_LOW[34] = {-5, -4, -4, -4, -4, -3, -3, -3, -2, -2, -2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 5};
    static const uint32_t THR[34] = {13743896, 16777216, 16777216, 16777216, 8589935, 16777216, 16777216, 10737419, 16777216, 16777216, 13421773,
                                     16777216, 16777216, 16777216, 16777216, 16777216, 16777216, 10485760, 16777216, 16777216, 13107200, 16777216,
                                     16777216, 16384000, 16777216, 16777216, 16777216, 10240000, 16777216, 16777216, 12800000, 16777216, 16777216, 16000000};
    static const uint64_t I10[11] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull, 10000000000ull};
    const int b2 = (int)(ab >> 23) - 127;
    const uint32_t m = (ab & 0x7FFFFFu) | 0x800000u;
    int e = E_LOW[b2 + 14] + (m >= THR[b2 + 14]);        // -4 .. 5
    const uint64_t scaled = (uint64_t)m * I10[5 - e];
    const int sh = 23 - b2;                              // 4 .. 37
    uint64_t n = scaled >> sh;
    const uint64_t rem = scaled & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    n += (uint64_t)((rem > half) | ((rem == half) & (n & 1)));
    if (n >= 1000000) { n = 100000; e++; if (e > 5) return km_fmt_g_slow(p, (double)f); }
    if (bits >> 31) *p++ = '-';
    // six significant digits, two per table lookup; d[6..15] pad the fixed-size copies below
    char d[16] = {'0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0', '0'};
    const uint32_t n32 = (uint32_t)n, g0 = n32 / 10000, lo = n32 % 10000, g1 = lo / 100, g2 = lo % 100;
    memcpy(d, KM_DIGIT_PAIRS + 2 * g0, 2); memcpy(d + 2, KM_DIGIT_PAIRS + 2 * g1, 2); memcpy(d + 4, KM_DIGIT_PAIRS + 2 * g2, 2);
    // index of the last non-zero digit (d[0] is never '0')
    int last;
    if (g2) last = (g2 % 10) ? 5 : 4;
    else if (g1) last = (g1 % 10) ? 3 : 2;
    else last = (g0 % 10) ? 1 : 0;
    // no per-digit loops (their trip counts are what the branch predictor cannot learn): whole-group copies, then the
    // returned pointer cuts the text to its length.  Callers leave >= 16 bytes of slack after every number.
    if (e >= 0) {
        memcpy(p, d, 8);                                 // integer part: digits 0 .. e (zeros included)
        if (last > e) { p[e + 1] = '.'; memcpy(p + e + 2, d + e + 1, 8); return p + last + 2; }
        return p + e + 1;
    }
    memcpy(p, "0.000000", 8);                            // "0." and -e - 1 zeros
    memcpy(p - e + 1, d, 8);
    return p - e + 2 + last;
}
static inline char *km_fmt_str(char *p, const char *s) { while (*s) *p++ = *s++; return p; }
// "<tid> <score>" with the text of the previous score reused when the value repeats (lineage ancestors share scores).
// The memo points at the previous text inside the output buffer (text is only ever appended); the copy is a fixed 16
// bytes (a %g text is at most 12), loaded before it is stored because source and destination may be closer than that.
struct KmScoreMemo { uint32_t bits = 0; const char *s = nullptr; int n = 0; };
static inline char *km_fmt_pair(char *p, uint32_t tid, float score, KmScoreMemo &m) {
    p = km_fmt_u32(p, tid);
    *p++ = ' ';
    uint32_t a;
    memcpy(&a, &score, 4);
    if (m.s && a == m.bits) {
        uint64_t w0, w1;
        memcpy(&w0, m.s, 8); memcpy(&w1, m.s + 8, 8);
        memcpy(p, &w0, 8); memcpy(p + 8, &w1, 8);
        return p + m.n;
    }
    char *q = km_fmt_g(p, score);
    m.bits = a; m.s = p; m.n = (int)(q - p);
    return q;
}

extern "C" int kmat_format_tail(const kmat_read_result *r, const kmat_pair *cands, const kmat_pair *lineage, int prn_all, char *buf, size_t cap) {
    // worst case per number: 16 bytes; every fixed part below is < 64 bytes
    auto need = [&](size_t pairs) { return 160 + pairs * 40; };
    char *p = buf;
    switch (r->status) {
        case KMAT_ST_SHORT_LEN: case KMAT_ST_SHORT_VALID:
            if (cap < need(0)) return KMAT_ERR_OVERFLOW;
            p = km_fmt_str(p, "-1 -1 -1\t-1 -1\t"); p = km_fmt_i32(p, r->n1); *p++ = ' '; p = km_fmt_i32(p, r->n2); p = km_fmt_str(p, " ReadTooShort\n");
            break;
        case KMAT_ST_NODBHITS:
            if (cap < need(0)) return KMAT_ERR_OVERFLOW;
            p = km_fmt_str(p, "-1 -1 "); p = km_fmt_i32(p, r->valid_kmers); p = km_fmt_str(p, "\t-1 -1\t"); p = km_fmt_i32(p, r->n1); *p++ = ' ';
            p = km_fmt_i32(p, r->n2); p = km_fmt_str(p, " NoDbHits\n");
            break;
        case KMAT_ST_SILENT: break;
        case KMAT_ST_PHIX:
            if (cap < need(0)) return KMAT_ERR_OVERFLOW;
            p = km_fmt_str(p, "-1 -1 "); p = km_fmt_i32(p, r->cand_kmer_cnt); *p++ = '\t';
            p = km_fmt_u32(p, r->tid); *p++ = ' '; p = km_fmt_g(p, r->score); *p++ = '\t';
            p = km_fmt_u32(p, r->tid); *p++ = ' '; p = km_fmt_g(p, r->score); *p++ = ' '; p = km_fmt_str(p, match_str(KMAT_DIRECT)); *p++ = '\n';
            break;
        case KMAT_ST_LABELED: {
            const size_t pairs = prn_all ? r->n_cand : ((r->match == KMAT_MULTI || r->match == KMAT_PARTIAL) ? r->n_lin : 0);
            if (cap < need(pairs)) return KMAT_ERR_OVERFLOW;
            KmScoreMemo memo;
            p = km_fmt_g(p, r->log_avg); *p++ = ' '; p = km_fmt_g(p, r->stdev); *p++ = ' '; p = km_fmt_i32(p, r->cand_kmer_cnt); *p++ = '\t';
            if (prn_all) {
                if (!cands && r->n_cand) return KMAT_ERR_ARG;
                bool prn = false;
                for (int i = (int)r->n_cand - 1; i >= 0; --i) {
                    const kmat_pair &c = cands[r->cand_off + (uint64_t)i];
                    if (c.score >= 0 || prn_all > 1) { *p++ = ' '; p = km_fmt_pair(p, c.tid, c.score, memo); prn = true; }   // prn_all = 2: -p under -y (:901)
                }
                if (!prn) p = km_fmt_str(p, "-1 -1");
                *p++ = '\t';
            }
            if (r->match == KMAT_DIRECT) { p = km_fmt_pair(p, r->tid, r->score, memo); *p++ = ' '; p = km_fmt_str(p, match_str(r->match)); }
            else if (r->match == KMAT_MULTI || r->match == KMAT_PARTIAL) {
                if (!prn_all) {
                    if (!lineage && r->n_lin) return KMAT_ERR_ARG;
                    for (uint32_t i = 0; i < r->n_lin; i++) { *p++ = ' '; p = km_fmt_pair(p, lineage[r->lin_off + i].tid, lineage[r->lin_off + i].score, memo); }
                    if (!r->n_lin) p = km_fmt_str(p, "-1 -1");
                    *p++ = '\t';
                }
                p = km_fmt_pair(p, r->tid, r->score, memo); *p++ = ' '; p = km_fmt_str(p, match_str(r->match));
            } else if (r->match == KMAT_NOMATCH) { p = km_fmt_str(p, "-1 -1 "); p = km_fmt_str(p, match_str(r->match)); }
            else p = km_fmt_str(p, "-1 -1 Unmatched");
            *p++ = '\n';
            break;
        }
        default: return KMAT_ERR_ARG;
    }
    if ((size_t)(p - buf) < cap) *p = 0;
    return (int)(p - buf);
}

// ---------------------------------------------------------------------------------------------
// rand_read_label: merge phase and the .rand_lst writer (src/rand_read_label.cpp:702-755)
// ---------------------------------------------------------------------------------------------
extern "C" int kmat_null_write(const char *path, int n_sets, const uint32_t *const *tids, const float *const *max_frac,
                               const uint64_t *const *counts, const uint32_t *n_rows) {
    if (!path || n_sets < 0 || (n_sets && (!tids || !max_frac || !counts || !n_rows))) { kmat_set_error("kmat_null_write: bad argument"); return KMAT_ERR_ARG; }
    struct Row { float mx[KMAT_NULL_BUCKETS]; uint64_t cnt[KMAT_NULL_BUCKETS]; };
    std::map<uint32_t, Row> merged;                                  // merge_score / merge_count (:703-735)
    for (int s = 0; s < n_sets; s++)
        for (uint32_t i = 0; i < n_rows[s]; i++) {
            auto it = merged.find(tids[s][i]);
            if (it == merged.end()) { Row z; memset(&z, 0, sizeof z); it = merged.emplace(tids[s][i], z).first; }
            for (int b = 0; b < KMAT_NULL_BUCKETS; b++) {
                const float v = max_frac[s][(size_t)i * KMAT_NULL_BUCKETS + b];
                if (v > it->second.mx[b]) it->second.mx[b] = v;
                it->second.cnt[b] += counts[s][(size_t)i * KMAT_NULL_BUCKETS + b];
            }
        }
    FILE *f = fopen(path, "w");
    if (!f) { kmat_set_error("Could not open for writing %s", path); return KMAT_ERR_IO; }
    char buf[64 * KMAT_NULL_BUCKETS + 32];
    for (const auto &kv : merged) {                                  // sum_ofs<<tid; " "<<max_score[val]<<" "<<cnt[val]; endl (:745-754)
        char *p = km_fmt_u32(buf, kv.first);
        for (int b = 0; b < KMAT_NULL_BUCKETS; b++) {
            *p++ = ' '; p = km_fmt_g(p, kv.second.mx[b]);
            *p++ = ' '; p += snprintf(p, 24, "%llu", (unsigned long long)kv.second.cnt[b]);
        }
        *p++ = '\n';
        if (fwrite(buf, 1, (size_t)(p - buf), f) != (size_t)(p - buf)) { fclose(f); kmat_set_error("write to %s failed", path); return KMAT_ERR_IO; }
    }
    if (fclose(f) != 0) { kmat_set_error("write to %s failed", path); return KMAT_ERR_IO; }
    return KMAT_OK;
}


// ---------------------------------------------------------------------------------------------
// compact interface, host side: ASCII -> 2-bit code words + positions of the invalid bases (include/kmat.h)
// ---------------------------------------------------------------------------------------------
extern "C" uint64_t kmat_pack_words(uint64_t total_bases) { return (total_bases + 15) / 16; }
extern "C" int kmat_pack_reads(const char *bases, uint64_t total_bases, int threads, uint32_t *codes, uint64_t *inv_pos, uint64_t inv_cap, uint64_t *n_inv) {
    if ((total_bases && (!bases || !codes)) || !n_inv) { kmat_set_error("kmat_pack_reads: bad argument"); return KMAT_ERR_ARG; }
    static const std::vector<uint8_t> lut = [] {                       // ENCODE (read_label.cpp:943-950): everything else is 4 = invalid
        std::vector<uint8_t> t(256, 4);
        t['a'] = t['A'] = 0; t['c'] = t['C'] = 1; t['g'] = t['G'] = 2; t['t'] = t['T'] = 3;
        return t; }();
    const uint64_t n_words = kmat_pack_words(total_bases);
    if (threads < 1) threads = 1;
    threads = (int)std::min<uint64_t>((uint64_t)threads, std::max<uint64_t>(1, n_words / 4096));
    std::vector<std::vector<uint64_t>> inv((size_t)threads);
    auto work = [&](int t) {
        const uint64_t w0 = n_words * (uint64_t)t / (uint64_t)threads, w1 = n_words * (uint64_t)(t + 1) / (uint64_t)threads;
        const uint8_t *L = lut.data();
        for (uint64_t w = w0; w < w1; w++) {
            const uint64_t b0 = w * 16, nb = std::min<uint64_t>(16, total_bases - b0);
            uint32_t v = 0;
            for (uint64_t i = 0; i < nb; i++) {
                const uint8_t c = L[(uint8_t)bases[b0 + i]];
                if (c > 3) inv[(size_t)t].push_back(b0 + i); else v |= (uint32_t)c << (2 * i);
            }
            codes[w] = v;
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    uint64_t total = 0;
    for (auto &v : inv) total += v.size();
    *n_inv = total;
    if (total > inv_cap || (total && !inv_pos)) { kmat_set_error("kmat_pack_reads: %llu invalid bases, room for %llu", (unsigned long long)total, (unsigned long long)inv_cap); return KMAT_ERR_OVERFLOW; }
    uint64_t at = 0;
    for (auto &v : inv) { if (!v.empty()) memcpy(inv_pos + at, v.data(), v.size() * 8); at += v.size(); }
    return KMAT_OK;
}
