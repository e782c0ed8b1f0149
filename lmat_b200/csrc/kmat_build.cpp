// kmat_build.cpp -- table construction from tax_histo files: the logical content that make_db_table /
// SortedDb<tid_T>::add_data (src/make_db_table.cpp:105-433, src/kmerdb/SortedDb.cpp:84-751) put into the PERM heap,
// built directly as a kmat_table (ascending k-mers + CSR lists of stored ids), ready for kmat_table_save /
// kmat_db_upload.  SURVEY.md 8(f-2).
//
// What is restated is what a reader sees through begin_/next, k-mer by k-mer:
//   * k-mers must arrive strictly ascending across all input files (SortedDb.cpp:164-167);
//   * optional sorted ASCII human k-mer stream (-j): human-only k-mers become singletons (9606, or the adaptor id when
//     the k-mer is an adaptor), shared ones get 9606 added to their list (:170-233, 339-357, 432-476, 662-706); the
//     stream's pending k-mer is re-read -- i.e. one k-mer is dropped -- at the start of every input file (:106-110);
//   * optional adaptor k-mer set (-u): such k-mers are singletons with id 32630 (:275-291; ILLU_TAXID, make_db_table.cpp:26);
//   * optional pruning (-g N -m ranks): lists longer than N go through a priority queue ordered by numeric rank only,
//     whole equal-rank batches are popped from the top until <= N remain; nothing left -> taxid 1; the survivors are
//     stored in pop order (:296-409, 570-596).  std::priority_queue<MyPair> (SortedDb.hpp:128-139) is used as is, so
//     the order equals the reference's for the same libstdc++;
//   * -g without -m (empty map): the list is cut to the single stored id 1 (:298-303, 556-560);
//   * optional 32->16-bit id map (-f): every stored id goes through it; a missing id is the reference's assert
//     ("bad read/single/set", :455-458 etc.) and KMAT_ERR_BAD_TAXID here.
// The physical layout (top tier, pages, the kmer % 4096 echo) is not reproduced: libkmat has its own.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <queue>
#include <thread>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "kmat_internal.h"

namespace {

struct MyPair {                                   // SortedDb.hpp:128-139
    MyPair(unsigned int f, uint32_t s) : first(f), second(s) {}
    bool operator<(const MyPair &mp) const { return first < mp.first; }
    unsigned int first;
    uint32_t second;
};

int tokbits(char c) {                             // kencode.hpp:26-39: unknown characters encode as 0
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}

struct KmerStream {                               // read_encode, SortedDb.cpp:38-60: fscanf("%s") tokens
    FILE *fp = nullptr;
    int k = 0;
    bool bad = false;
    uint64_t next() {
        if (!fp) return ~0ull;
        char buf[256];
        if (fscanf(fp, "%255s", buf) == EOF) return ~0ull;
        if ((int)strlen(buf) < k) { bad = true; return ~0ull; }     // the reference reads past the terminator here
        uint64_t v = 0;
        for (int i = 0; i < k; i++) v = (v << 2) | (uint64_t)tokbits(buf[i]);
        return v;
    }
};

struct Builder {
    const kmat_build_opts &o;
    std::unordered_map<uint32_t, uint16_t> br_map;    // -f
    bool has_br = false;
    std::unordered_map<uint32_t, uint32_t> species_map;   // -m (numeric ranks), loaded only with -g
    std::unordered_set<uint64_t> adaptor;
    bool has_adaptor = false;
    KmerStream human;
    uint64_t last_human = ~0ull, last_kmer = 0;
    uint64_t first_kmer = 0; bool any_kmer = false;      // first raw k-mer this builder saw (the parallel build checks the order across files)
    uint16_t HUMAN_16 = 0, ADAPTOR_16 = 0;
    kmat_table *t;
    std::string err;
    int err_code = KMAT_OK;

    Builder(const kmat_build_opts &opts, kmat_table *tab) : o(opts), t(tab) {}

    bool fail(int code, const std::string &m) { err_code = code; err = m; return false; }

    // (*p_br_map)[tid] with the reference's sanity check
    bool stored(uint32_t tid, uint32_t *out, const char *what) {
        if (!has_br) { *out = tid; return true; }
        auto it = br_map.find(tid);
        const uint16_t v = it == br_map.end() ? 0 : it->second;
        if (v == 0 || (size_t)v > br_map.size() + 1)
            return fail(KMAT_ERR_BAD_TAXID, std::string("bad ") + what + ": taxid " + std::to_string(tid) + " -> " + std::to_string(v) +
                                                " (not in the 32->16-bit id map; the reference asserts)");
        *out = v;
        return true;
    }
    void push_single(uint64_t kmer, uint32_t sid) {
        t->own_kmers.push_back(kmer);
        t->own_ids.push_back(sid);
        t->own_offs.push_back(t->own_ids.size());
    }
    void close_list(uint64_t kmer) {
        t->own_kmers.push_back(kmer);
        t->own_offs.push_back(t->own_ids.size());
    }

    bool add_file(const char *fn) {
        FILE *in = fopen(fn, "rb");
        if (!in) return fail(KMAT_ERR_IO, std::string("cannot open ") + fn);
        struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{in};
        fseek(in, 0, SEEK_END);
        const long fsize = ftell(in);
        fseek(in, 0, SEEK_SET);
        // KmerFileMetaData::read (KmerFileMetaData.cpp:44-94)
        uint32_t data_start = 0, version = 0, klen = 0; uint64_t kmer_ct = 0, test = 0; char loc = 0;
        if (fread(&data_start, 4, 1, in) != 1 || fread(&kmer_ct, 8, 1, in) != 1 || fread(&test, 8, 1, in) != 1 || fread(&version, 4, 1, in) != 1 ||
            fread(&loc, 1, 1, in) != 1 || fread(&klen, 4, 1, in) != 1)
            return fail(KMAT_ERR_FORMAT, std::string(fn) + ": truncated header");
        if (test != ~0ull) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": kmer data file is invalid; should have read 64 1s");
        if (loc != 'N') return fail(loc == 'Y' ? KMAT_ERR_UNSUPPORTED : KMAT_ERR_FORMAT, std::string(fn) + ": data file with genome locations");
        if (ftell(in) != (long)data_start) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": header length mismatch");
        if (o.tax_histo_format && version != 999) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": not a tax_histo file (version " + std::to_string(version) + ")");
        // the pending human k-mer is replaced by a fresh read at the start of every file (:106-110)
        last_human = human.fp ? human.next() : ~0ull;
        const uint64_t stopper = o.stopper ? o.stopper : ~0ull;
        const int tid_cutoff = o.tid_cutoff;
        std::vector<uint32_t> tids;
        for (uint64_t i = 0; i < kmer_ct; i++) {
            if (i > stopper) break;
            if (ftell(in) == fsize) break;
            uint64_t kmer;
            if (fread(&kmer, 8, 1, in) != 1) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": truncated record");
            if (last_kmer > 0 && kmer <= last_kmer)
                return fail(KMAT_ERR_FORMAT, "Kmers arriving out of order.  New: " + std::to_string(kmer) + " last: " + std::to_string(last_kmer));
            if (!any_kmer) { any_kmer = true; first_kmer = kmer; }
            while (last_human < kmer) {                                   // new human-only k-mers (:170-226)
                const bool ad = has_adaptor && adaptor.count(last_human);
                push_single(last_human, ad ? (ADAPTOR_16 ? ADAPTOR_16 : 32630u) : (HUMAN_16 ? HUMAN_16 : 9606u));
                last_human = human.next();
            }
            bool add_human = false;
            if (last_human == kmer) { add_human = true; last_human = human.next(); }
            uint16_t tid_count;
            if (o.tax_histo_format) { if (fread(&tid_count, 2, 1, in) != 1) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": truncated record"); }
            else { uint32_t c32; if (fread(&c32, 4, 1, in) != 1) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": truncated record"); tid_count = (uint16_t)c32; }
            tids.resize(tid_count);
            if (tid_count && fread(tids.data(), 4, tid_count, in) != tid_count) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": truncated record");
            uint32_t sid;
            if (has_adaptor && adaptor.count(kmer)) {                      // :275-291
                push_single(kmer, ADAPTOR_16 ? ADAPTOR_16 : 32630u);
            } else {
                uint16_t tmp_tid_count = tid_count;
                std::priority_queue<MyPair> q;
                if (tid_cutoff > 0 && (int)tid_count > tid_cutoff) {
                    if (species_map.empty()) tmp_tid_count = 0;            // :298-303
                    else {                                                 // :339-404
                        for (uint32_t tid : tids) {
                            if (add_human && tid == 9606) add_human = false;
                            q.push(MyPair(species_map[tid], tid));        // operator[]: unknown taxids get rank 0
                        }
                        if (add_human) q.push(MyPair(species_map[9606], 9606));
                        while (!q.empty()) {
                            const unsigned int cur = q.top().first;
                            while (q.top().first == cur) { q.pop(); if (q.empty()) break; }
                            if ((int)q.size() <= tid_cutoff) { tmp_tid_count = (uint16_t)q.size(); break; }
                        }
                        if (q.empty()) { tmp_tid_count = 1; q.push(MyPair(1, 1)); }
                    }
                }
                if (tmp_tid_count > 1) {
                    if (q.size() > 1) {                                    // pruned list, pop order (:570-609)
                        for (int j = 0; j < (int)tmp_tid_count; j++) {
                            const uint32_t tid = q.top().second;
                            q.pop();
                            if (!stored(tid, &sid, "set")) return false;
                            t->own_ids.push_back(sid);
                        }
                        close_list(kmer);
                    } else {                                               // no reduction: file order (+ 9606) (:613-706)
                        for (uint32_t tid : tids) {
                            if (tid == 9606) add_human = false;
                            if (!stored(tid, &sid, "read")) return false;
                            t->own_ids.push_back(sid);
                        }
                        if (add_human) { if (!stored(9606, &sid, "read")) return false; t->own_ids.push_back(sid); }
                        close_list(kmer);
                    }
                } else if (tid_count == 1) {                               // :425-517
                    const uint32_t tid = tids[0];
                    if (add_human && tid != 9606) {                        // "doubles": [tid, human]
                        if (!stored(tid, &sid, "read")) return false;
                        t->own_ids.push_back(sid);
                        t->own_ids.push_back(has_br ? (uint32_t)HUMAN_16 : 9606u);
                        close_list(kmer);
                    } else {
                        if (!stored(tid, &sid, "single")) return false;
                        push_single(kmer, sid);
                    }
                } else if (tmp_tid_count == 1) {                           // pruned to one (:519-535)
                    if (!stored(q.top().second, &sid, "single")) return false;
                    push_single(kmer, sid);
                } else push_single(kmer, 1u);                              // cut to taxid 1, stored unconverted (:537-541)
            }
            // sanity word after every 1500th (tax_histo) / 1000th (kmerPrefixCounter) record (:717-727)
            const uint64_t period = o.tax_histo_format ? 1500 : 1000;
            if ((i + 1) % period == 0) {
                uint64_t s;
                if (fread(&s, 8, 1, in) != 1 || s != ~0ull) return fail(KMAT_ERR_FORMAT, std::string(fn) + ": sanity word missing after record " + std::to_string(i + 1));
            }
            last_kmer = kmer;
        }
        if (human.bad) return fail(KMAT_ERR_FORMAT, "human k-mer file: a line shorter than k");
        return true;
    }
};

bool load_pairs(const char *fn, std::vector<std::pair<long long, long long>> &out) {
    std::ifstream f(fn);
    if (!f) return false;
    long long a, b;
    while (f >> a >> b) out.emplace_back(a, b);
    return true;
}

}  // namespace

extern "C" void kmat_build_opts_default(kmat_build_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->kmer_length = 20;
    o->tax_histo_format = 1;
}

extern "C" int kmat_table_build(const char *const *files, int n_files, const kmat_build_opts *opts, kmat_table **out) {
    if (!files || n_files < 1 || !opts || !out) { kmat_set_error("kmat_table_build: bad argument"); return KMAT_ERR_ARG; }
    if (opts->kmer_length < 1 || opts->kmer_length > 32) { kmat_set_error("kmat_table_build: k-mer length %d", opts->kmer_length); return KMAT_ERR_ARG; }
    kmat_table *t = new kmat_table();
    t->kmer_len = opts->kmer_length;
    t->own_offs.push_back(0);
    Builder b(*opts, t);
    std::vector<std::pair<long long, long long>> pairs;
    if (opts->map16 && opts->map16[0]) {                                   // make_db_table.cpp:259-273
        if (!load_pairs(opts->map16, pairs)) { delete t; kmat_set_error("cannot open %s", opts->map16); return KMAT_ERR_IO; }
        for (auto &p : pairs) b.br_map[(uint32_t)p.first] = (uint16_t)p.second;
        b.has_br = !b.br_map.empty();
    }
    if (b.has_br) { b.HUMAN_16 = b.br_map[9606]; b.ADAPTOR_16 = b.br_map[32630]; }   // SortedDb.cpp:145-148 (operator[] inserts)
    if (opts->tid_cutoff > 0 && opts->numrank && opts->numrank[0]) {       // :303-315
        pairs.clear();
        if (!load_pairs(opts->numrank, pairs)) { delete t; kmat_set_error("cannot open %s", opts->numrank); return KMAT_ERR_IO; }
        for (auto &p : pairs) b.species_map[(uint32_t)p.first] = (uint32_t)p.second;
    }
    FILE *hfp = nullptr, *afp = nullptr;
    if (opts->human_kmers && opts->human_kmers[0]) hfp = fopen(opts->human_kmers, "r");      // a missing file is "No human k-mer file." (:277-289)
    if (opts->adaptor_kmers && opts->adaptor_kmers[0]) afp = fopen(opts->adaptor_kmers, "r");
    b.human.fp = hfp; b.human.k = opts->kmer_length;
    int rc = KMAT_OK;
    // Without a human k-mer stream (the only state that runs across files besides the ascending-order check) every input
    // file can be parsed on its own: one worker per file, partial tables appended in file order.  The reference's build is
    // one thread; KMAT_BUILD_SERIAL=1 keeps ours serial too.
    if (!hfp && n_files > 1 && !getenv("KMAT_BUILD_SERIAL")) {
        if (afp) {
            KmerStream as; as.fp = afp; as.k = opts->kmer_length;
            for (uint64_t v = as.next(); v != ~0ull; v = as.next()) b.adaptor.insert(v);
            b.has_adaptor = true;
            if (as.bad) { fclose(afp); kmat_set_error("adaptor k-mer file: a line shorter than k"); delete t; return KMAT_ERR_FORMAT; }
            fclose(afp);
        }
        struct Part { std::unique_ptr<kmat_table> tab; std::unique_ptr<Builder> bld; bool ok = true; };
        std::vector<Part> parts((size_t)n_files);
        std::atomic<int> next{0};
        auto worker = [&] {
            for (int i = next++; i < n_files; i = next++) {
                Part &pt = parts[(size_t)i];
                pt.tab.reset(new kmat_table());
                pt.tab->kmer_len = opts->kmer_length;
                pt.tab->own_offs.push_back(0);
                pt.bld.reset(new Builder(*opts, pt.tab.get()));
                Builder &bi = *pt.bld;
                bi.br_map = b.br_map; bi.has_br = b.has_br; bi.species_map = b.species_map; bi.adaptor = b.adaptor; bi.has_adaptor = b.has_adaptor;
                bi.HUMAN_16 = b.HUMAN_16; bi.ADAPTOR_16 = b.ADAPTOR_16;
                pt.ok = bi.add_file(files[i]);
            }
        };
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n_thr = (int)std::min<unsigned>({(unsigned)n_files, hw, 32u});
        std::vector<std::thread> th;
        for (int i = 0; i < n_thr; i++) th.emplace_back(worker);
        for (auto &x : th) x.join();
        uint64_t last = 0;
        for (int i = 0; i < n_files && rc == KMAT_OK; i++) {
            Part &pt = parts[(size_t)i];
            Builder &bi = *pt.bld;
            // the order check of add_data across the file boundary (SortedDb.cpp:164-167); a failure inside an earlier file wins
            if (bi.any_kmer && last > 0 && bi.first_kmer <= last) {
                // the serial build meets this record before anything that fails later in file i
                rc = KMAT_ERR_FORMAT; b.err = "Kmers arriving out of order.  New: " + std::to_string(bi.first_kmer) + " last: " + std::to_string(last);
                break;
            }
            if (!pt.ok) { rc = bi.err_code; b.err = bi.err; break; }
            const uint64_t base = t->own_ids.size();
            t->own_kmers.insert(t->own_kmers.end(), pt.tab->own_kmers.begin(), pt.tab->own_kmers.end());
            t->own_ids.insert(t->own_ids.end(), pt.tab->own_ids.begin(), pt.tab->own_ids.end());
            for (size_t j = 1; j < pt.tab->own_offs.size(); j++) t->own_offs.push_back(pt.tab->own_offs[j] + base);
            if (bi.any_kmer) last = bi.last_kmer;
            pt.bld.reset(); pt.tab.reset();
        }
        if (rc != KMAT_OK) { kmat_set_error("%s", b.err.c_str()); delete t; return rc; }
        t->tid_bytes = b.has_br ? 2 : 4;
        t->n_kmers = t->own_kmers.size(); t->n_ids = t->own_ids.size();
        t->kmers = t->own_kmers.data(); t->offs = t->own_offs.data(); t->ids = t->own_ids.data();
        *out = t;
        return KMAT_OK;
    }
    for (int i = 0; i < n_files && rc == KMAT_OK; i++) {
        if (i == 0 && afp) {                                               // get_kmer_set on the first add_data call (:112-116)
            KmerStream as; as.fp = afp; as.k = opts->kmer_length;
            // the stream is read after the first human k-mer of the first file; order does not matter for a set
            for (uint64_t v = as.next(); v != ~0ull; v = as.next()) b.adaptor.insert(v);
            b.has_adaptor = true;
            if (as.bad) { rc = KMAT_ERR_FORMAT; b.err = "adaptor k-mer file: a line shorter than k"; break; }
        }
        if (!b.add_file(files[i])) rc = b.err_code;
    }
    if (hfp) fclose(hfp);
    if (afp) fclose(afp);
    if (rc != KMAT_OK) { kmat_set_error("%s", b.err.c_str()); delete t; return rc; }
    t->tid_bytes = b.has_br ? 2 : 4;
    t->n_kmers = t->own_kmers.size(); t->n_ids = t->own_ids.size();
    t->kmers = t->own_kmers.data(); t->offs = t->own_offs.data(); t->ids = t->own_ids.data();
    *out = t;
    return KMAT_OK;
}
