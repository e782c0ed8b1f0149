// rand_read_label -- drop-in host for LMAT's null-model generator (src/rand_read_label.cpp main(), :410-759) over
// libkmat's C ABI.
//
// Same getopt string and option meanings (bin/gen_rand_mod.sh:137 is the canonical invocation); writes
// <ofbase>.rand_lst in the reference's format.  `-t T -g N` means T "threads" of N reads each (:687-701): the run has
// T * N reads, read i of thread t being run index t * N + i, so every thread's reads cycle through the ten GC buckets
// from bucket 0 like the reference's.  The per-read work runs on every visible GPU (KMAT_DEVICES=0,2 selects),
// the run indices being dealt to the GPUs in contiguous ranges; per-GPU accumulators are merged by max / sum like the
// reference merges its per-thread maps (:702-735).
//
// Reads: by default they are drawn on the GPU from a counter-based generator keyed by (seed, thread, index within the
// thread) -- seed from $KMAT_RAND_SEED, else time(0) like the reference's srand (:412); the output then depends on the
// seed, -t and -g only, not on the number of GPUs or on scheduling.  KMAT_RAND_COMPAT=glibc reproduces the reference's own draw sequence instead: glibc rand()
// + genRandRead + std::random_shuffle on ONE host thread (what `-t 1` does under a fixed time(0)), reads copied to the GPU.
#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "kmat.h"

static void usage(const char *exe) {
    std::cout << "Usage:\n";
    std::cout << exe << " -d <input db file (list)> -i <read length> -g <reads per thread> -t <number of threads> -o <output path>\n";
    std::cout << "-c <tax tree file> -e <depth file> -f <32to16 id map> [-w <rank/tid-map-file>] [-h <tid-cutoff> -r <numeric rank table>] [-k <kmer size>]\n";
}

int main(int argc, char *argv[]) {
    signed char c;
    int k_size = -1;
    unsigned n_threads = 0, num_reads = 0, read_len = 0;
    uint16_t max_count = (uint16_t)~0;
    std::string rank_map_file, kmer_db_fn, ofbase, tax_tree_fn, depth_file, rank_table_file, id_bit_conv_fn;
    while ((c = getopt(argc, argv, "u:ah:n:j:b:ye:w:mpk:c:v:k:i:d:l:t:s:r:o:x:f:g:z:q:")) != -1) {      // :442
        switch (c) {
            case 'f': id_bit_conv_fn = optarg; break;
            case 'e': depth_file = optarg; break;
            case 'w': rank_map_file = optarg; break;
            case 'h': max_count = (uint16_t)atoi(optarg); break;
            case 'r': rank_table_file = optarg; break;
            case 't': n_threads = (unsigned)atoi(optarg); break;
            case 'c': tax_tree_fn = optarg; break;
            case 'k': k_size = atoi(optarg); break;
            case 'g': num_reads = (unsigned)atoi(optarg); break;
            case 'i': read_len = (unsigned)atoi(optarg); break;
            case 'd': kmer_db_fn = optarg; break;
            case 'o': ofbase = optarg; break;
            case 'j': case 'u': case 'x': case 'a': case 'n': case 'y': case 'p': case 'v': case 's': case 'b': case 'l': case 'z': case 'q': case 'm':
                break;                                         // parsed by the reference, without effect on the .rand_lst
            default: std::cout << "Unrecognized option: " << c << ", ignore." << std::endl;
        }
    }
    std::cout << "Total reads to evaluate: " << (unsigned long long)num_reads * n_threads << std::endl;
    if (ofbase.empty()) std::cout << "ofbase\n";
    if (n_threads == 0) std::cout << "n_threads\n";
    if (kmer_db_fn.empty()) std::cout << "kmer_db_fn\n";
    if (ofbase.empty() || n_threads == 0 || kmer_db_fn.empty()) {
        std::cout << ofbase << " " << n_threads << " " << kmer_db_fn << " " << depth_file << std::endl;
        usage(argv[0]);
        return -1;
    }
    if (read_len == 0) { std::cerr << "ERROR! -i <read length> is required" << std::endl; return -1; }

    std::cout << "Start kmer DB load\n";
    const char *tb = getenv("KMAT_TID_BYTES");
    const int tid_bytes = tb ? atoi(tb) : 2;
    if (tid_bytes == 2 && id_bit_conv_fn.empty()) { usage(argv[0]); return -1; }          // TID_SIZE == 16 needs -f (:519-524)
    kmat_table *table = nullptr;
    if (kmat_table_open(kmer_db_fn.c_str(), tid_bytes, &table) != KMAT_OK) {
        std::cerr << "Unable to open [" << kmer_db_fn << "]: " << kmat_last_error() << std::endl;
        return -1;
    }
    if (k_size < 1) k_size = kmat_table_kmer_length(table);
    std::cout << "num kmers: " << kmat_table_size(table) << " - " << k_size << std::endl;
    if (k_size <= 0) { std::cerr << "Unable to read database, k_size=" << k_size << std::endl; return -1; }

    std::vector<int> devs;
    if (const char *dv = getenv("KMAT_DEVICES")) {
        std::stringstream ss(dv);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
    } else for (int i = 0; i < kmat_device_count(); i++) devs.push_back(i);
    if (devs.empty()) { std::cerr << "ERROR! No CUDA device: this build has no CPU path (" << kmat_last_error() << ")" << std::endl; return -1; }

    if (!id_bit_conv_fn.empty()) std::cout << "Loading map file: " << id_bit_conv_fn << std::endl;
    std::cout << "Read taxonomy tree: " << tax_tree_fn << std::endl;
    std::cout << "Read taxonomy depth: " << depth_file << std::endl;
    if (depth_file.empty()) { std::cerr << "unable to open: " << depth_file << std::endl; return -1; }
    auto nz = [](const std::string &s) { return s.empty() ? nullptr : s.c_str(); };
    kmat_inputs *inputs = nullptr;
    // -r is the numeric rank table of the run-time pruning (tid_rank_map, :636-652); no plasmid list, no null models here
    if (kmat_inputs_load(nz(tax_tree_fn), depth_file.c_str(), nz(rank_map_file), nz(id_bit_conv_fn), nz(rank_table_file), nullptr, nullptr, nullptr, &inputs) != KMAT_OK) {
        std::cerr << "ERROR! " << kmat_last_error() << std::endl;
        return -1;
    }
    kmat_opts opt;
    kmat_opts_default(&opt);
    opt.rkmer_mode = 1;
    opt.max_count = max_count;

    const char *compat = getenv("KMAT_RAND_COMPAT");
    const bool glibc_compat = compat && strcmp(compat, "glibc") == 0;
    const char *seed_env = getenv("KMAT_RAND_SEED");
    const uint64_t seed = seed_env ? strtoull(seed_env, nullptr, 10) : (uint64_t)(unsigned)time(nullptr);          // std::srand(unsigned(std::time(0))) (:412)
    if (glibc_compat && n_threads != 1)
        std::cout << "KMAT_RAND_COMPAT=glibc: the reference's draw order is only defined for one thread; drawing all " << n_threads << " threads' reads in sequence" << std::endl;
    for (unsigned i = 0; i < 10; i++) std::cout << "gc check " << i << " " << i * 10 << " " << i * 10 + 9 << std::endl;     // :677-685

    const auto t0 = std::chrono::steady_clock::now();
    const size_t nd = devs.size();
    std::vector<kmat_db *> dbs(nd, nullptr);
    std::vector<kmat_ctx *> ctxs(nd, nullptr);
    std::vector<int> rcs(nd, 0);
    std::vector<std::string> errs(nd);
    std::vector<std::vector<uint32_t>> tids(nd);
    std::vector<std::vector<float>> mx(nd);
    std::vector<std::vector<uint64_t>> cnt(nd);
    std::vector<uint64_t> nerr(nd, 0);
    const uint64_t total = (uint64_t)num_reads * n_threads;
    // run index i = thread * num_reads + j belongs to GC bucket j % 10 (:693): the device API takes the bucket as
    // (first_index + r) % 10, so the GPUs' ranges are cut at thread boundaries and passed with first_index = j
    auto work = [&](size_t d) {
        auto fail = [&](int rc) { rcs[d] = rc; errs[d] = kmat_last_error(); };
        int rc;
        if ((rc = kmat_db_upload(table, devs[d], 0, 1, &dbs[d])) != KMAT_OK) return fail(rc);
        if ((rc = kmat_ctx_create(dbs[d], inputs, &opt, &ctxs[d])) != KMAT_OK) return fail(rc);
        const uint64_t lo = total * d / nd, hi = total * (d + 1) / nd;
        if (glibc_compat) {
            if (d != 0) return;                                   // one sequential rand() stream: device 0 takes all of it
            // glibc TYPE_3 rand() (stdlib/random_r.c) + genRandRead (:85-103) + libstdc++ random_shuffle, on the host
            int32_t st[31]; int f = 3, r = 0;
            auto next = [&]() -> int {
                const uint32_t v = (uint32_t)st[f] + (uint32_t)st[r];
                st[f] = (int32_t)v;
                if (++f >= 31) { f = 0; ++r; } else if (++r >= 31) r = 0;
                return (int)(v >> 1);
            };
            {
                int32_t word = (int32_t)((unsigned)seed ? (unsigned)seed : 1u);
                st[0] = word;
                for (int i = 1; i < 31; i++) { const long h = word / 127773, l = word % 127773; word = (int32_t)(16807 * l - 2836 * h); if (word < 0) word += 2147483647; st[i] = word; }
                for (int i = 0; i < 310; i++) (void)next();
            }
            const uint32_t batch = 1u << 18;
            std::string bases; std::vector<uint64_t> offs;
            for (unsigned t = 0; t < n_threads; t++)
                for (uint32_t j0 = 0; j0 < num_reads; j0 += batch) {
                    const uint32_t n = std::min<uint32_t>(batch, num_reads - j0);
                    bases.assign((size_t)n * read_len, '\0'); offs.assign((size_t)n + 1, 0);
                    for (uint32_t q = 0; q < n; q++) {
                        char *rb = &bases[(size_t)q * read_len];
                        const int bucket = (int)((j0 + q) % KMAT_NULL_BUCKETS);
                        const int gc_draw = (next() % 10) + bucket * 10;
                        const float gc_pcnt = (float)(gc_draw / 100.0);
                        const unsigned num_gc = (unsigned)(gc_pcnt * (float)read_len);
                        for (unsigned i = 0; i < num_gc; ++i) rb[i] = (next() % 100) < 50 ? 'g' : 'c';
                        for (unsigned i = num_gc; i < read_len; ++i) rb[i] = (next() % 100) < 50 ? 'a' : 't';
                        for (unsigned i = 1; i < read_len; ++i) { const unsigned jj = (unsigned)next() % (i + 1); if (jj != i) std::swap(rb[i], rb[jj]); }
                        offs[q + 1] = (uint64_t)(q + 1) * read_len;
                    }
                    if ((rc = kmat_null_batch(ctxs[d], bases.data(), offs.data(), n, j0)) != KMAT_OK) return fail(rc);
                }
        } else {
            for (uint64_t i = lo; i < hi;) {
                const uint64_t t = i / num_reads, j = i % num_reads;
                const uint64_t n = std::min<uint64_t>({hi - i, (uint64_t)num_reads - j, (uint64_t)1 << 30});
                // every "thread" has its own key space; within it the generator is keyed by j, which also gives the bucket
                if ((rc = kmat_null_random(ctxs[d], seed + 0x632BE59BD9B4E019ull * t, j, (uint32_t)n, read_len)) != KMAT_OK) return fail(rc);
                i += n;
            }
        }
        uint32_t rows = 0;
        rc = kmat_null_fetch(ctxs[d], nullptr, nullptr, nullptr, 0, &rows, &nerr[d]);
        if (rc != KMAT_OK && rc != KMAT_ERR_OVERFLOW) return fail(rc);
        tids[d].resize(rows); mx[d].resize((size_t)rows * KMAT_NULL_BUCKETS); cnt[d].resize((size_t)rows * KMAT_NULL_BUCKETS);
        if ((rc = kmat_null_fetch(ctxs[d], tids[d].data(), mx[d].data(), cnt[d].data(), rows, &rows, &nerr[d])) != KMAT_OK) return fail(rc);
    };
    {
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; d++) th.emplace_back(work, d);
        for (auto &t : th) t.join();
    }
    for (size_t d = 0; d < nd; d++)
        if (rcs[d] != KMAT_OK) { std::cerr << "ERROR! device " << devs[d] << ": " << errs[d] << std::endl; return -1; }
    uint64_t bad = 0;
    for (size_t d = 0; d < nd; d++) bad += nerr[d];
    if (bad) std::cerr << "WARNING! " << bad << " reads had more than 64 candidate taxids and were left out" << std::endl;

    std::cout << "Merge phase" << std::endl;
    std::vector<const uint32_t *> tp(nd); std::vector<const float *> mp(nd); std::vector<const uint64_t *> cp(nd); std::vector<uint32_t> nr(nd);
    for (size_t d = 0; d < nd; d++) { tp[d] = tids[d].data(); mp[d] = mx[d].data(); cp[d] = cnt[d].data(); nr[d] = (uint32_t)tids[d].size(); }
    const std::string out = ofbase + ".rand_lst";
    if (kmat_null_write(out.c_str(), (int)nd, tp.data(), mp.data(), cp.data(), nr.data()) != KMAT_OK) {
        std::cout << "Could not open for writing " << out << std::endl;
        return -1;
    }
    for (size_t d = 0; d < nd; d++) { kmat_ctx_destroy(ctxs[d]); kmat_db_free(dbs[d]); }
    kmat_inputs_free(inputs);
    kmat_table_free(table);
    std::cout << "query time: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << std::endl;
    return 0;
}
