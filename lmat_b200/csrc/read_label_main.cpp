// read_label -- drop-in host for LMAT's read_label (src/read_label.cpp main(), :1328-1871) over libkmat's C ABI.
//
// Same getopt string, same required options, same <ofbase><i>.out / .fastsummary / .nomatchsum outputs and the
// same stdout milestones ("Total query time").  The per-read work (proc_line, :1211-1279) runs on the GPUs:
//   reader thread  : kmat_reader_next -> batches of reads                       (replaces the thread-0 parser, :1651-1713)
//   device workers : one per GPU, kmat_label_batch on its replica (or shard set) (replaces the OpenMP proc_line loop)
//   writer threads : one per -t "thread": batch i goes to file i mod t, in input order, so -t 1 reproduces the
//                    reference's single-thread output byte for byte and any -t gives the same multiset of lines;
//                    each writer keeps the reference's per-thread tallies (float sums in file order), merged in
//                    thread order like :1760-1800.
//   (two device workers per GPU, each with its own context; batches live in page-locked memory; the tail of every line comes
//   formatted from the device -- kmat_label_batch_text -- and the writers paste header, read and tail together)
// Environment: KMAT_DEVICES="0,2,.." (default: every visible GPU), KMAT_BATCH_READS (default 131072),
// KMAT_WORKERS_PER_GPU=1 / KMAT_NO_PINNED=1 / KMAT_HOST_FORMAT=1 (switch the three off), KMAT_CLI_TRACE=1 (busy time per stage),
// KMAT_TABLE_MODE=sharded | exchange (table split over the GPUs: probes read the owner's memory over NVLink | query k-mers
// travel over NCCL, kmat_shard_label_batch; default: replicated, or sharded when one GPU cannot hold the table),
// KMAT_READER_THREADS (FASTA files are parsed in parallel segments; default hw threads / 4, 2..8),
// KMAT_TID_BYTES (2|4: sizeof(DBTID_T) of the DB, default 2), LMAT_DIR as in the reference (:555-560).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <unordered_map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "kmat.h"

#ifndef KMAT_LMAT_VERSION
#define KMAT_LMAT_VERSION "1.2.4_2018a"     /* include/version.h:1 of the reference this build tracks */
#endif

namespace {

// Result buffers of a batch in page-locked memory (kmat_host_alloc; plain memory when pinning fails): the copies back from the
// device are then DMA transfers that overlap the kernels instead of staged pageable copies (~6 GB/s, the device worker's time).
template <typename T>
struct PinBuf {
    T *p = nullptr; size_t cap = 0; bool pinned = false;
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
    ~PinBuf() { release(); }
    void release() { if (p) { if (pinned) kmat_host_free(p); else free(p); } p = nullptr; cap = 0; }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return cap; }
    T &operator[](size_t i) { return p[i]; }
    void reserve(size_t n) {               // contents are not kept
        if (n <= cap) return;
        release();
        n += n / 8;
        p = (T *)kmat_host_alloc(n * sizeof(T)); pinned = p != nullptr;
        if (!p) p = (T *)malloc(n * sizeof(T));
        cap = p ? n : 0;
    }
};
struct Batch {
    uint64_t seq = 0;
    kmat_read_batch *rb = nullptr;
    PinBuf<kmat_read_result> res;
    PinBuf<kmat_pair> cands, lin;
    PinBuf<char> text; PinBuf<uint64_t> tref; bool has_text = false;    // K5: the tails formatted on the device (kmat_label_batch_text)
    int rc = 0;
    std::string err;
};

template <typename T>
class Channel {            // bounded FIFO; close() wakes everybody
  public:
    explicit Channel(size_t cap) : cap_(cap) {}
    bool push(T v) {
        std::unique_lock<std::mutex> l(m_);
        cv_space_.wait(l, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return false;
        q_.push_back(std::move(v));
        cv_item_.notify_one();
        return true;
    }
    bool pop(T &out) {
        std::unique_lock<std::mutex> l(m_);
        cv_item_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.erase(q_.begin());
        cv_space_.notify_one();
        return true;
    }
    void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_item_.notify_all(); cv_space_.notify_all(); }
  private:
    std::mutex m_;
    std::condition_variable cv_item_, cv_space_;
    std::vector<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// Per output "thread": batches arrive in any order (several GPUs), are written in seq order.
struct Writer {
    std::mutex m;
    std::condition_variable cv;
    std::map<uint64_t, Batch *> ready;
    uint64_t next_seq = 0;       // first seq this writer expects (index, then += n_writers)
    bool done = false;
    std::map<uint32_t, int> track_match;          // read_label.cpp:1606-1608
    std::map<uint32_t, float> track_tscore;
    std::map<int, int> track_nomatch;             // 1 ReadTooShort, 2 NoDbHits, 3 LowScore
};

void usage(const char *exe) {
    std::cout << "==============================================\n"
                 "  kmat read_label -- B200-native drop-in for\n"
                 "  the Livermore Metagenomics Analysis Toolkit\n"
                 "==============================================\n\n"
                 "Taxonomic classification module usage:\n"
              << exe << " -d <input db file> -i <query fasta file | -> -t <number of output threads>\n"
                 "-o <output path> -e <depth file> -c <tax tree file> [-f <32-to-16-bit id map>] [-w <rank map>]\n"
                 "[-u <rank/name table>] [-n <null model list>] [-m <numeric rank file>] [-g <tid-cutoff>]\n"
                 "[-r <low-number plasmid list>] [-x <min score>] [-j <min valid k-mers>] [-z <min found k-mers>]\n"
                 "[-b <sdiff>] [-l <human bias>] [-k <kmer size>] [-v <pct cutoff>] [-q:fastq] [-p:print all candidates]\n"
                 "[-a:hide read] [-s:permissive] [-h:turn phiX screening off] [-y:verbose]\n"
                 "[-V:print version and exit] [-H:print this usage help and exit]\n";
}

std::string fmt_g(float v) {                 // ostream << float at default precision == "%g"
    char b[64];
    snprintf(b, sizeof b, "%g", (double)v);
    return b;
}

struct SimpleCmpDesc {                        // SimpleCmp, read_label.cpp:153-157
    bool operator()(const std::pair<uint32_t, float> &a, const std::pair<uint32_t, float> &b) const { return a.second > b.second; }
};

}  // namespace

int main(int argc, char *argv[]) {
    int k_size = -1, n_threads = 0;
    float min_score = 0.0f;
    std::string rank_map_file, rank_ids, kmer_db_fn, query_fn, ofbase, tax_tree_fn, depth_file, rand_hits_file, rank_table_file,
        id_bit_conv_fn, low_num_plasmid_file;
    bool fastq = false, prn_read = true, prn_all = false, verbose = false;
    kmat_opts opt;
    kmat_opts_default(&opt);
    int c;
    while ((c = getopt(argc, argv, "u:ahn:j:b:ye:w:pk:c:v:k:i:d:l:t:r:sm:o:x:f:g:z:qVH")) != -1) {     // :1351
        switch (c) {
            case 'h': opt.phix_screen = 0; break;
            case 'r': low_num_plasmid_file = optarg; break;
            case 'f': id_bit_conv_fn = optarg; break;
            case 'j': opt.min_kmer = atoi(optarg); break;
            case 'z': opt.min_fnd_kmer = atoi(optarg); break;
            case 'u': rank_ids = optarg; break;
            case 'x': min_score = (float)atof(optarg); break;
            case 'a': prn_read = false; break;
            case 'w': rank_map_file = optarg; break;
            case 's': opt.permissive = 1; break;
            case 'n': rand_hits_file = optarg; break;
            case 'b': opt.sdiff = (float)atof(optarg); break;
            case 'l': opt.hbias = (float)atof(optarg); break;
            case 'y': verbose = true; break;
            case 'e': depth_file = optarg; break;
            case 'q': fastq = true; break;
            case 'p': prn_all = true; break;
            case 'm': rank_table_file = optarg; break;
            case 't': n_threads = atoi(optarg); break;
            case 'v': break;                                   // threshold: parsed and never used by the reference (:1336,1409)
            case 'c': tax_tree_fn = optarg; break;
            case 'k': k_size = atoi(optarg); break;
            case 'g': opt.max_count = (int32_t)(uint16_t)atoi(optarg); break;      // uint16_t max_count (:1346,1422)
            case 'i': query_fn = optarg; break;
            case 'd': kmer_db_fn = optarg; break;
            case 'o': ofbase = optarg; break;
            case 'V': std::cout << "LMAT version " << KMAT_LMAT_VERSION << " (kmat B200 host, ABI " << kmat_abi_version() << ")" << std::endl; return 0;
            case 'H': usage(argv[0]); return 0;
            default: std::cerr << "WARNING! Unrecognized option " << (char)c << " to ignore." << std::endl;
        }
    }
    if (depth_file.empty()) std::cerr << "ERROR! Missing depth_file" << std::endl;
    if (ofbase.empty()) std::cerr << "ERROR! Missing ofbase" << std::endl;
    if (n_threads == 0) std::cerr << "ERROR! Missing n_threads" << std::endl;
    if (kmer_db_fn.empty()) std::cerr << "ERROR! Missing kmer_db_fn" << std::endl;
    if (query_fn.empty()) std::cerr << "ERROR! Missing query_fn" << std::endl;
    if (depth_file.empty() || ofbase.empty() || n_threads <= 0 || kmer_db_fn.empty() || query_fn.empty()) {
        std::cerr << "Params: " << ofbase << " " << n_threads << " " << kmer_db_fn << " " << query_fn << " " << depth_file << std::endl;
        usage(argv[0]);
        return -1;
    }
    opt.min_score = min_score;
    opt.want_lineage = prn_all ? 0 : 1;
    if (verbose) std::cerr << "WARNING! -y: the per-k-mer debug traces on stdout are not produced by the GPU path; its effect on the -p list (candidates with a negative score are printed too, read_label.cpp:901) is." << std::endl;
    const int prn_mode = prn_all ? (verbose ? 2 : 1) : 0;

    std::cout << "=== LMAT === read_label === ver. " << KMAT_LMAT_VERSION << " === kmat/B200 ===" << std::endl;
    std::cout << "Start kmer DB load..." << std::endl;
    const char *tb = getenv("KMAT_TID_BYTES");
    const int tid_bytes = tb ? atoi(tb) : 2;
    kmat_table *table = nullptr;
    if (kmat_table_open(kmer_db_fn.c_str(), tid_bytes, &table) != KMAT_OK) {
        std::cerr << "Error: unable to open kmer db [" << kmer_db_fn << "]: " << kmat_last_error() << std::endl;
        return -1;
    }
    const int db_k = kmat_table_kmer_length(table);
    if (k_size < 1) k_size = db_k;
    if (k_size != db_k) std::cerr << "WARNING! -k " << k_size << " differs from the database's k-mer length " << db_k << "; using " << db_k << std::endl;
    std::cout << "Mapping flat table. Num of k-mers: " << kmat_table_size(table) << " of size " << db_k << std::endl;
    if (db_k <= 0) { std::cerr << "ERROR! Unable to read database, k-mer size=" << db_k << std::endl; return -1; }

    // devices
    std::vector<int> devs;
    if (const char *dv = getenv("KMAT_DEVICES")) {
        std::stringstream ss(dv);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
    } else for (int i = 0; i < kmat_device_count(); i++) devs.push_back(i);
    if (devs.empty()) { std::cerr << "ERROR! No CUDA device: this build has no CPU path (" << kmat_last_error() << ")" << std::endl; return -1; }

    std::cout << "Reading taxonomy tree " << tax_tree_fn << std::endl;
    std::cout << "Reading taxonomy depth " << depth_file << std::endl;
    if (!id_bit_conv_fn.empty()) std::cout << "Loading map file " << id_bit_conv_fn << "... ";
    kmat_inputs *inputs = nullptr;
    const char *lmat_dir = getenv("LMAT_DIR");
    auto nz = [](const std::string &s) { return s.empty() ? nullptr : s.c_str(); };
    int rc = kmat_inputs_load(nz(tax_tree_fn), depth_file.c_str(), nz(rank_map_file), nz(id_bit_conv_fn), nz(rank_table_file),
                              nz(low_num_plasmid_file), nz(rand_hits_file), lmat_dir, &inputs);
    if (rc != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }
    if (!id_bit_conv_fn.empty()) std::cout << "OK!" << std::endl;

    const auto t_start = std::chrono::steady_clock::now();                       // StopWatch clock (:1604-1605)
    std::vector<kmat_db *> dbs(devs.size(), nullptr);
    std::vector<kmat_ctx *> ctxs(devs.size(), nullptr);
    std::vector<kmat_comm *> comms(devs.size(), nullptr);
    // replicated table: a second context per GPU (its own streams and batch buffers over the same table) driven by a second worker
    // thread, so that one batch's copies overlap the other's kernels (KMAT_WORKERS_PER_GPU=1 switches it off)
    std::vector<kmat_ctx *> ctxs2(devs.size(), nullptr);
    bool exchange = false;
    {
        std::vector<std::thread> up;
        std::vector<int> urc(devs.size(), 0);
        std::vector<std::string> uerr(devs.size());
        // KMAT_TABLE_MODE=sharded (or a table that one GPU cannot hold): GPU d keeps the k-mers with kmat_shard_of() == d
        // and every GPU maps the others' shards (peer access over NVLink); the probe kernel then reads each bucket from
        // its owner.  Default: the whole table on every GPU.
        const char *tm = getenv("KMAT_TABLE_MODE");
        exchange = tm && strcmp(tm, "exchange") == 0;            // sharded table, query k-mers exchanged over NCCL (kmat_shard_label_batch)
        bool split = tm && (strcmp(tm, "sharded") == 0 || exchange);
        if (!tm && devs.size() > 1) {
            uint64_t free_b = 0, total_b = 0;
            if (kmat_device_memory(devs[0], &free_b, &total_b) == KMAT_OK) {
                const double need = (double)kmat_table_device_bytes(table, 1);            // both table levels from the real geometry, list pools, upload temporaries, batch buffers
                if (need > (double)free_b) { split = true; std::cout << "Table does not fit one GPU (" << need / 1e9 << " GB needed): sharding it over " << devs.size() << " GPUs" << std::endl; }
            }
        }
        if (split && devs.size() > 16) { std::cerr << "ERROR! at most 16 table shards" << std::endl; return -1; }
        const int n_sh = split ? (int)devs.size() : 1;
        for (size_t d = 0; d < devs.size(); d++)
            up.emplace_back([&, d] {
                urc[d] = kmat_db_upload(table, devs[d], split ? (int)d : 0, n_sh, &dbs[d]);
                if (urc[d] == KMAT_OK) urc[d] = kmat_ctx_create(dbs[d], inputs, &opt, &ctxs[d]);
                const char *wpg = getenv("KMAT_WORKERS_PER_GPU");
                if (urc[d] == KMAT_OK && !split && !(wpg && atoi(wpg) == 1)) urc[d] = kmat_ctx_create(dbs[d], inputs, &opt, &ctxs2[d]);
                if (urc[d] != KMAT_OK) uerr[d] = kmat_last_error();
            });
        for (auto &t : up) t.join();
        for (size_t d = 0; d < devs.size(); d++)
            if (urc[d] != KMAT_OK) { std::cerr << "ERROR! device " << devs[d] << ": " << uerr[d] << std::endl; return -1; }
        if (split && exchange) {
            // one NCCL communicator per GPU thread; the unique id is shared through this process's memory
            unsigned char uid[128];
            if (kmat_comm_unique_id(uid) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }
            std::vector<std::thread> ci;
            std::vector<int> crc(devs.size(), 0);
            std::vector<std::string> cerr_(devs.size());
            for (size_t d = 0; d < devs.size(); d++)
                ci.emplace_back([&, d] { crc[d] = kmat_comm_init(devs[d], (int)d, n_sh, uid, &comms[d]); if (crc[d] != KMAT_OK) cerr_[d] = kmat_last_error(); });
            for (auto &t : ci) t.join();
            for (size_t d = 0; d < devs.size(); d++)
                if (crc[d] != KMAT_OK) { std::cerr << "ERROR! device " << devs[d] << ": " << cerr_[d] << std::endl; return -1; }
            std::cout << "Table sharded over " << n_sh << " GPUs (query k-mers exchanged over NCCL)" << std::endl;
        } else if (split && n_sh > 1) {
            std::vector<kmat_peer_info> blobs(devs.size());
            for (size_t d = 0; d < devs.size(); d++)
                if (kmat_ctx_peer_export(ctxs[d], &blobs[d]) != KMAT_OK) { std::cerr << "ERROR! device " << devs[d] << ": " << kmat_last_error() << std::endl; return -1; }
            for (size_t d = 0; d < devs.size(); d++)
                if (kmat_ctx_peer_attach(ctxs[d], n_sh, blobs.data()) != KMAT_OK) { std::cerr << "ERROR! device " << devs[d] << ": " << kmat_last_error() << std::endl; return -1; }
            std::cout << "Table sharded over " << n_sh << " GPUs (direct peer reads)" << std::endl;
        }
    }
    kmat_table_free(table);
    const auto t_query = std::chrono::steady_clock::now();

    kmat_reader *reader = nullptr;
    const char *rt = getenv("KMAT_READER_THREADS");
    const int reader_threads = rt ? std::max(1, atoi(rt)) : (int)std::max(2u, std::min(8u, std::thread::hardware_concurrency() / 4));
    if (kmat_reader_open_mt(query_fn.c_str(), fastq ? 1 : 0, reader_threads, &reader) != KMAT_OK) {
        std::cerr << "ERROR! Did not open for reading: " << query_fn << std::endl;
        return -1;
    }
    std::cout << "Classifing reads on " << devs.size() << " GPU(s), writing " << n_threads << " .out files..." << std::endl;

    const char *be = getenv("KMAT_BATCH_READS");
    const uint32_t batch_reads = be ? (uint32_t)std::max(1, atoi(be)) : 131072u;
    const uint64_t batch_bases = (uint64_t)batch_reads * 400;
    // batches in flight: one being read, two per GPU (one queued), one per writer + one waiting for it; the buffers are
    // reused, so a small pool also keeps the working set (and its first-touch page faults) small
    const size_t n_inflight = 4 * devs.size() + (size_t)n_threads + 3;
    Channel<Batch *> free_q(n_inflight + 1), work_q(n_inflight + 1);
    std::vector<Batch> pool(n_inflight);
    for (auto &b : pool) { b.rb = getenv("KMAT_NO_PINNED") ? kmat_read_batch_new() : kmat_read_batch_new_pinned(); free_q.push(&b); }
    std::vector<Writer> writers(n_threads);
    for (int w = 0; w < n_threads; w++) writers[w].next_seq = (uint64_t)w;
    std::atomic<uint64_t> reads_loaded{0};
    // KMAT_CLI_TRACE=1: busy seconds of every stage (reader: parsing; device workers: inside kmat_label_batch; writers:
    // formatting + fwrite), printed to stderr at the end -- which stage bounds the file-to-file rate
    const bool cli_trace = getenv("KMAT_CLI_TRACE") != nullptr;
    std::atomic<uint64_t> ns_reader{0}, ns_device{0}, ns_format{0}, ns_write{0};
    auto now_ns = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::atomic<int> failed{0};
    std::string fail_msg;
    std::mutex fail_m;
    auto fail = [&](const std::string &m) { std::lock_guard<std::mutex> l(fail_m); if (!failed.exchange(1)) fail_msg = m; };

    std::thread reader_thr([&] {
        uint64_t seq = 0;
        for (;;) {
            Batch *b;
            if (!free_q.pop(b)) break;
            const uint64_t t0 = cli_trace ? now_ns() : 0;
            const int64_t n = kmat_reader_next(reader, batch_reads, batch_bases, b->rb);
            if (cli_trace) ns_reader += now_ns() - t0;
            if (n < 0) { fail(kmat_last_error()); break; }
            if (n == 0) break;
            reads_loaded += (uint64_t)n;
            b->seq = seq++;
            if (!work_q.push(b)) break;
        }
        std::cout << "Total reads loaded: " << reads_loaded.load() << std::endl;
        work_q.close();
    });

    auto deliver = [&](Batch *b) {
        Writer &w = writers[b->seq % (uint64_t)n_threads];
        std::lock_guard<std::mutex> l(w.m);
        w.ready[b->seq] = b;
        w.cv.notify_one();
    };
    // exchange mode: the GPU workers move in lockstep -- every super-step each takes one batch (or none) and all of them
    // enter the collective pass; they stop together when no worker got a batch
    struct Lockstep {
        std::mutex m; std::condition_variable cv; int waiting = 0, got = 0; uint64_t gen = 0; bool last_any = false;
        bool any(bool mine, int n) {             // barrier + OR
            std::unique_lock<std::mutex> l(m);
            const uint64_t g = gen;
            got += mine ? 1 : 0;
            if (++waiting == n) { last_any = got > 0; waiting = 0; got = 0; gen++; cv.notify_all(); return last_any; }
            cv.wait(l, [&] { return gen != g; });
            return last_any;
        }
    } lockstep;
    std::vector<std::thread> dev_thr;
    // K5: the tails of the output lines come formatted from the device (KMAT_HOST_FORMAT=1: the host formatter for every read)
    const bool device_text = !exchange && !getenv("KMAT_HOST_FORMAT");
    const size_t n_dev_workers = devs.size() * 2;
    for (size_t dw = 0; dw < n_dev_workers; dw++) {
        const size_t d = dw / 2;
        kmat_ctx *const my_ctx = (dw & 1) ? ctxs2[d] : ctxs[d];
        if (!my_ctx) continue;
        dev_thr.emplace_back([&, d, my_ctx] {
            Batch *b;
            for (;;) {
                bool have = work_q.pop(b);
                if (exchange) { if (!lockstep.any(have, (int)devs.size())) break; }
                else if (!have) break;
                const char *bases = nullptr; const uint64_t *offs = nullptr; uint32_t n = 0;
                static const uint64_t zero_off[1] = {0};
                if (have) {
                    kmat_read_batch_view(b->rb, &bases, &offs, nullptr, nullptr, &n, nullptr);
                    b->res.reserve(n);
                    b->cands.reserve((size_t)n * 20 + 1024);
                    if (opt.want_lineage) b->lin.reserve((size_t)n * 20 + 1024);
                    b->has_text = false;
                    if (device_text) { b->text.reserve((size_t)n * 256 + 4096); b->tref.reserve((size_t)n + 1); if (!b->text.data() || !b->tref.data()) { std::cerr << "ERROR! out of host memory" << std::endl; _exit(1); } }
                    if (!b->res.data() || !b->cands.data() || (opt.want_lineage && !b->lin.data())) { std::cerr << "ERROR! out of host memory" << std::endl; _exit(1); }
                }
                uint64_t nc = 0, nl = 0;
                int rc = KMAT_OK;
                const uint64_t t_dev0 = cli_trace ? now_ns() : 0;
                for (int attempt = 0; attempt < 3; attempt++) {
                    if (exchange) {
                        rc = kmat_shard_label_batch(my_ctx, comms[d], have ? bases : "", have ? offs : zero_off, n, have ? b->res.data() : nullptr,
                                                    have ? b->cands.data() : nullptr, have ? b->cands.size() : 0, &nc,
                                                    have && opt.want_lineage ? b->lin.data() : nullptr, have ? b->lin.size() : 0, &nl);
                        // a retry is collective too: every worker learns whether any of them overflowed
                        const bool again = lockstep.any(rc == KMAT_ERR_OVERFLOW, (int)devs.size());
                        if (have && rc == KMAT_ERR_OVERFLOW) { b->cands.reserve(nc); b->lin.reserve(nl); }
                        if (!again) break;
                    } else {
                        if (device_text) {
                            uint64_t nt = 0;
                            rc = kmat_label_batch_text(my_ctx, bases, offs, n, b->res.data(), b->cands.data(), b->cands.size(), &nc,
                                                       opt.want_lineage ? b->lin.data() : nullptr, b->lin.size(), &nl,
                                                       prn_mode, b->text.data(), b->text.size(), &nt, b->tref.data());
                            b->has_text = true;
                        } else
                        rc = kmat_label_batch(my_ctx, bases, offs, n, b->res.data(), b->cands.data(), b->cands.size(), &nc,
                                              opt.want_lineage ? b->lin.data() : nullptr, b->lin.size(), &nl);
                        if (rc != KMAT_ERR_OVERFLOW) break;
                        b->cands.reserve(nc);
                        b->lin.reserve(nl);
                    }
                }
                if (cli_trace) ns_device += now_ns() - t_dev0;
                if (have) {
                    b->rc = rc;
                    if (b->rc != KMAT_OK) { b->err = kmat_last_error(); fail("device " + std::to_string(devs[d]) + ": " + b->err); }
                    deliver(b);
                } else if (rc != KMAT_OK && rc != KMAT_ERR_OVERFLOW) fail("device " + std::to_string(devs[d]) + ": " + kmat_last_error());
            }
        });
    }

    std::vector<std::thread> wr_thr;
    for (int wi = 0; wi < n_threads; wi++)
        wr_thr.emplace_back([&, wi] {
            Writer &w = writers[wi];
            const std::string ofname = ofbase + std::to_string(wi) + ".out";         // :1642-1647
            FILE *ofs = fopen(ofname.c_str(), "w");
            if (!ofs) fail("could not open for writing " + ofname);
            std::vector<char> out;
            std::unordered_map<uint32_t, std::pair<int, float>> tally;               // track_taxids / track_tscores of this "thread"
            // the map's nodes never move: a small direct-mapped cache of tid -> node saves the hash lookup for the tids that keep coming
            struct TallySlot { uint32_t tid = 0; std::pair<int, float> *v = nullptr; };
            std::vector<TallySlot> tcache(4096);
            int nomatch[4] = {0, 0, 0, 0};                                            // track_nomatch
            for (;;) {
                Batch *b = nullptr;
                {
                    std::unique_lock<std::mutex> l(w.m);
                    w.cv.wait(l, [&] { return w.done || w.ready.count(w.next_seq); });
                    auto it = w.ready.find(w.next_seq);
                    if (it == w.ready.end()) break;
                    b = it->second;
                    w.ready.erase(it);
                    w.next_seq += (uint64_t)n_threads;
                }
                if (b->rc == KMAT_OK && ofs) {
                    const char *bases, *hdrs; const uint64_t *offs, *hoffs; uint32_t n;
                    kmat_read_batch_view(b->rb, &bases, &offs, &hdrs, &hoffs, &n, nullptr);
                    // upper bound of the text of this batch: header + read + tail (160 fixed + 40 per printed pair)
                    size_t bound = (size_t)hoffs[n] + (prn_read ? (size_t)offs[n] : (size_t)n) + (size_t)n * 164;
                    for (uint32_t i = 0; i < n; i++) bound += (size_t)(prn_all ? b->res[i].n_cand : b->res[i].n_lin) * 40;
                    if (out.size() < bound) out.resize(bound + bound / 8);
                    char *p = out.data();
                    const uint64_t t_f0 = cli_trace ? now_ns() : 0;
                    for (uint32_t i = 0; i < n; i++) {
                        const kmat_read_result &r = b->res[i];
                        const size_t hl = (size_t)(hoffs[i + 1] - hoffs[i]), rl = (size_t)(offs[i + 1] - offs[i]);
                        memcpy(p, hdrs + hoffs[i], hl); p += hl;                                  // :1733-1738
                        *p++ = '\t';
                        if (prn_read) { memcpy(p, bases + offs[i], rl); p += rl; } else *p++ = 'X';
                        *p++ = '\t';
                        if (r.status == KMAT_ST_ERROR) { fail("read " + std::string(hdrs + hoffs[i], hdrs + hoffs[i + 1]) + ": " + kmat_strerror(r.err)); continue; }
                        const uint64_t tr = b->has_text ? b->tref[i] : KMAT_TEXT_ON_HOST;
                        if (tr != KMAT_TEXT_ON_HOST) {                           // formatted on the device: paste it in
                            const size_t tl = (size_t)(tr & ((1ull << KMAT_TEXT_LEN_BITS) - 1));
                            memcpy(p, b->text.data() + (tr >> KMAT_TEXT_LEN_BITS), tl);
                            p += tl;
                        } else {
                            const int tn = kmat_format_tail(&r, b->cands.data(), b->lin.data(), prn_mode, p, (size_t)(out.data() + out.size() - p));
                            if (tn < 0) { fail("formatting failed"); continue; }
                            p += tn;
                        }
                        switch (kmat_tally_class(&r, min_score, opt.min_kmer)) {                 // :1217-1277
                            case 0: {
                                // per tid: count and the float sum of the scores in arrival order (hashed here, copied into
                                // the ordered maps of the merge step when the writer ends: same per-tid additions, same order)
                                TallySlot &ts = tcache[(r.tid * 0x9E3779B1u) >> 20];
                                if (ts.v && ts.tid == r.tid) { ts.v->first += 1; ts.v->second += r.score; break; }
                                auto ins = tally.try_emplace(r.tid, 1, r.score);
                                if (!ins.second) { ins.first->second.first += 1; ins.first->second.second += r.score; }
                                ts.tid = r.tid; ts.v = &ins.first->second;
                                break;
                            }
                            case 1: nomatch[1] += 1; break;
                            case 2: nomatch[2] += 1; break;
                            case 3: nomatch[3] += 1; break;
                            default: break;
                        }
                    }
                    const size_t out_n = (size_t)(p - out.data());
                    const uint64_t t_w0 = cli_trace ? now_ns() : 0;
                    if (fwrite(out.data(), 1, out_n, ofs) != out_n) fail("write failed: " + ofname);
                    if (cli_trace) { ns_format += t_w0 - t_f0; ns_write += now_ns() - t_w0; }
                }
                free_q.push(b);
            }
            for (const auto &kv : tally) { w.track_match[kv.first] = kv.second.first; w.track_tscore[kv.first] = kv.second.second; }
            for (int k = 1; k <= 3; k++) if (nomatch[k]) w.track_nomatch[k] = nomatch[k];
            if (ofs) fclose(ofs);
        });

    reader_thr.join();
    for (auto &t : dev_thr) t.join();
    for (auto &w : writers) { std::lock_guard<std::mutex> l(w.m); w.done = true; w.cv.notify_all(); }
    for (auto &t : wr_thr) t.join();
    free_q.close();
    kmat_reader_close(reader);
    for (auto &b : pool) kmat_read_batch_free(b.rb);
    if (failed.load()) { std::cerr << "ERROR! " << fail_msg << std::endl; return -1; }
    if (cli_trace)
        fprintf(stderr, "[kmat cli trace] busy seconds: reader %.3f (%d parser threads), device workers %.3f (%zu), writers: format %.3f + write %.3f (%d threads)\n",
                ns_reader.load() * 1e-9, reader_threads, ns_device.load() * 1e-9, dev_thr.size(), ns_format.load() * 1e-9, ns_write.load() * 1e-9, n_threads);

    std::cout << "Finished classifing reads, doing final steps sequentially..." << std::endl;
    // merge the per-thread tallies in thread order (:1760-1800)
    std::map<uint32_t, int> merge_count;
    std::map<uint32_t, float> merge_score;
    std::map<int, int> nomatch_merge_count;
    for (auto &w : writers) {
        for (auto &kv : w.track_tscore) { auto it = merge_score.find(kv.first); if (it == merge_score.end()) merge_score.insert(kv); else it->second += kv.second; }
        for (auto &kv : w.track_match) merge_count[kv.first] += kv.second;
        for (auto &kv : w.track_nomatch) nomatch_merge_count[kv.first] += kv.second;
    }
    std::vector<std::pair<uint32_t, float>> sort_val(merge_score.begin(), merge_score.end());
    std::set<uint32_t> cand_tid;
    for (auto &p : sort_val) cand_tid.insert(p.first);
    std::map<uint32_t, std::string> save_id;                                      // names from -u (:1812-1835)
    if (!rank_ids.empty()) {
        std::ifstream tax_strm(rank_ids.c_str());
        std::string proc;
        while (std::getline(tax_strm, proc)) {
            std::vector<char> buf(proc.begin(), proc.end());
            buf.push_back('\0');
            char *save = nullptr;
            for (char *val = strtok_r(buf.data(), "=,", &save); val; val = strtok_r(nullptr, "=,", &save)) {
                if (strcmp(val, "taxid") == 0) {
                    val = strtok_r(nullptr, "=,", &save);
                    if (!val) break;
                    const uint32_t cid = (uint32_t)strtoul(val, nullptr, 10);
                    if (cand_tid.count(cid)) {
                        const size_t pos = proc.rfind('\t');
                        save_id.insert(std::make_pair(cid, pos == std::string::npos ? proc : proc.substr(pos + 1)));
                    }
                    break;
                }
            }
        }
    }
    const std::string base = ofbase + "." + fmt_g(min_score) + "." + std::to_string(opt.min_kmer);
    {
        std::ofstream sum_ofs((base + ".fastsummary").c_str());
        if (!sum_ofs) { std::cerr << "ERROR! Could not open for writing " << base << ".fastsummary" << std::endl; return -1; }
        std::cout << "Writing FastSummary file in " << base << ".fastsummary" << std::endl;
        std::sort(sort_val.begin(), sort_val.end(), SimpleCmpDesc());                   // :1844
        for (auto &p : sort_val) sum_ofs << p.second << "\t" << merge_count[p.first] << "\t" << p.first << "\t" << save_id[p.first] << std::endl;
    }
    {
        std::ofstream nom_ofs((base + ".nomatchsum").c_str());
        if (!nom_ofs) { std::cerr << "ERROR! Could not open for writing " << base << ".nomatchsum" << std::endl; return -1; }
        std::cout << "Writing NoMatchSum file in " << base << ".nomatchsum" << std::endl;
        static const char *names[] = {"Error", "ReadTooShort", "NoDbHits", "LowScore"};
        for (auto &kv : nomatch_merge_count) nom_ofs << names[kv.first] << "\t" << kv.second << std::endl;
    }
    for (size_t d = 0; d < devs.size(); d++) { kmat_comm_free(comms[d]); kmat_ctx_destroy(ctxs[d]); kmat_ctx_destroy(ctxs2[d]); kmat_db_free(dbs[d]); }
    kmat_inputs_free(inputs);
    const auto t_end = std::chrono::steady_clock::now();
    const double q = std::chrono::duration<double>(t_end - t_query).count(), up = std::chrono::duration<double>(t_query - t_start).count();
    std::cout << "Table upload time: " << up << " sec (" << devs.size() << " device(s))" << std::endl;
    std::cout << "DONE! Total query time: " << q << " sec = " << q / 60 << " min" << std::endl;
    return 0;
}
