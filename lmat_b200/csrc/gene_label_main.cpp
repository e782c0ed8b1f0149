// gene_label -- drop-in host for LMAT's gene_label (src/gene_label.cpp main(), :378-713) over libkmat's C ABI.
//
// Same getopt string; input = a list (-l) of read_label .out files, one output file <ofbase><i>.out per input file
// (the reference runs one OpenMP thread per listed file, :343-366,549-575), plus <ofbase>.<x>.<q>.genesummary and
// .genesummary.min_tax_score.<b>.  The per-read work (retrieve_kmer_labels + top gene, :217-301) runs on the GPU
// through kmat_gene_batch; parsing of the five tab columns, thresholds, tallies and summaries follow the reference
// line by line (same istringstream extraction, same std::map iteration orders).
#include <getopt.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "kmat.h"

typedef std::map<uint32_t, uint32_t> hmap_t;
typedef std::map<uint32_t, float> hfmap_t;

struct Pending { std::string hdr, read; uint32_t taxid; float tax_score; };

static void usage(const char *exe) {
    std::cout << "Usage:\n" << exe << " -d <gene db file> -l <list of read_label .out files> -o <output path> -g <gene annotation table (gz)>\n"
              << "[-x <min gene score>] [-q <min k-mers>] [-b <min tax score>] [-t <threads>] [-k <kmer size>] [-V]\n";
}

int main(int argc, char *argv[]) {
    signed char c;
    int n_threads = 0, k_size = -1, min_kmer = 0;
    float min_score = 0.0f, min_tax_score = 0.0f;
    std::string genefile, kmer_db_fn, query_fn, query_fn_lst, ofbase;
    while ((c = getopt(argc, argv, "b:h:n:jye:wmpk:c:v:k:i:d:l:t:s:r o:x:f:g:z:q:aV")) != -1) {       // :392
        switch (c) {
            case 'b': min_tax_score = (float)atof(optarg); break;
            case 'g': genefile = optarg; break;
            case 'h': case 's': case 'j': case 'y': case 'p': case 'a': break;      // max_count / heap size / verbose / prn_all / ascii: no effect on the output
            case 'l': query_fn_lst = optarg; break;
            case 't': n_threads = atoi(optarg); break;
            case 'x': min_score = (float)atof(optarg); break;
            case 'q': min_kmer = atoi(optarg); break;
            case 'k': k_size = atoi(optarg); break;
            case 'i': query_fn = optarg; break;
            case 'd': kmer_db_fn = optarg; break;
            case 'o': ofbase = optarg; break;
            case 'V': std::cout << "LMAT version 1.2.4_2018a (kmat gene_label, ABI " << kmat_abi_version() << ")\n"; return 0;
            default: std::cout << "Unrecognized option: " << c << ", ignore." << std::endl;
        }
    }
    if (ofbase.empty() || kmer_db_fn.empty()) {
        std::cout << "essential arguments missing: [" << ofbase << "] [" << n_threads << "] [" << kmer_db_fn << "] [" << query_fn << "] " << std::endl;
        usage(argv[0]);
        return -1;
    }
    if (!query_fn.empty()) { std::cout << "Sorry fasta input file not yet supported" << std::endl; return 0; }       // :556-559
    std::cout << "Start kmer DB load\n";
    kmat_table *table = nullptr;
    const char *tb = getenv("KMAT_TID_BYTES");
    if (kmat_table_open(kmer_db_fn.c_str(), tb ? atoi(tb) : 4, &table) != KMAT_OK) {
        std::cout << "Error opening db file, must exit:" << kmer_db_fn << std::endl;
        std::cerr << kmat_last_error() << std::endl;
        return -1;
    }
    if (k_size < 1) k_size = kmat_table_kmer_length(table);
    std::cout << "num kmers: " << kmat_table_size(table) << " - " << k_size << std::endl;
    if (kmat_device_count() < 1) { std::cerr << "ERROR! No CUDA device: this build has no CPU path" << std::endl; return -1; }
    kmat_db *db = nullptr;
    if (kmat_db_upload(table, 0, 0, 1, &db) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }
    kmat_table_free(table);

    std::vector<std::string> files;                                                  // split_file_names (:343-366)
    {
        std::ifstream ifs(query_fn_lst.c_str());
        std::string fn;
        while (ifs >> fn) files.push_back(fn);
    }
    if (n_threads != 0 && n_threads != (int)files.size())
        std::cout << "warning, thread count overwritten (for now assume when a list of LMAT taxonomy classification files are given, a thread is created for each file)" << std::endl;
    n_threads = (int)files.size();
    std::cout << "set threads=" << n_threads << std::endl;

    std::vector<std::map<uint32_t, hfmap_t>> score_gtrackall(n_threads), score_gtrackall_tax(n_threads);
    std::vector<std::map<uint32_t, hmap_t>> gtrackall(n_threads), gtrackall_tax(n_threads);
    const size_t batch_reads = 1u << 18;
    for (int th = 0; th < n_threads; th++) {
        std::ifstream ifs(files[th].c_str());
        if (!ifs) { std::cerr << "did not open for reading: [" << files[th] << "] tid: [" << th << "]" << std::endl; return -1; }
        std::ofstream ofs((ofbase + std::to_string(th) + ".out").c_str());
        std::vector<Pending> batch;
        std::string bases; std::vector<uint64_t> offs(1, 0);
        std::vector<kmat_gene_result> res;
        auto flush = [&]() -> bool {
            if (batch.empty()) return true;
            res.resize(batch.size());
            if (kmat_gene_batch(db, bases.data(), offs.data(), (uint32_t)batch.size(), res.data()) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return false; }
            for (size_t i = 0; i < batch.size(); i++) {
                const kmat_gene_result &g = res[i];
                const Pending &p = batch[i];
                // track[taxid] etc. are created by operator[] before proc_line whatever it prints (:624-633)
                hmap_t &gtrack = gtrackall[th][p.taxid], &gtrack_tax = gtrackall_tax[th][p.taxid];
                hfmap_t &score_gtrack = score_gtrackall[th][p.taxid], &score_gtrack_tax = score_gtrackall_tax[th][p.taxid];
                if (g.status < 0) { std::cerr << "ERROR! read " << p.hdr << ": " << kmat_strerror(g.status) << std::endl; return false; }
                if (g.status != 1) continue;
                const uint32_t cnt = g.valid_kmers, gl = g.gene;
                const float gscore = g.score;
                ofs << p.hdr << "\t" << p.read << "\t" << p.taxid << " " << p.tax_score << "\t";                  // :299-300
                ofs << "\t" << -1 << " " << g.count << " " << cnt << "\t" << gl << " " << gscore << " GL" << std::endl;
                if (gscore > min_score && (signed)cnt > min_kmer) { ++gtrack[gl]; score_gtrack[gl] += gscore; }
                if (p.tax_score >= min_tax_score && gscore > min_score && (signed)cnt > min_kmer) { ++gtrack_tax[gl]; score_gtrack_tax[gl] += gscore; }
            }
            batch.clear(); bases.clear(); offs.assign(1, 0);
            return true;
        };
        bool finished = false;
        std::string line;
        while (!finished) {                                                          // :586-640
            std::getline(ifs, line);
            const long long pos = (long long)ifs.tellg();
            if (pos == -1) finished = true;
            uint32_t taxid = 0;
            const size_t p1 = line.find('\t');
            const std::string hdr = line.substr(0, p1);
            const size_t p2 = line.find('\t', p1 + 1);
            const std::string read_buff = line.substr(p1 + 1, p2 - p1 - 1);
            const size_t p3 = line.find('\t', p2 + 1);
            const std::string stats = line.substr(p2 + 1, p3 - p2 - 1);
            std::istringstream istrm2(stats.c_str());
            float score1 = 0, score2 = 0, score3 = 0;       // uninitialised in the reference; 0 here when the column does not parse
            istrm2 >> score1 >> score2 >> score3;
            if (score3 == -1) continue;                     // this read lacks valid k-mers
            const size_t p4 = line.find('\t', p3 + 1);
            const size_t p5 = line.find('\t', p4 + 1);
            const std::string taxid_w_scores = line.substr(p4 + 1, p5 - p4);
            std::istringstream istrm(taxid_w_scores.c_str());
            float tax_score = 0.0f;
            std::string match_type;
            istrm >> taxid >> tax_score >> match_type;
            if (!match_type.empty() && (match_type[0] == 'N' || match_type[0] == 'R')) taxid = 0;
            batch.push_back(Pending{hdr, read_buff, taxid, tax_score});
            bases += read_buff;
            offs.push_back(bases.size());
            if (batch.size() >= batch_reads && !flush()) return -1;
        }
        if (!flush()) return -1;
    }
    // doMerge / doMergeF (:112-185): gene id -> taxid -> count / score, threads in order
    std::map<uint32_t, std::map<uint32_t, uint32_t>> merge_cnt, merge_cnt_tax;
    std::map<uint32_t, std::map<uint32_t, float>> score_merge_cnt, score_merge_cnt_tax;
    auto merge_u = [](const std::vector<std::map<uint32_t, hmap_t>> &all, std::map<uint32_t, std::map<uint32_t, uint32_t>> &m) {
        for (auto &gt : all) for (auto &kv : gt) for (auto &gc : kv.second) m[gc.first][kv.first] += gc.second;
    };
    auto merge_f = [](const std::vector<std::map<uint32_t, hfmap_t>> &all, std::map<uint32_t, std::map<uint32_t, float>> &m) {
        for (auto &gt : all) for (auto &kv : gt) for (auto &gc : kv.second) {
            auto &slot = m[gc.first];
            auto it = slot.find(kv.first);
            if (it == slot.end()) slot[kv.first] = gc.second; else it->second += gc.second;
        }
    };
    merge_u(gtrackall, merge_cnt); merge_u(gtrackall_tax, merge_cnt_tax);
    merge_f(score_gtrackall, score_merge_cnt); merge_f(score_gtrackall_tax, score_merge_cnt_tax);

    gzFile zf = gzopen(genefile.c_str(), "rb");
    if (!zf) { std::cout << "Unable to unzip gene annotation table: " << genefile << std::endl; return -1; }
    std::ostringstream output, output_tax;
    output << ofbase << "." << min_score << "." << min_kmer << ".genesummary";
    output_tax << ofbase << "." << min_score << "." << min_kmer << ".genesummary.min_tax_score." << min_tax_score;
    std::ofstream sum_ofs(output.str().c_str()), sum_ofs_tax(output_tax.str().c_str());
    if (!sum_ofs || !sum_ofs_tax) { std::cerr << "Can't write to " << output.str() << std::endl; return -1; }
    static char buff[20000];
    while (gzgets(zf, buff, sizeof buff)) {                                          // :681-709
        size_t bl = strlen(buff);
        if (bl && buff[bl - 1] == '\n') buff[--bl] = 0;
        std::istringstream istrm(buff);
        uint32_t tid = 0, gid = 0;
        istrm >> tid >> gid;
        auto emit = [&](std::map<uint32_t, std::map<uint32_t, uint32_t>> &mc, std::map<uint32_t, std::map<uint32_t, float>> &ms, std::ofstream &o) {
            auto f = mc.find(gid);
            if (f == mc.end()) return;
            for (auto &ti : f->second) {
                const float score = ms[gid][ti.first];
                const float avg = score / (float)ti.second;
                o << avg << "\t" << ti.second << "\t" << ti.first << "\t" << buff << std::endl;
            }
        };
        emit(merge_cnt, score_merge_cnt, sum_ofs);
        emit(merge_cnt_tax, score_merge_cnt_tax, sum_ofs_tax);
    }
    gzclose(zf);
    kmat_db_free(db);
    std::cout << "query time: done" << std::endl;
    return 0;
}
