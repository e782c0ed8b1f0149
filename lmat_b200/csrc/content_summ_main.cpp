// content_summ -- drop-in host for LMAT's content_summ (src/content_summ.cpp main(), :235-524) over libkmat's C ABI.
//
// Same getopt string and meanings (bin/run_cs.sh:148 is the canonical invocation): -l <.fastsummary> -f <list of read_label
// .out files> -c <taxonomy> -r <rank table> -k <k values> -a <ranks to count k-mers for> [-p <plasmid ids>] [-v <min score>]
// [-s skip human] [-n human regions] -o <output>.  Writes <output> (the indented tree of called taxids with read counts)
// and <output>.<rank>_kmer_cov (distinct k-mer counts and count histograms per taxid and k).
//
// The per-read k-mer work (every distinct canonical k-mer of a read counts once for the read's taxid, for every k of the
// list; :114-160) and the merge behind compKmerCov (:527-571) run on the GPU through kmat_kcov_*; parsing, the choice of
// the taxid a read counts for, the tree walk and the output formats follow the reference, its quirks included (the
// shadowed stream pointer at :503 leaves out the coverage of the first taxid of every rank).
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "kmat.h"

typedef uint32_t TID;

struct Tree {                                   // what TaxTree gives content_summ: parents and names (TaxTree.hpp:24-57, TaxNode.hpp:131-147)
    std::unordered_map<TID, TID> parent;
    std::unordered_map<TID, std::string> name;
    bool load(const char *fn) {
        std::ifstream in(fn);
        if (!in.is_open()) return false;
        std::string line;
        std::getline(in, line); std::getline(in, line);
        int count; in >> count; std::getline(in, line);
        while (true) {
            const std::streampos p = in.tellg();
            if (in.eof() || !in.good() || (int)p == -1) break;
            TID id = 0, ct = 0, ch = 0, par = 0;
            in >> id >> ct;
            for (TID j = 0; j < ct; j++) in >> ch;
            in >> par;
            std::string nm;
            std::getline(in, nm); std::getline(in, nm);
            parent[id] = par; name[id] = nm;
        }
        return true;
    }
    void path_to_root(TID tid, std::vector<TID> &out) const {     // strict ancestors, nearest first (TaxTree.hpp:60-91)
        out.clear();
        auto it = parent.find(tid);
        if (it == parent.end()) return;
        TID cur = tid;
        while (it->second != cur) {
            cur = it->second;
            out.push_back(cur);
            it = parent.find(cur);
            if (it == parent.end()) { std::cerr << "failed to find parent TaxNode for taxid " << cur << std::endl; exit(-1); }
        }
    }
    std::string get_name(TID t) const { auto it = name.find(t); return it == name.end() ? "" : it->second; }
};

static std::unordered_set<int> g_plasmids;
static bool is_plasmid(TID t) { return (t >= 10000000 && t < 11000000) || g_plasmids.count((int)t); }   // :46
static bool is_human(TID t) { return t == 9606 || t == 63221 || t == 741158; }                            // tid_checks.hpp:15-28

int main(int argc, char *argv[]) {
    signed char c;
    float threshold = 0.0f;
    std::string query_fn_lst, lmat_sum, ofbase, tax_tree_fn, rank_table_file, low_num_plasmid_file, k_size_str, rank_check_str;
    bool skipHuman = false, doHumanReg = false;
    while ((c = getopt(argc, argv, "m:f:a:h:njb:ye:wp:k:c:v:k:i:d:l:t:sr:o:x:f:q:V")) != -1) {               // :245
        switch (c) {
            case 'n': doHumanReg = true; break;
            case 'a': rank_check_str = optarg; break;
            case 'p': low_num_plasmid_file = optarg; break;
            case 's': skipHuman = true; break;
            case 'r': rank_table_file = optarg; break;
            case 'y': break;
            case 'l': lmat_sum = optarg; break;
            case 'v': threshold = (float)atof(optarg); break;
            case 'c': tax_tree_fn = optarg; break;
            case 'k': k_size_str = optarg; break;
            case 'f': query_fn_lst = optarg; break;
            case 'i': break;
            case 'o': ofbase = optarg; break;
            case 'V': std::cout << "LMAT version 1.2.4_2018a (kmat content_summ, ABI " << kmat_abi_version() << ")\n"; return 0;
            default: std::cout << "Unrecognized option: " << c << ", ignore." << std::endl;
        }
    }
    std::vector<int32_t> k_size;
    if (k_size_str.empty()) k_size = {8, 10, 14, 20};                                                     // :297-302
    else {
        std::stringstream ss(k_size_str);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) { std::istringstream is(tok); unsigned v = 0; is >> v; k_size.push_back((int32_t)v); }
    }
    std::set<std::string> rank_check;
    {
        std::stringstream ss(rank_check_str);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) { std::cout << "rank store: [" << tok << "]" << std::endl; rank_check.insert(tok); }
    }
    for (size_t i = 0; i < k_size.size(); i++) std::cout << "track k size=" << k_size[i] << std::endl;
    if (!low_num_plasmid_file.empty()) {
        std::ifstream ifs(low_num_plasmid_file.c_str());
        if (!ifs) std::cerr << "Unexpected reading error: " << low_num_plasmid_file << std::endl;
        TID pid;
        while (ifs >> pid) g_plasmids.insert((int)pid);
    }
    std::unordered_map<TID, std::string> rank_table;
    if (!rank_table_file.empty()) {
        std::ifstream ifs(rank_table_file.c_str());
        TID tid; std::string rank;
        while (ifs >> tid >> rank) rank_table.insert(std::make_pair(tid, rank));
    }
    std::vector<std::string> files;
    {
        std::ifstream ifs(query_fn_lst.c_str());
        std::string fn;
        while (ifs >> fn) files.push_back(fn);
    }
    std::cout << "set threads=" << files.size() << std::endl;
    if (files.empty()) { std::cerr << "no input files in [" << query_fn_lst << "]" << std::endl; return -1; }
    std::cout << "Read taxonomy tree: " << tax_tree_fn << std::endl;
    Tree tree;
    if (!tree.load(tax_tree_fn.c_str())) { std::cerr << "failed to open " << tax_tree_fn << " for reading\n"; return -1; }
    std::cout << "Done Read taxonomy tree: " << tax_tree_fn << std::endl;

    // ---- the called taxids (:349-383)
    std::map<TID, float> weighted_readcnt;
    std::map<TID, int> read_cnts;
    std::ifstream call_ifs(lmat_sum.c_str());
    if (!call_ifs) { std::cerr << "Failed to open " << lmat_sum << " must exit now" << std::endl; return -1; }
    std::list<TID> clst;
    std::unordered_map<TID, TID> strain2spec;
    const char *want_rank = doHumanReg ? "region" : "species";
    {
        static char buff[2024];
        std::vector<TID> ptor;
        while (call_ifs.getline(buff, sizeof buff)) {
            const std::string s = buff;
            if (s.find("\tNULL\t") != std::string::npos) continue;
            std::istringstream is(buff);
            TID tid = 0; unsigned rc = 0; std::string descrip; float w = 0;
            is >> w >> rc >> tid >> descrip;
            weighted_readcnt.insert(std::make_pair(tid, w));
            read_cnts.insert(std::make_pair(tid, (int)rc));
            if (rank_table[tid] == want_rank) strain2spec.insert(std::make_pair(tid, tid));
            if (!is_plasmid(tid)) {
                tree.path_to_root(tid, ptor);
                for (TID a : ptor) if (rank_table[a] == want_rank) strain2spec.insert(std::make_pair(tid, a));
            }
            clst.push_back(tid);
        }
    }

    // ---- the reads: which taxid each one counts for (:405-441), k-mer work on the GPU in batches
    if (kmat_device_count() < 1) { std::cerr << "ERROR! No CUDA device: this build has no CPU path" << std::endl; return -1; }
    kmat_kcov *kc = nullptr;
    if (kmat_kcov_create(0, k_size.data(), (int)k_size.size(), &kc) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }
    std::unordered_map<TID, uint32_t> group_of;                       // taxid -> dense group index
    std::string bases; std::vector<uint64_t> offs(1, 0); std::vector<uint32_t> groups;
    auto flush = [&]() -> bool {
        if (groups.empty()) return true;
        if (kmat_kcov_add(kc, bases.data(), offs.data(), groups.data(), (uint32_t)groups.size()) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return false; }
        bases.clear(); offs.assign(1, 0); groups.clear();
        return true;
    };
    for (const std::string &fn : files) {
        std::ifstream ifs(fn.c_str());
        if (!ifs) { std::cerr << "did not open for reading: [" << fn << "]" << std::endl; return -1; }
        std::string line;
        bool finished = false;
        while (!finished) {
            if (!std::getline(ifs, line)) break;
            if ((long long)ifs.tellg() == -1) finished = true;
            const size_t p1 = line.find('\t'), p2 = line.find('\t', p1 + 1), p3 = line.find('\t', p2 + 1), p4 = line.find('\t', p3 + 1), p5 = line.find('\t', p4 + 1);
            const std::string read_buff = line.substr(p1 + 1, p2 - p1 - 1);
            const std::string tws = line.substr(p4 + 1, p5 - p4 - 1);
            if (tws[0] == 'N' || tws[0] == 'R') continue;
            std::istringstream is(tws.c_str());
            float score = 0; TID taxid = 0; std::string match_type;
            is >> taxid >> score >> match_type;
            if (is_human(taxid) && skipHuman) continue;
            if (score < threshold) continue;
            TID use_tid = taxid;
            auto s2 = strain2spec.find(taxid);
            if (s2 != strain2spec.end() && !is_plasmid(taxid)) use_tid = s2->second;
            auto rk = rank_table.find(use_tid);
            const std::string rnk = rk != rank_table.end() ? rk->second : "undef";
            if (rank_check.count(rnk) || is_plasmid(taxid)) {
                auto g = group_of.find(use_tid);
                if (g == group_of.end()) g = group_of.insert(std::make_pair(use_tid, (uint32_t)group_of.size())).first;
                bases += read_buff; offs.push_back(bases.size()); groups.push_back(g->second);
                if (bases.size() >= ((size_t)64 << 20) && !flush()) return -1;
            }
        }
    }
    if (!flush()) return -1;
    if (kmat_kcov_finish(kc) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }

    // ---- the tree of called taxids and the two kinds of output (:443-522)
    std::set<TID> seen;
    std::map<TID, std::list<TID>> child;
    {
        std::vector<TID> ptor;
        for (TID tid : clst) {
            tree.path_to_root(tid, ptor);
            TID node = tid;
            for (TID p : ptor) {
                if (!seen.count(node)) { seen.insert(node); child[p].push_back(node); }
                node = p;
            }
        }
    }
    std::ofstream ofs(ofbase.c_str());
    ofs << "Name\tTaxID\tReads\tWReads" << std::endl;
    std::map<TID, std::string> tabs;
    std::list<TID> open;
    open.push_back(1);
    std::map<std::string, std::ofstream *> rank_ofs;
    auto kmer_cov = [&](TID tid, std::ofstream &o) -> bool {           // compKmerCov (:527-571)
        auto g = group_of.find(tid);
        for (size_t ki = 0; ki < k_size.size(); ki++) {
            uint64_t distinct = 0, total = 0; uint32_t nh = 0;
            std::vector<uint32_t> hc; std::vector<uint64_t> hn;
            if (g != group_of.end()) {
                if (kmat_kcov_query(kc, (int)ki, g->second, &distinct, &total, nullptr, nullptr, 0, &nh) != KMAT_OK) return false;
                hc.resize(nh + 1); hn.resize(nh + 1);
                if (kmat_kcov_query(kc, (int)ki, g->second, &distinct, &total, hc.data(), hn.data(), nh + 1, &nh) != KMAT_OK) return false;
            }
            o << "taxid=" << tid << " distinct_kmer_cnt=" << distinct << " k_size=" << k_size[ki] << " tot_kmer_cnt=" << (int)total << std::endl;
            for (uint32_t i = 0; i < nh; i++) o << tid << " " << k_size[ki] << " " << hc[i] << " " << (unsigned)hn[i] << std::endl;
        }
        return true;
    };
    while (!open.empty()) {
        const TID tid = open.front();
        open.pop_front();
        const std::string chk = tabs[tid] + "\t";
        for (TID ch : child[tid]) { tabs[ch] = chk; open.push_front(ch); }
        const unsigned tot_read_cnt = (unsigned)read_cnts[tid];
        float wrdc = 0;
        if (tot_read_cnt > 0) {
            wrdc = weighted_readcnt[tid];
            std::string rank = rank_table[tid];
            if (rank != "no_rank") {
                if (is_plasmid(tid)) rank = "plasmid";
                std::ofstream *kos = nullptr;
                auto it = rank_ofs.find(rank);
                if (it != rank_ofs.end()) kos = it->second;
                else {
                    // the reference assigns the new stream to a local that shadows `kos` (:503): the file is created, the
                    // first taxid of the rank gets no coverage block
                    std::ofstream *fresh = new std::ofstream((ofbase + "." + rank + "_kmer_cov").c_str());
                    if (!(*fresh)) std::cout << "Unable to write to " << ofbase << "." << rank << "_kmer_cov will try to continue" << std::endl;
                    rank_ofs.insert(std::make_pair(rank, fresh));
                }
                if (kos && tot_read_cnt > 1 && !kmer_cov(tid, *kos)) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return -1; }
            }
        }
        ofs << tabs[tid] << tree.get_name(tid) << "\t" << tid << "\t" << tot_read_cnt << "\t" << wrdc << std::endl;
    }
    for (auto &kv : rank_ofs) { kv.second->close(); delete kv.second; }
    kmat_kcov_free(kc);
    std::cout << "query time: done" << std::endl;
    return 0;
}
