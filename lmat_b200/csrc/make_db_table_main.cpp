// make_db_table -- drop-in host for LMAT's make_db_table (src/make_db_table.cpp main(), :105-433) over libkmat.
//
// Same getopt string and meaning of the options; the output (-o) is libkmat's flat ".kmat" table image (what the
// kmat read_label host opens with -d) instead of a PERM heap, so -s (heap size) and -c (extra records) are accepted
// and ignored.  The content is what the reference's SortedDb::add_data would have stored (kmat_table_build).
#include <getopt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "kmat.h"

static void usage() {
    std::cout << "Usage:\n"
                 "  -i <fn>  - input tax_histo file, or a file listing them with -l   [required]\n"
                 "  -o <fn>  - output table image (.kmat)                              [required]\n"
                 "  -k <int> - k-mer length                                            [required]\n"
                 "  -l       - the -i file is a list of input files (ascending k-mer order)\n"
                 "  -f <fn>  - 32-to-16-bit taxid map        -g <int> - prune lists longer than this   -m <fn> - numeric rank table\n"
                 "  -j <fn>  - sorted human k-mers (ASCII)   -u <fn>  - adaptor k-mers (ASCII)\n"
                 "  -h       - input is kmerPrefixCounter output, not tax_histo       -q <n> - stop after record n of every file\n"
                 "  -s <GiB>, -c <n> - accepted for compatibility (PERM heap sizing), ignored       -V - version\n";
}

int main(int argc, char *argv[]) {
    std::string inputfn, outputfn, species_map_fn, id_bit_conv_fn, human_kmer_fn, illu_kmer_fn;
    kmat_build_opts o;
    kmat_build_opts_default(&o);
    o.kmer_length = 0;
    bool list = false, strainspecies = false;
    int count = 0, c;
    std::cout << "invocation: ";
    for (int j = 0; j < argc; j++) std::cout << argv[j] << " ";
    std::cout << std::endl;
    while ((c = getopt(argc, argv, "g:q:k:i:o:s: l h m:f:wj:c:u:V")) != -1) {              // make_db_table.cpp:150
        switch (c) {
            case 'j': human_kmer_fn = optarg; break;
            case 'u': illu_kmer_fn = optarg; break;
            case 'w': strainspecies = true; break;
            case 'f': id_bit_conv_fn = optarg; break;
            case 'q': o.stopper = strtoull(optarg, nullptr, 10); break;
            case 'k': ++count; o.kmer_length = atoi(optarg); break;
            case 'l': list = true; break;
            case 'i': ++count; inputfn = optarg; break;
            case 'c': break;
            case 'o': ++count; outputfn = optarg; break;
            case 's': break;
            case 'h': o.tax_histo_format = 0; break;
            case 'g': o.tid_cutoff = atoi(optarg); break;
            case 'm': species_map_fn = optarg; break;
            case 'V': std::cout << "LMAT version 1.2.4_2018a (kmat make_db_table, ABI " << kmat_abi_version() << ")\n"; return 0;
            default: usage(); return 1;
        }
    }
    if (count != 3) { usage(); return 1; }
    if (strainspecies) { std::cout << "functionality disabled!\n"; return 1; }             // SortedDb.cpp:306-309
    std::vector<std::string> files;
    if (list) {
        std::ifstream ifs(inputfn.c_str());
        std::string line;
        while (ifs >> line) files.push_back(line);
    } else files.push_back(inputfn);
    if (files.empty()) { std::cerr << "no input files\n"; return 1; }
    std::vector<const char *> fp;
    for (auto &f : files) fp.push_back(f.c_str());
    o.map16 = id_bit_conv_fn.empty() ? nullptr : id_bit_conv_fn.c_str();
    o.numrank = species_map_fn.empty() ? nullptr : species_map_fn.c_str();
    o.human_kmers = human_kmer_fn.empty() ? nullptr : human_kmer_fn.c_str();
    o.adaptor_kmers = illu_kmer_fn.empty() ? nullptr : illu_kmer_fn.c_str();
    if (human_kmer_fn.empty()) std::cout << "No human k-mer file.\n";
    if (illu_kmer_fn.empty()) std::cout << "No Illumina k-mer file.\n";
    const auto t0 = std::chrono::steady_clock::now();
    kmat_table *t = nullptr;
    if (kmat_table_build(fp.data(), (int)fp.size(), &o, &t) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return 1; }
    std::cout << "kmer count: " << kmat_table_size(t) << "\n";
    if (kmat_table_save(t, outputfn.c_str()) != KMAT_OK) { std::cerr << "ERROR! " << kmat_last_error() << std::endl; return 1; }
    kmat_table_free(t);
    std::cout << "KmerDB load time: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << std::endl;
    return 0;
}
