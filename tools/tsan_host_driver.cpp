#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "kmat.h"
// this driver links the host sources only: no CUDA, so no page-locked memory (the batches fall back to plain memory)
extern "C" void *kmat_host_alloc(size_t) { return nullptr; }
extern "C" void kmat_host_free(void *) {}

int main(int argc, char **argv) {
    // parallel reader
    for (int rep = 0; rep < 3; rep++) {
        for (int fq = 0; fq < 2; fq++) {
            kmat_reader *r = nullptr;
            if (kmat_reader_open_mt(argv[1 + fq], fq, 4, &r) != 0) { fprintf(stderr, "open failed: %s\n", kmat_last_error()); return 1; }
            kmat_read_batch *b = kmat_read_batch_new_pinned();
            uint64_t n = 0; int64_t got;
            while ((got = kmat_reader_next(r, 1000, 1 << 20, b)) > 0) { n += (uint64_t)got; if (rep == 2 && n > 5000) break; }   // rep 2: close early with workers busy
            printf("fq=%d reads=%llu\n", fq, (unsigned long long)n);
            kmat_read_batch_free(b);
            kmat_reader_close(r);
        }
    }
    // parallel table build
    if (argc > 4) {
        std::vector<const char *> files;
        for (int i = 4; i < argc; i++) files.push_back(argv[i]);
        kmat_build_opts o; kmat_build_opts_default(&o);
        o.kmer_length = 20; o.map16 = argv[3];
        kmat_table *t = nullptr;
        int rc = kmat_table_build(files.data(), (int)files.size(), &o, &t);
        printf("build rc=%d n=%llu\n", rc, rc == 0 ? (unsigned long long)kmat_table_size(t) : 0ull);
        if (t) kmat_table_free(t);
    }
    return 0;
}
