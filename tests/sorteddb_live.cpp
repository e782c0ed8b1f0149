// TEST INFRASTRUCTURE.  The reference-side binding of INTEGRATION.md B.1, run inside a process that holds a LIVE reference
// SortedDb object: the DB file is opened exactly like read_label does (perm() + mopen(), read_label.cpp:1481-1489, through the
// reference's own headers and oracle/_ref/libmetag.a), its three arrays are handed to kmat_table_from_sorteddb, and the
// resulting logical table is checked against the reference's OWN lookup API on the same object: for every k-mer
// SortedDb::begin_ must find it with the same count and next() must yield the same stored ids in the same order
// (SortedDb.hpp:188-385); absent k-mers must fail begin_.  Then the table is saved as a .kmat image for the Python side.
// `private` is opened up instead of adding the accessor the integration note proposes, so the reference sources stay
// untouched.  Built by tests/test_sorteddb_live_cpu.py where /root/reference exists.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#define private public
#include "all_headers.hpp"
#undef private
#include "kmat.h"

using namespace metag;
static SortedDb<DBTID_T> *taxtable;

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: sorteddb_live <reference db> <out.kmat>\n"); return 2; }
    perm(&taxtable, sizeof(taxtable));
    if (mopen(argv[1], "r", 0) != 0) { fprintf(stderr, "mopen failed\n"); return 2; }
    const int k = taxtable->get_kmer_length();
    kmat_table *t = nullptr;
    int rc = kmat_table_from_sorteddb(taxtable->top_tier_block, TT_BLOCK_COUNT, k == 18 ? BITS_PER_2ND_18 : BITS_PER_2ND_20, taxtable->kmer_table,
                                      taxtable->size(), taxtable->m_storage_space, UINT64_MAX, k, (int)sizeof(DBTID_T), &t);
    if (rc != KMAT_OK) { fprintf(stderr, "kmat_table_from_sorteddb: %s\n", kmat_last_error()); return 1; }
    const uint64_t *kmers, *offs; const uint32_t *ids; uint64_t n_ids;
    kmat_table_view(t, &kmers, &offs, &ids, &n_ids);
    const uint64_t n = kmat_table_size(t);
    if (n != taxtable->size()) { fprintf(stderr, "size: %llu vs SortedDb::size() %zu\n", (unsigned long long)n, taxtable->size()); return 1; }
    uint64_t lists = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint16_t count = 0; uint32_t offset = 0; uint8_t page = 0;
        if (!taxtable->begin_(kmers[i], count, offset, page)) { fprintf(stderr, "begin_ misses k-mer %llu\n", (unsigned long long)kmers[i]); return 1; }
        if (count != offs[i + 1] - offs[i]) { fprintf(stderr, "count of k-mer %llu: %u vs %llu\n", (unsigned long long)kmers[i], count, (unsigned long long)(offs[i + 1] - offs[i])); return 1; }
        lists += count > 1;
        for (uint16_t j = 0; j < count; j++) {
            DBTID_T tid = 0;
            taxtable->next(offset, page, tid);
            if ((uint32_t)tid != ids[offs[i] + j]) { fprintf(stderr, "k-mer %llu entry %u: %u vs %u\n", (unsigned long long)kmers[i], j, (unsigned)tid, ids[offs[i] + j]); return 1; }
        }
    }
    uint64_t absent = 0, state = 88172645463325252ull;
    for (int q = 0; q < 200000; q++) {
        state ^= state << 13; state ^= state >> 7; state ^= state << 17;
        const uint64_t km = state & ((1ull << (2 * k)) - 1);
        const bool mine = std::binary_search(kmers, kmers + n, km);
        uint16_t count = 0; uint32_t offset = 0; uint8_t page = 0;
        const bool theirs = taxtable->begin_(km, count, offset, page);
        if (mine != theirs) { fprintf(stderr, "k-mer %llu: table %d, begin_ %d\n", (unsigned long long)km, (int)mine, (int)theirs); return 1; }
        absent += !theirs;
    }
    if (kmat_table_save(t, argv[2]) != KMAT_OK) { fprintf(stderr, "kmat_table_save: %s\n", kmat_last_error()); return 1; }
    printf("{\"kmers\": %llu, \"ids\": %llu, \"lists\": %llu, \"absent_checked\": %llu, \"k\": %d}\n", (unsigned long long)n, (unsigned long long)n_ids,
           (unsigned long long)lists, (unsigned long long)absent, k);
    kmat_table_free(t);
    return 0;
}
