// Every float of the fast range of km_fmt_g (lmat_b200/csrc/kmat_host.cpp: [1e-4, 999999), both ends overshot) against
// printf("%g") of the promoted double, which is what the reference's ostream << float prints.  argv[1] = stride over the bit
// patterns (1 = all 279 M, ~20 s on 8 cores).  Built and run by tests/test_abi_cpu.py.
#include "../lmat_b200/csrc/kmat_host.cpp"
#include <thread>
#include <atomic>
int main(int argc, char **argv) {
    const uint64_t stride = argc > 1 ? (uint64_t)atoll(argv[1]) : 1;
    const uint32_t lo = 0x38D1B000u, hi = 0x49742800u;      // a little beyond both ends of the fast range
    std::atomic<uint64_t> bad{0}, done{0};
    std::vector<std::thread> th;
    const int T = 8;
    for (int t = 0; t < T; t++) th.emplace_back([&, t] {
        char a[64], b[64];
        for (uint64_t u = lo + (uint64_t)t * stride; u < hi; u += (uint64_t)T * stride) {
            for (int sgn = 0; sgn < 2; sgn++) {
                if (sgn && (u & 0xFFF)) continue;             // negative: every 4096th pattern
                uint32_t bits = (uint32_t)u | (sgn ? 0x80000000u : 0u);
                float f; memcpy(&f, &bits, 4);
                const int na = (int)(km_fmt_g(a, f) - a);
                const int nb = snprintf(b, sizeof b, "%g", (double)f);
                if (na != nb || memcmp(a, b, (size_t)na) != 0) { if (bad++ < 10) fprintf(stderr, "MISMATCH %08x: %.*s vs %s\n", bits, na, a, b); }
            }
            done++;
        }
    });
    for (auto &x : th) x.join();
    printf("checked %llu floats, mismatches %llu\n", (unsigned long long)done.load(), (unsigned long long)bad.load());
    return bad.load() ? 1 : 0;
}
