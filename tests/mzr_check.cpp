// Host check of lmat_b200/csrc/kmat_mzr.h (minimizer-ordered table layout, a study: see that header).  Built and run by
// tests/test_mzr_layout_cpu.py; exits non-zero on the first failed property and prints one JSON line of statistics.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../lmat_b200/csrc/kmat_mzr.h"

static uint64_t rng_state = 0x1234567ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s -- ", __FILE__, __LINE__, #c); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

static uint64_t kmer_at(const std::string &s, size_t p, int k) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) v = (v << 2) | (uint64_t)s[p + i];
    return v;
}
static std::string random_seq(size_t n, int style) {
    std::string s(n, 0);
    if (style == 0) for (auto &c : s) c = (char)(rnd() & 3);
    else if (style == 1) for (auto &c : s) c = (char)((rnd() & 1) ? 0 : 3);                       // A/T only: m-mers repeat, ties
    else { const size_t period = 2 + rnd() % 9; std::string u = random_seq(period, 0); for (size_t i = 0; i < n; i++) s[i] = u[i % period]; }   // tandem repeat
    return s;
}

struct Table {
    int k, m, b;
    std::vector<uint64_t> slots;
    std::unordered_map<uint64_t, uint32_t> stash;
    uint64_t displaced = 0;
    Table(int k_, int m_, int b_) : k(k_), m(m_), b(b_), slots((size_t)KM_MZR_SLOTS_PER_LINE << b_, 0) {}
    void insert(uint64_t canon, uint32_t payload) {
        const KmMzr z = km_mzr_of(canon, k, m);
        const uint64_t home = km_mzr_line(z, m, b);
        const uint32_t key = km_mzr_key(canon, z, k, m, b);
        for (int d = 0; d <= 3; d++) {
            uint64_t *line = &slots[((home + d) & ((1ull << b) - 1)) * KM_MZR_SLOTS_PER_LINE];
            for (int s = 0; s < KM_MZR_SLOTS_PER_LINE; s++)
                if (!line[s]) { line[s] = (1ull << 63) | ((uint64_t)d << 60) | ((uint64_t)key << 32) | payload; displaced += d != 0; return; }
        }
        stash[canon] = payload;
    }
    // returns lines visited; hw = payload or 0xFFFFFFFE
    int find(uint64_t fwd, uint32_t &hw, std::set<uint64_t> *touched) const {
        const uint64_t rc = km_mzr_revcomp(fwd, k);
        const bool fc = fwd < rc;
        const uint64_t canon = fc ? fwd : rc;
        const KmMzr z = km_mzr_of_fwd(fwd, fc, k, m);
        const uint64_t home = km_mzr_line(z, m, b);
        const uint32_t key = km_mzr_key(canon, z, k, m, b);
        for (int d = 0; d <= 3; d++) {
            const uint64_t ln = (home + d) & ((1ull << b) - 1);
            if (touched) touched->insert(ln);
            const int r = km_mzr_line_find(&slots[ln * KM_MZR_SLOTS_PER_LINE], key, d, hw);
            if (r != 2) return d + 1;
        }
        auto it = stash.find(canon);
        hw = it == stash.end() ? 0xFFFFFFFEu : it->second;
        return 5;
    }
};

int main(int argc, char **argv) {
    const int n_pairs = argc > 1 ? atoi(argv[1]) : 3;      // genome pairs of the table in part 3 (3: ~2.3 k-mers per line)
    // 1. the order is a bijection and km_mzr_unmix inverts it
    for (int m = 8; m <= 16; m++) {
        if (m <= 10) {
            std::vector<uint8_t> seen((size_t)1 << (2 * m), 0);
            for (uint32_t x = 0; x < (1u << (2 * m)); x++) {
                const uint32_t h = km_mzr_mix(x, m);
                CHECK(h < (1u << (2 * m)) && !seen[h], "m=%d x=%u", m, x);
                seen[h] = 1;
                CHECK(km_mzr_unmix(h, m) == x, "m=%d x=%u", m, x);
            }
        }
        for (int i = 0; i < 200000; i++) {
            const uint32_t x = (uint32_t)rnd() & (m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1));
            CHECK(km_mzr_unmix(km_mzr_mix(x, m), m) == x, "m=%d x=%u", m, x);
        }
    }
    // 2. strand independence, tie rule, and (line, key) -> k-mer
    const int geoms[][3] = {{20, 14, 28}, {20, 14, 16}, {20, 14, 22}, {20, 13, 26}, {18, 12, 24}, {18, 14, 12}, {22, 16, 32}, {23, 16, 22}};
    uint64_t ties = 0, checked = 0;
    for (auto &g : geoms) {
        const int k = g[0], m = g[1], b = g[2];
        CHECK(km_mzr_geometry_ok(k, m, b), "geometry %d %d %d", k, m, b);
        for (int style = 0; style < 3; style++) {
            for (int rep = 0; rep < 60; rep++) {
                const std::string s = random_seq(400, style);
                for (size_t p = 0; p + k <= s.size(); p++) {
                    const uint64_t fwd = kmer_at(s, p, k), rc = km_mzr_revcomp(fwd, k);
                    CHECK(km_mzr_revcomp(rc, k) == fwd, "revcomp");
                    const bool fc = fwd < rc;
                    const uint64_t canon = fc ? fwd : rc;
                    const KmMzr a = km_mzr_of(canon, k, m), f = km_mzr_of_fwd(fwd, fc, k, m), r = km_mzr_of_fwd(rc, rc < fwd, k, m);
                    CHECK(a.hmin == f.hmin && a.off == f.off && a.flip == f.flip, "fwd strand k=%d m=%d kmer=%llx: %u/%u/%u vs %u/%u/%u", k, m, (unsigned long long)fwd, a.hmin, a.off, a.flip, f.hmin, f.off, f.flip);
                    CHECK(a.hmin == r.hmin && a.off == r.off && a.flip == r.flip, "rev strand k=%d m=%d kmer=%llx", k, m, (unsigned long long)fwd);
                    int n_min = 0;
                    for (int j = 0; j + m <= k; j++) {
                        const uint32_t w = km_mzr_window(canon, k, m, j), wr = (uint32_t)km_mzr_revcomp(w, m);
                        n_min += km_mzr_mix(w < wr ? w : wr, m) == a.hmin;
                    }
                    ties += n_min > 1;
                    const uint64_t line = km_mzr_line(a, m, b);
                    const uint32_t key = km_mzr_key(canon, a, k, m, b);
                    CHECK(line < (1ull << b) && key < (1u << KM_MZR_KEY_BITS), "ranges");
                    CHECK(km_mzr_kmer_of(line, key, k, m, b) == canon, "round trip k=%d m=%d b=%d kmer=%llx", k, m, b, (unsigned long long)canon);
                    checked++;
                }
            }
        }
    }
    CHECK(ties > 1000, "the low-complexity sequences were meant to produce ties (%llu)", (unsigned long long)ties);
    // 3. a table: genomes with a diverged sibling each, every k-mer found with its payload, absent k-mers missed
    const int k = 20, m = 14, b = 16;
    std::vector<std::string> genomes;
    for (int g = 0; g < n_pairs; g++) {
        genomes.push_back(random_seq(40000, 0));
        std::string sib = genomes.back();
        for (auto &c : sib) if (rnd() % 50 == 0) c = (char)(rnd() & 3);
        genomes.push_back(sib);
    }
    genomes.push_back(random_seq(3000, 1));                                  // a low-complexity stretch: heavy minimizers
    std::unordered_map<uint64_t, uint32_t> truth;
    for (auto &s : genomes)
        for (size_t p = 0; p + k <= s.size(); p++) {
            const uint64_t f = kmer_at(s, p, k), r = km_mzr_revcomp(f, k);
            truth.emplace(f < r ? f : r, (uint32_t)truth.size() & 0x7FFFFFFFu);
        }
    Table T(k, m, b);
    for (auto &kv : truth) T.insert(kv.first, kv.second);
    uint64_t found_lines = 0;
    for (auto &kv : truth) {
        uint32_t hw;
        found_lines += T.find(km_mzr_revcomp(kv.first, k), hw, nullptr);     // probe with the other strand
        CHECK(hw == kv.second, "k-mer %llx: %u vs %u", (unsigned long long)kv.first, hw, kv.second);
    }
    uint64_t absent = 0;
    for (int i = 0; i < 300000; i++) {
        const uint64_t f = rnd() & ((1ull << (2 * k)) - 1), r = km_mzr_revcomp(f, k);
        if (truth.count(f < r ? f : r)) continue;
        uint32_t hw;
        T.find(f, hw, nullptr);
        CHECK(hw == 0xFFFFFFFEu, "absent k-mer %llx answered %u", (unsigned long long)f, hw);
        absent++;
    }
    // 3b. the line table on 4-slot buckets (-DKMAT_LINE_TABLE): key round trip, and the step order shared by insert and probe
    {
        const int k2 = 20, m2 = 15;
        for (int bl : {16, 20, 29, 30}) {
            CHECK(km_line_ok(k2, m2, bl), "line geometry %d", bl);
            for (int style = 0; style < 3; style++)
                for (int rep = 0; rep < 20; rep++) {
                    const std::string s = random_seq(300, style);
                    for (size_t p = 0; p + k2 <= s.size(); p++) {
                        const uint64_t f = kmer_at(s, p, k2), r = km_mzr_revcomp(f, k2), c = f < r ? f : r;
                        const uint64_t x = km_line_x(c, k2, m2, bl);
                        CHECK((x >> 30) < (1ull << bl), "line range");
                        CHECK(km_line_kmer_of(x, k2, m2, bl) == c, "line key round trip bl=%d kmer=%llx", bl, (unsigned long long)c);
                        const KmMzr z = km_mzr_of(c, k2, m2);
                        CHECK((x >> 30) == km_mzr_line(z, m2, bl) && (uint32_t)(x & 0x0FFFFFFFu) == km_mzr_key(c, z, k2, m2, bl), "line key fields");
                    }
                }
        }
        const int bl = 16;
        const uint64_t bmask = (4ull << bl) - 1;
        std::vector<uint64_t> slots((size_t)16 << bl, 0);
        std::unordered_map<uint64_t, uint32_t> stash2;
        uint64_t at_step[KM_LINE_STEPS + 1] = {0};
        auto match = [&](const uint64_t *bk, uint64_t rem, int d, uint32_t &hw) {      // km_bucket_match of kmat_device.cuh
            const uint64_t want = (1ull << 63) | ((uint64_t)d << 60) | (rem << 32), keymask = ~((1ull << 62) | 0xFFFFFFFFull);
            bool full = true;
            for (int q = 0; q < 4; q++) { if ((bk[q] & keymask) == want) { hw = (uint32_t)bk[q]; return 0; } if (!bk[q]) full = false; }
            hw = 0xFFFFFFFEu; return full ? 2 : 1;
        };
        for (auto &kv : truth) {
            const uint64_t x = km_line_x(kv.first, k2, m2, bl), home = x >> 28, rem = x & 0x0FFFFFFFull;
            int t = 0; bool done = false;
            for (; t < KM_LINE_STEPS && !done; t++) {
                uint64_t *bk = &slots[km_line_bucket_at(home, t, bmask) * 4];
                for (int q = 0; q < 4 && !done; q++) if (!bk[q]) { bk[q] = (1ull << 63) | ((uint64_t)(t >> 2) << 60) | (rem << 32) | kv.second; done = true; }
                if (done) break;
            }
            if (!done) { stash2[x] = kv.second; at_step[KM_LINE_STEPS]++; } else at_step[t]++;
        }
        auto find2 = [&](uint64_t canon, uint32_t &hw) {
            const uint64_t x = km_line_x(canon, k2, m2, bl), home = x >> 28, rem = x & 0x0FFFFFFFull;
            for (int t = 0; t < KM_LINE_STEPS; t++) {
                const int r = match(&slots[km_line_bucket_at(home, t, bmask) * 4], rem, t >> 2, hw);
                if (r != 2) return t + 1;
            }
            auto it = stash2.find(x); hw = it == stash2.end() ? 0xFFFFFFFEu : it->second;
            return KM_LINE_STEPS + 1;
        };
        uint64_t steps = 0;
        for (auto &kv : truth) { uint32_t hw; steps += find2(kv.first, hw); CHECK(hw == kv.second, "line table k-mer %llx", (unsigned long long)kv.first); }
        for (int i = 0; i < 200000; i++) {
            const uint64_t f = rnd() & ((1ull << (2 * k2)) - 1), r = km_mzr_revcomp(f, k2), c = f < r ? f : r;
            if (truth.count(c)) continue;
            uint32_t hw; find2(c, hw);
            CHECK(hw == 0xFFFFFFFEu, "line table: absent k-mer %llx answered", (unsigned long long)c);
        }
        fprintf(stderr, "line table (m=15, %.2f k-mers per line): home sector %.1f%%, rest of line %.1f%%, later lines %.1f%%, stash %.2f%%, %.2f sectors per stored lookup\n",
                (double)truth.size() / (1ull << bl), 100.0 * at_step[0] / truth.size(), 100.0 * (at_step[1] + at_step[2] + at_step[3]) / truth.size(),
                100.0 * (truth.size() - at_step[0] - at_step[1] - at_step[2] - at_step[3] - at_step[KM_LINE_STEPS]) / truth.size(),
                100.0 * at_step[KM_LINE_STEPS] / truth.size(), (double)steps / truth.size());
    }
    // 3c. the sliding-minimum form of the minimizer (one hash per base, distance-carrying keys), emulated for a warp:
    //     arrays stand for the 32 lanes, index arithmetic for the shuffles; chunks of 32 bases as in the probe kernel
    {
        uint64_t n_checked = 0;
        for (int m2 : {13, 14, 15, 16}) {                                     // w = 8, 7, 6, 5
            const int k2 = 20, w = k2 - m2 + 1, last_shift = w - 4;
            for (int style = 0; style < 3; style++)
                for (int rep = 0; rep < 60; rep++) {
                    const std::string s = random_seq(150 + rnd() % 100, style);
                    const int len = (int)s.size(), nch = (len + 31) / 32;
                    uint64_t pr[3][32], pl[3][32];                        // the previous chunk's keys at the three levels
                    for (auto &a : pr) for (auto &v : a) v = KM_SLIDE_NONE;
                    for (auto &a : pl) for (auto &v : a) v = KM_SLIDE_NONE;
                    for (int c = 0; c < nch; c++) {
                        uint64_t fwd[32], r0[32], l0[32], r1[32], l1[32], r2[32], l2[32], r3[32], l3[32];
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane;
                            uint64_t v = 0;                               // the k2 bases ending at j (zeros before the read, as in the kernel)
                            for (int q = j - k2 + 1; q <= j; q++) v = (v << 2) | (uint64_t)((q >= 0 && q < len) ? s[q] : 0);
                            fwd[lane] = v;
                            const uint32_t h = km_slide_hash(v, m2);
                            r0[lane] = km_slide_r0(h); l0[lane] = km_slide_l0(h);
                        }
                        auto fetch = [&](const uint64_t *cur, const uint64_t *prv, int lane, int sh) { return lane >= sh ? cur[lane - sh] : prv[lane - sh + 32]; };
                        for (int lane = 0; lane < 32; lane++) { r1[lane] = km_slide_r(r0[lane], fetch(r0, pr[0], lane, 1), 1); l1[lane] = km_slide_l(l0[lane], fetch(l0, pl[0], lane, 1), 1); }
                        for (int lane = 0; lane < 32; lane++) { r2[lane] = km_slide_r(r1[lane], fetch(r1, pr[1], lane, 2), 2); l2[lane] = km_slide_l(l1[lane], fetch(l1, pl[1], lane, 2), 2); }
                        for (int lane = 0; lane < 32; lane++) { r3[lane] = km_slide_r(r2[lane], fetch(r2, pr[2], lane, last_shift), last_shift); l3[lane] = km_slide_l(l2[lane], fetch(l2, pl[2], lane, last_shift), last_shift); }
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane;
                            if (j < k2 - 1 || j >= len) continue;         // no k-mer ends here
                            const uint64_t rc = km_mzr_revcomp(fwd[lane], k2);
                            const bool fc = fwd[lane] < rc;
                            const uint64_t canon = fc ? fwd[lane] : rc;
                            const KmMzr a = km_mzr_of(canon, k2, m2), z = km_slide_finish(r3[lane], l3[lane], fwd[lane], fc, k2, m2);
                            CHECK(a.hmin == z.hmin && a.off == z.off && a.flip == z.flip, "sliding minimum m=%d base %d: %u/%u/%u vs %u/%u/%u", m2, j, a.hmin, a.off, a.flip, z.hmin, z.off, z.flip);
                            CHECK(km_line_x_of(canon, z, k2, m2, 20) == km_line_x(canon, k2, m2, 20), "line key from the sliding form");
                            n_checked++;
                        }
                        for (int lane = 0; lane < 32; lane++) { pr[0][lane] = r0[lane]; pr[1][lane] = r1[lane]; pr[2][lane] = r2[lane]; pl[0][lane] = l0[lane]; pl[1][lane] = l1[lane]; pl[2][lane] = l2[lane]; }
                    }
                }
        }
        CHECK(n_checked > 50000, "sliding-minimum emulation checked too little");
    }
    // 4. what the layout is for: distinct lines touched by the k-mers of a 150 bp read with a few substitutions
    uint64_t reads = 0, kmers = 0, lines = 0;
    for (int i = 0; i < 2000; i++) {
        const std::string &g = genomes[rnd() % (2 * n_pairs)];
        std::string rd = g.substr(rnd() % (g.size() - 150), 150);
        for (auto &c : rd) if (rnd() % 100 == 0) c = (char)(rnd() & 3);
        std::set<uint64_t> touched, seen;
        for (size_t p = 0; p + k <= rd.size(); p++) {
            const uint64_t f = kmer_at(rd, p, k), r = km_mzr_revcomp(f, k);
            if (!seen.insert(f < r ? f : r).second) continue;
            uint32_t hw;
            T.find(f, hw, &touched);
            auto it = truth.find(f < r ? f : r);
            CHECK(hw == (it == truth.end() ? 0xFFFFFFFEu : it->second), "read k-mer");
            kmers++;
        }
        lines += touched.size(); reads++;
    }
    printf("{\"kmers_checked\": %llu, \"ties\": %llu, \"table_kmers\": %zu, \"lines\": %llu, \"mean_fill\": %.2f, \"displaced\": %llu, \"stash\": %zu, "
           "\"lines_per_hit_lookup\": %.3f, \"absent_checked\": %llu, \"read_kmers\": %.1f, \"read_lines\": %.1f}\n",
           (unsigned long long)checked, (unsigned long long)ties, truth.size(), 1ull << b, (double)truth.size() / (double)(1ull << b),
           (unsigned long long)T.displaced, T.stash.size(), (double)found_lines / (double)truth.size(), (unsigned long long)absent,
           (double)kmers / reads, (double)lines / reads);
    return 0;
}
