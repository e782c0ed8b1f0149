// Host check of lmat_b200/csrc/kmat_mzr.h (the minimizer-ordered first level of the k-mer table).  Built and run by
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

// host model of the two-level table: sectors of four slots + "second level" (a map) behind an overflow flag per sector
struct Table {
    int k, m, b;
    std::vector<uint64_t> slots;                       // 16 per line
    std::unordered_map<uint64_t, uint32_t> second;
    Table(int k_, int m_, int b_) : k(k_), m(m_), b(b_), slots((size_t)16 << b_, 0) {}
    uint64_t *sector(uint64_t x) { return &slots[((x >> KM_LINE_XSHIFT) * 4 + ((x >> KM_MZR_KEY_BITS) & 3)) * 4]; }
    void insert(uint64_t canon, uint32_t payload) {
        const uint64_t x = km_line_x(canon, k, m, b);
        uint64_t *s = sector(x);
        for (int q = 0; q < 4; q++) if (!s[q]) { s[q] = (1ull << 63) | ((x & 0x0FFFFFFFull) << 32) | payload; return; }
        second[canon] = payload;
        s[0] |= KM_LINE_OVF;
    }
    // returns requests made (1 or 2); hw = payload or 0xFFFFFFFE; the probe starts from a read's FORWARD k-mer
    int find(uint64_t fwd, uint32_t &hw, std::set<uint64_t> *lines) {
        const uint64_t rc = km_mzr_revcomp(fwd, k);
        const bool fc = fwd < rc;
        const uint64_t canon = fc ? fwd : rc;
        const uint64_t x = km_line_x_of(canon, km_mzr_of_fwd(fwd, fc, k, m), k, m, b);
        if (lines) lines->insert(x >> KM_LINE_XSHIFT);
        const uint64_t *s = sector(x);
        const uint64_t want = (1ull << 63) | ((x & 0x0FFFFFFFull) << 32), keymask = ~((1ull << 62) | KM_LINE_OVF | 0xFFFFFFFFull);
        for (int q = 0; q < 4; q++) if ((s[q] & keymask) == want) { hw = (uint32_t)s[q]; return 1; }
        hw = 0xFFFFFFFEu;
        if (!(s[0] & KM_LINE_OVF)) return 1;
        CHECK(km_line_kmer_of(x, k, m, b) == canon, "second-level key");
        auto it = second.find(canon);
        if (it != second.end()) hw = it->second;
        return 2;
    }
};

int main(int argc, char **argv) {
    const int n_pairs = argc > 1 ? atoi(argv[1]) : 3;
    // 1. both hashes are bijections with working inverses
    for (int m = 8; m <= 16; m++) {
        if (m <= 10) {
            std::vector<uint8_t> seen((size_t)1 << (2 * m), 0), seen2((size_t)1 << (2 * m), 0);
            for (uint32_t x = 0; x < (1u << (2 * m)); x++) {
                const uint32_t h = km_mzr_mix(x, m), g = km_mzr_mix2(x, m);
                CHECK(h < (1u << (2 * m)) && !seen[h] && g < (1u << (2 * m)) && !seen2[g], "m=%d x=%u", m, x);
                seen[h] = 1; seen2[g] = 1;
                CHECK(km_mzr_unmix(h, m) == x && km_mzr_unmix2(g, m) == x, "m=%d x=%u", m, x);
            }
        }
        for (int i = 0; i < 200000; i++) {
            const uint32_t x = (uint32_t)rnd() & (m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1));
            CHECK(km_mzr_unmix(km_mzr_mix(x, m), m) == x && km_mzr_unmix2(km_mzr_mix2(x, m), m) == x, "m=%d x=%u", m, x);
        }
    }
    // 2. strand independence, tie rule, key -> k-mer, owner of a key
    const int geoms[][3] = {{20, 16, 29}, {20, 16, 16}, {20, 16, 22}, {20, 14, 26}, {18, 16, 12}, {17, 16, 30}, {22, 16, 32}, {23, 16, 22}, {20, 13, 26}};
    uint64_t ties = 0, checked = 0;
    for (auto &g : geoms) {
        const int k = g[0], m = g[1], b = g[2];
        CHECK(km_line_geometry_ok(k, m, b), "geometry %d %d %d", k, m, b);
        for (int style = 0; style < 3; style++) {
            for (int rep = 0; rep < 60; rep++) {
                const std::string s = random_seq(400, style);
                for (size_t p = 0; p + k <= s.size(); p++) {
                    const uint64_t fwd = kmer_at(s, p, k), rc = km_mzr_revcomp(fwd, k);
                    CHECK(km_mzr_revcomp(rc, k) == fwd, "revcomp");
                    const bool fc = fwd < rc;
                    const uint64_t canon = fc ? fwd : rc;
                    const KmMzr a = km_mzr_of(canon, k, m), f = km_mzr_of_fwd(fwd, fc, k, m), r = km_mzr_of_fwd(rc, rc < fwd, k, m);
                    CHECK(a.c == f.c && a.off == f.off && a.flip == f.flip, "fwd strand k=%d m=%d kmer=%llx: %u/%u/%u vs %u/%u/%u", k, m, (unsigned long long)fwd, a.c, a.off, a.flip, f.c, f.off, f.flip);
                    CHECK(a.c == r.c && a.off == r.off && a.flip == r.flip, "rev strand k=%d m=%d kmer=%llx", k, m, (unsigned long long)fwd);
                    int n_min = 0;
                    for (int j = 0; j + m <= k; j++) {
                        const uint32_t w = km_mzr_window(canon, k, m, j), wr = km_mzr_revcomp_m(w, m);
                        n_min += km_mzr_order(w < wr ? w : wr, m) == km_mzr_order(a.c, m);
                    }
                    ties += n_min > 1;
                    const uint64_t x = km_line_x(canon, k, m, b);
                    CHECK((x >> KM_LINE_XSHIFT) < (1ull << b) && ((x >> KM_MZR_KEY_BITS) & 3) == (a.off & 3), "ranges");
                    CHECK(km_line_kmer_of(x, k, m, b) == canon, "round trip k=%d m=%d b=%d kmer=%llx", k, m, b, (unsigned long long)canon);
                    const uint32_t gg = km_mzr_mix2(a.c, m);
                    CHECK(km_line_g_of_x(x, k, m, b) == gg, "g of key");
                    for (uint32_t ns : {2u, 3u, 8u, 16u}) {
                        const uint32_t o = km_line_owner_of_g(gg, m, ns);
                        const uint64_t first = km_line_shard_first(o, ns, m, b), cnt = km_line_shard_count(o, ns, m, b), line = x >> KM_LINE_XSHIFT;
                        CHECK(o < ns && line >= first && line < first + cnt, "owner %u of %u: line %llu not in [%llu, +%llu)", o, ns, (unsigned long long)line, (unsigned long long)first, (unsigned long long)cnt);
                    }
                    checked++;
                }
            }
        }
    }
    CHECK(ties > 1000, "the low-complexity sequences were meant to produce ties (%llu)", (unsigned long long)ties);
    // shard ranges tile the line space (neighbours may share their boundary line)
    for (uint32_t ns : {1u, 2u, 3u, 5u, 8u, 16u})
        for (int b : {16, 29, 32}) {
            uint64_t next = 0;
            for (uint32_t o = 0; o < ns; o++) {
                const uint64_t first = km_line_shard_first(o, ns, 16, b), cnt = km_line_shard_count(o, ns, 16, b);
                CHECK(first <= next && first + cnt >= next && cnt >= 1, "shard %u of %u, b=%d", o, ns, b);
                next = first + cnt;
            }
            CHECK(next == (1ull << b), "shards of %u cover 2^%d lines", ns, b);
        }
    // 3. a two-level table: genomes with a diverged sibling each, every k-mer found with its payload, absent k-mers missed
    const int k = 20, m = km_line_m(20), b = 16;
    CHECK(m == 16, "km_line_m(20)");
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
    uint64_t found_req = 0;
    for (auto &kv : truth) {
        uint32_t hw;
        found_req += T.find(km_mzr_revcomp(kv.first, k), hw, nullptr);     // probe with the other strand
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
    // 3c. the sliding-minimum form of the minimizer (one hash per base, distance-carrying keys), emulated for a warp:
    //     arrays stand for the 32 lanes, index arithmetic for the shuffles; chunks of 32 bases as in the probe kernel
    {
        uint64_t n_checked = 0;
        const int km[][2] = {{20, 13}, {20, 14}, {20, 15}, {20, 16}, {17, 16}, {18, 16}, {19, 16}, {21, 16}, {22, 16}, {23, 16}};      // w = 8 .. 2
        for (auto &g2 : km) {
            const int k2 = g2[0], m2 = g2[1], w = k2 - m2 + 1;
            int s1, s2, s3;
            km_slide_shifts(w, s1, s2, s3);
            CHECK(1 + s1 + s2 + s3 == w, "shifts cover %d windows", w);
            for (int style = 0; style < 3; style++)
                for (int rep = 0; rep < 40; rep++) {
                    const std::string s = random_seq(150 + rnd() % 100, style);
                    const int len = (int)s.size(), nch = (len + 31) / 32;
                    uint32_t pr[3][32], pl[3][32];                        // the previous chunk's keys at the three levels
                    for (auto &a : pr) for (auto &v : a) v = KM_SLIDE_NONE;
                    for (auto &a : pl) for (auto &v : a) v = KM_SLIDE_NONE;
                    for (int c = 0; c < nch; c++) {
                        uint64_t fwd[32]; uint32_t r0[32], l0[32], r1[32], l1[32], r2[32], l2[32], r3[32], l3[32];
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane;
                            uint64_t v = 0;                               // the k2 bases ending at j (zeros before the read, as in the kernel)
                            for (int q = j - k2 + 1; q <= j; q++) v = (v << 2) | (uint64_t)((q >= 0 && q < len) ? s[q] : 0);
                            fwd[lane] = v;
                            const uint32_t h = km_slide_hash(v, m2);
                            r0[lane] = km_slide_r0(h); l0[lane] = km_slide_l0(h);
                        }
                        auto fetch = [&](const uint32_t *cur, const uint32_t *prv, int lane, int sh) { return lane >= sh ? cur[lane - sh] : prv[lane - sh + 32]; };
                        for (int lane = 0; lane < 32; lane++) { r1[lane] = km_slide_r(r0[lane], fetch(r0, pr[0], lane, s1), s1); l1[lane] = km_slide_l(l0[lane], fetch(l0, pl[0], lane, s1), s1); }
                        for (int lane = 0; lane < 32; lane++) { r2[lane] = km_slide_r(r1[lane], fetch(r1, pr[1], lane, s2), s2); l2[lane] = km_slide_l(l1[lane], fetch(l1, pl[1], lane, s2), s2); }
                        for (int lane = 0; lane < 32; lane++) { r3[lane] = km_slide_r(r2[lane], fetch(r2, pr[2], lane, s3), s3); l3[lane] = km_slide_l(l2[lane], fetch(l2, pl[2], lane, s3), s3); }
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane;
                            if (j < k2 - 1 || j >= len) continue;         // no k-mer ends here
                            const uint64_t rc = km_mzr_revcomp(fwd[lane], k2);
                            const bool fc = fwd[lane] < rc;
                            const uint64_t canon = fc ? fwd[lane] : rc;
                            const KmMzr a = km_mzr_of(canon, k2, m2), z = km_slide_finish(r3[lane], l3[lane], fwd[lane], fc, k2, m2);
                            CHECK(a.c == z.c && a.off == z.off && a.flip == z.flip, "sliding minimum k=%d m=%d base %d: %u/%u/%u vs %u/%u/%u", k2, m2, j, a.c, a.off, a.flip, z.c, z.off, z.flip);
                            CHECK(km_line_x_of(canon, z, k2, m2, 26) == km_line_x(canon, k2, m2, 26), "line key from the sliding form");
                            n_checked++;
                        }
                        for (int lane = 0; lane < 32; lane++) { pr[0][lane] = r0[lane]; pr[1][lane] = r1[lane]; pr[2][lane] = r2[lane]; pl[0][lane] = l0[lane]; pl[1][lane] = l1[lane]; pl[2][lane] = l2[lane]; }
                    }
                }
        }
        CHECK(n_checked > 100000, "sliding-minimum emulation checked too little");
    }
    // 3d. the stateless form the LONG-read probe kernel uses: a warp sees one 32-base chunk and the packed bases of the chunk
    //     before it, nothing else.  Level-0 keys of this chunk's lanes and of the previous chunk's lanes (its last w - 1 are
    //     needed; their m-mers lie inside the previous chunk's 64-bit word for m <= 16 and lanes >= m - 1), then w - 1 linear
    //     steps: min over d = 0 .. w - 1 of key(j - d) shifted by d.  Same minimizer records as the definition.
    {
        uint64_t n_checked = 0;
        const int km[][2] = {{20, 13}, {20, 14}, {20, 15}, {20, 16}, {17, 16}, {18, 16}, {19, 16}, {21, 16}, {22, 16}, {23, 16}};
        for (auto &g2 : km) {
            const int k2 = g2[0], m2 = g2[1], w = k2 - m2 + 1;
            const uint64_t kmask = (1ull << (2 * k2)) - 1;
            for (int style = 0; style < 3; style++)
                for (int rep = 0; rep < 40; rep++) {
                    const std::string s = random_seq(257 + rnd() % 300, style);
                    const int len = (int)s.size(), nch = (len + 31) / 32;
                    for (int c = 0; c < nch; c++) {
                        uint64_t cur = 0, prev = 0;                       // bases of chunk c / c - 1, base of lane 0 in the top two bits
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane, jp = j - 32;
                            cur |= (uint64_t)((j < len) ? s[j] : 0) << (62 - 2 * lane);
                            prev |= (uint64_t)((jp >= 0) ? s[jp] : 0) << (62 - 2 * lane);
                        }
                        uint64_t fwd[32]; uint32_t r0[32], l0[32], r0p[32], l0p[32], kr[32], kl[32];
                        for (int lane = 0; lane < 32; lane++) {
                            const int sh = 62 - 2 * lane;
                            fwd[lane] = ((cur >> sh) | (sh ? (prev << (64 - sh)) : 0ull)) & kmask;
                            const uint32_t h = km_slide_hash(fwd[lane], m2), hp = km_slide_hash(prev >> sh, m2);
                            r0[lane] = km_slide_r0(h); l0[lane] = km_slide_l0(h); r0p[lane] = km_slide_r0(hp); l0p[lane] = km_slide_l0(hp);
                        }
                        auto fetch = [&](const uint32_t *cu, const uint32_t *pv, int lane, int sh) { return lane >= sh ? cu[lane - sh] : pv[lane - sh + 32]; };
                        for (int lane = 0; lane < 32; lane++) {
                            uint32_t a = r0[lane], b = l0[lane];
                            for (int d = 1; d < w; d++) { a = km_slide_r(a, fetch(r0, r0p, lane, d), d); b = km_slide_l(b, fetch(l0, l0p, lane, d), d); }
                            kr[lane] = a; kl[lane] = b;
                        }
                        for (int lane = 0; lane < 32; lane++) {
                            const int j = 32 * c + lane;
                            if (j < k2 - 1 || j >= len) continue;
                            const uint64_t rc = km_mzr_revcomp(fwd[lane], k2);
                            const bool fc = fwd[lane] < rc;
                            const uint64_t canon = fc ? fwd[lane] : rc;
                            const KmMzr a = km_mzr_of(canon, k2, m2), z = km_slide_finish(kr[lane], kl[lane], fwd[lane], fc, k2, m2);
                            CHECK(a.c == z.c && a.off == z.off && a.flip == z.flip, "stateless sliding minimum k=%d m=%d base %d: %u/%u/%u vs %u/%u/%u", k2, m2, j, a.c, a.off, a.flip, z.c, z.off, z.flip);
                            CHECK(km_line_x_of(canon, z, k2, m2, 26) == km_line_x(canon, k2, m2, 26), "line key from the stateless sliding form");
                            n_checked++;
                        }
                    }
                }
        }
        CHECK(n_checked > 200000, "stateless sliding-minimum emulation checked too little");
    }
    // 4. what the layout is for: requests made by the k-mers of a 150 bp read with a few substitutions -- one per distinct
    //    line of each 32-lane chunk (the lanes of a chunk issue their sector loads in one instruction) + second-level probes
    uint64_t reads = 0, kmers = 0, lines = 0, second = 0;
    for (int i = 0; i < 2000; i++) {
        const std::string &g = genomes[rnd() % (2 * n_pairs)];
        std::string rd = g.substr(rnd() % (g.size() - 150), 150);
        for (auto &c : rd) if (rnd() % 100 == 0) c = (char)(rnd() & 3);
        std::set<uint64_t> seen;
        std::set<uint64_t> chunk_lines[5];
        for (size_t p = 0; p + k <= rd.size(); p++) {
            const uint64_t f = kmer_at(rd, p, k), r = km_mzr_revcomp(f, k);
            if (!seen.insert(f < r ? f : r).second) continue;
            uint32_t hw;
            second += T.find(f, hw, &chunk_lines[(p + k - 1) / 32]) - 1;
            auto it = truth.find(f < r ? f : r);
            CHECK(hw == (it == truth.end() ? 0xFFFFFFFEu : it->second), "read k-mer");
            kmers++;
        }
        for (auto &cl : chunk_lines) lines += cl.size();
        reads++;
    }
    printf("{\"kmers_checked\": %llu, \"ties\": %llu, \"table_kmers\": %zu, \"lines\": %llu, \"mean_fill\": %.2f, \"second_level\": %zu, "
           "\"requests_per_stored_lookup\": %.3f, \"absent_checked\": %llu, \"read_kmers\": %.1f, \"read_line_requests\": %.1f, \"read_second_level\": %.2f}\n",
           (unsigned long long)checked, (unsigned long long)ties, truth.size(), 1ull << b, (double)truth.size() / (double)(1ull << b),
           T.second.size(), (double)found_req / (double)truth.size(), (unsigned long long)absent,
           (double)kmers / reads, (double)lines / reads, (double)second / reads);
    return 0;
}
