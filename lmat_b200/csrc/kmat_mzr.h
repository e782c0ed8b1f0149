// kmat_mzr.h -- building blocks of the minimizer-ordered table layout (profiles/r01_minimizer_study.md).
//
// STATUS: layout study.  Nothing in libkmat uses this header yet; tests/test_mzr_layout_cpu.py compiles it on the host and
// checks the properties the probe kernel will rely on.  The functions are host+device (KM_HD) so that the build and probe
// kernels of the next round share them with that test.
//
// Idea: today every canonical k-mer lives in the 32-byte bucket picked by a hash of the k-mer itself, so the ~131 k-mers
// of a 150 bp read cost ~131 random DRAM line fetches, and the probe kernel sits at 92 % of the measured request-rate
// ceiling.  Consecutive k-mers of a read share their minimizer for several positions.  If a k-mer lives in the 128-byte
// line picked by a hash of its canonical minimizer, those neighbours fall into the same line and the lanes that hold them
// coalesce into one request.
//
// Definitions (k = k-mer length, m = minimizer length, w = k - m + 1 windows, 2m <= 32, w <= 8):
//   * window j of a k-mer: its bases j .. j+m-1 (base 0 = the two top bits of the 2k-bit value);
//   * canonical m-mer of a window: min(window, reverse complement of the window);
//   * order: h = km_mzr_mix(canonical m-mer), a bijection of the 2m-bit space (random order, so poly-A is not special);
//   * minimizer of a k-mer: the window with the smallest h.  The SET of canonical m-mers of a k-mer equals that of its
//     reverse complement, so hmin does not depend on the strand -- a read's k-mer and its canonical form agree on it, and
//     so do neighbouring k-mers as long as the minimal window lies in both;
//   * ties (the same canonical m-mer twice in one k-mer): the smallest offset j IN THE CANONICAL K-MER wins.
//
// Entry key.  With 2^b lines, line = hmin >> (2m - b) and the slot stores
//     key = [hmin & (2^(2m-b) - 1)] [j : 3 bits] [flip : 1 bit] [the k - m bases outside the window : 2(k-m) bits]
// where flip = 1 when the window as it stands in the canonical k-mer is the larger of (window, its reverse complement).
// (line, key) determine the canonical k-mer: line and the first field give hmin, the bijection gives the canonical m-mer,
// flip says which strand of it stands in the k-mer, j where, and the last field is every other base.  So equality of
// (line, key) is equality of k-mers: the probe is exact, as with today's (bucket, remainder) pair.
// Key width = 2k - b + 4 bits; the slot format of kmat_internal.h has 28 key bits, hence b >= 2k - 24 (b >= 16 for k = 20).
#ifndef KMAT_MZR_H
#define KMAT_MZR_H
#include <cstdint>

#ifndef KM_HD
#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#else
#define KM_HD inline
#endif
#endif

#define KM_MZR_SLOTS_PER_LINE 16        // 16 x 8-byte slots = one 128-byte line
#define KM_MZR_OFF_BITS 3               // w <= 8
#define KM_MZR_KEY_BITS 28              // as KM_REM_BITS

struct KmMzr {
    uint32_t hmin;      // km_mzr_mix of the canonical minimizer
    uint32_t off;       // offset j of the minimal window in the CANONICAL k-mer
    uint32_t flip;      // 1: that window is the reverse complement of the canonical m-mer
};

// bijection of the 2m-bit space (2m <= 32): two rounds of odd multiply + xor-shift by m
KM_HD uint32_t km_mzr_mix(uint32_t x, int m) {
    const uint32_t mask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    x = (x * 0x9E3779B1u) & mask;
    x ^= x >> m;
    x = (x * 0x85EBCA6Bu) & mask;
    x ^= x >> m;
    return x;
}
// its inverse (the multipliers' inverses modulo 2^32 are also their inverses modulo 2^(2m); x ^= x >> m undoes itself
// because 2m bits shifted by m twice are gone)
KM_HD uint32_t km_mzr_unmix(uint32_t x, int m) {
    const uint32_t mask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    x ^= x >> m;
    x = (x * 0xA5CB9243u) & mask;       // 0x85EBCA6B^-1 mod 2^32
    x ^= x >> m;
    x = (x * 0x0E8B2F51u) & mask;       // 0x9E3779B1^-1 mod 2^32
    return x;
}

// reverse complement of an n-base value held in the low 2n bits (n <= 32)
KM_HD uint64_t km_mzr_revcomp(uint64_t v, int n) {
    uint64_t x = ~v;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return x >> (64 - 2 * n);
}

// window j (m bases) of a k-base value
KM_HD uint32_t km_mzr_window(uint64_t kmer, int k, int m, int j) { return (uint32_t)((kmer >> (2 * (k - m - j))) & ((m >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * m)) - 1))); }

// The minimizer of a CANONICAL k-mer, straight from the definition (what the table build uses).
KM_HD KmMzr km_mzr_of(uint64_t canon, int k, int m) {
    KmMzr z; z.hmin = 0xFFFFFFFFu; z.off = 0; z.flip = 0;
    bool have = false;
    for (int j = 0; j + m <= k; j++) {
        const uint32_t f = km_mzr_window(canon, k, m, j), r = (uint32_t)km_mzr_revcomp(f, m);
        const uint32_t c = f < r ? f : r;
        const uint32_t h = km_mzr_mix(c, m);
        if (!have || h < z.hmin) { z.hmin = h; z.off = (uint32_t)j; z.flip = f > r; have = true; }     // strict <: the smallest j wins a tie
    }
    return z;
}

// The same from a read's FORWARD k-mer (what a lane of the probe kernel holds) without forming the canonical k-mer's
// windows: window i of fwd is window (k - m - i) of the reverse complement, reverse-complemented.  `fwd_is_canon` = the
// read_label comparison fwd < rc (read_label.cpp:1009); a palindromic k-mer counts as forward.
KM_HD KmMzr km_mzr_of_fwd(uint64_t fwd, bool fwd_is_canon, int k, int m) {
    KmMzr z; z.hmin = 0xFFFFFFFFu; z.off = 0; z.flip = 0;
    bool have = false;
    for (int i = 0; i + m <= k; i++) {
        const uint32_t f = km_mzr_window(fwd, k, m, i), r = (uint32_t)km_mzr_revcomp(f, m);
        const uint32_t c = f < r ? f : r;
        const uint32_t h = km_mzr_mix(c, m);
        // forward strand: the smallest i wins a tie; reverse strand: offset in the canonical k-mer is k - m - i, so the LARGEST i
        if (!have || h < z.hmin || (!fwd_is_canon && h == z.hmin)) {
            z.hmin = h; have = true;
            if (fwd_is_canon) { z.off = (uint32_t)i; z.flip = f > r; }
            else { z.off = (uint32_t)(k - m - i); z.flip = r > f; }      // the canonical k-mer shows r at that offset
        }
    }
    return z;
}

// number of line-address bits b -> (line, key).  Requires 16 <= ... see the header comment: 2k - b + 4 <= KM_MZR_KEY_BITS, b <= 2m.
KM_HD uint64_t km_mzr_line(const KmMzr &z, int m, int b) { return (uint64_t)(z.hmin >> (2 * m - b)); }
KM_HD uint32_t km_mzr_key(uint64_t canon, const KmMzr &z, int k, int m, int b) {
    const int j = (int)z.off, right_bases = k - m - j;
    const uint64_t left = j ? canon >> (2 * (k - j)) : 0ull;
    const uint64_t right = canon & ((1ull << (2 * right_bases)) - 1);
    const uint64_t flanks = (left << (2 * right_bases)) | right;                        // 2 (k - m) bits
    const uint32_t hrem = z.hmin & (uint32_t)((1ull << (2 * m - b)) - 1);
    return (uint32_t)((((uint64_t)hrem << (KM_MZR_OFF_BITS + 1) | (uint64_t)z.off << 1 | z.flip) << (2 * (k - m))) | flanks);
}
// the way back (used by the test to prove that (line, key) is injective, and by a future table dump)
KM_HD uint64_t km_mzr_kmer_of(uint64_t line, uint32_t key, int k, int m, int b) {
    const int fb = 2 * (k - m);
    const uint64_t flanks = key & ((1ull << fb) - 1);
    const uint32_t head = key >> fb;
    const uint32_t flip = head & 1, j = (head >> 1) & ((1u << KM_MZR_OFF_BITS) - 1), hrem = head >> (KM_MZR_OFF_BITS + 1);
    const uint32_t h = (uint32_t)(line << (2 * m - b)) | hrem;
    const uint32_t c = km_mzr_unmix(h, m);
    const uint64_t win = flip ? km_mzr_revcomp(c, m) : (uint64_t)c;
    const int right_bases = k - m - (int)j;
    const uint64_t left = flanks >> (2 * right_bases), right = flanks & ((1ull << (2 * right_bases)) - 1);
    return (j ? left << (2 * (k - (int)j)) : 0ull) | (win << (2 * right_bases)) | right;
}
KM_HD bool km_mzr_geometry_ok(int k, int m, int b) { return m >= 8 && 2 * m <= 32 && k - m + 1 <= (1 << KM_MZR_OFF_BITS) && k > m && b <= 2 * m && 2 * k - b + 4 <= KM_MZR_KEY_BITS; }

// ---- the line table on today's slot format (-DKMAT_LINE_TABLE=1, kmat_internal.h; experiment, replicated table only) ----
// The kernels address the table through a 64-bit key x: bucket = x >> rem_bits, remainder = the low rem_bits.  With
//     x = [line : bl bits] [s0 : 2 bits] [key : 28 bits],  rem_bits = 28,  4 x 2^bl buckets of 4 slots
// the home bucket is sector s0 = (a hash of the key) of the minimizer's 128-byte line, so the probe kernel's one LDG.256
// per k-mer stays as it is.  What changes is where a k-mer goes when that sector is full -- the rest of its line first,
// then the next three lines: step t = 0 .. 15 visits sector (s0 + t) & 3 of line home + (t >> 2), and the slot's two
// displacement bits hold t >> 2.  Insert and probe walk the same order and an insert takes the first free slot, so a free
// slot on the way proves absence.  (The key-function-only shortcut of profiles/r01_minimizer_study.md fails because it
// keeps a minimizer's k-mers in ONE 4-slot bucket; here they have the 16 slots of their line and 48 more behind it.)
KM_HD bool km_line_ok(int k, int m, int bl) { return km_mzr_geometry_ok(k, m, bl); }
KM_HD uint64_t km_line_x(uint64_t canon, int k, int m, int bl) {
    const uint64_t rck = km_mzr_revcomp(canon, k);         // window j of rck = reverse complement of window k-m-j of canon
    KmMzr z; z.hmin = 0xFFFFFFFFu; z.off = 0; z.flip = 0;
    for (int j = 0; j + m <= k; j++) {
        const uint32_t f = km_mzr_window(canon, k, m, j), r = km_mzr_window(rck, k, m, k - m - j);
        const uint32_t h = km_mzr_mix(f < r ? f : r, m);
        if (h < z.hmin || j == 0) { z.hmin = h; z.off = (uint32_t)j; z.flip = f > r; }    // strict <: the smallest j wins a tie
    }
    const uint32_t key = km_mzr_key(canon, z, k, m, bl);
    const uint32_t s0 = (key * 0x9E3779B1u) >> 30;
    return (km_mzr_line(z, m, bl) << 30) | ((uint64_t)s0 << 28) | key;
}
KM_HD uint64_t km_line_kmer_of(uint64_t x, int k, int m, int bl) { return km_mzr_kmer_of(x >> 30, (uint32_t)(x & 0x0FFFFFFFu), k, m, bl); }
// bucket visited at step t of the probe order (home_bucket = x >> 28)
KM_HD uint64_t km_line_bucket_at(uint64_t home_bucket, int t, uint64_t bucket_mask) {
    return ((((home_bucket >> 2) + (uint64_t)(t >> 2)) << 2) | ((home_bucket + (uint64_t)t) & 3)) & bucket_mask;
}
#define KM_LINE_STEPS 16

// ---- the minimizer of every k-mer of a read from ONE hash per base (-DKMAT_LINE_SHFL=1 on top of KMAT_LINE_TABLE) ---------
// Lane j hashes the m-mer that ENDS at its base; the k-mer ending at base j owns the w = k - m + 1 m-mers ending at bases
// j - w + 1 .. j, so its minimizer is a sliding minimum over the lanes to its left (and the tail of the previous 32-base
// chunk).  Keys carry the distance to the owning lane so that the argmin survives: R = h << 3 | dist (ties: nearest =
// rightmost window of the forward strand), L = h << 3 | (7 - dist) (ties: farthest = leftmost).  Three doubling steps
// cover 1 -> 2 -> 4 -> w windows (5 <= w <= 8).  The canonical k-mer's rule (smallest offset in the CANONICAL k-mer) is the
// leftmost window when the read's strand is the canonical one and the rightmost otherwise.
#define KM_SLIDE_NONE 0x7FFFFFFFFFFFFF00ull     // "no m-mer here" (before the read): loses every minimum, survives + shift
KM_HD uint64_t km_slide_r0(uint32_t h) { return (uint64_t)h << 3; }
KM_HD uint64_t km_slide_l0(uint32_t h) { return ((uint64_t)h << 3) | 7u; }
KM_HD uint64_t km_slide_r(uint64_t mine, uint64_t from_left, int shift) { const uint64_t c = from_left + (uint64_t)shift; return c < mine ? c : mine; }
KM_HD uint64_t km_slide_l(uint64_t mine, uint64_t from_left, int shift) { const uint64_t c = from_left - (uint64_t)shift; return c < mine ? c : mine; }
KM_HD uint32_t km_mzr_revcomp_m(uint32_t f, int m) {      // 32-bit twin of km_mzr_revcomp for m-mers (2m <= 32)
    uint32_t x = ~f;
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    return m >= 16 ? x : x >> (32 - 2 * m);
}
KM_HD uint32_t km_slide_hash(uint64_t fwd_kmer_ending_here, int m) {
    const uint32_t f = (uint32_t)(fwd_kmer_ending_here & (m >= 16 ? 0xFFFFFFFFull : ((1ull << (2 * m)) - 1))), r = km_mzr_revcomp_m(f, m);
    return km_mzr_mix(f < r ? f : r, m);
}
// after the last step: the minimizer record of the k-mer `fwd` that ends at this lane
KM_HD KmMzr km_slide_finish(uint64_t kr, uint64_t kl, uint64_t fwd, bool fwd_is_canon, int k, int m) {
    KmMzr z;
    z.hmin = (uint32_t)(kr >> 3);
    const int w = k - m + 1, dist = fwd_is_canon ? 7 - (int)(kl & 7) : (int)(kr & 7);
    const int i = (w - 1) - dist;                             // window index in the forward k-mer
    const uint32_t f = km_mzr_window(fwd, k, m, i), r = km_mzr_revcomp_m(f, m);
    if (fwd_is_canon) { z.off = (uint32_t)i; z.flip = f > r; }
    else { z.off = (uint32_t)(k - m - i); z.flip = r > f; }
    return z;
}
KM_HD uint64_t km_line_x_of(uint64_t canon, const KmMzr &z, int k, int m, int bl) {
    const uint32_t key = km_mzr_key(canon, z, k, m, bl);
    return (km_mzr_line(z, m, bl) << 30) | ((uint64_t)((key * 0x9E3779B1u) >> 30) << 28) | key;
}

// ---- one line: 16 slots of [63] occupied [62] is_list [61:60] displacement [59:32] key [31:0] payload (kmat_internal.h) ----
// 0 = found (payload and list flag in hw), 1 = absent for good (a free slot: keys are only displaced out of FULL lines),
// 2 = the line is full, look at the next one
KM_HD int km_mzr_line_find(const uint64_t *line, uint32_t key, int d, uint32_t &hw) {
    const uint64_t want = (1ull << 63) | ((uint64_t)d << 60) | ((uint64_t)key << 32);
    const uint64_t keymask = ~((1ull << 62) | 0xFFFFFFFFull);
    bool full = true;
    for (int s = 0; s < KM_MZR_SLOTS_PER_LINE; s++) {
        const uint64_t v = line[s];
        if ((v & keymask) == want) { hw = (uint32_t)v | (((v >> 62) & 1) ? 0x80000000u : 0u); return 0; }
        if (!v) full = false;
    }
    hw = 0xFFFFFFFEu;
    return full ? 2 : 1;
}
#endif
