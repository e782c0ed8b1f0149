// kmat_mzr.h -- the minimizer-ordered first level of the k-mer table ("line level"), host + device.
//
// Why.  Random table reads on B200 are limited by the NUMBER of L2 requests (~37 G/s), and a request is one 128-byte line
// per load instruction, however many lanes of the warp read from that line (profiles/r02a_line_gather.jsonl: 4 lanes that
// share a random line, each reading its own sector, run at 36.4 G LINES/s = 146 G lanes/s).  In the probe kernel a lane
// holds the k-mer that ends at its base, so neighbouring lanes hold overlapping k-mers -- and overlapping k-mers share their
// minimizer for several positions.  A k-mer therefore lives in the line picked by its canonical minimizer: the ~131 k-mers of
// a 150 bp read then ask for ~48 lines instead of 131 buckets.
//
// Definitions (k = k-mer length, m = minimizer length, w = k - m + 1 windows, 2 <= w <= 8, 2m <= 32):
//   * window j of a k-mer: its bases j .. j+m-1 (base 0 = the two top bits of the 2k-bit value);
//   * canonical m-mer of a window: min(window, reverse complement of the window);
//   * order: h = km_mzr_mix(canonical m-mer) >> 3, a random order (poly-A is not special); the low three bits are dropped so
//     that the probe kernel's sliding minimum fits (h, distance) into one 32-bit word per lane;
//   * minimizer of a k-mer: the window with the smallest h.  The SET of canonical m-mers of a k-mer equals that of its
//     reverse complement, so the minimum does not depend on the strand;
//   * ties (the same canonical m-mer twice in one k-mer, or two m-mers whose hashes differ in the dropped bits only): the
//     smallest offset j IN THE CANONICAL K-MER wins.
//   * m is as large as the arithmetic allows (km_line_m): the minimizers that get picked are the m-mers with SMALL hashes, so
//     each of them collects every occurrence of its m-mer in the database -- with 4^m / 2 well above the number of database
//     positions most picked minimizers occur once (m = 15 for k = 20 put ~8 k-mers on every picked value of the 0.96 G
//     k-mer bench table and 18 % of them overflowed; m = 16 leaves 4 %).
//
// Addressing.  g = km_mzr_mix2(canonical minimizer), a second bijection, spreads the picked m-mers over the table; with 2^b
// lines, line = g >> (2m - b); the 32-byte SECTOR inside the line is (offset j of the minimizer) & 3 -- the k-mers of one
// super-k-mer have distinct offsets, so they spread over the four sectors and the lanes that hold them read different
// sectors of one line in ONE request.  A sector holds four 8-byte slots:
//     [63] occupied  [62] payload is a list  [61] (slot 0 only) this sector has overflowed  [59:32] key  [31:0] payload
//     key = [g & (2^(2m-b) - 1)] [j : 3 bits] [flip : 1 bit] [the k - m bases outside the window : 2(k-m) bits]
// flip = 1 when the window as it stands in the canonical k-mer is the larger of (window, its reverse complement).
// (line, key) determine the canonical k-mer, so equality of (line, key) is equality of k-mers: the probe is exact.
// Key width = 2k - b + 4 <= 28, hence b >= 2k - 24.  A k-mer whose sector is full at build time goes to the second level
// (the k-mer-hashed bucket table of kmat_internal.h) and sets the sector's overflow flag; a lookup reads ONE sector and goes
// on to the second level only when the flag is set, so a probe is at most (1 + bucket probe) dependent requests long.
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

#define KM_MZR_OFF_BITS 3               // w <= 8
#define KM_MZR_KEY_BITS 28              // as KM_REM_BITS
#define KM_LINE_OVF (1ull << 61)        // slot 0 of a sector: keys of this sector live in the second level too
#define KM_LINE_XSHIFT 30               // table key x = [line] [sector : 2] [key : 28]

// minimizer length for k-mer length k (0: no line level, the whole table is the bucket table)
KM_HD int km_line_m(int k) { return (k >= 17 && k <= 23) ? 16 : 0; }

struct KmMzr {
    uint32_t c;         // the canonical minimizer (an m-mer)
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
// position of an m-mer in the minimizer order
KM_HD uint32_t km_mzr_order(uint32_t canonical_mmer, int m) { return km_mzr_mix(canonical_mmer, m) >> 3; }
// second bijection: canonical minimizer -> table address value
KM_HD uint32_t km_mzr_mix2(uint32_t x, int m) {
    const uint32_t mask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    x = (x * 0xC2B2AE35u) & mask;
    x ^= x >> m;
    x = (x * 0x27D4EB2Fu) & mask;
    x ^= x >> m;
    return x;
}
KM_HD uint32_t km_mzr_unmix2(uint32_t x, int m) {
    const uint32_t mask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    x ^= x >> m;
    x = (x * 0xA0FE3BCFu) & mask;       // 0x27D4EB2F^-1 mod 2^32
    x ^= x >> m;
    x = (x * 0x7ED1B41Du) & mask;       // 0xC2B2AE35^-1 mod 2^32
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
KM_HD uint32_t km_mzr_revcomp_m(uint32_t f, int m) {      // 32-bit twin for m-mers (2m <= 32)
    uint32_t x = ~f;
#if defined(__CUDA_ARCH__)
    x = __brev(x);                                         // reverses the bases and the two bits inside each
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    return m >= 16 ? x : x >> (32 - 2 * m);
#endif
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    return m >= 16 ? x : x >> (32 - 2 * m);
}

// window j (m bases) of a k-base value
KM_HD uint32_t km_mzr_window(uint64_t kmer, int k, int m, int j) { return (uint32_t)((kmer >> (2 * (k - m - j))) & ((m >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * m)) - 1))); }

// The minimizer of a CANONICAL k-mer, straight from the definition (what the table build and the generic probe use).
KM_HD KmMzr km_mzr_of(uint64_t canon, int k, int m) {
    KmMzr z; z.c = 0; z.off = 0; z.flip = 0;
    uint32_t best = 0;
    bool have = false;
    for (int j = 0; j + m <= k; j++) {
        const uint32_t f = km_mzr_window(canon, k, m, j), r = km_mzr_revcomp_m(f, m);
        const uint32_t c = f < r ? f : r;
        const uint32_t h = km_mzr_order(c, m);
        if (!have || h < best) { best = h; z.c = c; z.off = (uint32_t)j; z.flip = f > r; have = true; }     // strict <: the smallest j wins a tie
    }
    return z;
}

// The same from a read's FORWARD k-mer (what a lane of the probe kernel holds) without forming the canonical k-mer's
// windows: window i of fwd is window (k - m - i) of the reverse complement, reverse-complemented.  `fwd_is_canon` = the
// read_label comparison fwd < rc (read_label.cpp:1009); a palindromic k-mer counts as forward.
KM_HD KmMzr km_mzr_of_fwd(uint64_t fwd, bool fwd_is_canon, int k, int m) {
    KmMzr z; z.c = 0; z.off = 0; z.flip = 0;
    uint32_t best = 0;
    bool have = false;
    for (int i = 0; i + m <= k; i++) {
        const uint32_t f = km_mzr_window(fwd, k, m, i), r = km_mzr_revcomp_m(f, m);
        const uint32_t c = f < r ? f : r;
        const uint32_t h = km_mzr_order(c, m);
        // forward strand: the smallest i wins a tie; reverse strand: offset in the canonical k-mer is k - m - i, so the LARGEST i
        if (!have || h < best || (!fwd_is_canon && h == best)) {
            best = h; z.c = c; have = true;
            if (fwd_is_canon) { z.off = (uint32_t)i; z.flip = f > r; }
            else { z.off = (uint32_t)(k - m - i); z.flip = r > f; }      // the canonical k-mer shows r at that offset
        }
    }
    return z;
}

KM_HD bool km_line_geometry_ok(int k, int m, int b) { return m >= 8 && 2 * m <= 32 && k > m && k - m + 1 <= (1 << KM_MZR_OFF_BITS) && b >= 1 && b <= 2 * m && 2 * k - b + 4 <= KM_MZR_KEY_BITS; }
// smallest line count the key width allows
KM_HD int km_line_min_bits(int k) { const int b = 2 * k - 24; return b < 4 ? 4 : b; }

// table key of a canonical k-mer whose minimizer record is z: x = [line : b bits] [sector : 2] [key : 28]
KM_HD uint64_t km_line_x_of(uint64_t canon, const KmMzr &z, int k, int m, int b) {
    const uint32_t g = km_mzr_mix2(z.c, m);
    const int s = 2 * m - b;
    const uint64_t line = s >= 32 ? 0ull : (uint64_t)(g >> s);
    const uint32_t grem = s >= 32 ? g : (g & (uint32_t)((1ull << s) - 1));
    // the k - m bases outside the window: zero the window's 2m bits; what is left of the left flank sits at the top of the
    // 2k-bit value, the right flank at the bottom -- fold the top down (k - m <= 7 bases: the two parts cannot collide)
    const int right_bits = 2 * (k - m - (int)z.off), fb = 2 * (k - m);
    const uint64_t wmask = (m >= 16 ? 0xFFFFFFFFull : ((1ull << (2 * m)) - 1)) << right_bits;
    const uint64_t v = canon & ~wmask;
    const uint64_t flanks = ((v >> (2 * m)) | v) & ((1ull << fb) - 1);
    const uint64_t key = ((((uint64_t)grem << (KM_MZR_OFF_BITS + 1)) | ((uint64_t)z.off << 1) | z.flip) << fb) | flanks;
    return (line << KM_LINE_XSHIFT) | ((uint64_t)(z.off & 3u) << KM_MZR_KEY_BITS) | key;
}
KM_HD uint64_t km_line_x(uint64_t canon, int k, int m, int b) { return km_line_x_of(canon, km_mzr_of(canon, k, m), k, m, b); }
// value g = km_mzr_mix2(canonical minimizer) of a table key (owner shard, line)
KM_HD uint32_t km_line_g_of_x(uint64_t x, int k, int m, int b) {
    const int s = 2 * m - b;
    const uint32_t key = (uint32_t)x & ((1u << KM_MZR_KEY_BITS) - 1);
    const uint32_t grem = key >> (2 * (k - m) + KM_MZR_OFF_BITS + 1);
    return s >= 32 ? grem : (uint32_t)(((x >> KM_LINE_XSHIFT) << s) | grem);
}
// the way back: (line, key) -> canonical k-mer (second-level probe of a flagged sector, table dumps, tests)
KM_HD uint64_t km_line_kmer_of(uint64_t x, int k, int m, int b) {
    const int fb = 2 * (k - m);
    const uint32_t key = (uint32_t)x & ((1u << KM_MZR_KEY_BITS) - 1);
    const uint64_t flanks = key & ((1ull << fb) - 1);
    const uint32_t head = key >> fb;
    const uint32_t flip = head & 1, j = (head >> 1) & ((1u << KM_MZR_OFF_BITS) - 1);
    const uint32_t c = km_mzr_unmix2(km_line_g_of_x(x, k, m, b), m);
    const uint64_t win = flip ? (uint64_t)km_mzr_revcomp_m(c, m) : (uint64_t)c;
    const int right_bits = 2 * (k - m - (int)j);
    const uint64_t right = flanks & ((1ull << right_bits) - 1), left = flanks >> right_bits;       // left flank: the top j bases
    return (left << (2 * m + right_bits)) | (win << right_bits) | right;
}
// owner shard of a table key: the top of g, so a line belongs to one owner (multiply-shift: any shard count)
KM_HD uint32_t km_line_owner_of_g(uint32_t g, int m, uint32_t n_shards) { return (uint32_t)(((uint64_t)g * n_shards) >> (2 * m)); }
// first global line of shard `o` and the number of lines it spans (a boundary line may be shared by two neighbours: each
// keeps its own k-mers of it)
KM_HD uint64_t km_line_shard_first(uint32_t o, uint32_t n_shards, int m, int b) {
    const uint64_t span = 1ull << (2 * m);
    const uint64_t g_lo = ((uint64_t)o * span + n_shards - 1) / n_shards;              // smallest g with owner o
    return g_lo >> (2 * m - b);
}
KM_HD uint64_t km_line_shard_count(uint32_t o, uint32_t n_shards, int m, int b) {
    const uint64_t span = 1ull << (2 * m);
    const uint64_t g_hi = (((uint64_t)o + 1) * span + n_shards - 1) / n_shards - 1;    // largest g with owner o
    return (g_hi >> (2 * m - b)) - km_line_shard_first(o, n_shards, m, b) + 1;
}

// ---- the minimizer of every k-mer of a read from ONE hash per base (the fast probe kernel) ----------------------------------
// Lane j hashes the m-mer that ENDS at its base; the k-mer ending at base j owns the w = k - m + 1 m-mers ending at bases
// j - w + 1 .. j, so its minimizer is a sliding minimum over the lanes to its left (and the tail of the previous 32-base
// chunk).  Keys carry the distance to the owning lane in the three bits the order drops, so that the argmin survives:
// R = order << 3 | dist (ties: nearest = rightmost window of the forward strand), L = order << 3 | (7 - dist) (ties: farthest
// = leftmost).  Three doubling steps with shifts (1, 2, w - 4) clipped to what w needs cover 1 -> 2 -> 4 -> w windows
// (2 <= w <= 8; a shift of 0 is a no-op).  The canonical k-mer's rule (smallest offset in the CANONICAL k-mer) is the
// leftmost window when the read's strand is the canonical one and the rightmost otherwise.
#define KM_SLIDE_NONE 0xFFFFFFF0u       // "no m-mer here" (before the read); only ever competes inside k-mers that are not valid anyway
KM_HD uint32_t km_slide_r0(uint32_t mix) { return mix & ~7u; }
KM_HD uint32_t km_slide_l0(uint32_t mix) { return mix | 7u; }
KM_HD uint32_t km_slide_r(uint32_t mine, uint32_t from_left, int shift) { const uint32_t c = from_left + (uint32_t)shift; return c < mine ? c : mine; }
KM_HD uint32_t km_slide_l(uint32_t mine, uint32_t from_left, int shift) { const uint32_t c = from_left - (uint32_t)shift; return c < mine ? c : mine; }
KM_HD void km_slide_shifts(int w, int &s1, int &s2, int &s3) {      // windows covered: 1 -> 1 + s1 -> 1 + s1 + s2 -> w
    s1 = w >= 2 ? 1 : 0;
    s2 = w >= 4 ? 2 : (w == 3 ? 1 : 0);
    s3 = w > 4 ? w - 4 : 0;
}
// km_mzr_mix of the canonical form of the m-mer that ends here (the low 2m bits of the k-mer ending here)
KM_HD uint32_t km_slide_hash(uint64_t fwd_kmer_ending_here, int m) {
    const uint32_t f = (uint32_t)(fwd_kmer_ending_here & (m >= 16 ? 0xFFFFFFFFull : ((1ull << (2 * m)) - 1))), r = km_mzr_revcomp_m(f, m);
    return km_mzr_mix(f < r ? f : r, m);
}
// after the last step: the minimizer record of the k-mer `fwd` that ends at this lane
KM_HD KmMzr km_slide_finish(uint32_t kr, uint32_t kl, uint64_t fwd, bool fwd_is_canon, int k, int m) {
    KmMzr z;
    const int w = k - m + 1, dist = fwd_is_canon ? 7 - (int)(kl & 7) : (int)(kr & 7);
    const int i = (w - 1) - dist;                             // window index in the forward k-mer
    const uint32_t f = km_mzr_window(fwd, k, m, i), r = km_mzr_revcomp_m(f, m);
    z.c = f < r ? f : r;
    if (fwd_is_canon) { z.off = (uint32_t)i; z.flip = f > r; }
    else { z.off = (uint32_t)(k - m - i); z.flip = r > f; }
    return z;
}
#endif
