// kmat_format.cuh -- K5: the tail of every output line formatted on the device (included by kmat_label.cu).
//
// What the host writers spend most of their time on is kmat_format_tail (kmat_host.cpp; read_label.cpp:1218,1233,1271,
// 844-848,894-937): ~20 numbers per read, each a printf("%g") of a float or a decimal taxid.  The float formatter there is
// integer-exact by construction (km_fmt_g: the 24-bit significand times a power of ten, exact remainder, ties to even), so
// it moves to the device unchanged: km_format_kernel writes the same bytes, one thread per read into a shared-memory
// staging row, then the warp copies its 32 rows to one contiguous piece of the pass's text buffer (one atomic per warp).
// The host only pastes header, read and tail together.  Reads the kernel does not take -- a number outside km_fmt_g's fixed
// range (below 2^-70 or from 999999 on, inf, nan), a tail longer than the staging row, a full text buffer -- are marked
// KMAT_TEXT_ON_HOST and formatted by kmat_format_tail as before; tests/test_gpu_format.py compares the two byte for byte.
#ifndef KMAT_FORMAT_CUH
#define KMAT_FORMAT_CUH

#define KF_THREADS 128
#define KF_ROW 768                 // staging bytes per read: a tail of up to ~40 printed pairs
#define KF_STRIDE (KF_ROW + 4)     // row stride: the same byte of the 32 rows of a warp falls into 32 different banks

__constant__ int8_t kf_e_low[34] = {-5, -4, -4, -4, -4, -3, -3, -3, -2, -2, -2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 5};
__constant__ uint32_t kf_thr[34] = {13743896, 16777216, 16777216, 16777216, 8589935, 16777216, 16777216, 10737419, 16777216, 16777216, 13421773,
                                    16777216, 16777216, 16777216, 16777216, 16777216, 16777216, 10485760, 16777216, 16777216, 13107200, 16777216,
                                    16777216, 16384000, 16777216, 16777216, 16777216, 10240000, 16777216, 16777216, 12800000, 16777216, 16777216, 16000000};
__constant__ unsigned long long kf_i10[11] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull, 10000000000ull};

// values below 1e-4 print in exponent notation ("6.32203e-08": the standard deviation of equal scores is float noise of that
// size, so these are common): the same exact scheme with the binades 2^-70 .. 2^-14 (tables generated with exact rational
// arithmetic, checked against printf in tests/test_gpu_format.py) and a 128-bit product -- m * 10^(5 - e) needs up to 114 bits
__constant__ int8_t kf_e_low2[57] = {-22, -21, -21, -21, -20, -20, -20, -19, -19, -19, -19, -18, -18, -18, -17, -17, -17, -16, -16, -16, -16, -15, -15, -15, -14, -14, -14, -13, -13, -13, -13, -12, -12, -12, -11, -11, -11, -10, -10, -10, -10, -9, -9, -9, -8, -8, -8, -7, -7, -7, -7, -6, -6, -6, -5, -5, -5};
__constant__ uint32_t kf_thr2[57] = {9903521, 16777216, 16777216, 12379401, 16777216, 16777216, 15474251, 16777216, 16777216, 16777216, 9671407, 16777216, 16777216, 12089259, 16777216, 16777216, 15111573, 16777216, 16777216, 16777216, 9444733, 16777216, 16777216, 11805917, 16777216, 16777216, 14757396, 16777216, 16777216, 16777216, 9223373, 16777216, 16777216, 11529216, 16777216, 16777216, 14411519, 16777216, 16777216, 16777216, 9007200, 16777216, 16777216, 11259000, 16777216, 16777216, 14073749, 16777216, 16777216, 16777216, 8796094, 16777216, 16777216, 10995117, 16777216, 16777216, 16777216};
__constant__ unsigned long long kf_p10_hi[28] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 5ull, 54ull, 542ull, 5421ull, 54210ull, 542101ull, 5421010ull, 54210108ull};
__constant__ unsigned long long kf_p10_lo[28] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull, 7766279631452241920ull, 3875820019684212736ull, 1864712049423024128ull, 200376420520689664ull, 2003764205206896640ull, 1590897978359414784ull, 15908979783594147840ull, 11515845246265065472ull};

struct KfOut {                     // a staging row being written; `bad` = this read goes to the host formatter
    char *p, *end; bool bad;
    __device__ __forceinline__ void put(char c) { if (p < end) *p++ = c; else bad = true; }
    __device__ __forceinline__ void str(const char *s) { while (*s) put(*s++); }
};
__device__ __forceinline__ void kf_u32(KfOut &o, uint32_t v) {
    char d[10]; int n = 0;
    do { d[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (n) o.put(d[--n]);
}
__device__ __forceinline__ void kf_i32(KfOut &o, int32_t v) {
    if (v < 0) { o.put('-'); kf_u32(o, (uint32_t)(-(long long)v)); } else kf_u32(o, (uint32_t)v);
}
// 2^-70 <= |f| < 1e-4: six significant digits in exponent notation, "d[.ddddd]e-XX" (or "0.0001" when the value rounds up to it)
__device__ __forceinline__ void kf_g_small(KfOut &o, uint32_t bits) {
    const uint32_t ab = bits & 0x7FFFFFFFu;
    const int b2 = (int)(ab >> 23) - 127;                          // -70 .. -14
    const uint32_t m = (ab & 0x7FFFFFu) | 0x800000u;
    int e = kf_e_low2[b2 + 70] + (m >= kf_thr2[b2 + 70]);          // -22 .. -5
    const unsigned __int128 p10 = ((unsigned __int128)kf_p10_hi[5 - e] << 64) | kf_p10_lo[5 - e];
    const unsigned __int128 scaled = p10 * m;
    const int sh = 23 - b2;                                        // 37 .. 93
    unsigned long long n = (unsigned long long)(scaled >> sh);
    const unsigned __int128 rem = scaled & ((((unsigned __int128)1) << sh) - 1), half = ((unsigned __int128)1) << (sh - 1);
    n += (unsigned long long)((rem > half) | ((rem == half) & (n & 1)));
    if (n >= 1000000) { n = 100000; e++; }
    if (bits >> 31) o.put('-');
    char d[6];
    uint32_t n32 = (uint32_t)n;
#pragma unroll
    for (int i = 5; i >= 0; i--) { d[i] = (char)('0' + n32 % 10u); n32 /= 10u; }
    int last = 5;
    while (last > 0 && d[last] == '0') last--;
    if (e >= -4) { o.str("0.000"); for (int i = 0; i <= last; i++) o.put(d[i]); return; }      // rounded up to 1e-4: fixed notation
    o.put(d[0]);
    if (last > 0) { o.put('.'); for (int i = 1; i <= last; i++) o.put(d[i]); }
    o.put('e'); o.put('-'); o.put((char)('0' + (-e) / 10)); o.put((char)('0' + (-e) % 10));
}
// printf("%g") of a float: km_fmt_g of kmat_host.cpp (same tables, same integer arithmetic); values it hands to the C library
// there (tiny, huge, inf, nan) mark the read for the host instead
__device__ __forceinline__ void kf_g(KfOut &o, float f) {
    const uint32_t bits = __float_as_uint(f), ab = bits & 0x7FFFFFFFu;
    if (ab - 0x38D1B718u >= 0x497423F0u - 0x38D1B718u) {
        if (ab == 0) { if (bits >> 31) o.put('-'); o.put('0'); }
        else if (ab < 0x38D1B718u && ab >= ((127u - 70u) << 23)) kf_g_small(o, bits);
        else o.bad = true;
        return;
    }
    const int b2 = (int)(ab >> 23) - 127;
    const uint32_t m = (ab & 0x7FFFFFu) | 0x800000u;
    int e = kf_e_low[b2 + 14] + (m >= kf_thr[b2 + 14]);
    const unsigned long long scaled = (unsigned long long)m * kf_i10[5 - e];
    const int sh = 23 - b2;
    unsigned long long n = scaled >> sh;
    const unsigned long long rem = scaled & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    n += (unsigned long long)((rem > half) | ((rem == half) & (n & 1)));
    if (n >= 1000000) { n = 100000; e++; if (e > 5) { o.bad = true; return; } }
    if (bits >> 31) o.put('-');
    char d[6];
    uint32_t n32 = (uint32_t)n;
#pragma unroll
    for (int i = 5; i >= 0; i--) { d[i] = (char)('0' + n32 % 10u); n32 /= 10u; }
    int last = 5;
    while (last > 0 && d[last] == '0') last--;                     // d[0] is never '0'
    if (e >= 0) {
        for (int i = 0; i <= e; i++) o.put(d[i]);                  // integer part: digits 0 .. e (zeros included)
        if (last > e) { o.put('.'); for (int i = e + 1; i <= last; i++) o.put(d[i]); }
    } else {
        o.put('0'); o.put('.');
        for (int i = 0; i < -e - 1; i++) o.put('0');
        for (int i = 0; i <= last; i++) o.put(d[i]);
    }
}
__device__ __forceinline__ void kf_pair(KfOut &o, uint32_t tid, float score) { kf_u32(o, tid); o.put(' '); kf_g(o, score); }
__device__ __forceinline__ void kf_match(KfOut &o, int m) {
    switch (m) {
        case KMAT_DIRECT: o.str("DirectMatch"); break;
        case KMAT_MULTI: o.str("MultiMatch"); break;
        case KMAT_PARTIAL: o.str("PartialMultiMatch"); break;
        case KMAT_NOMATCH: o.str("NoMatch"); break;
        default: o.str("LCA_ERROR"); break;
    }
}
// kmat_format_tail, statement for statement
__device__ __forceinline__ void kf_tail(KfOut &o, const kmat_read_result &r, const kmat_pair *cands, const kmat_pair *lin, int prn_all) {
    switch (r.status) {
        case KMAT_ST_SHORT_LEN: case KMAT_ST_SHORT_VALID:
            o.str("-1 -1 -1\t-1 -1\t"); kf_i32(o, r.n1); o.put(' '); kf_i32(o, r.n2); o.str(" ReadTooShort\n");
            break;
        case KMAT_ST_NODBHITS:
            o.str("-1 -1 "); kf_i32(o, r.valid_kmers); o.str("\t-1 -1\t"); kf_i32(o, r.n1); o.put(' '); kf_i32(o, r.n2); o.str(" NoDbHits\n");
            break;
        case KMAT_ST_SILENT: break;
        case KMAT_ST_PHIX:
            o.str("-1 -1 "); kf_i32(o, r.cand_kmer_cnt); o.put('\t');
            kf_u32(o, r.tid); o.put(' '); kf_g(o, r.score); o.put('\t');
            kf_u32(o, r.tid); o.put(' '); kf_g(o, r.score); o.put(' '); kf_match(o, KMAT_DIRECT); o.put('\n');
            break;
        case KMAT_ST_LABELED: {
            kf_g(o, r.log_avg); o.put(' '); kf_g(o, r.stdev); o.put(' '); kf_i32(o, r.cand_kmer_cnt); o.put('\t');
            if (prn_all) {
                bool prn = false;
                for (int i = (int)r.n_cand - 1; i >= 0 && !o.bad; --i) {
                    const kmat_pair c = cands[r.cand_off + (uint64_t)i];
                    if (c.score >= 0 || prn_all > 1) { o.put(' '); kf_pair(o, c.tid, c.score); prn = true; }
                }
                if (!prn) o.str("-1 -1");
                o.put('\t');
            }
            if (r.match == KMAT_DIRECT) { kf_pair(o, r.tid, r.score); o.put(' '); kf_match(o, r.match); }
            else if (r.match == KMAT_MULTI || r.match == KMAT_PARTIAL) {
                if (!prn_all) {
                    for (uint32_t i = 0; i < r.n_lin && !o.bad; i++) { o.put(' '); kf_pair(o, lin[r.lin_off + i].tid, lin[r.lin_off + i].score); }
                    if (!r.n_lin) o.str("-1 -1");
                    o.put('\t');
                }
                kf_pair(o, r.tid, r.score); o.put(' '); kf_match(o, r.match);
            } else if (r.match == KMAT_NOMATCH) { o.str("-1 -1 "); kf_match(o, r.match); }
            else o.str("-1 -1 Unmatched");
            o.put('\n');
            break;
        }
        default: o.bad = true; break;          // KMAT_ST_ERROR and anything internal: the host reports it
    }
}

struct KmFormatParams {
    const kmat_read_result *out; uint32_t n_reads;
    const kmat_pair *cands, *lin; int prn_all;
    char *text; unsigned long long text_cap; unsigned long long *text_cursor;     // the pass's text buffer and its running cursor
    unsigned long long *ref;                                                       // per read: offset << KMAT_TEXT_LEN_BITS | length, or KMAT_TEXT_ON_HOST
};
__global__ void __launch_bounds__(KF_THREADS) km_format_kernel(KmFormatParams F) {
    extern __shared__ __align__(16) char kf_stage[];
    const int lane = threadIdx.x & 31;
    char *warp_rows = kf_stage + (size_t)(threadIdx.x & ~31) * KF_STRIDE;
    const uint32_t n_warps = (gridDim.x * KF_THREADS) >> 5;
    for (uint32_t base = ((blockIdx.x * KF_THREADS + threadIdx.x) >> 5) * 32u; base < F.n_reads; base += n_warps * 32u) {
        const uint32_t r = base + lane;
        KfOut o;
        o.p = warp_rows + (size_t)lane * KF_STRIDE; o.end = o.p + KF_ROW; o.bad = false;
        char *const row = o.p;
        if (r < F.n_reads) kf_tail(o, F.out[r], F.cands, F.lin, F.prn_all);
        uint32_t len = (r < F.n_reads && !o.bad) ? (uint32_t)(o.p - row) : 0u;
        // the warp's rows go to one contiguous piece of the text buffer
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(KM_FULL, incl, d); if (lane >= d) incl += v; }
        const uint32_t total = __shfl_sync(KM_FULL, incl, 31);
        unsigned long long at = 0;
        if (lane == 0 && total) at = atomicAdd(F.text_cursor, (unsigned long long)total);
        at = kb_shfl64(at, 0);
        const bool fits = at + total <= F.text_cap;                       // warp-uniform: a piece that does not fit is left to the host
        const unsigned long long mine = at + (incl - len);
        if (r < F.n_reads) F.ref[r] = (o.bad || !fits) ? KMAT_TEXT_ON_HOST : ((mine << KMAT_TEXT_LEN_BITS) | len);
        __syncwarp();
        if (fits && total) {
            for (int t = 0; t < 32; t++) {
                const uint32_t lt = __shfl_sync(KM_FULL, len, t);
                const unsigned long long ot = kb_shfl64(mine, t);
                const char *src = warp_rows + (size_t)t * KF_STRIDE;
                for (uint32_t j = lane; j < lt; j += 32) F.text[ot + j] = src[j];
            }
        }
        __syncwarp();
    }
}

// test hook: kf_g of n floats, 16 bytes of text each (NUL padded; first byte 0xFF = left to the host)
__global__ void km_format_floats_kernel(const float *v, uint32_t n, char *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char buf[16];
    for (int j = 0; j < 16; j++) buf[j] = 0;
    KfOut o; o.p = buf; o.end = buf + 16; o.bad = false;
    kf_g(o, v[i]);
    if (o.bad) buf[0] = (char)0xFF;
    for (int j = 0; j < 16; j++) out[(size_t)i * 16 + j] = buf[j];
}
extern "C" int kmat_test_format_floats(int device, const float *vals, uint32_t n, char *out16) {
    if (!vals || !out16) return KMAT_ERR_ARG;
    KM_CUDA(cudaSetDevice(device));
    float *d_v = nullptr; char *d_o = nullptr;
    KM_CUDA(cudaMalloc((void **)&d_v, (size_t)n * 4 + 4));
    if (cudaMalloc((void **)&d_o, (size_t)n * 16 + 16) != cudaSuccess) { cudaFree(d_v); cudaGetLastError(); return KMAT_ERR_NOMEM; }
    cudaMemcpy(d_v, vals, (size_t)n * 4, cudaMemcpyHostToDevice);
    km_format_floats_kernel<<<(n + 255) / 256, 256>>>(d_v, n, d_o);
    const cudaError_t e = cudaMemcpy(out16, d_o, (size_t)n * 16, cudaMemcpyDeviceToHost);
    cudaFree(d_v); cudaFree(d_o);
    if (e != cudaSuccess) { kmat_set_error("kmat_test_format_floats: %s", cudaGetErrorString(e)); cudaGetLastError(); return KMAT_ERR_CUDA; }
    return KMAT_OK;
}

// Queue K5 for the reads [0, n) of a finished pass (results at d_out) on `st`.
static int km_launch_format(kmat_ctx *c, const kmat_read_result *d_out, uint32_t n, int prn_all, uint64_t text_cap, unsigned long long *d_ref, cudaStream_t st) {
    KmFormatParams F;
    F.out = d_out; F.n_reads = n; F.cands = c->d_cands; F.lin = c->d_lin; F.prn_all = prn_all;
    F.text = c->d_text; F.text_cap = text_cap; F.text_cursor = c->d_cursors + 2; F.ref = d_ref;
    const int smem = KF_THREADS * KF_STRIDE;
    KM_CUDA(cudaFuncSetAttribute(km_format_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      // per device: cheap, so per launch
    const int grid = (int)std::min<uint32_t>((n + KF_THREADS - 1) / KF_THREADS, (uint32_t)c->sms * 2u);
    km_format_kernel<<<std::max(1, grid), KF_THREADS, smem, st>>>(F);
    g_km_launches++;
    KM_CUDA(cudaGetLastError());
    return KMAT_OK;
}
#endif
