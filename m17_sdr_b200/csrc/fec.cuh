// fec.cuh -- batched bit-domain primitives: CRC-16, Golay(24,12), convolutional encoder, puncture /
// de-puncture, QPP (de)interleave, (de)randomise, soft demap, sync-word correlator, PRBS9, and the
// stand-alone batched Viterbi decoder.  One thread per output element wherever the operation allows it,
// so global loads/stores coalesce.
#pragma once
#include "tables.cuh"

static inline unsigned grid_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---------------------------------------------------------------- CRC-16  (m17_crc.cpp:26-35)
__global__ void k_crc(const uint8_t *in, int64_t stride, int len, int64_t n, uint16_t *out, const uint16_t *g_tab) {
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_tab[i];
    __syncthreads();
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t *p = in + r * stride;
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < len; i++) crc = crc16_step(crc, p[i], tab);
    out[r] = crc;
}
extern "C" int m17b_crc_array_encode(m17b_ctx *ctx, const uint8_t *d_in, int64_t stride, int len, int64_t n, uint16_t *d_crc, void *stream) {
    if (!ctx || !d_in || !d_crc || len < 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_crc<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_in, stride, len, n, d_crc, ctx->d_crc);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- Golay(24,12)  (m17_golay.cpp:94-116)
__global__ void k_golay_enc(const uint16_t *in, int64_t n, uint32_t *out, const uint16_t *genc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t d = in[r];
    out[r] = (d << 12) | __ldg(&genc[d & 0xFFF]);
}
__device__ __forceinline__ int golay_decode_word(uint32_t word, const uint16_t *genc, const uint16_t *gerr, uint32_t *data_out) {
    uint32_t data = (word >> 12) & 0xFFF, parity = word & 0xFFF;
    uint32_t e = __ldg(&gerr[parity ^ __ldg(&genc[data])]);
    *data_out = data ^ (e & 0xFFF);
    return (int)(e >> 12);
}
__global__ void k_golay_dec(const uint32_t *in, int64_t n, uint16_t *data, uint8_t *err, const uint16_t *genc, const uint16_t *gerr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t d;
    int e = golay_decode_word(in[r], genc, gerr, &d);
    data[r] = (uint16_t)d;
    err[r] = (uint8_t)e;
}
extern "C" int m17b_golay_encode(m17b_ctx *ctx, const uint16_t *d_data, int64_t n, uint32_t *d_words, void *stream) {
    if (!ctx || !d_data || !d_words || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_golay_enc<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(d_data, n, d_words, ctx->d_genc);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_golay_decode(m17b_ctx *ctx, const uint32_t *d_words, int64_t n, uint16_t *d_data, uint8_t *d_err, void *stream) {
    if (!ctx || !d_words || !d_data || !d_err || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_golay_dec<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(d_words, n, d_data, d_err, ctx->d_genc, ctx->d_gerr);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- convolutional encoder (m17_conv.cpp:22-71)
// Register after shifting in bit t: reg = sum_k b[t-k] << (4-k).  Outputs G1 = reg bits {4,1,0}, G2 = bits {4,3,2,0}.
__device__ __forceinline__ uint32_t conv_pair(uint32_t reg) {
    uint32_t g1 = ((reg >> 4) ^ (reg >> 1) ^ reg) & 1u;
    uint32_t g2 = ((reg >> 4) ^ (reg >> 3) ^ (reg >> 2) ^ reg) & 1u;
    return (g1 << 1) | g2;
}
// bits are supplied by a functor: bit(t) for t in [0,nbits), 0 outside (the 4-step zero tail)
template <class BitFn> __device__ __forceinline__ uint32_t conv_reg(BitFn bit, int t, int nbits) {
    uint32_t reg = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int q = t - k;
        uint32_t b = (q >= 0 && q < nbits) ? bit(q) : 0u;
        reg |= b << (4 - k);
    }
    return reg;
}
__global__ void k_conv_enc(const uint8_t *in, int nin, int nbits, int bytes_mode, int64_t n, uint8_t *out) {
    const int steps = nbits + 4;
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * steps) return;
    int64_t f = gid / steps;
    int t = (int)(gid % steps);
    const uint8_t *p = in + f * nin;
    auto bit = [&](int q) -> uint32_t { return bytes_mode ? (uint32_t)((p[q >> 3] >> (7 - (q & 7))) & 1) : (uint32_t)(p[q] != 0); };
    uint32_t pr = conv_pair(conv_reg(bit, t, nbits));
    uint8_t *o = out + f * (2 * steps) + 2 * t;
    o[0] = (uint8_t)(pr >> 1);
    o[1] = (uint8_t)(pr & 1);
}
extern "C" int m17b_conv_encode_8(m17b_ctx *ctx, const uint8_t *d_in, int nbytes, int64_t n, uint8_t *d_out, void *stream) {
    if (!ctx || !d_in || !d_out || nbytes <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_conv_enc<<<grid_for(n * (8 * nbytes + 4), 256), 256, 0, as_stream(stream)>>>(d_in, nbytes, 8 * nbytes, 1, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_conv_encode_1(m17b_ctx *ctx, const uint8_t *d_in, int nbits, int64_t n, uint8_t *d_out, void *stream) {
    if (!ctx || !d_in || !d_out || nbits <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_conv_enc<<<grid_for(n * (nbits + 4), 256), 256, 0, as_stream(stream)>>>(d_in, nbits, nbits, 0, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- puncture / de-puncture (m17_puncture.cpp:4-79)
__device__ __forceinline__ bool d_punct_keeps(int pattern, int i) {
    if (pattern == 1) return ((i % 61) & 3) != 2;
    if (pattern == 2) return (i % 12) != 11;
    return (i % 8) != 7;
}
// number of kept positions among coded positions [0, i)
__device__ __host__ __forceinline__ int punct_kept_before(int pattern, int i) {
    if (pattern == 1) { int q = i / 61, r = i % 61; return q * 46 + r - (r + 1) / 4; }   // dropped at r = 2,6,..: (r+1)/4 of them below r
    if (pattern == 2) return i - i / 12;
    return i - i / 8;
}
__global__ void k_punc(int pattern, const uint8_t *in, int len, int out_len, int64_t n, uint8_t *out) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    int64_t f = gid / len;
    int i = (int)(gid % len);
    if (d_punct_keeps(pattern, i)) out[f * out_len + punct_kept_before(pattern, i)] = in[gid];
}
__global__ void k_depunc(int pattern, const float *in, int in_len, int len, int64_t n, float *out) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    int64_t f = gid / len;
    int i = (int)(gid % len);
    out[gid] = d_punct_keeps(pattern, i) ? in[f * in_len + punct_kept_before(pattern, i)] : 0.0f;
}
extern "C" int m17b_punc(m17b_ctx *ctx, int pattern, const uint8_t *d_in, int len, int64_t n, uint8_t *d_out, int *out_len, void *stream) {
    if (!ctx || !d_in || !d_out || pattern < 1 || pattern > 3 || len <= 0 || n < 0) return M17B_E_ARG;
    if (d_in == d_out) return M17B_E_ARG;   // the batched form is out-of-place
    int kept = punct_kept_before(pattern, len);
    if (out_len) *out_len = kept;
    if (n == 0) return M17B_OK;
    k_punc<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(pattern, d_in, len, kept, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_de_punc(m17b_ctx *ctx, int pattern, const float *d_in, int in_len, int len, int64_t n, float *d_out, void *stream) {
    if (!ctx || !d_in || !d_out || pattern < 1 || pattern > 3 || len <= 0 || n < 0 || d_in == d_out) return M17B_E_ARG;
    if (in_len < punct_kept_before(pattern, len)) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_depunc<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(pattern, d_in, in_len, len, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- QPP (de)interleave  (m17_interleave.cpp:3-12)
// The reference scatters out[pi(i)] = in[i]; pi is an involution, so the coalesced gather out[j] = in[pi(j)] is identical.
template <class T> __global__ void k_qpp(const T *in, int64_t n, T *out) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * 368) return;
    int64_t f = gid / 368;
    int j = (int)(gid % 368);
    out[gid] = in[f * 368 + c_tx.qpp[j]];
}
extern "C" int m17b_interleave(m17b_ctx *ctx, const uint8_t *d_in, int64_t n, uint8_t *d_out, void *stream) {
    if (!ctx || !d_in || !d_out || n < 0 || d_in == d_out) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_qpp<uint8_t><<<grid_for(n * 368, 256), 256, 0, as_stream(stream)>>>(d_in, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_de_interleave(m17b_ctx *ctx, const float *d_in, int64_t n, float *d_out, void *stream) {
    if (!ctx || !d_in || !d_out || n < 0 || d_in == d_out) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_qpp<float><<<grid_for(n * 368, 256), 256, 0, as_stream(stream)>>>(d_in, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- (de)randomiser  (m17_correlate.cpp:11-31)
__global__ void k_derand_bytes(uint8_t *io, int len, int64_t n) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    int i = (int)(gid % len) % 46;
    uint8_t key = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) key = (uint8_t)((key << 1) | c_tx.rnd[8 * i + b]);
    io[gid] ^= key;
}
__global__ void k_derand_u8(const uint8_t *in, uint8_t *out, int len, int64_t n) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    out[gid] = (in[gid] ^ c_tx.rnd[(int)(gid % len) % 368]) & 1;
}
__global__ void k_derand_f32(const float *in, float *out, int len, int64_t n) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    float v = in[gid];
    out[gid] = c_tx.rnd[(int)(gid % len) % 368] ? -v : v;
}
extern "C" int m17b_de_correlate_8(m17b_ctx *ctx, uint8_t *d_io, int len, int64_t n, void *stream) {
    if (!ctx || !d_io || len <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_derand_bytes<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(d_io, len, n);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_de_correlate_1_u8(m17b_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, int len, int64_t n, void *stream) {
    if (!ctx || !d_in || !d_out || len <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_derand_u8<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(d_in, d_out, len, n);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_de_correlate_1_f32(m17b_ctx *ctx, const float *d_in, float *d_out, int len, int64_t n, void *stream) {
    if (!ctx || !d_in || !d_out || len <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_derand_f32<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(d_in, d_out, len, n);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- soft demap  (m17_dsp.cpp:35-42,82-95)
// cor = 8.0/sum in double, rounded to float: by the 2p+2 theorem (53 >= 2*24+2) this equals the correctly
// rounded fp32 quotient 8.0f/sum, which is what operator/ gives without fast-math.
__device__ __forceinline__ float demap_cor(const float *sync8) {
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += fabsf(sync8[i]);
    return 8.0f / s;
}
__global__ void k_demap(const float *sym, int64_t n, float *soft) {
    // one warp per frame: lanes cover the 184 payload symbols; every lane recomputes cor from the 8 sync symbols
    int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (f >= n) return;
    const float *s = sym + f * 192;
    float hdr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) hdr[i] = s[i];
    float cor = demap_cor(hdr);
    float2 *o = (float2 *)(soft + f * 368);
    for (int k = lane; k < 184; k += 32) {
        float v = s[8 + k];
        o[k] = make_float2(demap_soft(v, cor, false), demap_soft(v, cor, true));
    }
}
extern "C" int m17b_demap_frame(m17b_ctx *ctx, const float *d_sym, int64_t n, float *d_soft, void *stream) {
    if (!ctx || !d_sym || !d_soft || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_demap<<<grid_for(n * 32, 256), 256, 0, as_stream(stream)>>>(d_sym, n, d_soft);
    KERNEL_CHECK();
    return M17B_OK;
}

// Exhaustive proof-by-enumeration that demap_lsb == demap_lsb_ieee for every float bit pattern of m (NaN results only have to
// be NaN in both), and that the sign-only hard decision of the frame decoder agrees with the sign of that value.
__global__ void k_selftest_demap(unsigned long long *mism, uint32_t lo, uint32_t count, uint32_t *dump, int dump_cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = 0;
    for (uint32_t k = i; k < count; k += gridDim.x * blockDim.x) {
        const float m = __uint_as_float(lo + k);
        const float a = demap_lsb(m), b = demap_lsb_ieee(m);
        const bool same = (a != a && b != b) || __float_as_uint(a) == __float_as_uint(b);
        const bool hard_ok = ((fabsf(m) >= 0.66660005f) == (b >= 0.0f)) && ((fabsf(m) < 0.66660005f) == (-b >= 0.0f));
        if (!same || !hard_ok) {
            bad++;
            if (dump) {
                const unsigned long long slot = atomicAdd(mism + 1, 1ull);
                if (3 * slot + 2 < (unsigned long long)dump_cap) { dump[3 * slot] = lo + k; dump[3 * slot + 1] = __float_as_uint(a); dump[3 * slot + 2] = __float_as_uint(b); }
            }
        }
    }
    if (bad) atomicAdd(mism, (unsigned long long)bad);
}
// h_dump (optional, dump_cap words) receives {bits(m), fast, reference} of the first mismatches found
extern "C" int m17b_selftest_demap(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream) {
    if (!ctx || !h_mismatches || first + count > (1ull << 32) || dump_cap < 0) return M17B_E_ARG;
    cudaStream_t st = as_stream(stream);
    unsigned long long *d;
    uint32_t *d_dump = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, 16));
    CUDA_TRY(cudaMemsetAsync(d, 0, 16, st));
    if (h_dump && dump_cap) { CUDA_TRY(cudaMalloc((void **)&d_dump, 4 * (size_t)dump_cap)); CUDA_TRY(cudaMemsetAsync(d_dump, 0, 4 * (size_t)dump_cap, st)); }
    for (uint64_t off = 0; off < count; off += (1ull << 30)) {
        const uint64_t n = count - off < (1ull << 30) ? count - off : (1ull << 30);
        k_selftest_demap<<<148 * 16, 256, 0, st>>>(d, (uint32_t)(first + off), (uint32_t)n, d_dump, dump_cap);
    }
    KERNEL_CHECK();
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, st));
    if (d_dump) CUDA_TRY(cudaMemcpyAsync(h_dump, d_dump, 4 * (size_t)dump_cap, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaFree(d));
    if (d_dump) CUDA_TRY(cudaFree(d_dump));
    *h_mismatches = h;
    return M17B_OK;
}

// ---------------------------------------------------------------- sync-word correlator (m17_rx_frame.cpp:22-81)
struct SyncResult { int type, votes; float variance; float spread, vmax; };   // spread = max|v| - min|v| (rounded), vmax = max|v|: variance = spread / vmax
// sync templates as sign masks, bit i set = template[i] is -1 (m17_rx_frame.cpp:5-12): preamble, LSF 0x55F7, stream 0xFF5D,
// packet 0x75FF, BERT 0xDF55, EOT 0x555D.  Compile-time so the six correlations are plain add/subtract chains.
__host__ __device__ constexpr unsigned sync_neg_mask(int t) {
    return t == 0 ? 0xAAu : t == 1 ? 0xB0u : t == 2 ? 0x4Fu : t == 3 ? 0xF2u : t == 4 ? 0x0Du : 0x40u;
}
template <int T_> __device__ __forceinline__ float sync_corr(const float *v) {
    constexpr unsigned m = sync_neg_mask(T_);
    float s = (m & 1u) ? -v[0] : v[0];                                       // vect[0]*sframe[t][0], exact
#pragma unroll
    for (int i = 1; i < 8; i++) s = ((m >> i) & 1u) ? s - v[i] : s + v[i];   // sums[t] += vect[i]*sframe[t][i]  (a + (-b) == a - b)
    return s;
}
// find_variance (m17_rx_frame.cpp:47-60) of a window, compared with a double limit exactly as the callers do
__device__ __forceinline__ bool sync_variance_lt(const float *v, double limit) {
    float mn = fabsf(v[0]), mx = mn;
#pragma unroll
    for (int i = 1; i < 8; i++) { float a = fabsf(v[i]); if (a > mx) mx = a; else if (a < mn) mn = a; }
    float var = (mx - mn) / mx;
    if (var != var) var = 1.0f;
    // the callers compare (double)var with 0.3 / 0.5 (m17_rx_frame.cpp:82-103).  No float lies in [0.3 (double), 0.3f), and 0.5 is
    // a float, so the comparison in float with the rounded limit decides identically
    return var < (float)limit;
}
// find_variance < 0.3 (the unlocked limit) without the divide: fl(d / m) < 0.3f <=> d / m < MID, the midpoint of 0.3f = 0x3E99999A and
// its predecessor (a tie rounds to the even 0.3f, which is not below the limit) <=> d < MID * m, and that product of a 25-bit and
// a 24-bit number is exact in double.  m = 0 or NaN: the reference's NaN -> 1.0 is not below the limit, the comparison is false too.
__device__ __forceinline__ bool sync_variance_lt03(const float *v) {
    float mn = fabsf(v[0]), mx = mn;
#pragma unroll
    for (int i = 1; i < 8; i++) { float a = fabsf(v[i]); if (a > mx) mx = a; else if (a < mn) mn = a; }
    return (double)(mx - mn) < 0x1.333333p-2 * (double)mx;
}
// variance < 0.5 (the locked limit) without waiting for the quotient: for floats d, m > 0 (m not subnormal), fl(d / m) < 0.5
// <=> d < 0.5 m: any float d below 0.5 m is at least one ulp below it, so d / m <= 0.5 - 2^-25 and rounds below 0.5; d >= 0.5 m
// gives a quotient >= 0.5.  m = 0 or NaN: the reference's variance is NaN -> 1.0 -> not below 0.5, and the comparison is false too.
__device__ __forceinline__ bool sync_spread_lt_half(float d, float m) { return m >= 1e-30f ? d < 0.5f * m : (d / m) < 0.5f; }
__device__ __forceinline__ float sync_variance_of(float d, float m) { float var = d / m; if (var != var) var = 1.0f; return var; }
// DEFER_VAR: leave r.variance unset (the caller divides later, off the path that decides lock / loss)
template <bool DEFER_VAR = false>
__device__ __forceinline__ SyncResult sync_check8(const float *v) {
    SyncResult r;
    // find_variance: the 'else' means a sample that raises the max is never tested against the min
    float mn = fabsf(v[0]), mx = mn;
#pragma unroll
    for (int i = 1; i < 8; i++) { float a = fabsf(v[i]); if (a > mx) mx = a; else if (a < mn) mn = a; }
    r.spread = mx - mn; r.vmax = mx;
    r.variance = DEFER_VAR ? 0.0f : sync_variance_of(r.spread, r.vmax);
    unsigned negm = 0, posm = 0;                                             // bit i: v[i] < 0 / v[i] > 0 (both clear for 0 and NaN)
#pragma unroll
    for (int i = 0; i < 8; i++) { negm |= (v[i] < 0) ? (1u << i) : 0u; posm |= (v[i] > 0) ? (1u << i) : 0u; }
    // Fast path, exact: no zero / NaN element, the sign pattern IS template t's, and variance < 0.5 (every |v[i]| > 0.5 max|v|).
    // Template t then correlates at sum|v| and any other template -- they differ from t in at least one position -- at sum|v|
    // minus at least 2 * 0.5 max|v| >= sum|v| / 8, far beyond the rounding of eight adds: t wins the arg-max, with zero votes.
    if ((negm ^ posm) == 0xFFu && sync_spread_lt_half(r.spread, r.vmax)) {
        int t = -1;
#pragma unroll
        for (int k = 0; k < 6; k++) if (negm == sync_neg_mask(k)) t = k;
        if (t >= 0) { r.type = t; r.votes = 0; return r; }
    }
    // six 8-term correlations, sequential adds; arg-max with strict '>' starting from (0, type 0)
    float best = 0;
    int type = 0;
    { const float s = sync_corr<0>(v); if (s > best) { best = s; type = 0; } }
    { const float s = sync_corr<1>(v); if (s > best) { best = s; type = 1; } }
    { const float s = sync_corr<2>(v); if (s > best) { best = s; type = 2; } }
    { const float s = sync_corr<3>(v); if (s > best) { best = s; type = 3; } }
    { const float s = sync_corr<4>(v); if (s > best) { best = s; type = 4; } }
    { const float s = sync_corr<5>(v); if (s > best) { best = s; type = 5; } }
    r.type = type;
    // votes: vect[i]*sframe[type][i] < 0  <=>  v[i] < 0 where the template is +1, v[i] > 0 where it is -1
    const unsigned m = (unsigned)((0x400DF24FB0AAull >> (8 * type)) & 0xFFu);
    r.votes = __popc((negm & ~m & 0xFFu) | (posm & m));
    return r;
}
// m17_unlocked_sync_check / m17_locked_sync_check (m17_rx_frame.cpp:82-103); variance compared as double
__device__ __forceinline__ bool sync_accept(const SyncResult &r, bool locked) {
    if (r.votes > (locked ? 1 : 0)) return false;
    if (r.type < 1 || r.type > 4) return false;
    return locked ? sync_spread_lt_half(r.spread, r.vmax) : r.variance < 0.3f;      // (double)variance < 0.5 / 0.3, see sync_variance_lt
}
// m17_unlocked_sync_check on one window, with an exact early-out.  Acceptance needs votes == 0, type in 1..4 and variance < 0.3.
// variance < 0.3 means every |v[i]| >= 0.7 max|v| > 0, so no element is zero; votes == 0 then means the sign pattern of the
// window IS the winning template's (and a window whose signs equal a template's correlates with it at the maximum possible
// sum |v|, so that template wins the arg-max).  Hence: unless the sign pattern is exactly one of the four frame sync words,
// the window cannot be accepted and the six correlations / variance need not be computed.  (A NaN element makes every
// correlation NaN -> type 0 -> rejected; it also fails the sign test.)
__device__ __forceinline__ bool sync_unlocked_ok(const float *v) {
    unsigned negm = 0, posm = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { negm |= (v[i] < 0) ? (1u << i) : 0u; posm |= (v[i] > 0) ? (1u << i) : 0u; }
    if ((negm ^ posm) != 0xFFu) return false;
    if (negm != sync_neg_mask(1) && negm != sync_neg_mask(2) && negm != sync_neg_mask(3) && negm != sync_neg_mask(4)) return false;
    // The signs are exactly those of frame sync word t.  If in addition variance < 0.3, every |v[i]| >= 0.7 max|v|, so any other
    // template's correlation is smaller by at least 2 * 0.7 max|v| >= 17 % of sum|v| -- far beyond the rounding of eight adds --
    // and the arg-max is t with zero votes: the variance test alone decides (find_variance, m17_rx_frame.cpp:47-60, compared as
    // double with 0.3, :82-92).  If variance >= 0.3 the window is rejected whatever the arg-max.
    return sync_variance_lt(v, 0.3);
}
__global__ void k_sync_check(const float *vec, int64_t n, uint8_t *type, uint8_t *votes, float *var) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = vec[r * 8 + i];
    SyncResult s = sync_check8(v);
    type[r] = (uint8_t)s.type; votes[r] = (uint8_t)s.votes; var[r] = s.variance;
}
extern "C" int m17b_sync_check(m17b_ctx *ctx, const float *d_vec, int64_t n, uint8_t *d_type, uint8_t *d_votes, float *d_var, void *stream) {
    if (!ctx || !d_vec || !d_type || !d_votes || !d_var || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_sync_check<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_vec, n, d_type, d_votes, d_var);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- PRBS9 (m17_prbs9.cpp:16-32)
__global__ void k_prbs9(const int32_t *start, int len, int64_t n, uint8_t *out, const uint8_t *seq) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * len) return;
    int64_t f = gid / len;
    int i = (int)(gid % len);
    int s = start ? start[f] : 0;
    out[gid] = __ldg(&seq[(s + i) % 511]);
}
extern "C" int m17b_prbs9_tx_load(m17b_ctx *ctx, const int32_t *d_start, int len, int64_t n, uint8_t *d_out, void *stream) {
    if (!ctx || !d_out || len <= 0 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_prbs9<<<grid_for(n * len, 256), 256, 0, as_stream(stream)>>>(d_start, len, n, d_out, ctx->d_prbs);
    KERNEL_CHECK();
    return M17B_OK;
}
