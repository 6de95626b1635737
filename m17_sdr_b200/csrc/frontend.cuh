// frontend.cuh -- 48 kHz RX front end, batched over (channel, 40 ms block) pairs:
//   int16 IQ -> float (x 0.00003 in double) -> hard limiter -> 2-sample-baseline cross-product FM
//   discriminator -> keep every 5th -> block mean (sequential fp32 sum of all 1920 values).
// Replaces dsp_short_to_float, dsp_limit, dsp_arctan_disc2 as chained by m17_dsp_rx
// (m17_dsp.cpp:136-141,412-419,194-222,461-472).
//
// With AFC off (the reference default, radio.cpp:146-155) nothing but two input samples and the /5 phase
// crosses a block boundary, so every (channel, block) item is independent.  The only serial work inside an
// item is the fp32 running sum, whose order must be kept for bit-exactness, hence:
//   mapping: one LANE per item, one warp per 32 items.  Each lane streams its own 7680-byte row with 16-byte
//   loads (two alternating register buffers, one 20-sample chunk ahead; 32-byte sectors fully consumed through L1), and
//   walks it sequentially with the discriminator history in registers.  Because 20 = lcm(4 samples per load,
//   5 = decimation), the kept sample of every group of five sits at a lane-constant position: no counters.
//   Kept outputs collect in a [32][33] shared tile that is flushed with coalesced 128-byte row stores every
//   160 samples.  HBM traffic per item: 7680 B in, 1536 + 4 B out (the algorithmic minimum of the staged design).
// The raw (not mean-removed) discriminator samples and the block mean are written; the consumer applies
// out[i] - mean (one fp32 subtract, identical to m17_dsp.cpp:217-219).
//
// Arithmetic (all verified EXHAUSTIVELY on the GPU by m17b_selftest_frontend over the 2^32 possible IQ samples):
//   * (float)((double)x * 0.00003) == fmaf(x, c_hi, x * c_lo) with c_hi + c_lo the two-float split of 0.00003;
//   * m = sqrtf(re^2 + im^2) and g = (float)(1.0 / m) (== correctly rounded fp32 reciprocal, 2p+2 theorem) are
//     produced from ONE rsqrt.approx seed with FMA residual corrections instead of the branchy library sequences.
#pragma once
#include "decode.cuh"

#define FE_WARPS 4

struct LimSample { float re, im; };

// reference formulation, kept for the exhaustive self-test
__device__ __forceinline__ LimSample fe_limit_ieee(uint32_t raw, float *mo = nullptr, float *go = nullptr) {
    const int re_i = (int)(int16_t)(raw & 0xFFFFu), im_i = (int)(int16_t)(raw >> 16);
    float re = __double2float_rn((double)re_i * 0.00003);     // dsp_short_to_float, m17_dsp.cpp:138-139
    float im = __double2float_rn((double)im_i * 0.00003);
    float m = sqrtf(re * re + im * im);                       // dsp_limit, m17_dsp.cpp:414-417
    float g = 1.0f / m;
    if (mo) { *mo = m; *go = g; }
    LimSample s;
    s.re = re * g;
    s.im = im * g;
    return s;
}

__device__ __forceinline__ LimSample fe_limit(uint32_t raw, float *mo = nullptr, float *go = nullptr) {
    // int16 -> float: one I2F.S16 per component, reading the half-register directly (exact); from here (re, im) travel as
    // one packed pair: FMUL2 / FFMA2 do the two-float scaling, the squares and the final limiter scaling in one issue slot each
    const f32x2 x = pack2((float)(short)(raw & 0xFFFFu), (float)(short)(raw >> 16));
    constexpr float c_hi = 0.00003f;
    constexpr float c_lo = (float)(0.00003 - (double)0.00003f);
    const f32x2 v = fma2(x, pack2(c_hi, c_hi), mul2(x, pack2(c_lo, c_lo)));      // (re, im) = fmaf(x, c_hi, x * c_lo)
    float q0, q1;
    unpack2(mul2(v, v), q0, q1);
    const float s = q0 + q1;                                   // two rounded products, one rounded (scalar) sum: no contraction
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
    // m = RN(sqrt(s)): Newton step on the residual
    float m = s * y;
    const float h = 0.5f * y;
    m = __fmaf_rn(__fmaf_rn(-m, m, s), h, m);
    // g = RN(1/m): two residual corrections starting from y ~ 1/m (Markstein).  The one input class this cannot round
    // correctly is a divisor whose significand is all ones: the Newton iterate then lands exactly on a rounding midpoint
    // while the true quotient 2^-(e+1) (1 + 2^-24 + 2^-48 + ..) lies just above it; its correctly rounded value is known in
    // closed form, 2^-(e+1) (1 + 2^-23), whose bit pattern is 0x7F000000 - bits(m).
    float g = __fmaf_rn(__fmaf_rn(-m, y, 1.0f), y, y);
    g = __fmaf_rn(__fmaf_rn(-m, g, 1.0f), g, g);
    const uint32_t mb = __float_as_uint(m);
    if (((mb + 1u) & 0x7FFFFFu) == 0u) g = __uint_as_float(0x7F000000u - mb);
    if (mo) { *mo = m; *go = g; }
    LimSample o;
    unpack2(mul2(v, pack2(g, g)), o.re, o.im);
    return o;
}

// 5-way select by a lane-constant position
__device__ __forceinline__ float sel5(float a0, float a1, float a2, float a3, float a4, int k) {
    float lo = (k == 0) ? a0 : a1;
    float hi = (k == 2) ? a2 : a3;
    float r = (k < 2) ? lo : hi;
    return (k == 4) ? a4 : r;
}

// Items are the (channel, block) pairs of blocks [t0, t0+Tc) of a call of T blocks per channel (T is the row pitch of iq,
// disc and mean); the whole call is t0 = 0, Tc = T.  Sub-ranges let the host pipeline the front end of one time slice with
// the timing loop of the previous one (rx.cuh).
__global__ void __launch_bounds__(FE_WARPS * 32) k_frontend(const uint32_t *__restrict__ iq, int64_t nchan, int64_t T, int64_t t0, int64_t Tc,
                                                            RxChanState *st, float *__restrict__ disc, float *__restrict__ mean) {
    __shared__ float tout[FE_WARPS][32][33];
    __shared__ int64_t gsl[FE_WARPS][32];                        // global (channel, block) index of each lane's item
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nitems = nchan * Tc;
    const int64_t item0 = ((int64_t)blockIdx.x * FE_WARPS + wid) * 32;
    if (item0 >= nitems) return;
    const bool live = item0 + lane < nitems;
    const int64_t item = live ? item0 + lane : nitems - 1;       // dead lanes shadow the last item (results discarded)
    const int64_t ch = item / Tc, t = t0 + item % Tc;
    const int64_t g = ch * T + t;
    gsl[wid][lane] = g;
    const uint4 *row = (const uint4 *)(iq + g * 1920);

    // carried discriminator state: z[0], z[1] are the two previous LIMITED samples (m17_dsp.cpp:196,205-206)
    float z0re, z0im, z1re, z1im;
    const int count0 = st[ch].disc_count;                        // 1920 % 5 == 0: the /5 phase is the same in every block
    if (t == 0) { z0re = st[ch].z0re; z0im = st[ch].z0im; z1re = st[ch].z1re; z1im = st[ch].z1im; }
    else {
        const uint32_t *prev = iq + g * 1920;
        LimSample a = fe_limit(__ldg(prev - 1)), b = fe_limit(__ldg(prev - 2));
        z0re = a.re; z0im = a.im; z1re = b.re; z1im = b.im;
    }
    const int keep = 4 - count0;                                 // count = (count+1)%5 hits 0 at samples = keep (mod 5)
    float acc = 0.0f;                                            // sum of u; sum of u*0.5 == 0.5*sum (exact power-of-two scaling)
    // 20 samples (five 16-byte loads) per chunk; two register buffers alternate so the next chunk's loads are in flight
    // while the current one is processed, without register copies
    auto process20 = [&](const uint4 (&w)[5], int slot) {
        float u[20];
#pragma unroll
        for (int s = 0; s < 20; s++) {
            const uint4 q = w[s >> 2];
            const uint32_t raw = (s & 3) == 0 ? q.x : (s & 3) == 1 ? q.y : (s & 3) == 2 ? q.z : q.w;
            const LimSample x = fe_limit(raw);
            // dsp_arctan_disc2 (m17_dsp.cpp:203-212)
            const float a = z0im * (x.re - z1re);
            const float b = z0re * (x.im - z1im);
            u[s] = b - a;
            z1re = z0re; z1im = z0im; z0re = x.re; z0im = x.im;
            acc += u[s];
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            tout[wid][lane][slot * 4 + j] = sel5(u[5 * j], u[5 * j + 1], u[5 * j + 2], u[5 * j + 3], u[5 * j + 4], keep) * 0.5f;
    };
    uint4 bufa[5], bufb[5];
#pragma unroll
    for (int q = 0; q < 5; q++) bufa[q] = __ldg(row + q);
    for (int grp = 0; grp < 12; grp++) {
#pragma unroll 1
        for (int chk = 0; chk < 8; chk += 2) {
            const int c20 = grp * 8 + chk;
#pragma unroll
            for (int q = 0; q < 5; q++) bufb[q] = __ldg(row + (c20 + 1) * 5 + q);
            process20(bufa, chk);
            if (c20 + 2 < 96) {
#pragma unroll
                for (int q = 0; q < 5; q++) bufa[q] = __ldg(row + (c20 + 2) * 5 + q);
            }
            process20(bufb, chk + 1);
        }
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < 32; r++)
            if (item0 + r < nitems) disc[gsl[wid][r] * 384 + grp * 32 + lane] = tout[wid][r][lane];
        __syncwarp();
    }
    if (live) {
        mean[g] = (acc * 0.5f) / 1920.0f;                     // offset/len (m17_dsp.cpp:214)
        if (t == T - 1) { st[ch].nz0re = z0re; st[ch].nz0im = z0im; st[ch].nz1re = z1re; st[ch].nz1im = z1im; }
    }
}

// Exhaustive proof-by-enumeration that fe_limit == fe_limit_ieee for every possible int16 IQ pair except (0,0)
// (which the reference itself turns into NaN).  Counts bitwise mismatches of either output component.
__global__ void k_selftest_frontend(unsigned long long *mism, uint32_t lo, uint32_t count, uint32_t *dump, int dump_cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = 0;
    for (uint32_t k = i; k < count; k += gridDim.x * blockDim.x) {
        const uint32_t raw = lo + k;
        if (raw == 0) continue;
        float m1, g1, m2, g2;
        const LimSample a = fe_limit(raw, &m1, &g1), b = fe_limit_ieee(raw, &m2, &g2);
        if ((__float_as_uint(a.re) != __float_as_uint(b.re)) | (__float_as_uint(a.im) != __float_as_uint(b.im))) {
            bad++;
            if (dump) {
                unsigned long long slot = atomicAdd(mism + 1, 1ull);
                if (5 * slot + 4 < (unsigned long long)dump_cap) {
                    dump[5 * slot] = raw; dump[5 * slot + 1] = __float_as_uint(m1); dump[5 * slot + 2] = __float_as_uint(m2);
                    dump[5 * slot + 3] = __float_as_uint(g1); dump[5 * slot + 4] = __float_as_uint(g2);
                }
            }
        }
    }
    if (bad) atomicAdd(mism, (unsigned long long)bad);
}
// h_dump (optional, dump_cap words) receives {raw, m_fast, m_ieee, g_fast, g_ieee} of the first mismatches found
extern "C" int m17b_selftest_frontend(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream) {
    if (!ctx || !h_mismatches || first + count > (1ull << 32) || dump_cap < 0) return M17B_E_ARG;
    cudaStream_t st = as_stream(stream);
    unsigned long long *d;
    uint32_t *d_dump = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, 16));
    CUDA_TRY(cudaMemsetAsync(d, 0, 16, st));
    if (h_dump && dump_cap) { CUDA_TRY(cudaMalloc((void **)&d_dump, 4 * (size_t)dump_cap)); CUDA_TRY(cudaMemsetAsync(d_dump, 0, 4 * (size_t)dump_cap, st)); }
    for (uint64_t off = 0; off < count; off += (1ull << 30)) {
        const uint64_t n = count - off < (1ull << 30) ? count - off : (1ull << 30);
        k_selftest_frontend<<<148 * 16, 256, 0, st>>>(d, (uint32_t)(first + off), (uint32_t)n, d_dump, dump_cap);
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
