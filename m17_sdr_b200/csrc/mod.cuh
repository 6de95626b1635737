// mod.cuh -- the 4FSK modulator as ONE kernel: polyphase RRC x os -> fp32 phase accumulator -> cos/sin -> int16 IQ.
// Replaces m17_mod_dibits / mod_filter / sub_filter / mod_fsk (m17_modulate.cpp:22-61,79-92).
//
// What is parallel and what is not.  A frequency sample is 31 products and 30 ordered adds of the last 31 symbol
// deviations (sub_filter :42-48) -- independent per sample.  The phase m_acc += f (:24) is a per-channel running fp32 sum
// whose rounding order must be kept, re-wrapped once per symbol through double (:33-37): a serial chain of
// os adds + one wrap per symbol and channel -- 107 cycles per symbol on B200 with the fp32 wrap below (22 dependent fp32
// instructions at 4.5 cycles; measured, benchmarks/lat_probe.cu), the floor of this kernel whatever the batch size.
// cos/sin/int16 are independent again.
//
// Mapping.  A CTA owns G consecutive channels for the whole call (G = 4..32, chosen so that two CTAs share an SM) and walks
// the time axis in chunks of S symbols (G*S <= 512 at os = 10).  Warp 0 is the SCAN warp (lane = channel); ten other warps are
// WORKERS.  Per chunk:
//   workers: symbols of the chunk (fetched one chunk ahead) -> deviations in smem; FIR -> work[b] (thread = (channel, phase,
//            16 symbols): 16 independent ordered sums, taps / deviation window streamed from smem); signal full[b];
//            wait scanned[b^1] of the PREVIOUS chunk; cos/sin of its phases -> 16-byte IQ stores.
//   scan   : wait full[b]; walk the rows of work[b] (frequency) into phase[b], four symbols per trip, 16-byte loads / stores,
//            only the add chain and the wrap on the critical path; signal scanned[b].
// Double buffers let the FIR of chunk i+1 and the cos/sin of chunk i-1 run beside the scan of chunk i; hand-over is by named
// barriers (bar.arrive / bar.sync).  Nothing leaves shared memory between the stages: HBM sees 1 byte per symbol in and
// 4*os bytes per symbol out.
//
// The wrap in fp32 (tx_wrap_fast): the reference computes a1 = (float)((double)acc / (2 pi)), f = modf(a1),
// acc = (float)((double)f * 2 pi).  Both roundings are reproduced with fp32 FMAs: the quotient by one Markstein-style correction
// with a two-float 2 pi and constants tuned by enumeration, the fraction by an add that rounds toward zero, the product by an
// error-free product plus the low word.  The result was compared with the double formulation for EVERY float with biased
// exponent 25..149 on the CPU (benchmarks/tx_wrap_enum.c) and is compared again on the GPU for all 2^32 bit patterns
// (m17b_selftest_tx_wrap, a -m gpu test): 0 mismatches for 2^-101 <= |acc| < 2^23 and +0.  The scan warp only flags values
// outside that range (they cannot occur with a wrapped accumulator and finite taps) and re-walks a flagged chunk with the
// double form.
#pragma once
#include "rx.cuh"

#define TXM_SEG 16                           // symbols per FIR task
#define TXM_NW 10                            // worker warps
#define TXM_NWT (TXM_NW * 32)
#define TXM_BAR (TXM_NWT + 32)               // threads that take part in the hand-over barriers: workers + scan warp
// Warp roles.  The scan warp's chain is latency-critical and issues one instruction every ~4.5 cycles; worker warps on the SAME
// scheduler (sub-partition = warp id mod 4) compete with it for issue slots.  Layout 1 therefore leaves the scan warp alone on
// its scheduler: 14 warps, warp 0 scans, warps 4 / 8 / 12 exit at once, the other ten work.
#ifndef TXM_LAYOUT
#define TXM_LAYOUT 1
#endif
#if TXM_LAYOUT == 1
#define TXM_THREADS (14 * 32)
__device__ __forceinline__ int txm_worker_index(int warp) { return (warp & 3) == 0 ? -1 : warp - 1 - (warp >> 2); }   // 1,2,3,5,6,7,9,10,11,13 -> 0..9
#else
#define TXM_THREADS (TXM_NWT + 32)
__device__ __forceinline__ int txm_worker_index(int warp) { return warp - 1; }
#endif

__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// clock read that cannot issue before `dep` (a value loaded after a barrier) is available: BAR.SYNC.DEFER_BLOCKING lets a plain
// clock read through, a shared-memory load blocks until the barrier completes
__device__ __forceinline__ long long clock_after(const float *smem_word) {
    const float x = *(const volatile float *)smem_word;
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(x));
    return t;
}
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// reference form of the per-symbol wrap (m17_modulate.cpp:33-37), every store rounding to fp32
__device__ __noinline__ float tx_wrap_ref(float acc) {
    acc = (float)((double)acc / (2.0 * M_PI));
    const double a = (double)acc, ip = trunc(a);
    acc = (float)copysign(a - ip, a);                    // modf: the fraction carries the sign of its argument (-0.0 for -3.0)
    return (float)((double)acc * 2.0 * M_PI);
}
// fp32 form; valid (== tx_wrap_ref, proven by enumeration) when tx_wrap_fast_ok(acc)
__device__ __forceinline__ bool tx_wrap_fast_ok(float acc) {
    const uint32_t bits = __float_as_uint(acc);
    return (((bits >> 23) & 0xFFu) - 26u) < 124u || bits == 0u;   // 2^-101 <= |acc| < 2^23, or +0 (-0.0: the quotient's sign is lost in the residual)
}
__device__ __forceinline__ float tx_wrap_fast(float acc) {
    // quotient: q0 = acc*r, residual against (C_hi + C_lo'), one correction with r' (constants tuned by enumeration, see header)
    const float r = __uint_as_float(0x3E22F983u), r2 = __uint_as_float(0x3E22F984u);
    const float chi = __uint_as_float(0x40C90FDBu), clo_q = __uint_as_float(0xB43BBD2Du), clo = __uint_as_float(0xB43BBD2Eu);
    const float q0 = acc * r;
    const float e1 = __fmaf_rn(-q0, chi, acc);
    const float e2 = __fmaf_rn(-q0, clo_q, e1);
    const float q1 = __fmaf_rn(e2, r2, q0);
    // fraction of |q1| without FRND (17 cycles on B200): (|q1| + 2^23) rounded toward zero is floor|q1| + 2^23 for |q1| < 2^23
    const float aq = fabsf(q1);
    const float fl = __fadd_rz(aq, 8388608.0f) - 8388608.0f;
    const float f = aq - fl;                             // exact, >= 0
    const float p = f * chi;
    const float ep = __fmaf_rn(f, chi, -p);              // exact error of the product
    const float t = __fmaf_rn(f, clo, ep);
    return __uint_as_float(__float_as_uint(p + t) | (__float_as_uint(q1) & 0x80000000u));   // modf: the fraction (even a zero one) carries the quotient's sign
}
__device__ __forceinline__ float tx_wrap(float acc) { return tx_wrap_fast_ok(acc) ? tx_wrap_fast(acc) : tx_wrap_ref(acc); }
__global__ void k_selftest_tx_wrap(uint64_t first, uint64_t count, unsigned long long *mism, uint32_t *dump, int dump_cap) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t bits = (uint32_t)(first + i);
    if (((bits >> 23) & 0xFFu) == 0xFFu) return;         // Inf / NaN never reach the accumulator
    const float x = __uint_as_float(bits);
    const uint32_t a = __float_as_uint(tx_wrap(x)), b = __float_as_uint(tx_wrap_ref(x));
    if (a != b) {
        const unsigned long long k = atomicAdd(mism, 1ull);
        if (k < (unsigned long long)dump_cap) { dump[3 * k] = bits; dump[3 * k + 1] = a; dump[3 * k + 2] = b; }
    }
}
extern "C" int m17b_selftest_tx_wrap(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream) {
    if (!ctx || !h_mismatches || dump_cap < 0 || (dump_cap && !h_dump)) return M17B_E_ARG;
    cudaStream_t st = as_stream(stream);
    unsigned long long *d_n; uint32_t *d_dump = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d_n, 8));
    CUDA_TRY(cudaMemsetAsync(d_n, 0, 8, st));
    if (dump_cap) { CUDA_TRY(cudaMalloc((void **)&d_dump, 12 * (size_t)dump_cap)); CUDA_TRY(cudaMemsetAsync(d_dump, 0, 12 * (size_t)dump_cap, st)); }
    for (uint64_t o = 0; o < count; o += (1ull << 30)) {
        const uint64_t n = count - o < (1ull << 30) ? count - o : (1ull << 30);
        k_selftest_tx_wrap<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(first + o, n, d_n, d_dump, dump_cap);
        KERNEL_CHECK();
    }
    unsigned long long n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, st));
    if (dump_cap) CUDA_TRY(cudaMemcpyAsync(h_dump, d_dump, 12 * (size_t)dump_cap, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *h_mismatches = n;
    cudaFree(d_n); cudaFree(d_dump);
    return M17B_OK;
}

// cos/sin for the phase accumulator (|x| < 17 by construction: one turn plus a symbol's worth of deviation): Cody-Waite reduction
// by pi/2 with the quadrant taken from a magic-number round (no F2I / I2F), degree-7 / degree-8 minimax kernels evaluated with
// FMAs.  Measured against double sin/cos over |x| <= 13 on the CPU: max abs error 6.6e-8 / 7.4e-8, and the truncated int16
// products differ from glibc's cosf / sinf in 0.009 % of the samples, always by one LSB -- the same +-1 LSB contract the CUDA
// library's sincosf gave (mod_fsk computes cos / sin through libm, m17_modulate.cpp:25-26; no two libms agree to the last bit).
__device__ __forceinline__ void tx_sincos(float x, float &sn, float &cs) {
    if (fabsf(x) > 8192.0f) { sincosf(x, &sn, &cs); return; }               // never with a wrapped accumulator
    const float t = __fmaf_rn(x, 0.636619747f, 12582912.0f);
    const uint32_t k = __float_as_uint(t);
    const float j = t - 12582912.0f;
    float r = __fmaf_rn(j, -1.57079601e+00f, x);
    r = __fmaf_rn(j, -3.13916473e-07f, r);
    r = __fmaf_rn(j, -5.39030253e-15f, r);
    const float s = r * r;
    float ps = __fmaf_rn(s, -1.95152959e-4f, 8.33216087e-3f);
    ps = __fmaf_rn(ps, s, -1.66666546e-1f);
    const float si = __fmaf_rn(ps, r * s, r);
    float pc = __fmaf_rn(s, 2.44331571e-5f, -1.38873163e-3f);
    pc = __fmaf_rn(pc, s, 4.16666418e-2f);
    pc = __fmaf_rn(pc, s, -0.5f);
    const float co = __fmaf_rn(pc, s, 1.0f);
    const float a = (k & 1u) ? co : si, b = (k & 1u) ? si : co;
    sn = __uint_as_float(__float_as_uint(a) ^ ((k & 2u) << 30));
    cs = __uint_as_float(__float_as_uint(b) ^ (((k + 1u) & 2u) << 30));
}
__device__ __forceinline__ int tx_iq_word(float ph) {                           // {re, im} = {(int16)(cos*0x3FFF), (int16)(sin*0x3FFF)}, truncating
    float sn, cs;
    tx_sincos(ph, sn, cs);
    const int re = __float2int_rz(cs * 16383.0f), im = __float2int_rz(sn * 16383.0f);
    return (re & 0xFFFF) | (im << 16);
}

struct ModGeom {
    int os, G, S;                // samples per symbol, channels per CTA (<= 32), symbols per chunk (multiple of TXM_SEG)
    int DP, WP;                  // row pitches (floats): deviation rows (30 + S), work / phase rows (S*os + pad, a multiple of 4, == 4 mod 32)
    int dbg;                     // builds with -DM17B_EXPERIMENTS only (M17B_MOD_DBG): 1 = workers skip the FIR, 2 = workers skip cos/sin + stores
    int64_t nsym;
};
__host__ __device__ inline size_t mod_smem_floats(const ModGeom &g) {
    // taps, deviation table, history, deviation rows x2, FIR rows x2, phase rows x2, a zero row (idle scan lanes) + read-ahead slack
    return (size_t)((31 * g.os + 3) & ~3) + 8 + (size_t)((g.G * 31 + 3) & ~3) + 2 * (size_t)g.G * g.DP + 4 * (size_t)g.G * g.WP + (size_t)g.WP + 32;
}

// one symbol of the phase chain at os = 10: ten ordered adds (mod_fsk, m17_modulate.cpp:23-24), then the wrap; `bad` collects
// accumulator values outside the range in which the fp32 wrap was proven
__device__ __forceinline__ void scan_symbol10(float &acc, const float *v, float *o, bool &bad) {
#pragma unroll
    for (int j = 0; j < 10; j++) { acc += v[j]; o[j] = acc; }
    bad |= !tx_wrap_fast_ok(acc);
    acc = tx_wrap_fast(acc);
}

// OS > 0: samples per symbol known at compile time (10 = the reference's 48 kHz rate); OS == 0: geom.os at run time (80 = Pluto)
template <int OS>
__global__ void __launch_bounds__(TXM_THREADS) k_mod_fused(const uint8_t *__restrict__ syms, ModGeom gm, const float *__restrict__ taps, const float *__restrict__ devtab,
                                                           TxChanState *st, int64_t nchan, int16_t *__restrict__ iq, float *__restrict__ freq,
                                                           unsigned long long *dbg_clk) {
    extern __shared__ __align__(16) float sm[];
    const int os = OS > 0 ? OS : gm.os;
    const int G = gm.G, S = gm.S, DP = gm.DP, WP = gm.WP;
    const int64_t nsym = gm.nsym;
    float *taps_s = sm;                                  // [31*os]
    float *dev_tab = taps_s + ((31 * os + 3) & ~3);      // [8]
    float *hist_s = dev_tab + 8;                         // [G][31]  deviation history as the call found it
    float *dev_s = hist_s + ((G * 31 + 3) & ~3);         // [2][G][DP]
    float *work_s = dev_s + 2 * G * DP;                  // [2][G][WP]  FIR output (frequency samples)
    float *phase_s = work_s + 2 * G * WP;                // [2][G][WP]  scan output (phases)
    float *zero_s = phase_s + 2 * G * WP;                // [WP + 32]   what the idle lanes of the scan warp read; also absorbs read-ahead
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t c0 = (int64_t)blockIdx.x * G;
    const int nchunks = (int)((nsym + S - 1) / S);

    for (int i = tid; i < 31 * os; i += TXM_THREADS) taps_s[i] = taps[i];
    if (tid < 8) dev_tab[tid] = tid < 5 ? devtab[tid] : 0.0f;
    for (int i = tid; i < WP + 32; i += TXM_THREADS) zero_s[i] = 0.0f;
    for (int i = tid; i < G * 31; i += TXM_THREADS) {
        const int64_t c = c0 + i / 31;
        hist_s[i] = c < nchan ? st[c].hist[i % 31] : 0.0f;
    }
    __syncthreads();

    if (warp == 0) {
        // ------------------------------------------------------------ scan warp: lane = channel
        // All 32 lanes walk (idle lanes read the zero row and store nothing), so the warp stays converged.  A warp alone on its
        // scheduler issues about one instruction every two cycles (ncu, profiles/r02_mod_*), so the walk is kept short: four
        // symbols per trip, 16-byte loads / stores (a symbol pair is 20 floats = 5 quads), no branch per symbol -- accumulator
        // values outside the proven range of the fp32 wrap only set a flag, and a flagged chunk (never seen; it needs
        // |acc| < 2^-101 or >= 2^23) is walked again from its saved start with the double-precision wrap: the FIR output is
        // still there because the phases go to a separate buffer.
        const int64_t c = c0 + lane;
        const bool live = lane < G && c < nchan;
        float acc = live ? st[c].acc : 0.0f;
        long long t_busy = 0, t_wait = 0;
        for (int i = 0; i < nchunks; i++) {
            const int b = i & 1;
            const int Si = (int)(nsym - (int64_t)i * S < S ? nsym - (int64_t)i * S : S);
            __syncwarp();
            const long long tw = clock64();
            bar_sync_n(2 + b, TXM_BAR);
            const long long t0 = clock_after(zero_s);
            t_wait += t0 - tw;
            const float *src = live ? work_s + ((size_t)b * G + lane) * WP : zero_s;
            float *dst = phase_s + ((size_t)b * G + (live ? lane : 0)) * WP;
            const float acc_start = acc;
            bool bad = false;
            int k = 0;
            if (OS == 10) {
                float4 va[5], vb[5], oa[5];
#pragma unroll
                for (int q = 0; q < 5; q++) va[q] = ((const float4 *)src)[q];
                for (; k + 4 <= Si; k += 4) {
                    const float4 *ps = (const float4 *)(src + k * 10);
                    float4 *pd = (float4 *)(dst + k * 10);
#pragma unroll
                    for (int q = 0; q < 5; q++) vb[q] = ps[5 + q];
                    scan_symbol10(acc, (const float *)va, (float *)oa, bad);
                    scan_symbol10(acc, (const float *)va + 10, (float *)oa + 10, bad);
                    if (live) {
#pragma unroll
                        for (int q = 0; q < 5; q++) pd[q] = oa[q];
                    }
#pragma unroll
                    for (int q = 0; q < 5; q++) va[q] = ps[10 + q];      // at most 40 floats past the chunk: pad / next row / slack
                    scan_symbol10(acc, (const float *)vb, (float *)oa, bad);
                    scan_symbol10(acc, (const float *)vb + 10, (float *)oa + 10, bad);
                    if (live) {
#pragma unroll
                        for (int q = 0; q < 5; q++) pd[5 + q] = oa[q];
                    }
                }
            }
            for (; k < Si; k++) {                       // generic os, and the last symbols of a ragged chunk
                const float *ps = src + k * os;
                float *pd = dst + k * os;
                float nx = ps[0];
                for (int j = 0; j < os; j++) {
                    const float x = nx;
                    nx = ps[j + 1];
                    acc += x;
                    if (live) pd[j] = acc;
                }
                bad |= !tx_wrap_fast_ok(acc);
                acc = tx_wrap_fast(acc);
            }
            if (__any_sync(0xFFFFFFFFu, bad)) {         // the exact, slow walk of the same chunk
                acc = acc_start;
                for (k = 0; k < Si; k++) {
                    for (int j = 0; j < os; j++) { acc += src[k * os + j]; if (live) dst[k * os + j] = acc; }
                    acc = tx_wrap_ref(acc);
                }
            }
            __threadfence_block();
            __syncwarp();
            bar_arrive_n(4 + b, TXM_BAR);
            t_busy += clock64() - t0;
        }
        if (live) st[c].acc = acc;
        if (dbg_clk && lane == 0) {               // instrumentation: cycles the scan warp spent walking its rows (barrier waits may leak in)
            atomicAdd(dbg_clk, (unsigned long long)t_wait); atomicAdd(dbg_clk + 1, (unsigned long long)t_busy);
            atomicAdd(dbg_clk + 2, (unsigned long long)nchunks); atomicAdd(dbg_clk + 3, 1ull);
        }
        return;
    }

    // ---------------------------------------------------------------- worker warps
    const int widx = txm_worker_index(warp);
    if (widx < 0) return;
    const int wt = widx * 32 + lane;
    const bool vec_ok = (((nsym * os) & 3) == 0) && ((((uintptr_t)iq) & 15) == 0);
    const int row_quads = (S * os) >> 2;              // sample quads per full row (S is a multiple of 16)
    auto emit = [&](int i) {                          // cos/sin of the phases of chunk i -> int16 IQ (m17_modulate.cpp:25-26, truncating)
        const int b = i & 1;
        const int64_t k0 = (int64_t)i * S;
        const int Si = (int)(nsym - k0 < S ? nsym - k0 : S);
        const int nsamp = Si * os;
        for (int it = wt; it < G * row_quads; it += TXM_NWT) {
            const int g = it / row_quads, n = 4 * (it - g * row_quads);
            const int64_t c = c0 + g;
            if (c >= nchan || n >= nsamp) continue;
            const float4 ph = *(const float4 *)(phase_s + ((size_t)b * G + g) * WP + n);
            int *o = (int *)iq + (c * nsym * os + k0 * os + n);
            int4 w;
            w.x = tx_iq_word(ph.x); w.y = tx_iq_word(ph.y); w.z = tx_iq_word(ph.z); w.w = tx_iq_word(ph.w);
            if (vec_ok && n + 3 < nsamp) *(int4 *)o = w;
            else {
                o[0] = w.x;
                if (n + 1 < nsamp) o[1] = w.y;
                if (n + 2 < nsamp) o[2] = w.z;
                if (n + 3 < nsamp) o[3] = w.w;
            }
        }
    };
    // symbols of a chunk are fetched one chunk ahead, four per thread (one 32-bit load when the rows allow it)
    const int sw = S >> 2;                            // 4-symbol groups per row
    const bool sym_vec = ((nsym & 3) == 0) && ((((uintptr_t)syms) & 3) == 0);
    auto fetch = [&](int i, int slot) -> uint32_t {   // group `slot` (< G*sw) of chunk i as 4 packed bytes (0 beyond the data)
        const int g = slot / sw, q = slot - g * sw;
        const int64_t c = c0 + g, kk = (int64_t)i * S + 4 * q;
        if (c >= nchan || kk >= nsym) return 0x04040404u;
        const uint8_t *p = syms + c * nsym + kk;
        if (sym_vec) return *(const uint32_t *)p;     // kk + 3 < nsym because nsym is a multiple of 4
        uint32_t w = 0x04040404u;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (kk + e < nsym) w = (w & ~(0xFFu << (8 * e))) | ((uint32_t)p[e] << (8 * e));
        return w;
    };
    constexpr int MAXF = 2;                           // G*S/4 <= 2*TXM_NWT groups per chunk (G*S <= 2560)
    uint32_t pre[MAXF];
#pragma unroll
    for (int u = 0; u < MAXF; u++) pre[u] = (wt + u * TXM_NWT < G * sw) ? fetch(0, wt + u * TXM_NWT) : 0u;
    const int segs = S / TXM_SEG;
    const int ntask = G * os * segs;
    long long w_fill = 0, w_fir = 0, w_wait = 0, w_emit = 0;      // instrumentation (worker warp 0)
    for (int i = 0; i < nchunks; i++) {
        const int b = i & 1;
        const int64_t k0 = (int64_t)i * S;
        const int Si = (int)(nsym - k0 < S ? nsym - k0 : S);
        const long long ta = clock64();
        // deviations of symbols k0-30 .. k0+S-1 of every channel of the CTA (m_tx_lu, m17_modulate.cpp:9; 4 = blank carrier):
        // the 30 of history from the previous chunk's row (the caller's history for the first chunk), the new ones from `pre`
        float *dv = dev_s + (size_t)b * G * DP;
        const float *dvp = dev_s + (size_t)(b ^ 1) * G * DP;
        for (int idx = wt; idx < G * 30; idx += TXM_NWT) {
            const int g = idx / 30, j = idx - g * 30;
            dv[g * DP + j] = i == 0 ? hist_s[g * 31 + 1 + j] : dvp[g * DP + S + j];
        }
#pragma unroll
        for (int u = 0; u < MAXF; u++) {
            const int slot = wt + u * TXM_NWT;
            if (slot < G * sw) {
                const int g = slot / sw, q = slot - g * sw;
                float *d4 = dv + g * DP + 30 + 4 * q;
#pragma unroll
                for (int e = 0; e < 4; e++) d4[e] = dev_tab[(pre[u] >> (8 * e)) & 7u];
                if (i + 1 < nchunks) pre[u] = fetch(i + 1, slot);
            }
        }
        __syncwarp();
        bar_sync_n(1, TXM_NWT);
        const long long tb = clock_after(zero_s);
        if (gm.dbg & 1) {                          // experiment: no FIR, a benign constant in the work rows
            float *w = work_s + (size_t)b * G * WP;
            for (int idx = wt; idx < G * WP; idx += TXM_NWT) w[idx] = 0.0123f;
        }
        // polyphase RRC (sub_filter, m17_modulate.cpp:42-48): sum = s[0]*c[0]; sum += s[j]*c[j*os], c = taps + (os-1-ph)
        for (int task = wt; task < ntask && !(gm.dbg & 1); task += TXM_NWT) {
            const int ph = task % os, gq = task / os;
            const int seg = gq / G, g = gq - seg * G;
            const int ks = seg * TXM_SEG;
            if (ks >= Si) continue;
            float tp[31], d[30 + TXM_SEG], sum[TXM_SEG];
            const float *tsrc = taps_s + (os - 1 - ph);
#pragma unroll
            for (int j = 0; j < 31; j++) tp[j] = tsrc[j * os];
            const float *dsrc = dv + g * DP + ks;
#pragma unroll
            for (int j = 0; j < 30 + TXM_SEG; j++) d[j] = dsrc[j];
#pragma unroll
            for (int j = 0; j < 31; j++) {
#pragma unroll
                for (int o = 0; o < TXM_SEG; o++) {
                    const float prod = d[o + j] * tp[j];
                    sum[o] = j == 0 ? prod : sum[o] + prod;
                }
            }
            float *w = work_s + ((size_t)b * G + g) * WP + ks * os + ph;
            const int64_t c = c0 + g;
            float *fq = (freq && c < nchan) ? freq + c * nsym * os + (k0 + ks) * os + ph : nullptr;
#pragma unroll
            for (int o = 0; o < TXM_SEG; o++) {
                if (ks + o < Si) {
                    w[o * os] = sum[o];
                    if (fq) fq[(int64_t)o * os] = sum[o];
                }
            }
        }
        __threadfence_block();
        __syncwarp();
        bar_arrive_n(2 + b, TXM_BAR);
        const long long tc = clock64();
        w_fill += tb - ta; w_fir += tc - tb;
        if (i > 0) {
            bar_sync_n(4 + (b ^ 1), TXM_BAR);
            const long long td = clock_after(zero_s);
            if (!(gm.dbg & 2)) emit(i - 1);
            const long long te = clock64();
            w_wait += td - tc; w_emit += te - td;
        }
    }
    if (dbg_clk && wt == 0) {
        atomicAdd(dbg_clk + 4, (unsigned long long)w_fill); atomicAdd(dbg_clk + 5, (unsigned long long)w_fir);
        atomicAdd(dbg_clk + 6, (unsigned long long)w_wait); atomicAdd(dbg_clk + 7, (unsigned long long)w_emit);
    }
    __syncwarp();
    bar_sync_n(4 + ((nchunks - 1) & 1), TXM_BAR);
    if (!(gm.dbg & 2)) emit(nchunks - 1);
    // m_tx_s after the call: deviations of the last 31 symbols (older ones from the history the call found)
    for (int idx = wt; idx < G * 31; idx += TXM_NWT) {
        const int g = idx / 31, j = idx - g * 31;
        const int64_t c = c0 + g, kk = nsym - 31 + j;
        if (c >= nchan) continue;
        st[c].hist[j] = kk >= 0 ? dev_tab[syms[c * nsym + kk] & 7] : hist_s[g * 31 + 31 + (int)kk];
    }
}

static inline ModGeom mod_geometry(int64_t nchan, int os, int64_t nsym) {
    ModGeom g;
    g.os = os; g.nsym = nsym;
    g.dbg = 0;
#ifdef M17B_EXPERIMENTS
    if (const char *e = getenv("M17B_MOD_DBG")) g.dbg = atoi(e);
#endif
    // channels per CTA: two CTAs per SM when the batch allows it (the workers are throughput-bound and a second CTA's warps fill
    // the issue slots the first leaves; the scan is one lane per channel, so up to 32)
    int G = (int)((nchan + 2 * 148 - 1) / (2 * 148));
    if (G < 4) G = 4;
    if (G > 32) G = 32;
#ifdef M17B_EXPERIMENTS
    if (const char *eg = getenv("M17B_MOD_G")) G = atoi(eg);
#endif
    // symbols per chunk: G*S*os <= 5120 samples per buffer (512 channel-symbols at os = 10), S a multiple of 16
    int S = (5120 / (G * os)) / TXM_SEG * TXM_SEG;
    while (S < TXM_SEG && G > 1) { G--; S = (5120 / (G * os)) / TXM_SEG * TXM_SEG; }
    if (S < TXM_SEG) S = TXM_SEG;
    if (S > 128) S = 128;
    while (G * S > 2048 && S > TXM_SEG) S -= TXM_SEG;       // the symbol prefetch holds 2 x 4 symbols per worker thread
    g.G = G; g.S = S;
    g.DP = 30 + S;
    g.WP = S * os + ((4 - (S * os) % 32) + 32) % 32;      // == 4 (mod 32): 16-byte row starts, lanes 4 banks apart
    if (g.WP < S * os + 4) g.WP += 32;
    return g;
}
