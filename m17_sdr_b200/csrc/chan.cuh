// chan.cuh -- wideband channeliser ahead of the path (second half): one int16 IQ capture at
// 1.2 MS/s (= 25 x 48 kS/s) -> 96 channels spaced 12.5 kHz, each at 48 kS/s int16 IQ, laid out as m17b_dsp_rx wants them.
// It generalises the reference's Pluto receive decimator (sub_filter / rx_decimate_filter, radio.cpp:18-40: int16 taps, int32
// accumulate, >> 15) from one channel to M: the same integer FIR folded into M = 96 polyphase branches, a 96-point DFT in
// fixed point (prime-factor split 3 x 32, radix-2 decimation in time, Q30 twiddles, products (int64 a*w) >> 30, so the twiddles
// 1 and -j act exactly), the phase of the sliding window, the reference's >> 15.  Integer arithmetic throughout, so the output is a
// well-defined function of the input: oracle/m17_oracle.c (m17o_chan_run) states it in plain C -- pinned at M = 1, D = 8 against
// radio.cpp itself -- and the kernel below must equal it bit for bit (tests).
//
// Why it matters: per-channel 48 kS/s IQ costs 192 kB per channel-second on the host link; a 12.5 kHz raster out of a
// wideband capture costs 50 kB, so the end-to-end rate from host memory is no longer bound at ~280 k channel-s/s by PCIe.
//
// Mapping.  CTA = (capture, tile of 64 output times); the tile's 63*25 + L input samples are staged in shared memory once.
// A warp takes one output time at a time: lane l owns the DFT inputs n2 = bitrev5(l) of the three interleaved 32-point
// sub-sequences (p = (32 a + 3 n2) mod 96, a = 0..2): it folds its 3 x P taps (taps in registers), does the 3-point DFT in
// registers, then the five radix-2 stages across the warp (one shuffle exchange per stage: the lower lane of a butterfly sends
// its twiddled value, the upper one its own), rotates by the window phase and drops three bins into a [96][65] tile; the CTA
// writes the tile out as 96 rows of 256 bytes.  HBM: 100 B in, 384 B out per output time; the kernel is integer-ALU bound.
#pragma once
#include "dec.cuh"

#define CH_M 96
#define CH_D 25
#define CH_TT 64                   // output times per CTA
#define CH_PMAX 16                 // taps per polyphase branch, at most
#define CH_THREADS 128

struct m17b_chan {
    m17b_ctx *ctx;
    int64_t ncap;
    int P, L;                      // taps per branch, L = 96 P
    int16_t *d_taps;               // [L]
    int32_t *d_tab;                // [16][2] twiddles of the 32-point DFT, [96][2] window-phase table, [1] sqrt(3)/2
    uint32_t *d_hist;              // [ncap][L] last L input samples of the previous call
    uint32_t *d_hist2;             // scratch for the update
    long long n_done;              // outputs produced so far (per capture): fixes the window phase
    int16_t h_taps[CH_M * CH_PMAX];
};

__device__ __forceinline__ int32_t chq_mul(int32_t a, int32_t w) { return (int32_t)(((long long)a * (long long)w) >> 30); }   // Q30 twiddles: +-1.0 exact
__device__ __forceinline__ void chq_cmul(int32_t ar, int32_t ai, int32_t wr, int32_t wi, int32_t &or_, int32_t &oi) {
    or_ = chq_mul(ar, wr) - chq_mul(ai, wi);
    oi = chq_mul(ar, wi) + chq_mul(ai, wr);
}

template <int P>
__global__ void __launch_bounds__(CH_THREADS) k_chan96(const uint32_t *__restrict__ in, int64_t nin, const uint32_t *__restrict__ hist, const int16_t *__restrict__ taps,
                                                        const int32_t *__restrict__ tab, long long n_done, int64_t nout, uint32_t *__restrict__ out, int64_t out_pitch) {
    constexpr int L = CH_M * P, NW = (CH_TT - 1) * CH_D + L;
    __shared__ uint32_t xs[NW];
    __shared__ uint32_t outs[CH_M][CH_TT + 1];
    __shared__ int32_t tw_s[16][2], rot_s[CH_M][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t cap = blockIdx.y, n0 = (int64_t)blockIdx.x * CH_TT;
    const uint32_t *xin = in + cap * nin;
    // window of output n: x[n D - L .. n D - 1]; the tile needs samples [n0 D - L, n0 D - L + NW)
    for (int i = tid; i < NW; i += CH_THREADS) {
        const int64_t s = n0 * CH_D - L + i;
        uint32_t v = 0;
        if (s < 0) v = hist[cap * L + (L + s)];
        else if (s < nin) v = __ldg(xin + s);
        xs[i] = v;
    }
    if (tid < 16) { tw_s[tid][0] = tab[2 * tid]; tw_s[tid][1] = tab[2 * tid + 1]; }
    for (int i = tid; i < CH_M; i += CH_THREADS) { rot_s[i][0] = tab[32 + 2 * i]; rot_s[i][1] = tab[32 + 2 * i + 1]; }
    const int32_t s3 = tab[32 + 2 * CH_M];
    // the lane's three polyphase branches and their taps
    const int n2 = (int)(__brev((unsigned)lane) >> 27);
    int pa[3];
    int32_t h[3][P];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        pa[a] = (32 * a + 3 * n2) % CH_M;
#pragma unroll
        for (int q = 0; q < P; q++) h[a][q] = taps[pa[a] + CH_M * q];
    }
    __syncthreads();
    for (int t = warp; t < CH_TT; t += CH_THREADS / 32) {
        if (n0 + t >= nout) break;
        const uint32_t *w = xs + t * CH_D;
        // ---- fold: z[p] = sum_q h[p + 96 q] x[p + 96 q]   (int32, wraps like the reference's accumulator)
        int32_t zr[3], zi[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int32_t sr = 0, si = 0;
#pragma unroll
            for (int q = 0; q < P; q++) {
                const uint32_t v = w[pa[a] + CH_M * q];
                sr += h[a][q] * (int32_t)(int16_t)(v & 0xFFFFu);
                si += h[a][q] * ((int32_t)v >> 16);
            }
            zr[a] = sr; zi[a] = si;
        }
        // ---- 3-point DFTs (prime-factor split: no twiddles towards the 32-point part)
        int32_t br[3], bi[3];
        {
            const int32_t t1r = zr[1] + zr[2], t1i = zi[1] + zi[2], t2r = zr[1] - zr[2], t2i = zi[1] - zi[2];
            const int32_t m1r = zr[0] - (t1r >> 1), m1i = zi[0] - (t1i >> 1);
            const int32_t m2r = chq_mul(t2r, s3), m2i = chq_mul(t2i, s3);
            br[0] = zr[0] + t1r; bi[0] = zi[0] + t1i;
            br[1] = m1r + m2i;   bi[1] = m1i - m2r;          // m1 - j m2
            br[2] = m1r - m2i;   bi[2] = m1i + m2r;          // m1 + j m2
        }
        // ---- 32-point DFT across the warp, radix-2 decimation in time (inputs are in bit-reversed lane order)
#pragma unroll
        for (int half = 1; half < 32; half <<= 1) {
            const bool lower = (lane & half) != 0;
            const int ti = (lane & (half - 1)) * (16 / half);               // twiddle index j * (32 / m), m = 2 half
            const int32_t wr = tw_s[ti][0], wi = tw_s[ti][1];
            const bool mj = ti == 8;                                        // W = -j (the only other twiddle of stage 2)
#pragma unroll
            for (int a = 0; a < 3; a++) {
                int32_t xr, xi;
                if (half == 1) { xr = br[a]; xi = bi[a]; }                                    // W = 1
                else if (half == 2) { xr = mj ? bi[a] : br[a]; xi = mj ? -br[a] : bi[a]; }    // W = 1 or -j: what the Q30 product gives, without the multiply
                else chq_cmul(br[a], bi[a], wr, wi, xr, xi);
                const int32_t sr = lower ? xr : br[a], si = lower ? xi : bi[a];
                const int32_t rr = __shfl_xor_sync(0xffffffffu, sr, half), ri = __shfl_xor_sync(0xffffffffu, si, half);
                br[a] = lower ? rr - xr : br[a] + rr;
                bi[a] = lower ? ri - xi : bi[a] + ri;
            }
        }
        // ---- window phase, scaling, bin k = (64 k1 + 33 k2) mod 96
        long long i0 = ((n_done + n0 + t) * CH_D - L) % CH_M;
        if (i0 < 0) i0 += CH_M;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int k = (64 * a + 33 * lane) % CH_M;
            const int r = (k * (int)i0) % CH_M;
            int32_t yr, yi;
            chq_cmul(br[a], bi[a], rot_s[r][0], rot_s[r][1], yr, yi);                  // r = 0: the identity, exactly (Q30)
            outs[k][t] = ((uint32_t)(yr >> 15) & 0xFFFFu) | ((uint32_t)(yi >> 15) << 16);
        }
    }
    __syncthreads();
    const int nt = (int)(nout - n0 < CH_TT ? nout - n0 : CH_TT);
    for (int idx = tid; idx < CH_M * CH_TT; idx += CH_THREADS) {
        const int k = idx / CH_TT, t = idx - k * CH_TT;
        if (t < nt) out[(cap * CH_M + k) * out_pitch + n0 + t] = outs[k][t];
    }
}
// the last L input samples become the history of the next call (older history fills in when the call was shorter than L)
__global__ void k_chan_hist(const uint32_t *__restrict__ in, int64_t nin, const uint32_t *__restrict__ hist, uint32_t *__restrict__ hist_new, int L, int64_t ncap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncap * L) return;
    const int64_t cap = i / L, j = i % L, s = nin - L + j;
    hist_new[i] = s >= 0 ? in[cap * nin + s] : hist[cap * L + (L + s)];
}

static double chan_bessel_i0(double x) { double s = 1, t = 1; for (int k = 1; k < 60; k++) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; } return s; }
extern "C" int m17b_chan_destroy(m17b_chan *c) {
    if (!c) return M17B_E_ARG;
    cudaFree(c->d_taps); cudaFree(c->d_tab); cudaFree(c->d_hist); cudaFree(c->d_hist2);
    free(c);
    return M17B_OK;
}
extern "C" int m17b_chan_reset(m17b_chan *c, void *stream) {
    if (!c) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    CUDA_TRY(cudaMemsetAsync(c->d_hist, 0, sizeof(uint32_t) * c->ncap * c->L, as_stream(stream)));
    c->n_done = 0;
    return M17B_OK;
}
static int32_t chan_q31(double v) { double s = v * 1073741824.0; s = s < 0 ? s - 0.5 : s + 0.5; return (int32_t)s; }      // Q30 (the name is historical)
extern "C" int m17b_chan_create(m17b_ctx *ctx, int64_t ncap, int taps_per_branch, m17b_chan **out) {
    if (!ctx || !out || ncap <= 0 || (taps_per_branch != 4 && taps_per_branch != 8 && taps_per_branch != 12 && taps_per_branch != 16)) return M17B_E_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    m17b_chan *c = (m17b_chan *)calloc(1, sizeof(m17b_chan));
    if (!c) return M17B_E_NOMEM;
    c->ctx = ctx; c->ncap = ncap; c->P = taps_per_branch; c->L = CH_M * taps_per_branch;
    // prototype: Kaiser-windowed sinc, cut-off 6.25 kHz (half the raster) at 1.2 MS/s, DC gain 0.9 in Q15 like
    // build_pluto_rx_dec_filter (radio.cpp:44-51); all in double, one rounding to int16 per tap
    {
        const int L = c->L;
        const double fc = 6250.0 / 1200000.0, beta = 7.0;
        double *d = (double *)malloc(sizeof(double) * L), sum = 0;
        for (int i = 0; i < L; i++) {
            const double t = i - (L - 1) / 2.0, x = 2.0 * fc * t;
            const double sinc = fabs(x) < 1e-12 ? 1.0 : sin(M_PI * x) / (M_PI * x);
            const double r = 2.0 * i / (L - 1) - 1.0;
            d[i] = sinc * chan_bessel_i0(beta * sqrt(1.0 - r * r)) / chan_bessel_i0(beta);
            sum += d[i];
        }
        for (int i = 0; i < L; i++) c->h_taps[i] = (int16_t)lrint(d[i] / sum * 0.9 * 32767.0);
        free(d);
    }
    int32_t tab[32 + 2 * CH_M + 1];
    for (int j = 0; j < 16; j++) { tab[2 * j] = chan_q31(cos(2.0 * M_PI * j / 32)); tab[2 * j + 1] = chan_q31(-sin(2.0 * M_PI * j / 32)); }
    for (int r = 0; r < CH_M; r++) { tab[32 + 2 * r] = chan_q31(cos(2.0 * M_PI * r / CH_M)); tab[32 + 2 * r + 1] = chan_q31(-sin(2.0 * M_PI * r / CH_M)); }
    tab[32 + 2 * CH_M] = chan_q31(sqrt(3.0) / 2.0);
    int rc = upload(&c->d_taps, c->h_taps, (size_t)c->L);
    if (!rc) rc = upload(&c->d_tab, tab, sizeof(tab) / sizeof(tab[0]));
    if (!rc && cudaMalloc((void **)&c->d_hist, sizeof(uint32_t) * ncap * c->L) != cudaSuccess) rc = M17B_E_NOMEM;
    if (!rc && cudaMalloc((void **)&c->d_hist2, sizeof(uint32_t) * ncap * c->L) != cudaSuccess) rc = M17B_E_NOMEM;
    if (!rc) rc = m17b_chan_reset(c, nullptr);
    if (rc) { m17b_chan_destroy(c); return rc; }
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    *out = c;
    return M17B_OK;
}
extern "C" int m17b_chan_get_taps(const m17b_chan *c, int16_t *h_taps, int *len) {
    if (!c || !h_taps || !len) return M17B_E_ARG;
    memcpy(h_taps, c->h_taps, sizeof(int16_t) * c->L);
    *len = c->L;
    return M17B_OK;
}
// d_in int16 [ncap][25 * nout][2] at 1.2 MS/s -> d_out int16 [ncap * 96][out_pitch][2] at 48 kS/s (row ncap_index * 96 + k = the
// channel centred k * 12.5 kHz above the capture's centre, k >= 48 below it); the filter history and the window phase carry
// over from call to call
extern "C" int m17b_chan_run(m17b_chan *c, const int16_t *d_in, int64_t nout, int16_t *d_out, int64_t out_pitch, void *stream) {
    if (!c || !d_in || !d_out || nout <= 0 || out_pitch < nout) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(c->ctx->device));
    cudaStream_t st = as_stream(stream);
    const dim3 grid((unsigned)((nout + CH_TT - 1) / CH_TT), (unsigned)c->ncap);
    const int64_t nin = nout * CH_D;
#define CHAN_LAUNCH(PP) k_chan96<PP><<<grid, CH_THREADS, 0, st>>>((const uint32_t *)d_in, nin, c->d_hist, c->d_taps, c->d_tab, c->n_done, nout, (uint32_t *)d_out, out_pitch)
    if (c->P == 4) CHAN_LAUNCH(4); else if (c->P == 8) CHAN_LAUNCH(8); else if (c->P == 12) CHAN_LAUNCH(12); else CHAN_LAUNCH(16);
#undef CHAN_LAUNCH
    KERNEL_CHECK();
    k_chan_hist<<<grid_for(c->ncap * c->L, 256), 256, 0, st>>>((const uint32_t *)d_in, nin, c->d_hist, c->d_hist2, c->L, c->ncap);
    KERNEL_CHECK();
    uint32_t *t = c->d_hist; c->d_hist = c->d_hist2; c->d_hist2 = t;
    c->n_done += nout;
    return M17B_OK;
}
