// dec.cuh -- the Pluto receive path's front-end decimator, batched: int16 IQ at 384 kS/s -> 31-tap symmetric low-pass
// (int16 taps, int32 accumulate, >> 15) -> every 8th sample -> int16 IQ at 48 kS/s, the input of m17_dsp_rx.
// Replaces sub_filter / rx_decimate_filter / build_pluto_rx_dec_filter and the chunk loop of radio_receive_samples
// (radio.cpp:18-51,157-177); filter design m17_dsp_build_lpf_filter / m17_dsp_float_to_short (m17_dsp.cpp:347-360,382-386).
// the step immediately before the hot path.
//
// Streaming form: output k of a channel is sum_j taps[j] * x[8k - 31 + j], j = 0..30 (the reference carries the last 31
// samples of each 1920-sample chunk at the head of its buffer, radio.cpp:167).  int32 products and sums wrap modulo 2^32
// exactly as the reference's int32_t arithmetic does, so the folded form coffs[i]*(in[i] + in[30-i]) and the plain 31-term
// sum used here are the same number.
//
// HBM-bound: 32 B in + 4 B out per output sample (69 120 B per channel-frame), ~90 integer instructions.  Mapping: one CTA
// per (channel, tile of 512 outputs); the tile's 4128 input samples are staged with 16-byte cp.async into shared memory
// (padded by 4 words every 32 so that the 16-byte column reads below are conflict-free); each thread produces four
// consecutive outputs from 56 staged samples (14 LDS.128) and stores them as one 16-byte word.
#pragma once
#include "framer.cuh"

#define DEC_NTAP 31
#define DEC_TILE 512                       // outputs per CTA
#define DEC_THREADS (DEC_TILE / 4)
#define DEC_STAGE (DEC_TILE * 8 + 32)      // staged samples: 32 of history (the first is not used) + 8 per output
#define DEC_PAD(s) ((s) + 4 * ((s) >> 5))

struct DecChanState { uint32_t hist[2][32]; };   // last 32 input samples of the previous call, double buffered by call parity

__constant__ int16_t c_dec_taps[32];

struct m17b_dec {
    m17b_ctx *ctx;
    int64_t nchan;
    DecChanState *d_state;
    int phase;
    int16_t h_taps[DEC_NTAP];
};

__global__ void __launch_bounds__(DEC_THREADS) k_decimate8(const uint32_t *__restrict__ in, int64_t nout, int64_t tiles, DecChanState *st, int phase,
                                                          uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t xs[DEC_PAD(DEC_STAGE) + 4];
    const int tid = threadIdx.x;
    const int64_t c = blockIdx.x / tiles;
    const int64_t o0 = (int64_t)(blockIdx.x % tiles) * DEC_TILE;    // first output of the tile
    const int64_t nin = nout * 8;
    const uint32_t *x = in + c * nin;
    const int64_t n0 = o0 * 8 - 32;                                 // input index of staged sample 0
    // ---- stage (16-byte chunks; history of the previous call for n < 0, zeros past the end)
    for (int q = tid; q < DEC_STAGE / 4; q += DEC_THREADS) {
        const int s = 4 * q;
        const int64_t n = n0 + s;
        uint32_t *dst = &xs[DEC_PAD(s)];
        if (n >= 0 && n + 3 < nin) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(x + n));
        } else if (n < 0) {
            *(uint4 *)dst = *(const uint4 *)&st[c].hist[phase][n + 32];
        } else {
            *(uint4 *)dst = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    // ---- the last tile leaves the channel's final 32 input samples for the next call (other buffer: tile 0 may still be reading)
    if (o0 + DEC_TILE >= nout && tid < 8) {
        const int64_t n = nin - 32 + 4 * tid;                       // nin >= 32 (at least 4 outputs per call)
        *(uint4 *)&st[c].hist[phase ^ 1][4 * tid] = *(const uint4 *)(x + n);
    }
    // ---- four outputs per thread: outputs o0 + 4 tid + j use staged samples 32 tid + 8 j + 1 .. + 31
    uint32_t w[56];
#pragma unroll
    for (int q = 0; q < 14; q++) {
        const uint4 v = *(const uint4 *)&xs[DEC_PAD(32 * tid + 4 * q)];
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    uint32_t res[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int re = 0, im = 0;
#pragma unroll
        for (int i = 0; i < DEC_NTAP; i++) {
            const uint32_t v = w[1 + 8 * j + i];
            const int t = c_dec_taps[i];
            re += t * (int)(short)(v & 0xFFFFu);                    // sub_filter, radio.cpp:18-33 (int32 wrap-around arithmetic)
            im += t * ((int)v >> 16);
        }
        res[j] = ((uint32_t)(re >> 15) & 0xFFFFu) | ((uint32_t)(im >> 15) << 16);   // sum.re = real >> 15 narrowed to int16
    }
    const int64_t o = o0 + 4 * tid;
    if (o + 3 < nout) *(uint4 *)(out + c * nout + o) = make_uint4(res[0], res[1], res[2], res[3]);
    else for (int j = 0; j < 4; j++) if (o + j < nout) out[c * nout + o + j] = res[j];
}

// m17_dsp_build_lpf_filter (m17_dsp.cpp:347-360): rectangular-window sinc, double math; the first tap time is the INTEGER
// division -(ntaps-1)/2 converted to double
extern "C" int m17b_build_lpf_filter(float *h_taps, float bw, int ntaps) {
    if (!h_taps || ntaps <= 0) return M17B_E_ARG;
    const double B = bw;
    double t = -(ntaps - 1) / 2;
    for (int i = 0; i < ntaps; i++) {
        const double a = (t == 0) ? 2.0 * B : 2.0 * B * sin(M_PI * t * B) / (M_PI * t * B);
        h_taps[i] = (float)a;
        t = t + 1.0;
    }
    return M17B_OK;
}
// m17_dsp_float_to_short (m17_dsp.cpp:382-386)
extern "C" int m17b_float_to_short(const float *h_in, int16_t *h_out, int len) {
    if (!h_in || !h_out || len < 0) return M17B_E_ARG;
    for (int i = 0; i < len; i++) h_out[i] = (int16_t)(h_in[i] * 0x7FFF);
    return M17B_OK;
}

extern "C" int m17b_dec_destroy(m17b_dec *d) {
    if (!d) return M17B_E_ARG;
    cudaFree(d->d_state);
    free(d);
    return M17B_OK;
}
extern "C" int m17b_dec_reset(m17b_dec *d, void *stream) {
    if (!d) return M17B_E_ARG;
    CUDA_TRY(cudaMemsetAsync(d->d_state, 0, sizeof(DecChanState) * d->nchan, as_stream(stream)));   // m_rx_buff is a zeroed static
    d->phase = 0;
    return M17B_OK;
}
// build_pluto_rx_dec_filter (radio.cpp:44-51): low-pass at 0.125 fs, gain 0.9, int16 taps
extern "C" int m17b_dec_create(m17b_ctx *ctx, int64_t nchan, m17b_dec **out) {
    if (!ctx || !out || nchan <= 0) return M17B_E_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    m17b_dec *d = (m17b_dec *)calloc(1, sizeof(m17b_dec));
    if (!d) return M17B_E_NOMEM;
    d->ctx = ctx; d->nchan = nchan;
    float f[DEC_NTAP];
    m17b_build_lpf_filter(f, 0.125f, DEC_NTAP);
    m17b_set_filter_gain(f, (float)0.9, 1, DEC_NTAP);
    m17b_float_to_short(f, d->h_taps, DEC_NTAP);
    int16_t t32[32] = {0};
    memcpy(t32, d->h_taps, sizeof(d->h_taps));
    CUDA_TRY(cudaMemcpyToSymbol(c_dec_taps, t32, sizeof(t32)));
    if (cudaMalloc((void **)&d->d_state, sizeof(DecChanState) * nchan) != cudaSuccess) { free(d); return M17B_E_NOMEM; }
    int rc = m17b_dec_reset(d, nullptr);
    if (rc) { m17b_dec_destroy(d); return rc; }
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    *out = d;
    return M17B_OK;
}
extern "C" int m17b_dec_get_taps(const m17b_dec *d, int16_t *h_taps31) {
    if (!d || !h_taps31) return M17B_E_ARG;
    memcpy(h_taps31, d->h_taps, sizeof(d->h_taps));
    return M17B_OK;
}
// radio_receive_samples, Pluto branch (radio.cpp:157-177), for nchan channels: d_in int16 [nchan][8*nout][2] at 384 kS/s
// -> d_out int16 [nchan][nout][2] at 48 kS/s; nout a multiple of 4 (the reference works in chunks of 240)
extern "C" int m17b_dec_run(m17b_dec *d, const int16_t *d_in, int64_t nout, int16_t *d_out, void *stream) {
    if (!d || !d_in || !d_out || nout < 4 || (nout & 3)) return M17B_E_ARG;
    const int64_t tiles = (nout + DEC_TILE - 1) / DEC_TILE;
    if (tiles * d->nchan > 0x7fffffffLL) return M17B_E_ARG;
    k_decimate8<<<(unsigned)(tiles * d->nchan), DEC_THREADS, 0, as_stream(stream)>>>((const uint32_t *)d_in, nout, tiles, d->d_state, d->phase, (uint32_t *)d_out);
    KERNEL_CHECK();
    d->phase ^= 1;
    return M17B_OK;
}
