// frontend.cuh -- 48 kHz RX front end, batched over (channel, 40 ms block) pairs:
//   int16 IQ -> float (x 0.00003 in double) -> hard limiter -> 2-sample-baseline cross-product FM
//   discriminator -> keep every 5th -> block mean (sequential fp32 sum of all 1920 values).
// Replaces dsp_short_to_float, dsp_limit, dsp_arctan_disc2 as chained by m17_dsp_rx
// (m17_dsp.cpp:136-141,412-419,194-222,461-472).
//
// With AFC off (the reference default, radio.cpp:146-155) nothing but two input samples and the /5 phase
// crosses a block boundary, so every (channel, block) item is independent.  The only serial work inside an
// item is the fp32 running sum, whose order must be kept for bit-exactness, hence:
//   mapping: one LANE per item, one warp per 32 items.  The warp loads a [32 items][32 samples] tile with
//   32 coalesced 128-byte row requests, stages it in shared memory (pitch 33), and each lane then walks its
//   own row sequentially.  Kept outputs collect in a second [32][33] tile that is flushed with coalesced row
//   stores every 160 samples (160 = 32 outputs x 5).  HBM traffic per item: 7680 B in, 1536 + 4 B out.
// The raw (not mean-removed) discriminator samples and the block mean are written; the consumer applies
// out[i] - mean (one fp32 subtract, identical to m17_dsp.cpp:217-219).
#pragma once
#include "decode.cuh"

#define FE_WARPS 4

struct LimSample { float re, im; };
__device__ __forceinline__ LimSample fe_limit(uint32_t raw) {
    // dsp_short_to_float: int16 * 0.00003 evaluated in double, rounded once to float (m17_dsp.cpp:138-139)
    const int re_i = (int)(int16_t)(raw & 0xFFFFu), im_i = (int)(int16_t)(raw >> 16);
    float re = __double2float_rn((double)re_i * 0.00003);
    float im = __double2float_rn((double)im_i * 0.00003);
    // dsp_limit: m = sqrtf(re*re + im*im); g = (float)(1.0 / m) == correctly rounded fp32 reciprocal (2p+2 theorem)
    float m = sqrtf(re * re + im * im);
    float g = 1.0f / m;
    LimSample s;
    s.re = re * g;
    s.im = im * g;
    return s;
}

__global__ void __launch_bounds__(FE_WARPS * 32) k_frontend(const uint32_t *__restrict__ iq, int64_t nchan, int64_t T, RxChanState *st,
                                                            float *__restrict__ disc, float *__restrict__ mean) {
    __shared__ uint32_t tin[FE_WARPS][32][33];
    __shared__ float tout[FE_WARPS][32][33];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nitems = nchan * T;
    const int64_t item0 = ((int64_t)blockIdx.x * FE_WARPS + wid) * 32;
    if (item0 >= nitems) return;
    const int64_t item = item0 + lane;
    const bool live = item < nitems;
    const int64_t ch = live ? item / T : 0;
    const int64_t t = live ? item % T : 0;

    // carried discriminator state: z[0], z[1] are the two previous LIMITED samples (m17_dsp.cpp:196,205-206)
    float z0re = 0, z0im = 0, z1re = 0, z1im = 0;
    int count = 0;
    if (live) {
        count = st[ch].disc_count;                       // 1920 % 5 == 0: the /5 phase is the same in every block
        if (t == 0) { z0re = st[ch].z0re; z0im = st[ch].z0im; z1re = st[ch].z1re; z1im = st[ch].z1im; }
        else {
            const uint32_t *prev = iq + item * 1920;
            LimSample a = fe_limit(__ldg(prev - 1)), b = fe_limit(__ldg(prev - 2));
            z0re = a.re; z0im = a.im; z1re = b.re; z1im = b.im;
        }
    }
    float offset = 0;
    int nout = 0;                                        // outputs in the current 160-sample group
    for (int tile = 0; tile < 60; tile++) {
        // coalesced loads: row r = item0 + r, 32 consecutive samples
        uint32_t v[32];
#pragma unroll
        for (int r = 0; r < 32; r++) v[r] = (item0 + r < nitems) ? __ldg(iq + (item0 + r) * 1920 + tile * 32 + lane) : 0x00010001u;
#pragma unroll
        for (int r = 0; r < 32; r++) tin[wid][r][lane] = v[r];
        __syncwarp();
#pragma unroll 8
        for (int s = 0; s < 32; s++) {
            LimSample x = fe_limit(tin[wid][lane][s]);
            // dsp_arctan_disc2 (m17_dsp.cpp:203-212)
            float a = z0im * (x.re - z1re);
            float b = z0re * (x.im - z1im);
            float u = b - a;
            z1re = z0re; z1im = z0im; z0re = x.re; z0im = x.im;
            float uc = u * 0.5f;
            count = (count + 1 == 5) ? 0 : count + 1;
            if (count == 0) tout[wid][lane][nout++] = uc;
            offset += uc;
        }
        __syncwarp();
        if (tile % 5 == 4) {                             // 160 samples done: exactly 32 outputs per lane
            const int grp = tile / 5;
#pragma unroll 4
            for (int r = 0; r < 32; r++)
                if (item0 + r < nitems) disc[(item0 + r) * 384 + grp * 32 + lane] = tout[wid][r][lane];
            nout = 0;
            __syncwarp();
        }
    }
    if (live) {
        mean[item] = offset / 1920.0f;                   // offset/len (m17_dsp.cpp:214)
        if (t == T - 1) { st[ch].nz0re = z0re; st[ch].nz0im = z0im; st[ch].nz1re = z1re; st[ch].nz1im = z1im; }
    }
}
