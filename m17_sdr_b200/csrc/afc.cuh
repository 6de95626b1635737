// afc.cuh -- 48 kHz RX front end WITH automatic frequency control: dsp_short_to_float -> dsp_nco_mixer -> dsp_limit ->
// dsp_arctan_disc2 -> radio_afc as chained by m17_dsp_rx when radio_get_afc_status() is true
// (m17_dsp.cpp:136-141,390-408,412-419,194-222,461-472; radio.cpp:196-208).
//
// AFC closes a loop around the whole chain: the NCO step of block t is the loop state after block t-1, and it only moves
// while the framer is in a frame (m17_db_in_frame(), set by m17_aos / cleared by m17_los, i.e. the framer's lock flag at
// the block boundary).  Blocks of one channel are therefore serial -- but channels are not coupled, so the front end of a block
// runs INSIDE the timing-loop kernel's block loop (k_sync_frame<.., AFC = true>, sync.cuh): the same warp that owns the channel
// mixes / limits / discriminates the block into the shared-memory row the timing loop reads, one launch for all blocks
// instead of two launches per block.  Inside a block the work is spread over the WARP:
//   1. the NCO phase chain acc[i+1] = acc[i] + delta (sequential double adds in the reference) is resolved exactly:
//      lane l hypothesises acc[60 l] = RN(base + 60 l * delta) -- exact as long as no add was rounded since `base` --
//      runs its 60 adds, and compares its end value bit for bit with the next lane's start; the first mismatch becomes
//      the new base for the lanes behind it (adds only round where |acc| crosses a binade upwards: a handful per block);
//   2. every lane mixes and limits its 60 samples (double sincos, IEEE sqrt / divide) into shared memory;
//   3. the discriminator runs sample-parallel over the limited samples (coalesced stores of the kept fifth);
//   4. lane 0 adds the 1920 discriminator values in the reference's order (the fp32 sum is not associative) and updates
//      the loop: delta -= 0.1 * mean while in a frame, phase wrapped by modf.
// cos/sin: a lane's phases are covered by one double sincos and a double rotation recurrence (step 2), good to ~1e-14; after
// rounding to float the values equal the reference's except when the double result falls within that distance of a float
// rounding boundary (about one sample in a few million), where the float differs by one ulp.  Parity of this path is therefore stated as: decoded records exact, symbols within 1e-5 relative
// RMS (they are bit-identical whenever no such sample occurred, which the tests also report).
#pragma once
#include "frontend.cuh"

#define AFC_PER_LANE 60            // 1920 / 32

struct AfcWarpSmem {
    float2 lim[M17B_BLOCK_SAMPLES + 2];     // [0], [1] = z[1], z[0] carried in; [2 + i] = limited sample i
    float u[M17B_BLOCK_SAMPLES];            // discriminator values before the x0.5 (m17_dsp.cpp:209)
};

// One 40-ms block of one channel, by one warp.  Loop state (NCO phase, AFC delta) is carried in registers by the caller (every
// lane holds the same values); the discriminator history z[1], z[0] lives in sm.lim[0..1] between blocks.  Writes the block's
// 384 kept discriminator values (not mean-removed) and its mean to out384 / mean_out (shared memory of the timing loop) and
// to the global rows drow / mrow (the view the tests read).
__device__ __forceinline__ void afc_block(AfcWarpSmem &sm, const uint32_t *__restrict__ iqrow, int lane, int in_frame, int count0,
                                          float &afc_delta, double &nco_acc, float *out384, float *mean_out, float *__restrict__ drow, float *__restrict__ mrow) {
    if (!in_frame) afc_delta = 0.0f;                                // radio_get_afc_delta, radio.cpp:201-208
    const double dl = (double)afc_delta;
    const double acc0 = nco_acc;

    // ---- 1. exact NCO phase at the start of each lane's 60 samples
    double start = acc0, end = acc0, base = acc0;
    int j0 = 0, first = 0;
    for (int it = 0; it < 33; it++) {
        if (lane >= first) {
            start = base + (double)(AFC_PER_LANE * lane - j0) * dl;
            end = start;
#pragma unroll 4
            for (int k = 0; k < AFC_PER_LANE; k++) end += dl;       // acc += delta (m17_dsp.cpp:395)
        }
        const long long nxt = __shfl_down_sync(0xffffffffu, __double_as_longlong(start), 1);
        const bool bad = lane >= first && lane < 31 && __double_as_longlong(end) != nxt;
        const unsigned m = __ballot_sync(0xffffffffu, bad);
        if (!m) break;
        const int f = __ffs(m) - 1;                                 // lanes <= f started from the true phase
        base = __longlong_as_double(__shfl_sync(0xffffffffu, __double_as_longlong(end), f));
        j0 = AFC_PER_LANE * (f + 1);
        first = f + 1;
    }
    const double acc_end = __longlong_as_double(__shfl_sync(0xffffffffu, __double_as_longlong(end), 31));

    // ---- 2. int16 -> float, NCO mixer, limiter
    // cos / sin of the lane's 60 phases: one double sincos at the lane's (exactly resolved) start phase, then the rotation by
    // (cos delta, sin delta) in double with FMAs.  The recurrence drifts by ~1e-16 per step, i.e. < 1e-14 after 60 steps --
    // the same order as the distance between CUDA's and glibc's double sincos -- so the values ROUNDED TO FLOAT, which is all
    // the reference uses (float c = cos(acc), m17_dsp.cpp:393-394), differ from a per-sample sincos only when a double result
    // lies within ~1e-14 of a float rounding boundary: about one sample in a few million, by one float ulp (see header).
    {
        const uint4 *row = (const uint4 *)(iqrow + AFC_PER_LANE * lane);
        double sd, cd, rs, rc;
        sincos(start, &sd, &cd);
        sincos(dl, &rs, &rc);
        for (int q = 0; q < AFC_PER_LANE / 4; q++) {
            const uint4 w = __ldg(row + q);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t raw = j == 0 ? w.x : j == 1 ? w.y : j == 2 ? w.z : w.w;
                const int re_i = (int)(int16_t)(raw & 0xFFFFu), im_i = (int)(int16_t)(raw >> 16);
                const float re = __double2float_rn((double)re_i * 0.00003);          // dsp_short_to_float :138-139
                const float im = __double2float_rn((double)im_i * 0.00003);
                const float cs = __double2float_rn(cd), sn = __double2float_rn(sd);  // float c = cos(acc); float s = sin(acc);
                const double nc = fma(cd, rc, -(sd * rs)), ns = fma(sd, rc, cd * rs); // acc += delta
                cd = nc; sd = ns;
                const float nre = (re * cs) - (im * sn);                               // :396-397 (no contraction)
                const float nim = (re * sn) + (im * cs);
                const float m = sqrtf(nre * nre + nim * nim);                          // dsp_limit :414-417
                const float g = 1.0f / m;                                              // == (float)(1.0 / m) (2p+2 theorem)
                sm.lim[2 + AFC_PER_LANE * lane + 4 * q + j] = make_float2(nre * g, nim * g);
            }
        }
    }
    __syncwarp();

    // ---- 3. discriminator, sample-parallel (m17_dsp.cpp:203-212); every 5th value is kept
    const int keep = 4 - count0;
    for (int i = lane; i < M17B_BLOCK_SAMPLES; i += 32) {
        const float2 x = sm.lim[2 + i], z0 = sm.lim[1 + i], z1 = sm.lim[i];
        const float a = z0.y * (x.x - z1.x);
        const float b = z0.x * (x.y - z1.y);
        const float u = b - a;
        sm.u[i] = u;
        if (i % 5 == keep) { const float v = u * 0.5f; out384[i / 5] = v; drow[i / 5] = v; }
    }
    __syncwarp();

    // ---- 4. block mean in the reference's order, AFC loop update
    float mu = 0.0f;
    if (lane == 0) {
        float acc = 0.0f;
        const float4 *u4 = (const float4 *)sm.u;
#pragma unroll 4
        for (int i = 0; i < M17B_BLOCK_SAMPLES / 4; i++) { const float4 v = u4[i]; acc += v.x; acc += v.y; acc += v.z; acc += v.w; }
        mu = (acc * 0.5f) / 1920.0f;                                 // sum of u*0.5 == 0.5 * sum of u (exact scaling); offset/len :214
        *mean_out = mu;
        *mrow = mu;
    }
    mu = __shfl_sync(0xffffffffu, mu, 0);
    if (in_frame) afc_delta = __double2float_rn((double)afc_delta - (double)mu * 0.1);        // radio_afc, radio.cpp:196-200
    double a = acc_end / (2.0 * M_PI), ip;                           // :401-407
    a = modf(a, &ip);
    a = a * 2.0 * M_PI;
    if (a != a) a = 0;
    nco_acc = a;
    // z[1], z[0] for the next block
    float2 y0 = make_float2(0, 0), y1 = y0;
    if (lane == 0) { y0 = sm.lim[M17B_BLOCK_SAMPLES + 1]; y1 = sm.lim[M17B_BLOCK_SAMPLES]; }
    __syncwarp();
    if (lane == 0) { sm.lim[0] = y1; sm.lim[1] = y0; }
    __syncwarp();
}

__global__ void k_afc_off(RxChanState *st, int64_t nchan) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchan) st[c].afc_delta = 0.0f;                          // radio_set_afc_off, radio.cpp:149-152
}
