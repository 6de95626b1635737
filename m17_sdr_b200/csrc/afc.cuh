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
//   2. every lane mixes, limits and discriminates its 60 samples (IEEE sqrt / divide; the two samples of history a lane's first
//      two values need come from the previous lane by shuffle) -- only the 1920 discriminator values go to shared memory;
//   3. the kept fifth is stored lane-parallel (coalesced);
//   4. lane 0 adds the 1920 discriminator values in the reference's order (the fp32 sum is not associative) and updates
//      the loop: delta -= 0.1 * mean while in a frame, phase wrapped by modf.
// cos/sin: a lane's phases are covered by one double sincos and a double rotation recurrence (step 2), good to ~1e-14; after
// rounding to float the values equal the reference's except when the double result falls within that distance of a float
// rounding boundary (about one sample in a few million), where the float differs by one ulp.  Parity of this path is therefore stated as: decoded records exact, symbols within 1e-5 relative
// RMS (they are bit-identical whenever no such sample occurred, which the tests also report).
#pragma once
#include "frontend.cuh"

#define AFC_PER_LANE 60            // 1920 / 32

// 15.4 KB per warp (with the timing loop's 6.2 KB: 8 warps = two CTAs per SM, so that 1024 channels are resident at once).
struct AfcWarpSmem {
    uint4 raw[M17B_BLOCK_SAMPLES / 4];      // the block's int16 IQ row, fetched with cp.async while the previous block's mean is summed and
                                            // its timing loop runs (read straight from global memory the 15 loads of a lane cost a quarter
                                            // of the kernel's time: profiles/r02_afc_lines_before.txt)
    float u[M17B_BLOCK_SAMPLES];            // discriminator values before the x0.5 (m17_dsp.cpp:209)
};

// cp.async of one block's IQ row into sm.raw: 480 pieces of 16 bytes, piece p by lane p % 32 (coalesced); lane l later reads
// pieces 15 l .. 15 l + 14 (its 60 samples; the 240-byte lane stride is conflict-free for 16-byte accesses).
__device__ __forceinline__ void afc_prefetch(AfcWarpSmem &sm, const uint32_t *__restrict__ iqrow, int lane) {
#pragma unroll
    for (int k = 0; k < M17B_BLOCK_SAMPLES / 4 / 32; k++) {
        const int p = lane + 32 * k;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.raw[p]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(iqrow + 4 * p));
    }
    asm volatile("cp.async.commit_group;");
}

// m17_dsp.cpp:203-206: x = sample i, z0 = sample i-1, z1 = sample i-2
__device__ __forceinline__ float afc_disc(float2 x, float2 z0, float2 z1) {
    const float a = z0.y * (x.x - z1.x);
    const float b = z0.x * (x.y - z1.y);
    return b - a;
}

// One 40-ms block of one channel, by one warp.  Loop state (NCO phase, AFC delta, the discriminator history zc1 = z[1], zc0 = z[0])
// is carried in registers by the caller (every lane holds the same values).  The block's IQ row is already on its way into
// sm.raw (afc_prefetch); iq_next (or null) is the row to fetch for the next call.  Writes the block's 384 kept discriminator
// values (not mean-removed) and its mean to out384 / mean_out (shared memory of the timing loop) and to the global rows
// drow / mrow (the view the tests read).
__device__ __forceinline__ void afc_block(AfcWarpSmem &sm, const uint32_t *__restrict__ iq_next, int lane, int in_frame, int count0,
                                          float &afc_delta, double &nco_acc, float2 &zc1, float2 &zc0,
                                          float *out384, float *mean_out, float *__restrict__ drow, float *__restrict__ mrow) {
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

    // ---- 2. int16 -> float, NCO mixer, limiter, discriminator: each lane walks its 60 samples
    // cos / sin of the lane's 60 phases: one double sincos at the lane's (exactly resolved) start phase, then the rotation by
    // (cos delta, sin delta) in double with FMAs.  The recurrence drifts by ~1e-16 per step, i.e. < 1e-14 after 60 steps --
    // the same order as the distance between CUDA's and glibc's double sincos -- so the values ROUNDED TO FLOAT, which is all
    // the reference uses (float c = cos(acc), m17_dsp.cpp:393-394), differ from a per-sample sincos only when a double result
    // lies within ~1e-14 of a float rounding boundary: about one sample in a few million, by one float ulp (see header).
    // The discriminator value of sample i needs the limited samples i-1 and i-2: inside the lane's run they are the two
    // registers behind the walk; the lane's first two values wait for the previous lane's last two samples (one shuffle pair
    // after the walk; lane 0 takes the history carried in from the previous block).  Nothing but the 1920 discriminator values
    // goes through shared memory.
    asm volatile("cp.async.wait_group 0;");
    __syncwarp();
    float2 x0 = make_float2(0, 0), x1 = x0, p0 = x0, p1 = x0;       // x0, x1: the lane's first two samples; p0, p1: samples i-1, i-2
    {
        const uint4 *row = sm.raw + (AFC_PER_LANE / 4) * lane;
        float4 *urow = (float4 *)(sm.u + AFC_PER_LANE * lane);
        double sd, cd, rs, rc;
        sincos(start, &sd, &cd);
        sincos(dl, &rs, &rc);
        constexpr float c_hi = 0.00003f;
        constexpr float c_lo = (float)(0.00003 - (double)0.00003f);
        for (int q = 0; q < AFC_PER_LANE / 4; q++) {
            const uint4 w = row[q];
            float uo[4];
#pragma unroll
            for (int jp = 0; jp < 2; jp++) {                                         // two samples at a time (packed normaliser)
                float nre[2], nim[2], ss[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t raw = jp == 0 ? (h == 0 ? w.x : w.y) : (h == 0 ? w.z : w.w);
                    const float xr = (float)(short)(raw & 0xFFFFu), xi = (float)(short)(raw >> 16);
                    // dsp_short_to_float :138-139: fmaf(x, c_hi, x * c_lo) == (float)((double)x * 0.00003) for every int16 x
                    // (part of the exhaustive front-end self-test, frontend.cuh)
                    const float re = fmaf(xr, c_hi, xr * c_lo);
                    const float im = fmaf(xi, c_hi, xi * c_lo);
                    const float cs = __double2float_rn(cd), sn = __double2float_rn(sd);  // float c = cos(acc); float s = sin(acc);
                    const double nc = fma(cd, rc, -(sd * rs)), ns = fma(sd, rc, cd * rs); // acc += delta
                    cd = nc; sd = ns;
                    nre[h] = (re * cs) - (im * sn);                                    // :396-397 (no contraction)
                    nim[h] = (re * sn) + (im * cs);
                    ss[h] = nre[h] * nre[h] + nim[h] * nim[h];
                }
                // dsp_limit :414-417: m = sqrtf(s), 1.0 / m -- the front end's fast normaliser, equal to the IEEE forms for every
                // float s >= 2^-102 (m17b_selftest_limiter; here s >= 8e-10: a rotated nonzero int16 sample)
                float nma, nmb, ga, gb;
                fe_norm_pair(pack2(ss[0], ss[1]), nma, nmb, ga, gb);
                const float2 xa = make_float2(nre[0] * ga, nim[0] * ga), xb = make_float2(nre[1] * gb, nim[1] * gb);
                uo[2 * jp] = afc_disc(xa, p0, p1);                                     // (the lane's first two: placeholders)
                uo[2 * jp + 1] = afc_disc(xb, xa, p0);
                p1 = xa; p0 = xb;
                if (q == 0 && jp == 0) { x0 = xa; x1 = xb; }
            }
            urow[q] = make_float4(uo[0], uo[1], uo[2], uo[3]);
        }
    }
    {
        float2 m1, m2;                                              // samples 60 l - 1, 60 l - 2
        m1.x = __shfl_up_sync(0xffffffffu, p0.x, 1); m1.y = __shfl_up_sync(0xffffffffu, p0.y, 1);
        m2.x = __shfl_up_sync(0xffffffffu, p1.x, 1); m2.y = __shfl_up_sync(0xffffffffu, p1.y, 1);
        if (lane == 0) { m1 = zc0; m2 = zc1; }
        *(float2 *)(sm.u + AFC_PER_LANE * lane) = make_float2(afc_disc(x0, m1, m2), afc_disc(x1, x0, m1));
        zc0.x = __shfl_sync(0xffffffffu, p0.x, 31); zc0.y = __shfl_sync(0xffffffffu, p0.y, 31);     // z[0], z[1] for the next block
        zc1.x = __shfl_sync(0xffffffffu, p1.x, 31); zc1.y = __shfl_sync(0xffffffffu, p1.y, 31);
    }
    __syncwarp();
    if (iq_next) afc_prefetch(sm, iq_next, lane);                   // every lane is done with sm.raw

    // ---- 3. every 5th value is kept (m17_dsp.cpp:207-211)
    const int keep = 4 - count0;
#pragma unroll
    for (int k = 0; k < M17B_DISC_PER_BLOCK / 32; k++) {
        const int m = lane + 32 * k;
        const float v = sm.u[5 * m + keep] * 0.5f;
        out384[m] = v;
        drow[m] = v;
    }

    // ---- 4. block mean in the reference's order, AFC loop update
    float mu = 0.0f;
    if (lane == 0) {
        // 1920 dependent adds; the loads run one batch of 16 values ahead so that the chain never waits for shared memory
        float acc = 0.0f;
        const float4 *u4 = (const float4 *)sm.u;
        float4 a0 = u4[0], a1 = u4[1], a2 = u4[2], a3 = u4[3], b0, b1, b2, b3;
#define AFC_ADD16(v0, v1, v2, v3) acc += v0.x; acc += v0.y; acc += v0.z; acc += v0.w; acc += v1.x; acc += v1.y; acc += v1.z; acc += v1.w; \
                                  acc += v2.x; acc += v2.y; acc += v2.z; acc += v2.w; acc += v3.x; acc += v3.y; acc += v3.z; acc += v3.w
#pragma unroll 1
        for (int i = 4; i < M17B_BLOCK_SAMPLES / 4; i += 8) {       // two batches per trip (ping-pong: no register moves)
            b0 = u4[i]; b1 = u4[i + 1]; b2 = u4[i + 2]; b3 = u4[i + 3];
            AFC_ADD16(a0, a1, a2, a3);
            if (i + 4 < M17B_BLOCK_SAMPLES / 4) { a0 = u4[i + 4]; a1 = u4[i + 5]; a2 = u4[i + 6]; a3 = u4[i + 7]; }
            AFC_ADD16(b0, b1, b2, b3);
        }
#undef AFC_ADD16
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
    __syncwarp();
}

__global__ void k_afc_off(RxChanState *st, int64_t nchan) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchan) st[c].afc_delta = 0.0f;                          // radio_set_afc_off, radio.cpp:149-152
}
