// frontend.cuh -- 48 kHz RX front end, batched over (channel, 40 ms block) pairs:
//   int16 IQ -> float (x 0.00003 in double) -> hard limiter -> 2-sample-baseline cross-product FM
//   discriminator -> keep every 5th -> block mean (sequential fp32 sum of all 1920 values).
// Replaces dsp_short_to_float, dsp_limit, dsp_arctan_disc2 as chained by m17_dsp_rx
// (m17_dsp.cpp:136-141,412-419,194-222,461-472).
//
// With AFC off (the reference default, radio.cpp:146-155) nothing but two input samples and the /5 phase
// crosses a block boundary, so every (channel, block) item is independent.  The only serial work inside an
// item is the fp32 running sum, whose order must be kept for bit-exactness, hence:
//   mapping: one LANE per item, one warp per 32 items.  The warp fetches the rows cooperatively in 20-sample chunks
//   (coalesced 16-byte pieces, two chunks ahead, through a small shared-memory tile), and every lane walks its own row
//   sequentially with the discriminator history in registers.  Because 20 = lcm(4 samples per load,
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

#ifndef FE_WARPS
#define FE_WARPS 4
#endif

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

// g = bits(m) has an all-ones significand ? 0x7F000000 - bits(m) : g, given the bits of -m (0x7F000000 - bits(m) = 0xFF000000 -
// bits(-m)), in two instructions: a LOP3 (~x & 0x7FFFFF) that only writes its "result != 0" predicate, and the predicated
// subtract (the compiler's own form is LOP3 + ISETP + subtract).
__device__ __forceinline__ float fe_patch_all_ones(float g, uint32_t nmbits) {
    uint32_t gb = __float_as_uint(g);
    asm("{ .reg .pred p; .reg .b32 t; lop3.or.b32 t|p, %1, 0x7FFFFF, 0, 0x0C, 0; @!p sub.u32 %0, 0xFF000000, %1; }" : "+r"(gb) : "r"(nmbits));
    return __uint_as_float(gb);
}

// The limiter's normaliser for a pair of samples: from S = (s_a, s_b), s = re^2 + im^2, to -m = -RN(sqrt(s)) (returned as bit
// patterns nma, nmb) and g = RN(1 / m) (ga, gb).  Everything here is a function of s alone, so its equality with sqrtf / IEEE
// division does not depend on where s came from: m17b_selftest_limiter enumerates EVERY normal float s (used by the AFC front
// end, whose limiter input is the mixer output and not an int16 grid point; the int16 path has its own end-to-end enumeration,
// m17b_selftest_frontend).
__device__ __forceinline__ void fe_norm_pair(f32x2 S, float &nma, float &nmb, float &ga, float &gb) {
    float sa, sb, ya, yb;
    unpack2(S, sa, sb);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ya) : "f"(sa));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yb) : "f"(sb));
    const f32x2 Y = pack2(ya, yb);
    const f32x2 NEG1 = pack2(-1.0f, -1.0f), ONE = pack2(1.0f, 1.0f);
    // -m = -RN(sqrt(s)): Newton step on the residual s - m0*m0, carried out on the negated iterate (round-to-nearest is
    // symmetric, so fma(r, -y/2, -m0) is exactly -(m0 + r*y/2)); only -m is needed below
    const f32x2 M0 = mul2(S, Y);
    const f32x2 NM0 = mul2(M0, NEG1);
    // -y/2 on the ALU pipe instead of a packed multiply on the FMA pipe (which the packed arithmetic of this kernel keeps busiest):
    // one integer add per half flips the sign and decrements the exponent.  Exact whenever y and y/2 are normal: y = rsqrt(s) has
    // an exponent of about -e(s)/2, so that holds for every normal s (m17b_selftest_limiter enumerates them).  0.523 -> 0.516 ms on
    // the bench workload; doing the same for -m0 (a sign flip by LOP3) gains as much alone (0.518) and nothing on top (0.518).
    const f32x2 NHY = pack2(__uint_as_float(__float_as_uint(ya) + 0x7F800000u), __uint_as_float(__float_as_uint(yb) + 0x7F800000u));
    const f32x2 NM = fma2(fma2(NM0, M0, S), NHY, NM0);
    // g = RN(1/m): two residual corrections starting from y ~ 1/m (Markstein): fma(-m, g, 1) = 1 - m*g, g + (1 - m*g)*g.
    // The one input class this cannot round correctly is a divisor whose significand is all ones: the Newton iterate then
    // lands exactly on a rounding midpoint while the true quotient 2^-(e+1) (1 + 2^-24 + ..) lies just above it; its correctly
    // rounded value is known in closed form, 2^-(e+1) (1 + 2^-23) = bits 0x7F000000 - bits(m).
    f32x2 G = fma2(fma2(NM, Y, ONE), Y, Y);
    G = fma2(fma2(NM, G, ONE), G, G);
    unpack2(NM, nma, nmb);
    unpack2(G, ga, gb);
    // (Tried in round 2: leave the patch out of the walking loop -- a running VIMNMX3 of bits(m) | 0xFF800000 instead, one warp
    // vote per 80-sample segment, and a second pass with the patch over the segments that need it, about 1 % of the warp-units
    // of the bench workload.  Exact (tests/gpu_check.py check_rx_chain_limiter_patch plants the class), 1.5 instructions per
    // sample fewer, and no faster.  Redoing whole warp-units instead was 0.72 ms -- a unit is half the kernel's duration, so
    // any unit that runs twice in the second wave extends the kernel by a quarter.)
    ga = fe_patch_all_ones(ga, __float_as_uint(nma));
    gb = fe_patch_all_ones(gb, __float_as_uint(nmb));
}

// Two samples at a time, packed ACROSS the two samples: RE = (re_a, re_b), IM = (im_a, im_b).  Scaling, squares, the sum of
// squares, the Newton / Markstein residual steps for sqrt and reciprocal and the final limiter scaling are then all packed
// operations on naturally aligned pairs (one FMUL2 / FFMA2 / FADD2 serves both samples), and the outputs XRE = (x_a.re, x_b.re),
// XIM = (x_a.im, x_b.im) are the form the discriminator wants.  Every half of every packed operation is an independent IEEE
// round-to-nearest operation, so the results are those of the scalar formulation, which m17b_selftest_frontend compares with
// fe_limit_ieee over all 2^32 raw words.
// `one` = (1.0f, 1.0f) from a kernel argument: re^2 + im^2 must be two rounded products and one rounded sum, but ptxas contracts
// mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (common.cuh); fma2(im^2, one, re^2) is the same rounded sum and cannot be
// contracted because the multiplier is not a compile-time 1.
__device__ __forceinline__ void fe_limit_pair(uint32_t raw_a, uint32_t raw_b, f32x2 one, f32x2 &XRE, f32x2 &XIM, float *mo = nullptr, float *go = nullptr) {
    // int16 -> float (exact either way): I2F.S16 reads the half-register directly on the XU pipe.  (Tried twice in round 2: the
    // conversion on the ALU / FMA pipes instead -- (x ^ 0x8000) in the low mantissa bits of 2^23, minus 2^23 + 32768, exact and
    // verified over all 2^32 inputs -- to unload the XU pipe, where I2F draws twice its share of the stall samples (mio /
    // short_sb): 0.606 -> 0.631 ms on the first form of this kernel, 0.526 -> 0.553 ms on this one.  The four extra ALU / packed
    // instructions per sample cost more than the two I2F they replace.)
    // (Round 2, last: the real component alone converted on the ALU pipe -- PRMT sign-extends the low half, I2FP.F32.S32 converts --
    // while the imaginary one keeps its I2F.S16: one instruction more per sample, two XU slots per pair fewer, 0.5125 -> 0.5095 ms.
    // The imaginary component alone the same way: no change; both: 0.5117.)
    auto cvt_lo = [](uint32_t raw) -> float { int v; asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(v) : "r"(raw)); return __int2float_rn(v); };
    auto cvt_hi = [](uint32_t raw) -> float { return (float)(short)(raw >> 16); };
    const f32x2 xr = pack2(cvt_lo(raw_a), cvt_lo(raw_b));
    const f32x2 xi = pack2(cvt_hi(raw_a), cvt_hi(raw_b));
    constexpr float c_hi = 0.00003f;
    constexpr float c_lo = (float)(0.00003 - (double)0.00003f);
    const f32x2 CH = pack2(c_hi, c_hi), CL = pack2(c_lo, c_lo);
    const f32x2 RE = fma2(xr, CH, mul2(xr, CL));               // fmaf(x, c_hi, x * c_lo) == (float)((double)x * 0.00003)
    const f32x2 IM = fma2(xi, CH, mul2(xi, CL));
    const f32x2 S = fma2(mul2(IM, IM), one, mul2(RE, RE));     // two rounded products, one rounded sum (see above)
    float nma, nmb, ga, gb;
    fe_norm_pair(S, nma, nmb, ga, gb);
    if (mo) { mo[0] = -nma; mo[1] = -nmb; go[0] = ga; go[1] = gb; }
    const f32x2 G = pack2(ga, gb);
    XRE = mul2(RE, G);
    XIM = mul2(IM, G);
}
__device__ __forceinline__ void fe_limit2(uint32_t raw_a, uint32_t raw_b, f32x2 one, LimSample &oa, LimSample &ob, float *mo = nullptr, float *go = nullptr) {
    f32x2 XRE, XIM;
    fe_limit_pair(raw_a, raw_b, one, XRE, XIM, mo, go);
    unpack2(XRE, oa.re, ob.re);
    unpack2(XIM, oa.im, ob.im);
}
__device__ __forceinline__ LimSample fe_limit(uint32_t raw, f32x2 one, float *mo = nullptr, float *go = nullptr) {
    LimSample a, b;
    float m2[2], g2[2];
    fe_limit2(raw, raw, one, a, b, mo ? m2 : nullptr, mo ? g2 : nullptr);
    if (mo) { *mo = m2[0]; *go = g2[0]; }
    return a;
}

// The /5 decimation phase: dsp_arctan_disc2 keeps the sample at which count = (count+1)%5 reaches 0 (m17_dsp.cpp:207-211).
// The reference only ever feeds whole 1920-sample blocks (m17_dsp_rx, m17_tx_rx.cpp), as does this ABI, and 1920 % 5 == 0, so
// `count` is 0 at every block boundary and the kept samples are those at positions 4 (mod 5): a compile-time pattern.
#define FE_KEEP 4

// Items are the (channel, block) pairs of blocks [t0, t0+Tc) of a call of T blocks per channel (T is the row pitch of iq,
// disc and mean); the whole call is t0 = 0, Tc = T.  Sub-ranges let the host pipeline the front end of one time slice with
// the timing loop of the previous one (rx.cuh).
__global__ void __launch_bounds__(FE_WARPS * 32, 28 / FE_WARPS) k_frontend(const uint32_t *__restrict__ iq, int64_t nchan, int64_t T, int64_t t0, int64_t Tc,
                                                            RxChanState *st, float *__restrict__ disc, float *__restrict__ mean, f32x2 one) {
    __shared__ float tout[FE_WARPS][32][17];
    __shared__ __align__(16) uint4 stage[FE_WARPS][2][160];      // two 20-sample chunks of the warp's 32 rows
    __shared__ int64_t gsl[FE_WARPS][32];                        // global (channel, block) index of each lane's item
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nitems = nchan * Tc;
    const int64_t item0 = ((int64_t)blockIdx.x * FE_WARPS + wid) * 32;      // one unit of 32 items per warp, channel-major
    if (item0 >= nitems) return;
    {
    const bool live = item0 + lane < nitems;
    const int64_t item = live ? item0 + lane : nitems - 1;       // dead lanes shadow the last item (results discarded)
    const int64_t ch = item / Tc, t = t0 + item % Tc;
    const int64_t g = ch * T + t;
    gsl[wid][lane] = g;
    __syncwarp();

    // carried discriminator state: z[0], z[1] are the two previous LIMITED samples (m17_dsp.cpp:196,205-206)
    float z0re, z0im, z1re, z1im;
    if (t == 0) { z0re = st[ch].z0re; z0im = st[ch].z0im; z1re = st[ch].z1re; z1im = st[ch].z1im; }
    else {
        const uint32_t *prev = iq + g * 1920;
        LimSample a = fe_limit(__ldg(prev - 1), one), b = fe_limit(__ldg(prev - 2), one);
        z0re = a.re; z0im = a.im; z1re = b.re; z1im = b.im;
    }
    float acc = 0.0f;                                            // sum of u; sum of u*0.5 == 0.5*sum (exact power-of-two scaling)
    // the two previous limited samples as the pair the packed discriminator subtracts: (x[n-2], x[n-1]) per component
    f32x2 PRE = pack2(z1re, z0re), PIM = pack2(z1im, z0im);
    auto process20 = [&](const uint4 *w, int slot) {              // w: the lane's five 16-byte pieces (registers or shared memory)
#pragma unroll
        for (int s = 0; s < 20; s += 2) {
            const uint4 q = w[s >> 2];
            const uint32_t raw0 = (s & 3) == 0 ? q.x : q.z, raw1 = (s & 3) == 0 ? q.y : q.w;
            f32x2 XRE, XIM;
            fe_limit_pair(raw0, raw1, one, XRE, XIM);
            // dsp_arctan_disc2 (m17_dsp.cpp:203-212), two samples: u[n] = x[n-1].re * (x[n].im - x[n-2].im) - x[n-1].im * (x[n].re - x[n-2].re).
            // The differences of both samples are one packed subtract per component (the subtrahend pair IS the previous
            // output pair), the four products are scalar (their factors sit in different halves), b - a is packed again.
            const f32x2 DRE = sub2(XRE, PRE), DIM = sub2(XIM, PIM);
            float dre0, dre1, dim0, dim1, pre1, pim1, xre0, xim0, pre0_, pim0_, xre1_, xim1_;
            unpack2(DRE, dre0, dre1); unpack2(DIM, dim0, dim1);
            unpack2(PRE, pre0_, pre1); unpack2(PIM, pim0_, pim1);
            unpack2(XRE, xre0, xre1_); unpack2(XIM, xim0, xim1_);
            const float a0 = pim1 * dre0, b0 = pre1 * dim0;
            const float a1 = xim0 * dre1, b1 = xre0 * dim1;
            float u0, u1;
            unpack2(sub2(pack2(b0, b1), pack2(a0, a1)), u0, u1);
            acc += u0;
            acc += u1;
            if (s % 5 == FE_KEEP) tout[wid][lane][slot * 4 + s / 5] = u0 * 0.5f;
            if ((s + 1) % 5 == FE_KEEP) tout[wid][lane][slot * 4 + (s + 1) / 5] = u1 * 0.5f;
            PRE = XRE; PIM = XIM;
        }
    };
    // Loads.  A warp-wide LDG.128 whose lanes each walk their own row (7680 B apart) costs 32 L1 tag wavefronts for 512 B, so
    // the warp fetches each 20-sample chunk of its 32 rows COOPERATIVELY instead: the chunk is 160 16-byte pieces (5 per row);
    // piece p = lane + 32 k is loaded by `lane`, so one load instruction covers 6.4 rows x 80 contiguous bytes (about 8
    // wavefronts).  The pieces land in a shared-memory tile (piece p at byte 16 p: the rows come out contiguous at an 80-byte
    // pitch = 4 x 5 words, conflict-free for 16-byte row reads), and each lane reads its own row back.  Everything is sized
    // for 28 resident warps per SM (<= 72 registers, 7.5 KB of shared memory per warp): the 8000 warp-units of the 1024 x 250
    // workload then fit in two full waves of 148 x 28.
    uint32_t off[5];                                               // piece offsets in bytes from the warp's first row (rows ascend with the item)
    const char *base = (const char *)(iq + gsl[wid][0] * 1920);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int p = lane + 32 * k, r = p / 5;
        int64_t it = item0 + r;
        if (it >= nitems) it = nitems - 1;
        const int64_t gr = (it / Tc) * T + t0 + it % Tc;
        off[k] = (uint32_t)((gr - gsl[wid][0]) * 7680 + (p - 5 * r) * 16);
    }
    // two tiles of 160 pieces (one 20-sample chunk of the warp's 32 rows each), filled two chunks ahead with 16-byte cp.async:
    // completion is tracked by the async-copy group, NOT by a register scoreboard -- with plain loads into registers the
    // compiler's scoreboard sharing made the first arithmetic instruction of every chunk wait for that chunk's own prefetch
    // (17 % of all stall samples, profiles/r01b_ncu_hotspots.txt)
    auto fetch = [&](int c20) {
        if (c20 < 96) {
            uint4 *tl = stage[wid][c20 & 1];
            const char *bc = base + c20 * 80;                        // warp-uniform; the lane adds its 32-bit piece offset
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(tl + lane + 32 * k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(bc + off[k]));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    const int64_t g0 = gsl[wid][0];
    const bool contig = item0 + 31 < nitems && gsl[wid][31] - g0 == 31;
    fetch(0);
    fetch(1);
    for (int seg = 0; seg < 24; seg++) {                           // 24 segments of 4 chunks = 80 samples -> 16 kept values per row
#pragma unroll 1
        for (int chk = 0; chk < 4; chk++) {
            const int c20 = seg * 4 + chk;
            asm volatile("cp.async.wait_group 1;");                // chunk c20 has landed (chunk c20 + 1 may still be in flight)
            __syncwarp();
            process20(stage[wid][c20 & 1] + lane * 5, chk);        // row pieces are read from the tile as they are needed
            __syncwarp();                                          // every lane is done with its row: the tile may be refilled
            fetch(c20 + 2);
        }
        // flush: 16 kept values per row, half a 128-byte line per row and store
        {
            const int r0 = lane >> 4, col = lane & 15;
            if (contig) {
                // the unit's 32 rows are consecutive rows of disc (always, when the call covers whole channels): one pointer,
                // compile-time row offsets
                float *d = disc + (g0 + r0) * 384 + seg * 16 + col;
                const float *tsrc = &tout[wid][r0][col];
#pragma unroll
                for (int r = 0; r < 32; r += 2) d[r * 384] = tsrc[r * 17];
            } else {
#pragma unroll 4
                for (int r = 0; r < 32; r += 2)
                    if (item0 + r + r0 < nitems) disc[gsl[wid][r + r0] * 384 + seg * 16 + col] = tout[wid][r + r0][col];
            }
        }
        __syncwarp();
    }
    if (live) {
        mean[g] = (acc * 0.5f) / 1920.0f;                     // offset/len (m17_dsp.cpp:214)
        if (t == T - 1) {
            unpack2(PRE, z1re, z0re); unpack2(PIM, z1im, z0im);
            st[ch].nz0re = z0re; st[ch].nz0im = z0im; st[ch].nz1re = z1re; st[ch].nz1im = z1im;
        }
    }
    }
}

// Exhaustive proof-by-enumeration that fe_limit == fe_limit_ieee for every possible int16 IQ pair except (0,0)
// (which the reference itself turns into NaN).  Counts bitwise mismatches of either output component.
__global__ void k_selftest_frontend(unsigned long long *mism, uint32_t lo, uint32_t count, uint32_t *dump, int dump_cap, f32x2 one) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = 0;
    for (uint32_t k = i; k < count; k += gridDim.x * blockDim.x) {
        const uint32_t raw = lo + k;
        if (raw == 0) continue;
        float m1, g1, m2, g2;
        const LimSample a = fe_limit(raw, one, &m1, &g1), b = fe_limit_ieee(raw, &m2, &g2);
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
// The same proof for the normaliser alone, over float bit patterns of s = re^2 + im^2 (fe_norm_pair): m must equal sqrtf(s) and
// g must equal 1.0f / m (IEEE) bit for bit.
__global__ void k_selftest_limiter(unsigned long long *mism, uint32_t lo, uint32_t count, uint32_t *dump, int dump_cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = 0;
    for (uint32_t k = i; k < count; k += gridDim.x * blockDim.x) {
        const uint32_t sb = lo + k;
        const float s = __uint_as_float(sb);
        float nma, nmb, ga, gb;
        fe_norm_pair(pack2(s, s), nma, nmb, ga, gb);
        const float m = __fsqrt_rn(s), g = __frcp_rn(m);
        if ((__float_as_uint(-nma) != __float_as_uint(m)) | (__float_as_uint(ga) != __float_as_uint(g)) | (__float_as_uint(nmb) != __float_as_uint(nma)) | (__float_as_uint(gb) != __float_as_uint(ga))) {
            bad++;
            if (dump) {
                unsigned long long slot = atomicAdd(mism + 1, 1ull);
                if (5 * slot + 4 < (unsigned long long)dump_cap) {
                    dump[5 * slot] = sb; dump[5 * slot + 1] = __float_as_uint(-nma); dump[5 * slot + 2] = __float_as_uint(m);
                    dump[5 * slot + 3] = __float_as_uint(ga); dump[5 * slot + 4] = __float_as_uint(g);
                }
            }
        }
    }
    if (bad) atomicAdd(mism, (unsigned long long)bad);
}
extern "C" int m17b_selftest_limiter(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream) {
    if (!ctx || !h_mismatches || first + count > (1ull << 32) || dump_cap < 0) return M17B_E_ARG;
    cudaStream_t st = as_stream(stream);
    unsigned long long *d;
    uint32_t *d_dump = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, 16));
    CUDA_TRY(cudaMemsetAsync(d, 0, 16, st));
    if (h_dump && dump_cap) { CUDA_TRY(cudaMalloc((void **)&d_dump, 4 * (size_t)dump_cap)); CUDA_TRY(cudaMemsetAsync(d_dump, 0, 4 * (size_t)dump_cap, st)); }
    for (uint64_t off = 0; off < count; off += (1ull << 30)) {
        const uint64_t n = count - off < (1ull << 30) ? count - off : (1ull << 30);
        k_selftest_limiter<<<148 * 16, 256, 0, st>>>(d, (uint32_t)(first + off), (uint32_t)n, d_dump, dump_cap);
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
        k_selftest_frontend<<<148 * 16, 256, 0, st>>>(d, (uint32_t)(first + off), (uint32_t)n, d_dump, dump_cap, F32X2_ONE);
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
