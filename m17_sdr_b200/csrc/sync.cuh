// sync.cuh -- RX matched filter + symbol-timing loop + sync-word correlator / framer, one WARP per channel.
// Replaces m17_rx_sync_samples (+ rx_sync_filter, sync_update, m17_sync_adjust: m17_rx_sync.cpp:25-99) and
// m17_rx_symbols / m17_rx_sym / m17_sync_check (m17_rx_frame.cpp:47-177).
//
// The timing loop is recursive (the polyphase branch m_index depends on an up/down counter fed by every
// symbol), but the counter moves by at most 1 per symbol and the branch only changes when |m_thr| crosses the
// threshold (10 unlocked / 80 locked).  So the warp SPECULATES: 32 lanes compute the next 64 symbols, two
// consecutive ones per lane (matched + derivative 31-tap dot products, sequential fp32 adds exactly as
// rx_sync_filter does, four independent chains per lane, the 33 window samples loaded once), with the
// current branch; a warp prefix-sum over the +-1 votes finds the first symbol at which the threshold would
// trip, symbols up to there are committed, the branch is stepped and speculation restarts.  Results are
// identical to the serial loop, including the forward bit-slip (a zero symbol inserted, one sample skipped),
// the backward slip (a symbol dropped) and votes that straddle a block boundary.
// While locked -- and while unlocked on a tracked signal (the previous block tripped at most once) -- the round is the whole
// 40-ms block: six consecutive symbols per lane, twelve independent 31-tap chains.
// The framer's unlocked search evaluates the 8-symbol sync window at 32 positions per step (exact sign-pattern pre-filter, then
// the variance test; warp ballot, first hit wins); the locked path only counts symbols and checks each completed frame's head,
// in straight-line code for the common case of one frame boundary per block.
// The next block's samples are prefetched as 16-byte cp.async pieces and staged with vector loads.
#pragma once
#include "eq.cuh"

#ifndef SY_WARPS
#define SY_WARPS 2            // channels (warps) per CTA.  1024 channels are 6.9 warps per SM: with 4-warp CTAs 108 SMs hold 8 warps and 40
                              // hold 4; 2-warp CTAs spread them 6..8 (bench step 1.496 -> 1.484 ms; one warp per CTA: 1.510)
#endif
#define SY_XQ   106           // (30 + 384 + 2 + 8 pad) / 4 entries per residue class
#define SY_HIST  (8 + 208)
#define SY_RECQ  32            // record queue per channel; a block completes at most 2 frames, flushed above SY_RECQ - 3

struct SyncWarpSmem {
    float x[4][SY_XQ];                  // discriminator samples incl. 30 of history: sample n lives at x[n & 3][n >> 2], so the
                                        // stride-4 windows of the 32 lanes (two symbols per lane) hit 32 different banks
    float hist[SY_HIST];                // [0,8): sliding sync window carried in; [8, 8+n): symbols emitted in this block
    float head[8];                      // m_f_sym[0..7] of the frame being collected
    float pre[2][384 + 4];              // cp.async landing zone for the NEXT block's raw samples (+ its mean), double buffered
    uint4 rec[SY_RECQ];                 // completed frames whose records are not written yet: (sym_off, type | flags << 8 | votes << 16 |
                                        // frame_errors << 24, spread, max) -- see flush_records
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
    return v;
}

// m17_sync_adjust (m17_rx_sync.cpp:45-72).  clk is the value m_clk has before the NEXT sample is processed.
__device__ __forceinline__ void sync_adjust(int TH, int &thr, int &index, int &clk, int &m_idx, float *out, int lane, float *mid = nullptr) {
    if (thr > TH) {
        index = (index + 1 == M17B_NF) ? 0 : index + 1;
        thr = 0;
        if (index == 0) { clk = 1; if (m_idx >= 0 && lane == 0) { out[m_idx] = 0.0f; if (mid) mid[m_idx] = 0.0f; } m_idx++; }
    }
    if (thr < -TH) {
        thr = 0;
        index = (index == 0) ? M17B_NF - 1 : index - 1;
        if (index == M17B_NF - 1) { clk = 1; m_idx--; }
    }
}

// Two consecutive symbols for one lane: windows xs[n0 .. n0+30] and xs[n0+2 .. n0+32], n0 = i + 4*lane.  R = i & 3 is
// warp-uniform, so array / offset of every tap are compile-time and the 32 lanes read consecutive words.
// The window load depends on the residue R of the first sample (compile-time indices into the split array); the arithmetic
// does not.  Each kernel calls load_window<R> under its switch on R and then ONE copy of the arithmetic: with dot2<0..3> and
// dot6<0..1> inlined whole, the packed dot products alone were 1240 of the kernel's 4500 SASS instructions, and instruction
// fetch is one of this kernel's stall reasons (no_instruction, profiles/r02f_ncu_hotspots.txt).
template <int R, int NX>
__device__ __forceinline__ void load_window(const float (*X)[SY_XQ], int base, float (&x)[NX]) {
#pragma unroll
    for (int k = 0; k < NX; k++) x[k] = X[(R + k) & 3][base + ((R + k) >> 2)];
}
// Two consecutive symbols for one lane: windows xs[n0 .. n0+30] and xs[n0+2 .. n0+32], n0 = i + 4*lane.  R = i & 3 is
// warp-uniform, so array / offset of every tap are compile-time and the 32 lanes read consecutive words.
// tp[k] = (matched tap k, derivative tap k): one FMUL2 with the sample broadcast gives both rounded products, and one
// FFMA2 (product * 1.0 + running pair) adds them to the two running sums -- each half rounds exactly like the reference's
// sum += in[i]*c[i] (m17_rx_sync.cpp:25-31).  `one` comes from memory: with a literal 1.0 ptxas folds the pair into a
// contracted FFMA2 (one rounding), -fmad=false notwithstanding.
__device__ __forceinline__ void dot2_core(const float (&x)[M17B_FN + 2], const f32x2 (&tp)[M17B_FN], f32x2 one, float &sa, float &da, float &sb, float &db) {
    f32x2 acc0 = mul2(tp[0], pack2(x[0], x[0])), acc1 = mul2(tp[0], pack2(x[2], x[2]));
#pragma unroll
    for (int k = 1; k < M17B_FN; k++) {
        acc0 = fma2(mul2(tp[k], pack2(x[k], x[k])), one, acc0);
        acc1 = fma2(mul2(tp[k], pack2(x[k + 2], x[k + 2])), one, acc1);
    }
    unpack2(acc0, sa, da);
    unpack2(acc1, sb, db);
}
template <int R>
__device__ __forceinline__ void dot2(const float (*X)[SY_XQ], int base, const f32x2 (&tp)[M17B_FN], f32x2 one, float &sa, float &da, float &sb, float &db) {
    float x[M17B_FN + 2];
    load_window<R>(X, base, x);
    dot2_core(x, tp, one, sa, da, sb, db);
}
// Six consecutive symbols for one lane (a whole 40-ms block in one warp round): windows xs[n0 + 2m .. n0 + 2m + 30], m = 0..5,
// n0 = i + 12*lane, so the lane loads 41 samples once.  With the residue-split layout consecutive lanes (stride 12 samples =
// 3 words per residue array) read 32 different banks.
__device__ __forceinline__ void dot6_core(const float (&x)[M17B_FN + 10], const f32x2 (&tp)[M17B_FN], f32x2 one, float (&s)[6], float (&d)[6]) {
    f32x2 acc[6];
#pragma unroll
    for (int m = 0; m < 6; m++) acc[m] = mul2(tp[0], pack2(x[2 * m], x[2 * m]));
#pragma unroll
    for (int k = 1; k < M17B_FN; k++) {
#pragma unroll
        for (int m = 0; m < 6; m++) acc[m] = fma2(mul2(tp[k], pack2(x[2 * m + k], x[2 * m + k])), one, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < 6; m++) unpack2(acc[m], s[m], d[m]);
}
template <int R>
__device__ __forceinline__ void dot6(const float (*X)[SY_XQ], int base, const f32x2 (&tp)[M17B_FN], f32x2 one, float (&s)[6], float (&d)[6]) {
    float x[M17B_FN + 10];
    load_window<R>(X, base, x);
    dot6_core(x, tp, one, s, d);
}

// Equaliser option (EQ = true): the matched filter's output HALF A SYMBOL before a symbol instant -- the same polyphase branch on the
// window one sample earlier, xm1 = that window's first sample and x[0 .. 29] the rest -- summed in tap order like rx_sync_filter
// (m17_rx_sync.cpp:25-31).  tp[k]'s low half is the matched tap.
template <int NX>
__device__ __forceinline__ float dot_mid(float xm1, const float (&x)[NX], int first, const f32x2 (&tp)[M17B_FN]) {
    float t, td;
    unpack2(tp[0], t, td);
    float acc = (first == 0 ? xm1 : x[first - 1]) * t;
#pragma unroll
    for (int k = 1; k < M17B_FN; k++) { unpack2(tp[k], t, td); acc += x[first - 1 + k] * t; }
    return acc;
}

// AFC = true (m17b_rx_set_afc): the block's discriminator samples do not come from memory but from the AFC front end run by
// the same warp at the top of the block loop (afc.cuh): iq = int16 IQ rows, disc / mean are then OUTPUTS (the raw samples and
// block means the caller may inspect).  15.4 KB more shared memory per warp: still two CTAs per SM.
// EQ = true (m17b_rx_set_equaliser; not together with AFC): every symbol is paired with the matched filter's output half a symbol
// earlier (dot_mid), the block's pairs go through eq_train_unknown (eq.cuh) and the framer sees the equaliser's output.  Upstream has
// no call site for its equaliser; this is the wiring SURVEY 8f rank 3 describes, checked against the oracle's seam flag 64, which is
// itself pinned to the reference's functions (oracle/ref/eq_shim.cpp).  The half-symbol values live in dynamic shared memory.
template <bool HAS_MEAN, bool AFC = false, bool EQ = false>
__global__ void __launch_bounds__(SY_WARPS * 32, 2) k_sync_frame(const float *__restrict__ disc, const float *__restrict__ mean, int64_t nchan, int64_t T,
                                                              int t0, int t1, int2 *frame_rng, RxChanState *st, const float *__restrict__ g_mf, const float *__restrict__ g_md,
                                                              float *syms, int64_t sym_pitch, int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base,
                                                              m17b_frame_rec *frames, int64_t fcap, int32_t *__restrict__ nframes,
                                                              m17b_event_rec *events, int64_t ecap, int32_t *__restrict__ nevents,
                                                              unsigned long long *stats, int commit_fe, f32x2 one,
                                                              const uint32_t *__restrict__ iq = nullptr, float *disc_out = nullptr, float *mean_out = nullptr,
                                                              RxEqState *eqs = nullptr) {
    __shared__ __align__(16) SyncWarpSmem sm_all[SY_WARPS];
    extern __shared__ __align__(16) unsigned char afc_smem_raw[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * SY_WARPS + wid;
    if (c >= nchan) return;
    SyncWarpSmem &sm = sm_all[wid];
    RxChanState *S = st + c;
    float *out = sm.hist + 8;
    float *mid = EQ ? (float *)afc_smem_raw + wid * SY_HIST + 8 : nullptr;   // mid[q] pairs with out[q]
    float mid_carry = 0.0f;                                                   // the half-symbol value in front of the next block's sample 0
    if (EQ) mid_carry = eqs[c].mid;
    const long long clk_start = clock64();
    unsigned dbg_rounds = 0;
#ifdef M17B_PHASE_CLOCKS
    long long ph[5] = {0, 0, 0, 0, 0}, pt = clk_start;
    bool ph_on = true;
    long long ph_blocks = 0;
#define PHASE(i) do { const long long now__ = clock64(); if (ph_on) ph[i] += now__ - pt; pt = now__; } while (0)
#else
#define PHASE(i) do {} while (0)
#endif

    // ---- load state (uniform loads)
    // blocks [t0, t1) of a call of T blocks: t0 = 0 starts the call (symbol carry, record / event counts from zero),
    // t0 > 0 appends to what the earlier slices of the same call produced
    if (commit_fe && lane == 0 && t1 == T) { S->z0re = S->nz0re; S->z0im = S->nz0im; S->z1re = S->nz1re; S->z1im = S->nz1im; }
    int clk = S->clk, thr = S->thr, index = S->index;
    float sumc = S->sum, difc = S->dif;
    int flock = S->flock, fclk = S->fclk, ferr = S->ferr, frame_start = S->frame_start, sym_total = S->sym_total;
    const int base_g = t0 == 0 ? sym_total : sym_base[c];
    const int sym_entry = sym_total;
    if (lane < 30) sm.x[lane & 3][lane >> 2] = S->tail[lane];
    if (lane < 8) { sm.hist[lane] = S->win[lane]; sm.head[lane] = S->head[lane]; }
    // carry: the last 192 symbols of the previous call move in front of the new ones
    float *sbuf = syms + c * sym_pitch;
    if (t0 == 0) {
        const int prev_n = S->prev_n;
        float tmp[6];
#pragma unroll
        for (int k = 0; k < 6; k++) tmp[k] = sbuf[prev_n + lane + 32 * k];     // = sbuf[CARRY + prev_n - 192 + idx]
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 6; k++) sbuf[lane + 32 * k] = tmp[k];
    }
    if (lane == 0 && t0 == 0) sym_base[c] = base_g;
    int nfr = t0 == 0 ? 0 : nframes[c], nev = t0 == 0 ? 0 : nevents[c], n_aos = 0, n_los = 0;
    const int nfr_entry = nfr;
    // Records are queued, not written, where the frame completes: the serial path per block keeps only what the next block's
    // decisions need (lock, error count, frame clock).  The queue is drained 30 frames at a time with the lanes in parallel --
    // the variance quotient (an IEEE divide) once per lane instead of once per block, the 64-byte record headers two per store.
    int nq = 0, nfr_out = nfr;
    auto push_record = [&](int type, int flags, int votes, int fe, float spread, float vmax) {
        if (lane == 0) sm.rec[nq] = make_uint4((uint32_t)frame_start, (uint32_t)type | ((uint32_t)flags << 8) | ((uint32_t)votes << 16) | ((uint32_t)fe << 24),
                                               __float_as_uint(spread), __float_as_uint(vmax));
        nq++; nfr++;
    };
    auto flush_records = [&]() {
        __syncwarp();
        if (lane < nq) {
            const uint4 e = sm.rec[lane];
            ((uint32_t *)&sm.rec[lane])[2] = __float_as_uint(sync_variance_of(__uint_as_float(e.z), __uint_as_float(e.w)));
        }
        __syncwarp();
        // words of m17b_frame_rec: 0 = sym_off, 1 = type | flags << 8 (golay_err, nbytes zero), 11 = votes << 16 | frame_errors << 24
        // (crc zero), 12 = variance; the rest zero until the decoder fills them
        const int wi = lane & 15;
        const int slot = wi == 0 ? 0 : (wi == 1 || wi == 11) ? 1 : wi == 12 ? 2 : -1;
        const uint32_t msk = wi == 1 ? 0x0000FFFFu : wi == 11 ? 0xFFFF0000u : 0xFFFFFFFFu;
        uint32_t *dst = (uint32_t *)(frames + c * fcap + nfr_out);
        for (int f = lane >> 4; f < nq; f += 2) {
            const uint32_t v = slot >= 0 ? (((const uint32_t *)&sm.rec[f])[slot] & msk) : 0u;
            if (nfr_out + f < fcap) dst[16 * f + wi] = v;
        }
        nfr_out += nq; nq = 0;
        __syncwarp();
    };
    int nsym_keep = 0;                      // lane l: symbol count of block t0 + 32 k + l, stored 32 blocks at a time
    f32x2 tp[M17B_FN];                      // (matched, derivative) tap pairs of the current polyphase branch
    int tap_index = -1, trips_prev = 0;
    __syncwarp();
    // The block's samples are fetched one block ahead with cp.async straight into shared memory: completion is tracked by
    // the async-copy group, not by a register scoreboard, so nothing in the timing loop ever waits on the DRAM latency.
    auto prefetch = [&](int64_t tt, int buf) {
        const float *src = disc + (c * T + tt) * 384;
#pragma unroll
        for (int q = 0; q < 3; q++) {                                          // 96 pieces of 16 bytes (rows are 1536-byte aligned)
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][4 * (lane + 32 * q)]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + 4 * (lane + 32 * q)));
        }
        if (HAS_MEAN && lane == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][384]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(mean + c * T + tt));
        }
        asm volatile("cp.async.commit_group;");
    };
    // AFC loop state: every lane carries the same values (radio.cpp:8-10, m17_dsp.cpp:391)
    float afc_delta = 0.0f;
    double nco_acc = 0.0;
    int disc_count = 0;
    float2 afc_z1 = make_float2(0, 0), afc_z0 = afc_z1;      // dsp_arctan_disc2's z[1], z[0] (m17_dsp.cpp:195-196)
    AfcWarpSmem *afc_sm = nullptr;
    if (AFC) {
        afc_sm = (AfcWarpSmem *)afc_smem_raw + wid;
        afc_delta = S->afc_delta; nco_acc = S->nco_acc; disc_count = S->disc_count;
        afc_z1 = make_float2(S->z1re, S->z1im); afc_z0 = make_float2(S->z0re, S->z0im);
        afc_prefetch(*afc_sm, iq + (c * T + t0) * M17B_BLOCK_SAMPLES, lane);
    } else prefetch(t0, 0);

    for (int64_t t = t0; t < t1; t++) {
        // ---- stage the block's 384 discriminator samples behind the 30 of history
        const int buf = (int)((t - t0) & 1);
#ifdef M17B_PHASE_UNLOCKED
        ph_on = !flock; ph_blocks += ph_on;                         // profile the blocks entered unlocked only
#endif
        if (AFC) afc_block(*afc_sm, t + 1 < t1 ? iq + (c * T + t + 1) * M17B_BLOCK_SAMPLES : nullptr, lane, flock, disc_count, afc_delta, nco_acc, afc_z1, afc_z0, sm.pre[buf], &sm.pre[buf][384],
                           disc_out + (c * T + t) * M17B_DISC_PER_BLOCK, mean_out + c * T + t);
        else asm volatile("cp.async.wait_group 0;");
        __syncwarp();
        {
            const float pmu = !HAS_MEAN ? 0.0f : sm.pre[buf][384];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                // four consecutive samples 4m .. 4m+3 (m = lane + 32 q) land at n = 30 + 4m + j: residue (2 + j) & 3, slot
                // 7 + m + ((2 + j) >> 2) -- every store has consecutive lanes on consecutive words of one residue array
                const int mq = lane + 32 * q;
                float4 v = *(const float4 *)&sm.pre[buf][4 * mq];
                if (HAS_MEAN) { v.x = v.x - pmu; v.y = v.y - pmu; v.z = v.z - pmu; v.w = v.w - pmu; }   // m17_dsp.cpp:217-219
                sm.x[2][7 + mq] = v.x;
                sm.x[3][7 + mq] = v.y;
                sm.x[0][8 + mq] = v.z;
                sm.x[1][8 + mq] = v.w;
            }
        }
        if (!AFC && t + 1 < t1) prefetch(t + 1, buf ^ 1);
        __syncwarp();

        PHASE(0);
        // ---- timing loop (m17_rx_sync.cpp:77-99); m17_rx_lock() is constant inside a block
        const int TH = flock ? 80 : 10;
        int i = 0, m_idx = 0;
        // Round size is a pure scheduling choice (any speculation length gives the same result).  Locked (threshold 80) a trip is
        // rare and the whole block goes in one round.  Unlocked (threshold 10) it depends on the input: on a tracked signal the
        // counter trips less than once per block and the one-round form is ~1/3 cheaper than three 64-symbol rounds; on noise it
        // trips every few dozen symbols and short rounds waste less.  So: one round if the previous block tripped at most once.
        const bool whole_block = flock || trips_prev <= 1;
        int trips = 0;
        while (i < 384) {
            if (clk == 1) {
                // even-clock sample with no fresh symbol in this round: vote with the carried sum/dif (sync_update :38-42)
                float dd = (sumc < 0) ? -difc : difc;
                if (dd > 0) thr++;
                if (dd < 0) thr--;
                clk = 0;
                sync_adjust(TH, thr, index, clk, m_idx, out, lane, mid);
                i++;
                continue;
            }
            if (index != tap_index) {
#pragma unroll
                for (int k = 0; k < M17B_FN; k++) tp[k] = pack2(__ldg(g_mf + index * M17B_FN + k), __ldg(g_md + index * M17B_FN + k));
                tap_index = index;
            }
            dbg_rounds++;
            if (whole_block && i <= 1) {
                // Locked, at the start of a block: the threshold is 80, so a trip inside the block is rare -- speculate the
                // whole block in ONE round, six consecutive symbols per lane (192 symbols, all inside the block for i <= 1).
                float s6[6], d6[6], mid6[6];
                {
                    float xw[M17B_FN + 10];
                    if (i == 0) load_window<0>(sm.x, 3 * lane, xw); else load_window<1>(sm.x, 3 * lane, xw);
                    dot6_core(xw, tp, one, s6, d6);
                    if (EQ) {
                        // xw[0] is sample n0 = 12 lane + i of the staged array; the one before it (none for the block's first sample:
                        // that pair's half-symbol value was computed at the end of the previous block)
                        const float xm1 = i == 1 ? sm.x[0][3 * lane] : (lane > 0 ? sm.x[3][3 * lane - 1] : 0.0f);
#pragma unroll
                        for (int m = 0; m < 6; m++) mid6[m] = dot_mid(xm1, xw, 2 * m, tp);
                        if (i == 0 && lane == 0) mid6[0] = mid_carry;
                    }
                }
                // votes (sync_update, m17_rx_sync.cpp:38-42): every symbol votes on the NEXT sample, so the only symbol of the
                // block without a vote is the one at sample 383 (i == 1, lane 31, m == 5).  The common case needs no per-symbol
                // threshold values: the lane keeps the running sum and its prefix extremes, the warp scan supplies the offset,
                // and a trip exists iff some prefix leaves [-TH, TH].
                int vt[6], run = 0, pmax = -8, pmin = 8;
#pragma unroll
                for (int m = 0; m < 6; m++) {
                    const float dd = (s6[m] < 0) ? -d6[m] : d6[m];
                    int v = (dd > 0) - (dd < 0);
                    if (m == 5 && i == 1 && lane == 31) v = 0;
                    vt[m] = v;
                    run += v;
                    pmax = max(pmax, run); pmin = min(pmin, run);
                }
                // exclusive prefix of the lanes' sums without a shuffle chain: run + 6 is a 4-bit number, one ballot per bit
                const unsigned e6 = (unsigned)(run + 6);
                const unsigned q0 = __ballot_sync(0xffffffffu, e6 & 1u), q1 = __ballot_sync(0xffffffffu, e6 & 2u);
                const unsigned q2 = __ballot_sync(0xffffffffu, e6 & 4u), q3 = __ballot_sync(0xffffffffu, e6 & 8u);
                const unsigned lt = (1u << lane) - 1u;
                const int off = thr + __popc(q0 & lt) + 2 * __popc(q1 & lt) + 4 * __popc(q2 & lt) + 8 * __popc(q3 & lt) - 6 * lane;
                int fm = 6;                                                    // first symbol of this lane whose vote trips
                int th6[6];
                if (__any_sync(0xffffffffu, off + pmax > TH || off + pmin < -TH)) {
                    int r2 = off;
#pragma unroll
                    for (int m = 0; m < 6; m++) { r2 += vt[m]; th6[m] = r2; }
#pragma unroll
                    for (int m = 5; m >= 0; m--) {
                        const int j = i + 2 * (6 * lane + m);
                        if ((j + 1 < 384) && (th6[m] > TH || th6[m] < -TH)) fm = m;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 6; m++) th6[m] = 0;
                }
                const unsigned trip = __ballot_sync(0xffffffffu, fm < 6);
                if (!trip) {
#pragma unroll
                    for (int m = 0; m < 6; m++) if (m_idx + 6 * lane + m >= 0) { out[m_idx + 6 * lane + m] = s6[m]; if (EQ) mid[m_idx + 6 * lane + m] = mid6[m]; }
                    m_idx += 192;
                    thr = thr + __popc(q0) + 2 * __popc(q1) + 4 * __popc(q2) + 8 * __popc(q3) - 192;
                    sumc = __shfl_sync(0xffffffffu, s6[5], 31);
                    difc = __shfl_sync(0xffffffffu, d6[5], 31);
                    if (i == 0) clk = 0; else clk = 1;                         // last symbol at sample 382 (vote at 383) / 383 (vote in the next block)
                    i = 384;
                } else {
                    const int L = __ffs(trip) - 1;
                    const int fmL = __shfl_sync(0xffffffffu, fm, L);
                    const int P = 6 * L + fmL;
#pragma unroll
                    for (int m = 0; m < 6; m++) if (6 * lane + m <= P && m_idx + 6 * lane + m >= 0) { out[m_idx + 6 * lane + m] = s6[m]; if (EQ) mid[m_idx + 6 * lane + m] = mid6[m]; }
                    m_idx += P + 1;
                    int tsel = th6[0]; float ssel = s6[0], dsel = d6[0];
#pragma unroll
                    for (int m = 1; m < 6; m++) if (fmL == m) { tsel = th6[m]; ssel = s6[m]; dsel = d6[m]; }
                    thr = __shfl_sync(0xffffffffu, tsel, L);
                    sumc = __shfl_sync(0xffffffffu, ssel, L);
                    difc = __shfl_sync(0xffffffffu, dsel, L);
                    clk = 0;
                    trips++;
                    __syncwarp();
                    sync_adjust(TH, thr, index, clk, m_idx, out, lane, mid);
                    i = i + 2 * P + 2;
                }
                continue;
            }
            // speculate: lane l computes the symbols at samples ja = i + 4l and jb = ja + 2
            const int ja = i + 4 * lane, jb = ja + 2;
            const bool valid_a = ja < 384, valid_b = jb < 384;
            float sa = 0.0f, da = 0.0f, sb = 0.0f, db = 0.0f, mida = 0.0f, midb = 0.0f;
            if (valid_a) {
                const int base = (i >> 2) + lane;
                float xw[M17B_FN + 2];
                switch (i & 3) {
                    case 0: load_window<0>(sm.x, base, xw); break;
                    case 1: load_window<1>(sm.x, base, xw); break;
                    case 2: load_window<2>(sm.x, base, xw); break;
                    default: load_window<3>(sm.x, base, xw); break;
                }
                dot2_core(xw, tp, one, sa, da, sb, db);
                if (EQ) {
                    const int n0 = ja;                                     // staged index of xw[0]
                    const float xm1 = n0 > 0 ? sm.x[(n0 - 1) & 3][(n0 - 1) >> 2] : 0.0f;
                    mida = dot_mid(xm1, xw, 0, tp);
                    midb = dot_mid(xm1, xw, 2, tp);
                    if (n0 == 0) mida = mid_carry;
                }
            }
            // votes happen on the sample after each symbol (sync_update, m17_rx_sync.cpp:38-42) if it is in this block
            const bool vote_a = valid_a && (ja + 1 < 384), vote_b = valid_b && (jb + 1 < 384);
            int va = 0, vb = 0;
            if (vote_a) { float dd = (sa < 0) ? -da : da; va = (dd > 0) - (dd < 0); }
            if (vote_b) { float dd = (sb < 0) ? -db : db; vb = (dd > 0) - (dd < 0); }
            const unsigned e2 = (unsigned)(va + vb + 2);                       // 0..4: inclusive prefix from one ballot per bit
            const unsigned r0 = __ballot_sync(0xffffffffu, e2 & 1u), r1 = __ballot_sync(0xffffffffu, e2 & 2u), r2 = __ballot_sync(0xffffffffu, e2 & 4u);
            const unsigned le = 0xffffffffu >> (31 - lane);
            const int th_b = thr + __popc(r0 & le) + 2 * __popc(r1 & le) + 4 * __popc(r2 & le) - 2 * (lane + 1), th_a = th_b - vb;
            const unsigned ta = __ballot_sync(0xffffffffu, vote_a && (th_a > TH || th_a < -TH));
            const unsigned tb = __ballot_sync(0xffffffffu, vote_b && (th_b > TH || th_b < -TH));
            const int pa = ta ? 2 * (__ffs(ta) - 1) : 1 << 20, pb = tb ? 2 * (__ffs(tb) - 1) + 1 : 1 << 20;
            const int P = pa < pb ? pa : pb;                                  // first symbol (in stream order) whose vote trips
            if (P >= (1 << 20)) {
                const int nv = __popc(__ballot_sync(0xffffffffu, valid_a)) + __popc(__ballot_sync(0xffffffffu, valid_b));
                if (valid_a && m_idx + 2 * lane >= 0) { out[m_idx + 2 * lane] = sa; if (EQ) mid[m_idx + 2 * lane] = mida; }
                if (valid_b && m_idx + 2 * lane + 1 >= 0) { out[m_idx + 2 * lane + 1] = sb; if (EQ) mid[m_idx + 2 * lane + 1] = midb; }
                m_idx += nv;
                thr = __shfl_sync(0xffffffffu, th_b, 31);
                const int L = (nv - 1) >> 1;
                const float s0 = __shfl_sync(0xffffffffu, sa, L), s1 = __shfl_sync(0xffffffffu, sb, L);
                const float d0 = __shfl_sync(0xffffffffu, da, L), d1 = __shfl_sync(0xffffffffu, db, L);
                sumc = ((nv - 1) & 1) ? s1 : s0;
                difc = ((nv - 1) & 1) ? d1 : d0;
                const int last_j = i + 2 * (nv - 1);
                if (last_j + 1 < 384) { i = last_j + 2; clk = 0; } else { i = 384; clk = 1; }
            } else {
                const int L = P >> 1;
                if (2 * lane <= P && m_idx + 2 * lane >= 0) { out[m_idx + 2 * lane] = sa; if (EQ) mid[m_idx + 2 * lane] = mida; }
                if (2 * lane + 1 <= P && m_idx + 2 * lane + 1 >= 0) { out[m_idx + 2 * lane + 1] = sb; if (EQ) mid[m_idx + 2 * lane + 1] = midb; }
                m_idx += P + 1;
                const int thr0 = __shfl_sync(0xffffffffu, th_a, L), thr1 = __shfl_sync(0xffffffffu, th_b, L);
                const float s0 = __shfl_sync(0xffffffffu, sa, L), s1 = __shfl_sync(0xffffffffu, sb, L);
                const float d0 = __shfl_sync(0xffffffffu, da, L), d1 = __shfl_sync(0xffffffffu, db, L);
                thr = (P & 1) ? thr1 : thr0;
                sumc = (P & 1) ? s1 : s0;
                difc = (P & 1) ? d1 : d0;
                clk = 0;
                trips++;
                __syncwarp();
                sync_adjust(TH, thr, index, clk, m_idx, out, lane, mid);
                i = i + 2 * P + 2;
            }
        }
        const int n = m_idx < 0 ? 0 : m_idx;
        trips_prev = trips;
        __syncwarp();
        if (EQ) {
            // the next block's sample 0 is a symbol instant iff clk == 0 now; its half-symbol companion is this block's last sample
            // through the branch that symbol will use (the index cannot change in between)
            if (clk == 0) {
                const float *mfb = g_mf + index * M17B_FN;
                float acc = sm.x[383 & 3][383 >> 2] * __ldg(mfb);
                for (int k = 1; k < M17B_FN; k++) acc += sm.x[(383 + k) & 3][(383 + k) >> 2] * __ldg(mfb + k);
                mid_carry = acc;
            }
            eq_block(&eqs[c].e, out, mid, n, lane);
            __syncwarp();
        }
        PHASE(1);

        // ---- emit the block's symbols to the channel's stream
        {
            float *dst = sbuf + M17B_SYM_CARRY + (sym_total - base_g);
            if (n == 192) {                                                    // no slip in this block: six independent copies
#pragma unroll
                for (int k = 0; k < 6; k++) dst[lane + 32 * k] = out[lane + 32 * k];
            } else {
                for (int q = lane; q < n; q += 32) dst[q] = out[q];
            }
            const int tl = (int)(t - t0) & 31;
            if (lane == tl) nsym_keep = n;
            if (tl == 31 || t + 1 == t1) { if (lane <= tl) nsym[c * T + (t - tl) + lane] = nsym_keep; }
        }

        PHASE(2);
        // ---- framer (m17_rx_frame.cpp:126-172)
        int p = 0, reset_at = -8;
        bool have_bits = false;
        unsigned negw[7], posw[7];
        if (flock && fclk + n >= M17B_FRAME_SYMS && fclk + n < 2 * M17B_FRAME_SYMS) {
            // The common case while locked, straight-line: the frame being collected completes inside this block (after p1 of its
            // symbols) and the next one takes the rest.  Same steps as the general loop below (m17_rx_frame.cpp:126-157).
            const int p1 = M17B_FRAME_SYMS - fclk;
            float w[8];
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = (k < fclk) ? sm.head[k] : sm.hist[8 + k - fclk];      // m_f_sym[0..7]
            const SyncResult r = sync_check8<true>(w);                                 // (the quotient is only needed for the record)
            const bool ok = sync_accept(r, true);
            int flags = ok ? M17B_F_SYNC_OK : 0, fe;
            bool los = false;
            if (r.type == M17B_T_EOT) { los = true; fe = ferr; }                       // :137-140
            else if (ok) { flags |= M17B_F_PARSED; ferr = 0; fe = 0; }                 // :144-146
            else { ferr++; fe = ferr; if (ferr > 5) los = true; else flags |= M17B_F_PARSED; }   // :147-154
            if (los) flags |= M17B_F_LOS;
            push_record(r.type, flags, r.votes, fe, r.spread, r.vmax);
            const int fclk_in = fclk;
            fclk = 0;
            p = p1;
            frame_start = sym_total + p1;
            __syncwarp();                                                               // every lane has read the old head
            if (los) {
                if (lane < 8 && lane >= fclk_in) sm.head[lane] = sm.hist[8 + lane - fclk_in];   // (the head as the general loop leaves it)
                flock = 0;
                reset_at = p1;                                                          // reset_sync(): window reads as zeros
                if (lane == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + p1 - 1; events[c * ecap + nev].kind = M17B_EV_LOS; }
                nev++; n_los++;
            } else {
                const int rest = n - p1;                                                // < 192: the next frame stays open
                if (lane < 8 && lane < rest) sm.head[lane] = sm.hist[8 + p1 + lane];
                fclk = rest;
                p = n;
                __syncwarp();
            }
        }
        while (p < n) {
            if (!flock) {
                // Unlocked search (m17_rx_frame.cpp:158-170): acceptance needs the window's sign pattern to BE one of the four frame sync
                // words (sync_unlocked_ok, fec.cuh), so the signs of the block's symbols are taken once -- 7 x 2 ballots give the
                // 8 + n sign bits to every lane -- and the window ending at symbol q is 8 consecutive bits of that string.  All
                // positions of the block are screened with a few integer operations each; the variance test runs only on the
                // (rare) windows whose signs match.
                if (!have_bits) {
#pragma unroll
                    for (int k = 0; k < 7; k++) {
                        const int idx = lane + 32 * k;
                        const float v = idx < 8 + n ? sm.hist[idx] : 0.0f;
                        negw[k] = __ballot_sync(0xffffffffu, v < 0);
                        posw[k] = __ballot_sync(0xffffffffu, v > 0);
                    }
                    have_bits = true;
                }
                unsigned mine = 0;                                     // windows end at q = 32 k + lane - 1 (bits q + 1 .. q + 8)
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const int q = 32 * k + lane - 1;
                    const unsigned ng = __funnelshift_r(negw[k], k < 6 ? negw[k + 1] : 0u, lane) & 0xFFu;
                    const unsigned ps = __funnelshift_r(posw[k], k < 6 ? posw[k + 1] : 0u, lane) & 0xFFu;
                    const bool cnd = q >= p && q < n && q - 7 >= reset_at && (ng | ps) == 0xFFu &&
                                     (ng == sync_neg_mask(1) || ng == sync_neg_mask(2) || ng == sync_neg_mask(3) || ng == sync_neg_mask(4));
                    mine |= cnd ? (1u << k) : 0u;
                }
                // every lane tests its own candidates (ascending q; nearly always at most one per lane), the earliest accepted wins
                int myq = 1 << 20;
                while (__any_sync(0xffffffffu, mine != 0u)) {
                    if (mine) {
                        const int k = __ffs(mine) - 1;
                        mine &= mine - 1;
                        if (myq == (1 << 20)) {
                            float w[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) w[j] = sm.hist[32 * k + lane + j];             // hist[8 + q - 7 + j]
                            if (sync_variance_lt03(w)) myq = 32 * k + lane - 1;
                        }
                    }
                }
                const int fq = __reduce_min_sync(0xffffffffu, myq);
                const int found = fq < (1 << 20) ? fq : -1;
                if (found < 0) { p = n; break; }
                // acquisition: copy_sync(), m_fclk = 8 (m17_rx_frame.cpp:161-169)
                if (lane < 8) { int idx = found - 7 + lane; sm.head[lane] = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
                fclk = 8; ferr = 0; flock = 1;
                frame_start = sym_total + found - 7;
                if (lane == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + found; events[c * ecap + nev].kind = M17B_EV_AOS; }
                nev++; n_aos++;
                p = found + 1;
                __syncwarp();
            } else {
                const int need = M17B_FRAME_SYMS - fclk, avail = n - p;
                const int take = need < avail ? need : avail;
                if (fclk < 8 && lane < 8 && lane >= fclk && lane < fclk + take) sm.head[lane] = sm.hist[8 + p + lane - fclk];
                fclk += take;
                p += take;
                __syncwarp();
                if (fclk == M17B_FRAME_SYMS) {
                    fclk = 0;
                    float w[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) w[k] = sm.head[k];
                    const SyncResult r = sync_check8<true>(w);
                    const bool ok = sync_accept(r, true);
                    int flags = ok ? M17B_F_SYNC_OK : 0, fe;
                    bool los = false;
                    if (r.type == M17B_T_EOT) { los = true; fe = ferr; }                       // :137-140
                    else if (ok) { flags |= M17B_F_PARSED; ferr = 0; fe = 0; }                 // :144-146
                    else { ferr++; fe = ferr; if (ferr > 5) los = true; else flags |= M17B_F_PARSED; }   // :147-154
                    if (los) flags |= M17B_F_LOS;
                    push_record(r.type, flags, r.votes, fe, r.spread, r.vmax);
                    if (los) {
                        flock = 0;
                        reset_at = p;                                                           // reset_sync(): window reads as zeros
                        if (lane == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + p - 1; events[c * ecap + nev].kind = M17B_EV_LOS; }
                        nev++; n_los++;
                    }
                    frame_start = sym_total + p;
                    __syncwarp();
                }
            }
        }
        PHASE(3);
        // ---- carry: sliding window = last 8 symbols (zeros before a reset), filter history = last 30 samples
        {
            float wv = 0.0f, a = 0.0f;
            if (lane < 8) { int idx = n - 8 + lane; wv = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
            if (lane < 30) a = sm.x[lane & 3][96 + (lane >> 2)];            // sample 384 + lane -> slot lane
            __syncwarp();
            if (lane < 8) sm.hist[lane] = wv;
            if (lane < 30) sm.x[lane & 3][lane >> 2] = a;
        }
        sym_total += n;
        if (nq > SY_RECQ - 3 || t + 1 == t1) flush_records();       // (the only call site: the kernel's code footprint matters)
        __syncwarp();
        PHASE(4);
    }

    // ---- store state
    if (AFC && lane == 0) {
        S->afc_delta = afc_delta; S->nco_acc = nco_acc;
        S->z0re = afc_z0.x; S->z0im = afc_z0.y; S->z1re = afc_z1.x; S->z1im = afc_z1.y;
    }
    if (lane < 30) S->tail[lane] = sm.x[lane & 3][lane >> 2];
    if (lane < 8) { S->win[lane] = sm.hist[lane]; S->head[lane] = sm.head[lane]; }
    if (EQ && lane == 0) eqs[c].mid = mid_carry;
    if (lane == 0) {
        S->clk = clk; S->thr = thr; S->index = index; S->sum = sumc; S->dif = difc;
        S->flock = flock; S->fclk = fclk; S->ferr = ferr; S->frame_start = frame_start; S->sym_total = sym_total;
        S->prev_n = sym_total - base_g;
        S->dbg_cycles = (unsigned long long)(clock64() - clk_start); S->dbg_rounds = dbg_rounds;
#ifdef M17B_PHASE_CLOCKS
        for (int q5 = 0; q5 < 5; q5++) S->dbg_phase[q5] = (unsigned long long)ph[q5];
        S->dbg_phase[5] = (unsigned long long)ph_blocks;
#endif
        nframes[c] = nfr < fcap ? nfr : (int)fcap;
        nevents[c] = nev < ecap ? nev : (int)ecap;
        if (frame_rng) frame_rng[c] = make_int2(nfr_entry, nfr < fcap ? nfr : (int)fcap);   // records completed by this slice
        unsigned long long *q = stats + c * 8;
        q[0] += (unsigned long long)(nfr - nfr_entry); q[4] += (unsigned long long)n_aos; q[5] += (unsigned long long)n_los;
        q[7] += (unsigned long long)(sym_total - sym_entry);
    }
}
