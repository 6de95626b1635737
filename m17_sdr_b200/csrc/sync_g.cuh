// sync_g.cuh -- RX matched filter + symbol-timing loop + sync-word correlator / framer with G LANES PER CHANNEL
// (G = 32, 16 or 8; a warp serves 32/G channels).  Same arithmetic and the same speculation scheme as sync.cuh (see there),
// generalised in two ways:
//   * a speculation round covers NSL consecutive symbols per lane: NSL = 2 while unlocked (threshold 10: trips are frequent,
//     short rounds waste little), NSL = 6 / 12 while locked (threshold 80: a whole 40-ms block in one or two rounds);
//   * all warp-collective steps (vote prefix sum, first-trip ballot, sync-word search ballot, broadcasts) run on the G-lane
//     group's own member mask, so the 32/G channels of a warp are independent mini-warps that simply share an instruction
//     stream while they follow the same path.
// Why: with one warp per channel the kernel is bound by the LATENCY of its serial scalar code -- loop control, framer FSM,
// record writes are ~60 % of the instructions and ~80 % of the time (profiles/), executed by 32 lanes for the benefit of
// one channel, on ~2 warps per scheduler at 1024 channels.  Packing 2 or 4 channels into a warp lets that code serve 2 or 4
// channels per issued instruction; the dot products (6 or 12 or 24 symbols per lane per block) cost the same per channel.
// Replaces m17_rx_sync_samples (+ rx_sync_filter, sync_update, m17_sync_adjust: m17_rx_sync.cpp:25-99) and
// m17_rx_symbols / m17_rx_sym / m17_sync_check (m17_rx_frame.cpp:47-177).
#pragma once
#include "sync_cta.cuh"

#define SG_XQ 124             // (30 + 384 + 2 * 12 * 2 + pad) / 4 entries per residue class: windows of invalid symbols stay in bounds

struct SyncGroupSmem {
    float x[4][SG_XQ];                  // discriminator samples incl. 30 of history: sample n at x[n & 3][n >> 2]
    float hist[SY_HIST];                // [0,8): sliding sync window carried in; [8, 8+n): symbols emitted in this block
    float head[8];                      // m_f_sym[0..7] of the frame being collected
    float pre[2][384 + 4];              // cp.async landing zone for the NEXT block's raw samples (+ its mean), double buffered
    f32x2 taps[M17B_FN + 1];            // TAPS_SMEM variants: (matched, derivative) tap pairs of the current polyphase branch
};

template <int G>
__device__ __forceinline__ int group_incl_scan(unsigned gmask, int v, int gl) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) { int o = __shfl_up_sync(gmask, v, d, G); if (gl >= d) v += o; }
    return v;
}

// NSL consecutive symbols for one lane: windows xs[n0 + 2m .. n0 + 2m + 30], m = 0..NSL-1.  R = i & 3 is uniform in the group.
template <int R, int NSL>
__device__ __forceinline__ void dotn(const float (*X)[SG_XQ], int base, const f32x2 *tp, f32x2 one, float (&s)[NSL], float (&d)[NSL]) {
    float x[M17B_FN + 2 * NSL - 2];
#pragma unroll
    for (int k = 0; k < M17B_FN + 2 * NSL - 2; k++) x[k] = X[(R + k) & 3][base + ((R + k) >> 2)];
    // tp[k] = (matched tap k, derivative tap k): FMUL2 with the sample broadcast = both rounded products, FFMA2 (product * 1.0 +
    // running pair) = both running sums, each half rounding as sum += in[i]*c[i] does (m17_rx_sync.cpp:25-31; see sync.cuh dot2)
    f32x2 acc[NSL];
#pragma unroll
    for (int m = 0; m < NSL; m++) acc[m] = mul2(tp[0], pack2(x[2 * m], x[2 * m]));
#pragma unroll
    for (int k = 1; k < M17B_FN; k++) {
#pragma unroll
        for (int m = 0; m < NSL; m++) acc[m] = fma2(mul2(tp[k], pack2(x[2 * m + k], x[2 * m + k])), one, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < NSL; m++) unpack2(acc[m], s[m], d[m]);
}

// m17_sync_adjust (m17_rx_sync.cpp:45-72).  clk is the value m_clk has before the NEXT sample is processed.
__device__ __forceinline__ void sync_adjust_g(int TH, int &thr, int &index, int &clk, int &m_idx, float *out, int gl) {
    if (thr > TH) {
        index = (index + 1 == M17B_NF) ? 0 : index + 1;
        thr = 0;
        if (index == 0) { clk = 1; if (m_idx >= 0 && gl == 0) out[m_idx] = 0.0f; m_idx++; }
    }
    if (thr < -TH) {
        thr = 0;
        index = (index == 0) ? M17B_NF - 1 : index - 1;
        if (index == M17B_NF - 1) { clk = 1; m_idx--; }
    }
}

// One speculation round: lane gl computes the NSL symbols at samples i + 2 (NSL gl + m); commits up to the first threshold
// trip (or everything); updates the loop state.  All G lanes of the group call it together.
template <int G, int NSL>
__device__ __forceinline__ void sync_round(unsigned gmask, int gl, int gshift, const float (*X)[SG_XQ], float *out, const f32x2 *tp, f32x2 one, int TH,
                                           int &i, int &m_idx, int &thr, int &index, int &clk, float &sumc, float &difc) {
    float s[NSL], d[NSL];
    const int q0 = NSL * gl;                                       // first symbol of this lane within the round
    const int j0 = i + 2 * q0;
#pragma unroll
    for (int m = 0; m < NSL; m++) { s[m] = 0.0f; d[m] = 0.0f; }
    if (j0 < 384) {
        const int n0 = j0;                                         // window start in history coordinates (sample j sits at n = 30 + j)
        const int base = n0 >> 2;
        switch (n0 & 3) {
            case 0: dotn<0, NSL>(X, base, tp, one, s, d); break;
            case 1: dotn<1, NSL>(X, base, tp, one, s, d); break;
            case 2: dotn<2, NSL>(X, base, tp, one, s, d); break;
            default: dotn<3, NSL>(X, base, tp, one, s, d); break;
        }
    }
    // votes happen on the sample after each symbol (sync_update, m17_rx_sync.cpp:38-42) if it is in this block
    int th[NSL], run = 0;
#pragma unroll
    for (int m = 0; m < NSL; m++) {
        const int j = j0 + 2 * m;
        const float dd = (s[m] < 0) ? -d[m] : d[m];
        if (j + 1 < 384) run += (dd > 0) - (dd < 0);
        th[m] = run;
    }
    const int incl = group_incl_scan<G>(gmask, run, gl);
    const int off = thr + incl - run;
    int fm = NSL;                                                  // first symbol of this lane whose vote trips the threshold
#pragma unroll
    for (int m = NSL - 1; m >= 0; m--) {
        th[m] += off;
        const int j = j0 + 2 * m;
        if ((j + 1 < 384) && (th[m] > TH || th[m] < -TH)) fm = m;
    }
    const unsigned trip = (__ballot_sync(gmask, fm < NSL) & gmask) >> gshift;
    if (!trip) {
        // number of symbols of the round that lie inside the block: samples i, i+2, .. < 384
        const int rem = (385 - i) >> 1;
        const int nv = rem < G * NSL ? rem : G * NSL;
#pragma unroll
        for (int m = 0; m < NSL; m++) if (q0 + m < nv && m_idx + q0 + m >= 0) out[m_idx + q0 + m] = s[m];
        m_idx += nv;
        // state after the last committed symbol: its sum/dif (for a vote that falls into the next round / block) and the counter
        const int last = nv - 1, L = last / NSL, lm = last - NSL * L;
        int tsel = th[0]; float ssel = s[0], dsel = d[0];
#pragma unroll
        for (int m = 1; m < NSL; m++) if (lm == m) { tsel = th[m]; ssel = s[m]; dsel = d[m]; }
        thr = __shfl_sync(gmask, tsel, L, G);
        sumc = __shfl_sync(gmask, ssel, L, G);
        difc = __shfl_sync(gmask, dsel, L, G);
        const int last_j = i + 2 * last;
        if (last_j + 1 < 384) { i = last_j + 2; clk = 0; } else { i = 384; clk = 1; }
    } else {
        const int L = __ffs(trip) - 1;
        const int fmL = __shfl_sync(gmask, fm, L, G);
        const int P = NSL * L + fmL;                               // first symbol (in stream order) whose vote trips
#pragma unroll
        for (int m = 0; m < NSL; m++) if (q0 + m <= P && m_idx + q0 + m >= 0) out[m_idx + q0 + m] = s[m];
        m_idx += P + 1;
        int tsel = th[0]; float ssel = s[0], dsel = d[0];
#pragma unroll
        for (int m = 1; m < NSL; m++) if (fmL == m) { tsel = th[m]; ssel = s[m]; dsel = d[m]; }
        thr = __shfl_sync(gmask, tsel, L, G);
        sumc = __shfl_sync(gmask, ssel, L, G);
        difc = __shfl_sync(gmask, dsel, L, G);
        clk = 0;
        __syncwarp(gmask);
        sync_adjust_g(TH, thr, index, clk, m_idx, out, gl);
        i = i + 2 * P + 2;
    }
}

// TAPS_SMEM: the 31 tap pairs are read from shared memory (broadcast) instead of living in 62 registers: ~110 registers instead
// of ~170, i.e. 16 instead of 8 resident warps per SM.  No help while every channel is resident anyway (<= 1184 channels per
// GPU: the chain's latency rules), but above that the kernel runs in waves and twice the warps hide twice the latency.
template <bool HAS_MEAN, int G, bool TAPS_SMEM = false>
__global__ void __launch_bounds__(SY_WARPS * 32, TAPS_SMEM ? 4 : 2) k_sync_frame_g(const float *__restrict__ disc, const float *__restrict__ mean, int64_t nchan, int64_t T,
                                                                   int t0, int t1, int2 *frame_rng, RxChanState *st, const float *__restrict__ g_mf,
                                                                   const float *__restrict__ g_md, float *syms, int64_t sym_pitch,
                                                                   int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base, m17b_frame_rec *frames,
                                                                   int64_t fcap, int32_t *__restrict__ nframes, m17b_event_rec *events, int64_t ecap,
                                                                   int32_t *__restrict__ nevents, unsigned long long *stats, int commit_fe, f32x2 one) {
    constexpr int CPW = 32 / G;                                    // channels per warp
    constexpr int NSL_LOCKED = (G == 32) ? 6 : 12;                 // symbols per lane and round while locked
    extern __shared__ __align__(16) unsigned char sg_smem_raw[];
    SyncGroupSmem *sm_all = (SyncGroupSmem *)sg_smem_raw;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / G, gl = lane % G, gshift = sub * G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
    const int64_t c = ((int64_t)blockIdx.x * SY_WARPS + wid) * CPW + sub;
    // the groups of a warp are independent, but they only share issue slots while they are CONVERGED: after data-dependent
    // loops the warp is re-joined explicitly (all live lanes pass these points once per block)
    const unsigned wmask = __ballot_sync(0xffffffffu, c < nchan);
    if (c >= nchan) return;
    SyncGroupSmem &sm = sm_all[wid * CPW + sub];
    RxChanState *S = st + c;
    float *out = sm.hist + 8;

    // ---- load state (uniform loads within the group)
    // blocks [t0, t1) of a call of T blocks: t0 = 0 starts the call (symbol carry, record / event counts from zero),
    // t0 > 0 appends to what the earlier slices of the same call produced
    if (commit_fe && gl == 0 && t1 == T) { S->z0re = S->nz0re; S->z0im = S->nz0im; S->z1re = S->nz1re; S->z1im = S->nz1im; }
    int clk = S->clk, thr = S->thr, index = S->index;
    float sumc = S->sum, difc = S->dif;
    int flock = S->flock, fclk = S->fclk, ferr = S->ferr, frame_start = S->frame_start, sym_total = S->sym_total;
    const int base_g = t0 == 0 ? sym_total : sym_base[c];
    const int sym_entry = sym_total;
    for (int k = gl; k < 30; k += G) sm.x[k & 3][k >> 2] = S->tail[k];
    if (gl < 8) { sm.hist[gl] = S->win[gl]; sm.head[gl] = S->head[gl]; }
    // carry: the last 192 symbols of the previous call move in front of the new ones
    float *sbuf = syms + c * sym_pitch;
    if (t0 == 0) {
        const int prev_n = S->prev_n;
        float tmp[192 / G];
#pragma unroll
        for (int k = 0; k < 192 / G; k++) tmp[k] = sbuf[prev_n + gl + G * k];     // = sbuf[CARRY + prev_n - 192 + idx]
        __syncwarp(gmask);
#pragma unroll
        for (int k = 0; k < 192 / G; k++) sbuf[gl + G * k] = tmp[k];
    }
    if (gl == 0 && t0 == 0) sym_base[c] = base_g;
    int nfr = t0 == 0 ? 0 : nframes[c], nev = t0 == 0 ? 0 : nevents[c], n_aos = 0, n_los = 0;
    const int nfr_entry = nfr;
    f32x2 tp_reg[TAPS_SMEM ? 1 : M17B_FN];  // (matched, derivative) tap pairs of the current polyphase branch
    const f32x2 *tp = TAPS_SMEM ? sm.taps : tp_reg;
    int tap_index = -1;
    __syncwarp(gmask);
    // The block's samples are fetched one block ahead with cp.async straight into shared memory: completion is tracked by
    // the async-copy group, not by a register scoreboard, so nothing in the timing loop ever waits on the DRAM latency.
    auto prefetch = [&](int64_t tt, int buf) {
        const float *src = disc + (c * T + tt) * 384;
#pragma unroll
        for (int q = 0; q < 384 / G; q++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][gl + G * q]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + gl + G * q));
        }
        if (HAS_MEAN && gl == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][384]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(mean + c * T + tt));
        }
        asm volatile("cp.async.commit_group;");
    };
    prefetch(t0, 0);

    for (int64_t t = t0; t < t1; t++) {
        // ---- stage the block's 384 discriminator samples behind the 30 of history
        const int buf = (int)((t - t0) & 1);
        asm volatile("cp.async.wait_group 0;");
        __syncwarp(wmask);
        {
            const float pmu = !HAS_MEAN ? 0.0f : sm.pre[buf][384];
#pragma unroll
            for (int q = 0; q < 384 / G; q++) {
                float v = sm.pre[buf][gl + G * q];
                if (HAS_MEAN) v = v - pmu;                                  // m17_dsp.cpp:217-219
                const int n = 30 + gl + G * q;
                sm.x[n & 3][n >> 2] = v;
            }
        }
        if (t + 1 < t1) prefetch(t + 1, buf ^ 1);
        __syncwarp(gmask);

        // ---- timing loop (m17_rx_sync.cpp:77-99); m17_rx_lock() is constant inside a block
        const int TH = flock ? 80 : 10;
        int i = 0, m_idx = 0;
        while (i < 384) {
            while (clk == 1 && i < 384) {
                // even-clock sample with no fresh symbol in this round: vote with the carried sum/dif (sync_update :38-42)
                float dd = (sumc < 0) ? -difc : difc;
                if (dd > 0) thr++;
                if (dd < 0) thr--;
                clk = 0;
                sync_adjust_g(TH, thr, index, clk, m_idx, out, gl);
                i++;
            }
            if (i >= 384) break;
            if (index != tap_index) {
                if (TAPS_SMEM) {
                    for (int k = gl; k < M17B_FN; k += G) sm.taps[k] = pack2(__ldg(g_mf + index * M17B_FN + k), __ldg(g_md + index * M17B_FN + k));
                    __syncwarp(gmask);
                } else {
#pragma unroll
                    for (int k = 0; k < M17B_FN; k++) tp_reg[k] = pack2(__ldg(g_mf + index * M17B_FN + k), __ldg(g_md + index * M17B_FN + k));
                }
                tap_index = index;
            }
            if (flock) sync_round<G, NSL_LOCKED>(gmask, gl, gshift, sm.x, out, tp, one, TH, i, m_idx, thr, index, clk, sumc, difc);
            else       sync_round<G, 2>(gmask, gl, gshift, sm.x, out, tp, one, TH, i, m_idx, thr, index, clk, sumc, difc);
        }
        const int n = m_idx < 0 ? 0 : m_idx;
        __syncwarp(wmask);

        // ---- emit the block's symbols to the channel's stream
        {
            float *dst = sbuf + M17B_SYM_CARRY + (sym_total - base_g);
            for (int q = gl; q < n; q += G) dst[q] = out[q];
            if (gl == 0) nsym[c * T + t] = n;
        }

        // ---- framer (m17_rx_frame.cpp:126-172)
        int p = 0, reset_at = -8;
        while (p < n) {
            if (!flock) {
                int found = -1;
                for (int q0 = p; q0 < n && found < 0; q0 += G) {
                    const int q = q0 + gl;
                    bool ok = false;
                    if (q < n) {
                        float w[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) { int idx = q - 7 + k; w[k] = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
                        ok = sync_unlocked_ok(w);
                    }
                    const unsigned m = (__ballot_sync(gmask, ok) & gmask) >> gshift;
                    if (m) found = q0 + __ffs(m) - 1;
                }
                if (found < 0) { p = n; break; }
                // acquisition: copy_sync(), m_fclk = 8 (m17_rx_frame.cpp:161-169)
                if (gl < 8) { int idx = found - 7 + gl; sm.head[gl] = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
                fclk = 8; ferr = 0; flock = 1;
                frame_start = sym_total + found - 7;
                if (gl == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + found; events[c * ecap + nev].kind = M17B_EV_AOS; }
                nev++; n_aos++;
                p = found + 1;
                __syncwarp(gmask);
            } else {
                const int need = M17B_FRAME_SYMS - fclk, avail = n - p;
                const int take = need < avail ? need : avail;
                if (fclk < 8 && gl < 8 && gl >= fclk && gl < fclk + take) sm.head[gl] = sm.hist[8 + p + gl - fclk];
                fclk += take;
                p += take;
                __syncwarp(gmask);
                if (fclk == M17B_FRAME_SYMS) {
                    fclk = 0;
                    float w[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) w[k] = sm.head[k];
                    const SyncResult r = sync_check8(w);
                    const bool ok = sync_accept(r, true);
                    int flags = ok ? M17B_F_SYNC_OK : 0, fe;
                    bool los = false;
                    if (r.type == M17B_T_EOT) { los = true; fe = ferr; }                       // :137-140
                    else if (ok) { flags |= M17B_F_PARSED; ferr = 0; fe = 0; }                 // :144-146
                    else { ferr++; fe = ferr; if (ferr > 5) los = true; else flags |= M17B_F_PARSED; }   // :147-154
                    if (los) flags |= M17B_F_LOS;
                    if (nfr < fcap) {
                        for (int wd = gl; wd < 16; wd += G) {
                            uint32_t word = 0;
                            if (wd == 0) word = (uint32_t)frame_start;
                            else if (wd == 1) word = (uint32_t)r.type | ((uint32_t)flags << 8);
                            else if (wd == 11) word = ((uint32_t)r.votes << 16) | ((uint32_t)fe << 24);
                            else if (wd == 12) word = __float_as_uint(r.variance);
                            ((uint32_t *)(frames + c * fcap + nfr))[wd] = word;
                        }
                    }
                    nfr++;
                    if (los) {
                        flock = 0;
                        reset_at = p;                                                           // reset_sync(): window reads as zeros
                        if (gl == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + p - 1; events[c * ecap + nev].kind = M17B_EV_LOS; }
                        nev++; n_los++;
                    }
                    frame_start = sym_total + p;
                    __syncwarp(gmask);
                }
            }
        }
        // ---- carry: sliding window = last 8 symbols (zeros before a reset), filter history = last 30 samples
        {
            float wv = 0.0f, a[(30 + G - 1) / G];
            if (gl < 8) { int idx = n - 8 + gl; wv = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
#pragma unroll
            for (int k = 0; k < (30 + G - 1) / G; k++) { const int l = gl + G * k; a[k] = (l < 30) ? sm.x[l & 3][96 + (l >> 2)] : 0.0f; }   // sample 384 + l -> slot l
            __syncwarp(gmask);
            if (gl < 8) sm.hist[gl] = wv;
#pragma unroll
            for (int k = 0; k < (30 + G - 1) / G; k++) { const int l = gl + G * k; if (l < 30) sm.x[l & 3][l >> 2] = a[k]; }
        }
        sym_total += n;
        __syncwarp(wmask);
    }

    // ---- store state
    for (int k = gl; k < 30; k += G) S->tail[k] = sm.x[k & 3][k >> 2];
    if (gl < 8) { S->win[gl] = sm.hist[gl]; S->head[gl] = sm.head[gl]; }
    if (gl == 0) {
        S->clk = clk; S->thr = thr; S->index = index; S->sum = sumc; S->dif = difc;
        S->flock = flock; S->fclk = fclk; S->ferr = ferr; S->frame_start = frame_start; S->sym_total = sym_total;
        S->prev_n = sym_total - base_g;
        nframes[c] = nfr < fcap ? nfr : (int)fcap;
        nevents[c] = nev < ecap ? nev : (int)ecap;
        if (frame_rng) frame_rng[c] = make_int2(nfr_entry, nfr < fcap ? nfr : (int)fcap);   // records completed by this slice
        unsigned long long *q = stats + c * 8;
        q[0] += (unsigned long long)(nfr - nfr_entry); q[4] += (unsigned long long)n_aos; q[5] += (unsigned long long)n_los;
        q[7] += (unsigned long long)(sym_total - sym_entry);
    }
}
