// sync_xb.cuh -- RX matched filter + symbol-timing loop + framer with TWO WARPS PER CHANNEL that ALTERNATE BLOCKS.
// Same arithmetic and the same in-block speculation as sync.cuh / sync_g.cuh; what changes is the schedule along time:
//   warp r of a channel owns the 40-ms blocks t0 + r, t0 + r + 2, ...  For each of its blocks it first SPECULATES -- stages the
//   block's samples (the 30 samples of filter history are re-read from the previous block's input, so staging needs nothing from
//   the other warp), runs the 2 x 192 31-tap dot products with the polyphase branch / clock phase PREDICTED for the block and
//   turns them into votes and a warp prefix sum -- and only then waits for the other warp to hand over the exact loop and framer
//   state at the block boundary.  With the state in hand it RESOLVES the block: checks the prediction (branch, clock phase, lock
//   flag), adds the carried counter to the vote prefix, tests for a threshold trip, commits the symbols, runs the framer and
//   hands the state on.  A wrong prediction or a trip falls back to the ordinary speculation rounds from the exact state, so the
//   results are identical to the serial order (m17_rx_sync.cpp:77-99, m17_rx_frame.cpp:126-172).
// Why: at <= 1184 channels per GPU the one-warp kernel's time is the latency of each channel's serial chain (~6300 cycles per
// block, profiles/r01b_sync_per_channel.txt): staging + dot products + votes are about half of it and depend on the previous
// block only through (branch, clock phase, lock flag), which change in about one block out of five.  Here they leave the
// chain; what stays on it is hand-off + trip test + commit + emission + framer.
// The 31 tap pairs live in shared memory (broadcast reads) to stay under 128 registers: 16 resident warps per SM.
// Replaces m17_rx_sync_samples (+ rx_sync_filter, sync_update, m17_sync_adjust: m17_rx_sync.cpp:25-99) and
// m17_rx_symbols / m17_rx_sym / m17_sync_check (m17_rx_frame.cpp:47-177).
#pragma once
#include "sync_pc.cuh"

#define XB_CH 2                              // channels per CTA (4 warps)
#define XB_PRE (384 + 4 + 32 + 4)            // [0,384) raw block, [384] its mean, [388,418) raw tail of the block before, [420] that block's mean

struct XbHand {                              // loop + framer state at a block boundary, handed from the warp of block t-1 to the warp of block t
    int clk, thr, index, flock, fclk, ferr, frame_start, sym_total, nfr, nev, n_aos, n_los;
    float sumc, difc;
    float win[8];                            // sliding sync window = last 8 symbols of the previous block (zeros after a reset)
};
struct SyncXbWarp {
    float x[4][SG_XQ];                       // this warp's block: discriminator samples incl. 30 of history, residue-split (sync.cuh)
    float pre[XB_PRE];                       // cp.async landing zone for this warp's NEXT block
    float sym[SY_HIST];                      // [0,8): sliding window carried in; [8, 8+n): symbols of the block
    f32x2 taps[M17B_FN + 1];                 // (matched, derivative) tap pairs of the branch this warp last used
};
struct SyncXbSmem {
    SyncXbWarp w[2];
    float head[8];                           // m_f_sym[0..7] of the frame being collected (touched in resolve phases only)
    XbHand hand;
};

template <bool HAS_MEAN>
__global__ void __launch_bounds__(XB_CH * 64, 4) k_sync_frame_xb(const float *__restrict__ disc, const float *__restrict__ mean, int64_t nchan, int64_t T,
                                                                 int t0, int t1, int2 *frame_rng, RxChanState *st, const float *__restrict__ g_mf,
                                                                 const float *__restrict__ g_md, float *syms, int64_t sym_pitch,
                                                                 int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base, m17b_frame_rec *frames,
                                                                 int64_t fcap, int32_t *__restrict__ nframes, m17b_event_rec *events, int64_t ecap,
                                                                 int32_t *__restrict__ nevents, unsigned long long *stats, int commit_fe,
        const int * /*fe_done*/, int /*fe_slice*/, int * /*fe_err*/) {
    __shared__ __align__(16) SyncXbSmem sm_all[XB_CH];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = wid >> 1, r = wid & 1;
    const int64_t c = (int64_t)blockIdx.x * XB_CH + slot;
    if (c >= nchan || t0 + r >= t1) return;                        // (a warp without blocks never touches a barrier)
    SyncXbSmem &sm = sm_all[slot];
    SyncXbWarp &my = sm.w[r];
    XbHand &H = sm.hand;
    RxChanState *S = st + c;
    float *out = my.sym + 8;
    const int BAR = 1 + 2 * slot;                                  // BAR + ((t - t0) & 1): "the state at the start of block t is in H"
    const unsigned FULL = 0xffffffffu;

    // ---- entry state.  Both warps read what they need for addressing and for their first prediction; warp 0 owns the hand-off
    // record until it passes it on.
    const int sym_entry = S->sym_total;
    const int base_g = t0 == 0 ? sym_entry : sym_base[c];
    const int nfr_entry = t0 == 0 ? 0 : nframes[c];
    int p_index = S->index, p_clk = S->clk, p_flock = S->flock;    // prediction for this warp's next block: "nothing changed"
    float *sbuf = syms + c * sym_pitch;
    if (r == 0) {
        if (commit_fe && lane == 0 && t1 == T) { S->z0re = S->nz0re; S->z0im = S->nz0im; S->z1re = S->nz1re; S->z1im = S->nz1im; }
        if (t0 == 0) {
            // carry: the last 192 symbols of the previous call move in front of the new ones
            const int prev_n = S->prev_n;
            float tmp[6];
#pragma unroll
            for (int k = 0; k < 6; k++) tmp[k] = sbuf[prev_n + lane + 32 * k];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 6; k++) sbuf[lane + 32 * k] = tmp[k];
            if (lane == 0) sym_base[c] = base_g;
        }
        if (lane == 0) {
            H.clk = p_clk; H.thr = S->thr; H.index = p_index; H.flock = p_flock; H.fclk = S->fclk; H.ferr = S->ferr;
            H.frame_start = S->frame_start; H.sym_total = sym_entry; H.nfr = nfr_entry; H.nev = t0 == 0 ? 0 : nevents[c];
            H.n_aos = 0; H.n_los = 0; H.sumc = S->sum; H.difc = S->dif;
        }
        if (lane < 8) { H.win[lane] = S->win[lane]; sm.head[lane] = S->head[lane]; }
        __syncwarp();
    }
#ifdef M17B_PHASE_CLOCKS
    const long long clk_start = clock64();
    long long ph[6] = {0, 0, 0, 0, 0, 0}, pt = clk_start;
#define XBPH(i) do { const long long now__ = clock64(); ph[i] += now__ - pt; pt = now__; } while (0)
#else
#define XBPH(i) do {} while (0)
#endif
    unsigned n_full = 0, n_part = 0, n_miss = 0, n_unl = 0;         // instrumentation: how this warp's blocks were resolved
    int tap_index = -1;
    auto load_taps = [&](int idx) {
        if (lane < M17B_FN) my.taps[lane] = pack2(__ldg(g_mf + idx * M17B_FN + lane), __ldg(g_md + idx * M17B_FN + lane));
        tap_index = idx;
        __syncwarp();
    };
    // this warp's next block, fetched while it works on the current one (cp.async: completion by async group, no register scoreboard)
    auto prefetch = [&](int64_t tt) {
        const float *src = disc + (c * T + tt) * 384;
#pragma unroll
        for (int q = 0; q < 12; q++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&my.pre[lane + 32 * q]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + lane + 32 * q));
        }
        if (HAS_MEAN && lane == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&my.pre[384]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(mean + c * T + tt));
        }
        if (tt > t0) {                                                 // filter history = the last 30 samples of block tt-1 as that block saw them
            if (lane < 30) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&my.pre[388 + lane]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src - 30 + lane));
            }
            if (HAS_MEAN && lane == 0) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&my.pre[420]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(mean + c * T + tt - 1));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    prefetch(t0 + r);

    for (int64_t t = t0 + r; t < t1; t += 2) {
        // =========================================================== SPECULATE (off the channel's serial chain)
        asm volatile("cp.async.wait_group 0;");
        __syncwarp();
        {
            const float pmu = HAS_MEAN ? my.pre[384] : 0.0f;
#pragma unroll
            for (int q = 0; q < 12; q++) {
                float v = my.pre[lane + 32 * q];
                if (HAS_MEAN) v = v - pmu;                                  // m17_dsp.cpp:217-219
                const int n = 30 + lane + 32 * q;
                my.x[n & 3][n >> 2] = v;
            }
            if (lane < 30) {
                float v;
                if (t == t0) v = S->tail[lane];                             // history carried by the channel state
                else { v = my.pre[388 + lane]; if (HAS_MEAN) v = v - my.pre[420]; }
                my.x[lane & 3][lane >> 2] = v;
            }
        }
        __syncwarp();
        if (t + 2 < t1) prefetch(t + 2);
        XBPH(0);
        // whole-block round with the predicted branch and clock phase: symbol 6 lane + m sits at sample i_p + 2 (6 lane + m)
        bool spec = false;
        float s6[6], d6[6];
        int th6[6], run = 0, incl = 0;
        if (p_flock) {
            if (p_index != tap_index) load_taps(p_index);
            if (p_clk == 0) dotn<0, 6>(my.x, 3 * lane, my.taps, s6, d6); else dotn<1, 6>(my.x, 3 * lane, my.taps, s6, d6);
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const int j = p_clk + 2 * (6 * lane + m);
                const float dd = (s6[m] < 0) ? -d6[m] : d6[m];
                if (j + 1 < 384) run += (dd > 0) - (dd < 0);               // the vote happens on the next sample, if it is in this block
                th6[m] = run;
            }
            incl = warp_incl_scan(run, lane);
            spec = true;
        }

        XBPH(1);
        // =========================================================== WAIT for the state at the start of block t
        if (t > t0) pc_bar_sync(BAR + (int)((t - t0) & 1));
        XBPH(2);

        // =========================================================== RESOLVE (the serial chain)
        int clk = H.clk, thr = H.thr, index = H.index, flock = H.flock, fclk = H.fclk, ferr = H.ferr, frame_start = H.frame_start;
        int sym_total = H.sym_total, nfr = H.nfr, nev = H.nev, n_aos = H.n_aos, n_los = H.n_los;
        float sumc = H.sumc, difc = H.difc;
        if (lane < 8) my.sym[lane] = H.win[lane];
        __syncwarp();
        const int TH = flock ? 80 : 10;                                   // m17_rx_lock() is constant inside a block
        int i = 0, m_idx = 0;
        bool use = spec && flock && index == p_index && clk == p_clk;
        if (use && clk == 1) {
            // even-clock sample 0: vote with the carried sum/dif (sync_update :38-42); a trip here invalidates the prediction
            const float dd = (sumc < 0) ? -difc : difc;
            if (dd > 0) thr++;
            if (dd < 0) thr--;
            clk = 0;
            sync_adjust_g(TH, thr, index, clk, m_idx, out, lane);
            i = 1;
            if (index != p_index || clk != 0) use = false;
        }
        if (!flock) n_unl++; else if (!use) n_miss++;
        if (use) {
            const int off = thr + incl - run;
            int fm = 6;                                                    // first symbol of this lane whose vote trips
#pragma unroll
            for (int m = 5; m >= 0; m--) {
                th6[m] += off;
                const int j = i + 2 * (6 * lane + m);
                if ((j + 1 < 384) && (th6[m] > TH || th6[m] < -TH)) fm = m;
            }
            const unsigned trip = __ballot_sync(FULL, fm < 6);
            if (!trip) {
                n_full++;
#pragma unroll
                for (int m = 0; m < 6; m++) out[6 * lane + m] = s6[m];
                m_idx = 192;
                thr = __shfl_sync(FULL, th6[5], 31);
                sumc = __shfl_sync(FULL, s6[5], 31);
                difc = __shfl_sync(FULL, d6[5], 31);
                clk = i;                                                   // last symbol at sample 382 (vote at 383) / 383 (vote in the next block)
                i = 384;
            } else {
                n_part++;
                const int L = __ffs(trip) - 1;
                const int fmL = __shfl_sync(FULL, fm, L);
                const int P = 6 * L + fmL;
#pragma unroll
                for (int m = 0; m < 6; m++) if (6 * lane + m <= P) out[6 * lane + m] = s6[m];
                m_idx = P + 1;
                int tsel = th6[0]; float ssel = s6[0], dsel = d6[0];
#pragma unroll
                for (int m = 1; m < 6; m++) if (fmL == m) { tsel = th6[m]; ssel = s6[m]; dsel = d6[m]; }
                thr = __shfl_sync(FULL, tsel, L);
                sumc = __shfl_sync(FULL, ssel, L);
                difc = __shfl_sync(FULL, dsel, L);
                clk = 0;
                __syncwarp();
                sync_adjust_g(TH, thr, index, clk, m_idx, out, lane);
                i = i + 2 * P + 2;
            }
        }
        // whatever is left of the block (all of it after a wrong prediction or while unlocked): ordinary speculation rounds
        while (i < 384) {
            while (clk == 1 && i < 384) {
                const float dd = (sumc < 0) ? -difc : difc;
                if (dd > 0) thr++;
                if (dd < 0) thr--;
                clk = 0;
                sync_adjust_g(TH, thr, index, clk, m_idx, out, lane);
                i++;
            }
            if (i >= 384) break;
            if (index != tap_index) load_taps(index);
            if (flock) sync_round<32, 6>(FULL, lane, 0, my.x, out, my.taps, TH, i, m_idx, thr, index, clk, sumc, difc);
            else       sync_round<32, 2>(FULL, lane, 0, my.x, out, my.taps, TH, i, m_idx, thr, index, clk, sumc, difc);
        }
        const int n = m_idx < 0 ? 0 : m_idx;
        __syncwarp();
        XBPH(3);

        // ---- emit the block's symbols to the channel's stream
        {
            float *dst = sbuf + M17B_SYM_CARRY + (sym_total - base_g);
            for (int q = lane; q < n; q += 32) dst[q] = out[q];
            if (lane == 0) nsym[c * T + t] = n;
        }

        // ---- framer (m17_rx_frame.cpp:126-172)
        int p = 0, reset_at = -8;
        while (p < n) {
            if (!flock) {
                int found = -1;
                for (int q0 = p; q0 < n && found < 0; q0 += 32) {
                    const int q = q0 + lane;
                    bool ok = false;
                    if (q < n) {
                        float w[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) { int idx = q - 7 + k; w[k] = (idx >= reset_at) ? my.sym[8 + idx] : 0.0f; }
                        ok = sync_unlocked_ok(w);
                    }
                    const unsigned m = __ballot_sync(FULL, ok);
                    if (m) found = q0 + __ffs(m) - 1;
                }
                if (found < 0) { p = n; break; }
                // acquisition: copy_sync(), m_fclk = 8 (m17_rx_frame.cpp:161-169)
                if (lane < 8) { int idx = found - 7 + lane; sm.head[lane] = (idx >= reset_at) ? my.sym[8 + idx] : 0.0f; }
                fclk = 8; ferr = 0; flock = 1;
                frame_start = sym_total + found - 7;
                if (lane == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + found; events[c * ecap + nev].kind = M17B_EV_AOS; }
                nev++; n_aos++;
                p = found + 1;
                __syncwarp();
            } else {
                const int need = M17B_FRAME_SYMS - fclk, avail = n - p;
                const int take = need < avail ? need : avail;
                if (fclk < 8 && lane < 8 && lane >= fclk && lane < fclk + take) sm.head[lane] = my.sym[8 + p + lane - fclk];
                fclk += take;
                p += take;
                __syncwarp();
                if (fclk == M17B_FRAME_SYMS) {
                    fclk = 0;
                    float w[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) w[k] = sm.head[k];
                    const SyncResult rs = sync_check8(w);
                    const bool ok = sync_accept(rs, true);
                    int flags = ok ? M17B_F_SYNC_OK : 0, fe;
                    bool los = false;
                    if (rs.type == M17B_T_EOT) { los = true; fe = ferr; }                      // :137-140
                    else if (ok) { flags |= M17B_F_PARSED; ferr = 0; fe = 0; }                 // :144-146
                    else { ferr++; fe = ferr; if (ferr > 5) los = true; else flags |= M17B_F_PARSED; }   // :147-154
                    if (los) flags |= M17B_F_LOS;
                    if (nfr < fcap && lane < 16) {
                        uint32_t word = 0;
                        if (lane == 0) word = (uint32_t)frame_start;
                        else if (lane == 1) word = (uint32_t)rs.type | ((uint32_t)flags << 8);
                        else if (lane == 11) word = ((uint32_t)rs.votes << 16) | ((uint32_t)fe << 24);
                        else if (lane == 12) word = __float_as_uint(rs.variance);
                        ((uint32_t *)(frames + c * fcap + nfr))[lane] = word;
                    }
                    nfr++;
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
        XBPH(4);
        // ---- state at the end of the block
        float wv = 0.0f;
        if (lane < 8) { int idx = n - 8 + lane; wv = (idx >= reset_at) ? my.sym[8 + idx] : 0.0f; }
        sym_total += n;
        p_index = index; p_clk = clk; p_flock = flock;             // prediction for block t+2: block t+1 changes nothing
        if (t + 1 < t1) {
            // ---- hand over to the warp of block t+1
            if (lane == 0) {
                H.clk = clk; H.thr = thr; H.index = index; H.flock = flock; H.fclk = fclk; H.ferr = ferr; H.frame_start = frame_start;
                H.sym_total = sym_total; H.nfr = nfr; H.nev = nev; H.n_aos = n_aos; H.n_los = n_los; H.sumc = sumc; H.difc = difc;
            }
            if (lane < 8) H.win[lane] = wv;
            __syncwarp();
            pc_bar_arrive(BAR + (int)((t + 1 - t0) & 1));
            XBPH(5);
        } else {
            // ---- last block of the launch: store the channel state
            if (lane < 30) S->tail[lane] = my.x[lane & 3][96 + (lane >> 2)];                   // sample 384 + lane
            if (lane < 8) { S->win[lane] = wv; S->head[lane] = sm.head[lane]; }
            if (lane == 0) {
                S->clk = clk; S->thr = thr; S->index = index; S->sum = sumc; S->dif = difc;
                S->flock = flock; S->fclk = fclk; S->ferr = ferr; S->frame_start = frame_start; S->sym_total = sym_total;
                S->prev_n = sym_total - base_g;
                S->dbg_cycles = 0; S->dbg_rounds = n_full | (n_part << 8) | (n_miss << 16) | (n_unl << 24);   // of this warp's blocks
#ifdef M17B_PHASE_CLOCKS
                S->dbg_cycles = (unsigned long long)(clock64() - clk_start);
                for (int q6 = 0; q6 < 6; q6++) S->dbg_phase[q6] = (unsigned long long)ph[q6];
#endif
                nframes[c] = nfr < fcap ? nfr : (int)fcap;
                nevents[c] = nev < ecap ? nev : (int)ecap;
                if (frame_rng) frame_rng[c] = make_int2(nfr_entry, nfr < fcap ? nfr : (int)fcap);   // records completed by this slice
                unsigned long long *q = stats + c * 8;
                q[0] += (unsigned long long)(nfr - nfr_entry); q[4] += (unsigned long long)n_aos; q[5] += (unsigned long long)n_los;
                q[7] += (unsigned long long)(sym_total - sym_entry);
            }
        }
    }
}
