// sync_pc.cuh -- RX matched filter + symbol-timing loop + framer with TWO WARPS PER CHANNEL in a producer / consumer pair.
// Same arithmetic, same speculation scheme (sync.cuh, sync_g.cuh); what changes is who runs what:
//   warp A (producer): stages the block's samples, runs the timing loop (matched + derivative filters, votes, threshold
//                      trips) and leaves the block's symbols in one of two shared-memory buffers;
//   warp B (consumer): copies the symbols to the channel's stream, runs the sync-word correlator / framer FSM on them and
//                      writes records and events.
// A works on block t+1 while B works on block t; they meet once per block at two named barriers (FULL: A -> B, DONE: B -> A).
// Why: with 1024 channels the one-warp kernel is bound by the latency of its serial instruction stream (~1400 instructions
// per block at ~0.2 IPC, profiles/), of which the timing loop is ~55 % and emission + framer ~45 %; splitting the stream
// across two warps shortens the per-block chain to the longer half and doubles the warps available to the schedulers.
// The only thing A needs from B is the framer's lock flag at the block boundary (it selects the loop threshold 10 / 80,
// m17_rx_sync.cpp:91-94).  A SPECULATES that the flag did not change in the block B is still working on, and re-runs its
// block from the saved loop state in the rare case (acquisition, loss) that it did -- results are identical to the serial order.
// The 31 tap pairs of the current polyphase branch live in shared memory (broadcast reads) to keep A under 128 registers.
// Replaces m17_rx_sync_samples (+ rx_sync_filter, sync_update, m17_sync_adjust: m17_rx_sync.cpp:25-99) and
// m17_rx_symbols / m17_rx_sym / m17_sync_check (m17_rx_frame.cpp:47-177).
#pragma once
#include "sync_g.cuh"

#define PC_CH 2                              // channels per CTA (4 warps)

struct SyncPcSmem {
    float x[4][SG_XQ];                       // A: discriminator samples incl. 30 of history, residue-split (sync.cuh)
    float pre[2][384 + 4];                   // A: cp.async landing zone for the next block's samples (+ mean)
    float sym[2][8 + 208];                   // [0,8): sliding sync window carried in (B); [8, 8+n): the block's symbols (A)
    f32x2 taps[M17B_FN + 1];                 // A: (matched, derivative) tap pairs of the current polyphase branch
    float head[8];                           // B: m_f_sym[0..7] of the frame being collected
    int n[2];                                // A -> B: symbols in sym[buf]
    int flock_pub;                           // B -> A: framer lock flag after the last block B finished
};

__device__ __forceinline__ void pc_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// (no __threadfence_block here: a memory fence would also wait for the warp's outstanding GLOBAL traffic -- A's cp.async prefetch,
//  B's symbol / record stores -- once per block; the barrier itself orders the shared-memory accesses of its participants)
__device__ __forceinline__ void pc_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <bool HAS_MEAN>
__global__ void __launch_bounds__(PC_CH * 64, 4) k_sync_frame_pc(const float *__restrict__ disc, const float *__restrict__ mean, int64_t nchan, int64_t T,
                                                                 int t0, int t1, int2 *frame_rng, RxChanState *st, const float *__restrict__ g_mf,
                                                                 const float *__restrict__ g_md, float *syms, int64_t sym_pitch,
                                                                 int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base, m17b_frame_rec *frames,
                                                                 int64_t fcap, int32_t *__restrict__ nframes, m17b_event_rec *events, int64_t ecap,
                                                                 int32_t *__restrict__ nevents, unsigned long long *stats, int commit_fe,
        const int * /*fe_done*/, int /*fe_slice*/, int * /*fe_err*/) {
    __shared__ __align__(16) SyncPcSmem sm_all[PC_CH];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warps 2s, 2s+1 serve channel slot s.  Warp w of a CTA runs on scheduler w % 4, so the heavy A role alternates between
    // even and odd warps from one CTA to the next: every scheduler gets A and B warps instead of two schedulers getting all A's
    const int slot = wid >> 1, role = (wid + blockIdx.x) & 1;
    const int64_t c = (int64_t)blockIdx.x * PC_CH + slot;
    if (c >= nchan) return;                                        // both warps of the pair leave together
    SyncPcSmem &sm = sm_all[slot];
    RxChanState *S = st + c;
    const int BAR_FULL = 1 + 2 * slot, BAR_DONE = 2 + 2 * slot;

    if (role == 0) {
        // =========================================================== A: timing loop
        if (commit_fe && lane == 0 && t1 == T) { S->z0re = S->nz0re; S->z0im = S->nz0im; S->z1re = S->nz1re; S->z1im = S->nz1im; }
        int clk = S->clk, thr = S->thr, index = S->index;
        float sumc = S->sum, difc = S->dif;
        int flock = S->flock;                                      // the flag as A knows it
        if (lane < 30) sm.x[lane & 3][lane >> 2] = S->tail[lane];
        int tap_index = -1;
        __syncwarp();
#ifdef M17B_PHASE_CLOCKS
        long long ph[4] = {0, 0, 0, 0}, pt = clock64();
#define PCPH(i) do { const long long now__ = clock64(); ph[i] += now__ - pt; pt = now__; } while (0)
#else
#define PCPH(i) do {} while (0)
#endif
        auto prefetch = [&](int64_t tt, int buf) {
            const float *src = disc + (c * T + tt) * 384;
#pragma unroll
            for (int q = 0; q < 12; q++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][lane + 32 * q]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + lane + 32 * q));
            }
            if (HAS_MEAN && lane == 0) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.pre[buf][384]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(mean + c * T + tt));
            }
            asm volatile("cp.async.commit_group;");
        };
        prefetch(t0, 0);
        for (int64_t t = t0; t < t1; t++) {
            const int buf = (int)((t - t0) & 1);
            asm volatile("cp.async.wait_group 0;");
            __syncwarp();
            {
                const float pmu = HAS_MEAN ? sm.pre[buf][384] : 0.0f;
#pragma unroll
                for (int q = 0; q < 12; q++) {
                    float v = sm.pre[buf][lane + 32 * q];
                    if (HAS_MEAN) v = v - pmu;                              // m17_dsp.cpp:217-219
                    const int n = 30 + lane + 32 * q;
                    sm.x[n & 3][n >> 2] = v;
                }
            }
            if (t + 1 < t1) prefetch(t + 1, buf ^ 1);
            __syncwarp();
            PCPH(0);
            float *out = sm.sym[buf] + 8;
            const int s_clk = clk, s_thr = thr, s_index = index;
            const float s_sum = sumc, s_dif = difc;
            int n = 0;
            for (int attempt = 0; attempt < 2; attempt++) {
                // ---- timing loop (m17_rx_sync.cpp:77-99); m17_rx_lock() is constant inside a block
                const int TH = flock ? 80 : 10;
                int i = 0, m_idx = 0;
                while (i < 384) {
                    if (clk == 1) {
                        // even-clock sample with no fresh symbol: vote with the carried sum/dif (sync_update :38-42)
                        float dd = (sumc < 0) ? -difc : difc;
                        if (dd > 0) thr++;
                        if (dd < 0) thr--;
                        clk = 0;
                        sync_adjust_g(TH, thr, index, clk, m_idx, out, lane);
                        i++;
                        continue;
                    }
                    if (index != tap_index) {
                        if (lane < M17B_FN) sm.taps[lane] = pack2(__ldg(g_mf + index * M17B_FN + lane), __ldg(g_md + index * M17B_FN + lane));
                        tap_index = index;
                        __syncwarp();
                    }
                    if (flock) sync_round<32, 6>(0xffffffffu, lane, 0, sm.x, out, sm.taps, TH, i, m_idx, thr, index, clk, sumc, difc);
                    else       sync_round<32, 2>(0xffffffffu, lane, 0, sm.x, out, sm.taps, TH, i, m_idx, thr, index, clk, sumc, difc);
                }
                n = m_idx < 0 ? 0 : m_idx;
                if (attempt == 0) {
                    PCPH(1);
                    pc_bar_sync(BAR_DONE);
                    PCPH(2);                                  // B has finished the previous block
                    const int fl = *(volatile int *)&sm.flock_pub;
                    if (fl == flock) break;                                 // the speculation held (almost always)
                    flock = fl;                                             // acquisition or loss in the previous block: run again
                    clk = s_clk; thr = s_thr; index = s_index; sumc = s_sum; difc = s_dif;
                }
            }
            if (lane == 0) sm.n[buf] = n;
            // filter history for the next block = last 30 samples
            {
                float a = 0.0f;
                if (lane < 30) a = sm.x[lane & 3][96 + (lane >> 2)];        // sample 384 + lane -> slot lane
                __syncwarp();
                if (lane < 30) sm.x[lane & 3][lane >> 2] = a;
            }
            pc_bar_arrive(BAR_FULL);
            __syncwarp();
            PCPH(3);
        }
#ifdef M17B_PHASE_CLOCKS
        if (lane == 0) for (int q4 = 0; q4 < 4; q4++) S->dbg_phase[q4] = (unsigned long long)ph[q4];
#endif
        pc_bar_sync(BAR_DONE);                                              // pairs with B's last arrival
        if (lane < 30) S->tail[lane] = sm.x[lane & 3][lane >> 2];
        if (lane == 0) { S->clk = clk; S->thr = thr; S->index = index; S->sum = sumc; S->dif = difc; }
        return;
    }

    // =============================================================== B: emission + framer
    int flock = S->flock, fclk = S->fclk, ferr = S->ferr, frame_start = S->frame_start, sym_total = S->sym_total;
    const int base_g = t0 == 0 ? sym_total : sym_base[c];
    const int sym_entry = sym_total;
    if (lane < 8) { sm.sym[0][lane] = S->win[lane]; sm.head[lane] = S->head[lane]; }
    float *sbuf = syms + c * sym_pitch;
    if (t0 == 0) {
        // carry: the last 192 symbols of the previous call move in front of the new ones
        const int prev_n = S->prev_n;
        float tmp[6];
#pragma unroll
        for (int k = 0; k < 6; k++) tmp[k] = sbuf[prev_n + lane + 32 * k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 6; k++) sbuf[lane + 32 * k] = tmp[k];
    }
    if (lane == 0 && t0 == 0) sym_base[c] = base_g;
    int nfr = t0 == 0 ? 0 : nframes[c], nev = t0 == 0 ? 0 : nevents[c], n_aos = 0, n_los = 0;
    const int nfr_entry = nfr;
    if (lane == 0) sm.flock_pub = flock;
    __syncwarp();
    pc_bar_arrive(BAR_DONE);                                                // initial credit: A may publish its first block
#ifdef M17B_PHASE_CLOCKS
    long long phb[2] = {0, 0}, ptb = clock64();
#endif
    for (int64_t t = t0; t < t1; t++) {
        const int buf = (int)((t - t0) & 1);
        pc_bar_sync(BAR_FULL);
#ifdef M17B_PHASE_CLOCKS
        { const long long now__ = clock64(); phb[0] += now__ - ptb; ptb = now__; }
#endif
        float *hist = sm.sym[buf];
        const int n = *(volatile int *)&sm.n[buf];
        // ---- emit the block's symbols to the channel's stream
        {
            float *dst = sbuf + M17B_SYM_CARRY + (sym_total - base_g);
            for (int q = lane; q < n; q += 32) dst[q] = hist[8 + q];
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
                        for (int k = 0; k < 8; k++) { int idx = q - 7 + k; w[k] = (idx >= reset_at) ? hist[8 + idx] : 0.0f; }
                        ok = sync_unlocked_ok(w);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (m) found = q0 + __ffs(m) - 1;
                }
                if (found < 0) { p = n; break; }
                // acquisition: copy_sync(), m_fclk = 8 (m17_rx_frame.cpp:161-169)
                if (lane < 8) { int idx = found - 7 + lane; sm.head[lane] = (idx >= reset_at) ? hist[8 + idx] : 0.0f; }
                fclk = 8; ferr = 0; flock = 1;
                frame_start = sym_total + found - 7;
                if (lane == 0 && nev < ecap) { events[c * ecap + nev].sym_idx = sym_total + found; events[c * ecap + nev].kind = M17B_EV_AOS; }
                nev++; n_aos++;
                p = found + 1;
                __syncwarp();
            } else {
                const int need = M17B_FRAME_SYMS - fclk, avail = n - p;
                const int take = need < avail ? need : avail;
                if (fclk < 8 && lane < 8 && lane >= fclk && lane < fclk + take) sm.head[lane] = hist[8 + p + lane - fclk];
                fclk += take;
                p += take;
                __syncwarp();
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
                    if (nfr < fcap && lane < 16) {
                        uint32_t word = 0;
                        if (lane == 0) word = (uint32_t)frame_start;
                        else if (lane == 1) word = (uint32_t)r.type | ((uint32_t)flags << 8);
                        else if (lane == 11) word = ((uint32_t)r.votes << 16) | ((uint32_t)fe << 24);
                        else if (lane == 12) word = __float_as_uint(r.variance);
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
        // ---- carry: sliding window = last 8 symbols (zeros before a reset), into the head of the OTHER buffer
        {
            float wv = 0.0f;
            if (lane < 8) { int idx = n - 8 + lane; wv = (idx >= reset_at) ? hist[8 + idx] : 0.0f; }
            __syncwarp();
            if (lane < 8) sm.sym[buf ^ 1][lane] = wv;
        }
        sym_total += n;
        if (lane == 0) sm.flock_pub = flock;
        __syncwarp();
        pc_bar_arrive(BAR_DONE);
#ifdef M17B_PHASE_CLOCKS
        { const long long now__ = clock64(); phb[1] += now__ - ptb; ptb = now__; }
#endif
    }
#ifdef M17B_PHASE_CLOCKS
    if (lane == 0) { S->dbg_phase[4] = (unsigned long long)phb[0]; S->dbg_phase[5] = (unsigned long long)phb[1]; S->dbg_cycles = (unsigned long long)(phb[0] + phb[1]); }
#endif
    // ---- store state
    {
        const int last = (int)((t1 - t0) & 1);                              // the window of the next block sits in this buffer
        if (lane < 8) { S->win[lane] = sm.sym[last][lane]; S->head[lane] = sm.head[lane]; }
    }
    if (lane == 0) {
        S->flock = flock; S->fclk = fclk; S->ferr = ferr; S->frame_start = frame_start; S->sym_total = sym_total;
        S->prev_n = sym_total - base_g;
        nframes[c] = nfr < fcap ? nfr : (int)fcap;
        nevents[c] = nev < ecap ? nev : (int)ecap;
        if (frame_rng) frame_rng[c] = make_int2(nfr_entry, nfr < fcap ? nfr : (int)fcap);
        unsigned long long *q = stats + c * 8;
        q[0] += (unsigned long long)(nfr - nfr_entry); q[4] += (unsigned long long)n_aos; q[5] += (unsigned long long)n_los;
        q[7] += (unsigned long long)(sym_total - sym_entry);
    }
}

// ---------------------------------------------------------------- symbol seam: m17_rx_symbols on its own
// The sync-word correlator / framer FSM (m17_rx_sym, m17_rx_frame.cpp:126-172) for symbols that did not come from this
// library's timing loop (another demodulator, an equaliser in front of the framer, a symbol file): one warp per channel walks
// the channel's n symbols in chunks of 192 through the same shared-memory window the fused kernels use.  Writes the symbol
// stream, records (type / flags / sync fields) and events exactly as they do; the frame decode and post stages follow.
struct FramerSmem { float hist[8 + 208]; float head[8]; };
__global__ void __launch_bounds__(128) k_framer(const float *__restrict__ in, int64_t in_pitch, const int32_t *__restrict__ nin, int64_t nchan, RxChanState *st,
                                                float *syms, int64_t sym_pitch, int64_t sym_cap, int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base,
                                                m17b_frame_rec *frames, int64_t fcap, int32_t *__restrict__ nframes, m17b_event_rec *events, int64_t ecap,
                                                int32_t *__restrict__ nevents, unsigned long long *stats) {
    __shared__ FramerSmem sm_all[4];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * 4 + wid;
    if (c >= nchan) return;
    FramerSmem &sm = sm_all[wid];
    RxChanState *S = st + c;
    int flock = S->flock, fclk = S->fclk, ferr = S->ferr, frame_start = S->frame_start, sym_total = S->sym_total;
    const int base_g = sym_total;
    if (lane < 8) { sm.hist[lane] = S->win[lane]; sm.head[lane] = S->head[lane]; }
    float *sbuf = syms + c * sym_pitch;
    {   // carry: the last 192 symbols of the previous call move in front of the new ones
        const int prev_n = S->prev_n;
        float tmp[6];
#pragma unroll
        for (int k = 0; k < 6; k++) tmp[k] = sbuf[prev_n + lane + 32 * k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 6; k++) sbuf[lane + 32 * k] = tmp[k];
    }
    if (lane == 0) sym_base[c] = base_g;
    int ntot = nin[c];
    if (ntot < 0) ntot = 0;
    if (ntot > sym_cap) ntot = (int)sym_cap;
    int nfr = 0, nev = 0, n_aos = 0, n_los = 0;
    __syncwarp();
    for (int off = 0; off < ntot; off += 192) {
        const int n = ntot - off < 192 ? ntot - off : 192;
        for (int q = lane; q < n; q += 32) { const float v = in[c * in_pitch + off + q]; sm.hist[8 + q] = v; sbuf[M17B_SYM_CARRY + off + q] = v; }
        __syncwarp();
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
                        for (int k = 0; k < 8; k++) { int idx = q - 7 + k; w[k] = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
                        ok = sync_unlocked_ok(w);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (m) found = q0 + __ffs(m) - 1;
                }
                if (found < 0) { p = n; break; }
                if (lane < 8) { int idx = found - 7 + lane; sm.head[lane] = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }   // copy_sync :161-169
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
                    const SyncResult r = sync_check8(w);
                    const bool ok = sync_accept(r, true);
                    int flags = ok ? M17B_F_SYNC_OK : 0, fe;
                    bool los = false;
                    if (r.type == M17B_T_EOT) { los = true; fe = ferr; }                       // :137-140
                    else if (ok) { flags |= M17B_F_PARSED; ferr = 0; fe = 0; }                 // :144-146
                    else { ferr++; fe = ferr; if (ferr > 5) los = true; else flags |= M17B_F_PARSED; }   // :147-154
                    if (los) flags |= M17B_F_LOS;
                    if (nfr < fcap && lane < 16) {
                        uint32_t word = 0;
                        if (lane == 0) word = (uint32_t)frame_start;
                        else if (lane == 1) word = (uint32_t)r.type | ((uint32_t)flags << 8);
                        else if (lane == 11) word = ((uint32_t)r.votes << 16) | ((uint32_t)fe << 24);
                        else if (lane == 12) word = __float_as_uint(r.variance);
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
        float wv = 0.0f;
        if (lane < 8) { int idx = n - 8 + lane; wv = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
        __syncwarp();
        if (lane < 8) sm.hist[lane] = wv;
        sym_total += n;
        __syncwarp();
    }
    if (lane < 8) { S->win[lane] = sm.hist[lane]; S->head[lane] = sm.head[lane]; }
    if (lane == 0) {
        S->flock = flock; S->fclk = fclk; S->ferr = ferr; S->frame_start = frame_start; S->sym_total = sym_total;
        S->prev_n = sym_total - base_g;
        nsym[c] = ntot;
        nframes[c] = nfr < fcap ? nfr : (int)fcap;
        nevents[c] = nev < ecap ? nev : (int)ecap;
        unsigned long long *q = stats + c * 8;
        q[0] += (unsigned long long)nfr; q[4] += (unsigned long long)n_aos; q[5] += (unsigned long long)n_los;
        q[7] += (unsigned long long)(sym_total - base_g);
    }
}
