// sync_cta.cuh -- RX matched filter + symbol-timing loop + framer, one CTA (NW warps) per channel.
// Same arithmetic and the same speculation scheme as sync.cuh (see there), spread over NW warps so that a channel's
// serial chain per 40-ms block is NW times shorter: 32*NW symbols are speculated per round (one per thread, matched +
// derivative dot products as two independent sequential chains), the +-1 votes are prefix-summed across the CTA
// (warp scan + one shared-memory exchange), the first threshold trip is found with per-warp ballots, and the
// polyphase branch is stepped by every thread redundantly so no scalar state has to be broadcast.  With 1024
// channels this puts 4096 warps on the GPU instead of 1024.  The framer runs on warp 0.
// Replaces m17_rx_sync_samples / m17_sync_adjust (m17_rx_sync.cpp:25-99) and m17_rx_symbols / m17_rx_sym /
// m17_sync_check (m17_rx_frame.cpp:47-177).
#pragma once
#include "sync.cuh"

#define SY_XHALF 208          // (30 + 384) / 2 = 207 entries per parity

struct SyncCtaSmem {
    float xe[SY_XHALF], xo[SY_XHALF];   // samples incl. 30 of history, split by parity: the stride-2 windows of consecutive
                                        // threads read consecutive words
    float hist[SY_HIST];                // [0,8): sliding sync window carried in; [8, 8+n): symbols emitted in this block
    float head[8];
    int wtot[8], wfirst[8];             // per-warp vote totals / first trip position
    float b_sum, b_dif;                 // sum/dif of the last committed symbol (m17_rx_sync.cpp:78 statics)
    int b_thr, b_flock;
};
#define SC_NONE (1 << 20)

template <int NW, bool HAS_MEAN>
__global__ void __launch_bounds__(NW * 32) k_sync_frame_cta(const float *__restrict__ disc, const float *__restrict__ mean, int64_t nchan, int64_t T,
                                                            int t0, int t1, int2 *frame_rng, RxChanState *st, const float *__restrict__ g_mf, const float *__restrict__ g_md,
                                                            float *syms, int64_t sym_pitch, int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base,
                                                            m17b_frame_rec *frames, int64_t fcap, int32_t *__restrict__ nframes,
                                                            m17b_event_rec *events, int64_t ecap, int32_t *__restrict__ nevents,
                                                            unsigned long long *stats, int commit_fe) {
    constexpr int NT = NW * 32;
    constexpr int NQ = (384 + NT - 1) / NT;                       // samples staged per thread and block
    __shared__ SyncCtaSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t c = blockIdx.x;
    if (c >= nchan) return;
    RxChanState *S = st + c;
    float *out = sm.hist + 8;

    // ---- load state: scalars are replicated in every thread (uniform loads), framer state is only used by warp 0
    if (commit_fe && tid == 0 && t1 == T) { S->z0re = S->nz0re; S->z0im = S->nz0im; S->z1re = S->nz1re; S->z1im = S->nz1im; }
    int clk = S->clk, thr = S->thr, index = S->index;
    float sumc = S->sum, difc = S->dif;
    int flock = S->flock, fclk = S->fclk, ferr = S->ferr, frame_start = S->frame_start, sym_total = S->sym_total;
    const int base_g = t0 == 0 ? sym_total : sym_base[c];     // blocks [t0, t1) of the call: see sync.cuh
    const int sym_entry = sym_total;
    if (tid < 30) ((tid & 1) ? sm.xo : sm.xe)[tid >> 1] = S->tail[tid];
    if (tid < 8) { sm.hist[tid] = S->win[tid]; sm.head[tid] = S->head[tid]; }
    float *sbuf = syms + c * sym_pitch;
    if (t0 == 0) {   // carry: the last 192 symbols of the previous call move in front of the new ones
        const int prev_n = S->prev_n;
        float tmp[(192 + NT - 1) / NT];
#pragma unroll
        for (int k = 0; k < (192 + NT - 1) / NT; k++) { const int idx = tid + NT * k; tmp[k] = idx < 192 ? sbuf[prev_n + idx] : 0.0f; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < (192 + NT - 1) / NT; k++) { const int idx = tid + NT * k; if (idx < 192) sbuf[idx] = tmp[k]; }
    }
    if (tid == 0 && t0 == 0) sym_base[c] = base_g;
    int nfr = t0 == 0 ? 0 : nframes[c], nev = t0 == 0 ? 0 : nevents[c], n_aos = 0, n_los = 0;
    const int nfr_entry = nfr;
    f32x2 tp[M17B_FN];                      // (matched, derivative) tap pairs of the current polyphase branch
    int tap_index = -1;
    // the block's samples are fetched one block ahead
    float pf[NQ], pmu = 0.0f;
#pragma unroll
    for (int q = 0; q < NQ; q++) { const int j = tid + NT * q; pf[q] = j < 384 ? __ldg(disc + (c * T + t0) * 384 + j) : 0.0f; }
    if (HAS_MEAN) pmu = mean[c * T + t0];
    // staging slot of this thread: sample j = tid + NT*q lives at n = 30 + j; 30 and NT are even, so the parity is tid & 1
    float *stg = ((tid & 1) ? sm.xo : sm.xe) + ((30 + tid) >> 1);
    __syncthreads();

    for (int64_t t = t0; t < t1; t++) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            float v = pf[q];
            if (HAS_MEAN) v = v - pmu;                                          // m17_dsp.cpp:217-219
            if (tid + NT * q < 384) stg[(NT / 2) * q] = v;
        }
        if (t + 1 < t1) {
            const float *src = disc + (c * T + t + 1) * 384;
#pragma unroll
            for (int q = 0; q < NQ; q++) { const int j = tid + NT * q; pf[q] = j < 384 ? __ldg(src + j) : 0.0f; }
            if (HAS_MEAN) pmu = mean[c * T + t + 1];
        }
        __syncthreads();

        // ---- timing loop (m17_rx_sync.cpp:77-99); m17_rx_lock() is constant inside a block
        const int TH = flock ? 80 : 10;
        int i = 0, m_idx = 0;
        while (i < 384) {
            if (clk == 1) {
                // even-clock sample with no fresh symbol in this round: vote with the carried sum/dif (sync_update :38-42)
                float dd = (sumc < 0) ? -difc : difc;
                if (dd > 0) thr++;
                if (dd < 0) thr--;
                clk = 0;
                sync_adjust(TH, thr, index, clk, m_idx, out, tid);            // thread 0 writes an inserted zero symbol
                i++;
                continue;
            }
            if (index != tap_index) {
#pragma unroll
                for (int k = 0; k < M17B_FN; k++) tp[k] = pack2(__ldg(g_mf + index * M17B_FN + k), __ldg(g_md + index * M17B_FN + k));
                tap_index = index;
            }
            // speculate: thread k computes the symbol at sample j = i + 2k
            const int j = i + 2 * tid;
            const bool valid = j < 384;
            float s = 0.0f, d = 0.0f;
            if (valid) {
                const float *A = (i & 1) ? sm.xo : sm.xe;
                const float *B = (i & 1) ? sm.xe : sm.xo;
                const int h = (i >> 1) + tid, ob = i & 1;
                float x = A[h];
                unpack2(mul2(tp[0], pack2(x, x)), s, d);                      // sum = in[0]*c[0]; sum += in[i]*c[i]  (m17_rx_sync.cpp:25-31)
#pragma unroll
                for (int k = 1; k < M17B_FN; k++) {
                    x = (k & 1) ? B[h + (k >> 1) + ob] : A[h + (k >> 1)];
                    float ps, pd;
                    unpack2(mul2(tp[k], pack2(x, x)), ps, pd);                  // one FMUL2: both rounded products
                    s += ps;
                    d += pd;
                }
            }
            const bool has_vote = valid && (j + 1 < 384);
            int v = 0;
            if (has_vote) { float dd = (s < 0) ? -d : d; v = (dd > 0) - (dd < 0); }
            const int incl = warp_incl_scan(v, lane);
            if (lane == 31) sm.wtot[wid] = incl;
            __syncthreads();                                                   // (A) warp totals visible
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) { const int x = sm.wtot[w]; tot += x; if (w < wid) off += x; }
            const int th = thr + off + incl;
            const unsigned trig = __ballot_sync(0xffffffffu, has_vote && (th > TH || th < -TH));
            if (lane == 0) sm.wfirst[wid] = trig ? wid * 32 + __ffs(trig) - 1 : SC_NONE;
            __syncthreads();                                                   // (B) first trip per warp visible
            int P = SC_NONE;
#pragma unroll
            for (int w = 0; w < NW; w++) P = min(P, sm.wfirst[w]);
            const int rem = (385 - i) >> 1;                                    // number of samples i, i+2, .. below 384
            const int nv = rem < NT ? rem : NT;
            if (P == SC_NONE) {
                if (valid && m_idx + tid >= 0) out[m_idx + tid] = s;
                if (tid == nv - 1) { sm.b_sum = s; sm.b_dif = d; }
                m_idx += nv;
                thr += tot;
                const int last_j = i + 2 * (nv - 1);
                if (last_j + 1 < 384) { i = last_j + 2; clk = 0; } else { i = 384; clk = 1; }
            } else {
                if (tid <= P && m_idx + tid >= 0) out[m_idx + tid] = s;
                if (tid == P) { sm.b_sum = s; sm.b_dif = d; sm.b_thr = th; }
                m_idx += P + 1;
                clk = 0;
                i = i + 2 * P + 2;
            }
            __syncthreads();                                                   // (C) committed symbol's sum/dif (and thr) visible
            sumc = sm.b_sum;
            difc = sm.b_dif;
            if (P != SC_NONE) {
                thr = sm.b_thr;
                sync_adjust(TH, thr, index, clk, m_idx, out, tid);
            }
        }
        const int n = m_idx < 0 ? 0 : m_idx;
        // history for the next block: read now, written after the barrier below (staging never touches slots < 30)
        float keep_tail = 0.0f;
        if (tid < 30) keep_tail = ((tid & 1) ? sm.xo : sm.xe)[192 + (tid >> 1)];
        __syncthreads();                                                       // all emitted symbols (incl. an inserted zero) visible

        // ---- emit the block's symbols to the channel's stream (all threads)
        {
            float *dst = sbuf + M17B_SYM_CARRY + (sym_total - base_g);
            for (int q = tid; q < n; q += NT) dst[q] = out[q];
            if (tid == 0) nsym[c * T + t] = n;
        }

        // ---- framer (m17_rx_frame.cpp:126-172) on warp 0
        if (wid == 0) {
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
            // carry: sliding window = last 8 symbols (zeros before a reset)
            float wv = 0.0f;
            if (lane < 8) { int idx = n - 8 + lane; wv = (idx >= reset_at) ? sm.hist[8 + idx] : 0.0f; }
            __syncwarp();
            if (lane < 8) sm.hist[lane] = wv;
            if (lane == 0) sm.b_flock = flock;
        }
        if (tid < 30) ((tid & 1) ? sm.xo : sm.xe)[tid >> 1] = keep_tail;
        sym_total += n;
        __syncthreads();
        flock = sm.b_flock;
    }

    // ---- store state
    if (tid < 30) S->tail[tid] = ((tid & 1) ? sm.xo : sm.xe)[tid >> 1];
    if (tid < 8) { S->win[tid] = sm.hist[tid]; S->head[tid] = sm.head[tid]; }
    if (tid == 0) {
        S->clk = clk; S->thr = thr; S->index = index; S->sum = sumc; S->dif = difc;
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
