// framer.cuh -- the symbol seam: m17_rx_symbols / m17_rx_sym (m17_rx_frame.cpp:126-177) as a stand-alone kernel.
#pragma once
#include "sync_g.cuh"

// ---------------------------------------------------------------- symbol seam: m17_rx_symbols on its own
// The sync-word correlator / framer FSM (m17_rx_sym, m17_rx_frame.cpp:126-172) for symbols that did not come from this
// library's timing loop (another demodulator, an equaliser in front of the framer, a symbol file): one warp per channel walks
// the channel's n symbols in chunks of 192 through the same shared-memory window the fused kernels use.  Writes the symbol
// stream, records (type / flags / sync fields) and events exactly as they do; the frame decode and post stages follow.
struct FramerSmem { float hist[8 + 208]; float head[8]; };
__global__ void __launch_bounds__(128) k_framer(const float *__restrict__ in, int64_t in_pitch, const int32_t *__restrict__ nin, int64_t nchan, RxChanState *st,
                                                float *syms, int64_t sym_pitch, int64_t sym_cap, int32_t *__restrict__ nsym, int32_t *__restrict__ sym_base,
                                                m17b_frame_rec *frames, int64_t fcap, int32_t *__restrict__ nframes, m17b_event_rec *events, int64_t ecap,
                                                int32_t *__restrict__ nevents, unsigned long long *stats, int *overflow) {
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
    if (ntot > sym_cap) { ntot = (int)sym_cap; if (lane == 0) atomicOr(overflow, 1); }
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
