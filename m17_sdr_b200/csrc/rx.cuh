// rx.cuh -- the batched RX chain object: per-channel state, stage buffers, kernel sequencing, and the
// sequential per-channel LICH / packet bookkeeping.  Replaces m17_dsp_rx (m17_dsp.cpp:461-476) and the
// stateful tail of m17_rx_parse.cpp (update_lich :71-85, parse_packet :34-51, delivery gate :148-158).
//
// Kernel sequence for one call over [nchan] x [nblocks]:
//   k_frontend      (items = channel x block, lane per item)      int16 IQ -> raw discriminator + block mean
//   k_sync_frame    (warp per channel, blocks in order)           -> symbol stream, frame records (type/flags)
//   k_decode_frames (thread per frame)                            -> decoded bytes, Golay, CRC into the records
//   k_post          (thread per channel, frames in order)         -> LICH cache, delivery / LSF-event flags, stats
#pragma once
#include "chan.cuh"

#define M17B_TIMING_RING 64
#define M17B_MAX_SLICES 16
#define M17B_HOST_TAIL_PIECES 6       // m17b_dsp_rx_host: time pieces of the last channel chunk
#define M17B_MAX_GROUPS 8
struct m17b_rx {
    m17b_ctx *ctx;
    int64_t nchan, max_blocks, last_blocks;
    int64_t sym_pitch, fcap, ecap;
    RxChanState *d_state;
    float *d_disc, *d_mean;
    float *d_syms;
    int32_t *d_nsym, *d_sym_base, *d_nframes, *d_nevents;
    m17b_frame_rec *d_frames;
    m17b_event_rec *d_events;
    unsigned long long *d_stats;
    int32_t *d_dlist;                  // [nchan][fcap + 1]: per launch over channels [c0, c0 + nc) the region at c0 * (fcap + 1) holds {count, entries} of the
                                      // LSF / packet / BERT frames k_decode_frames has to take (decode.cuh)
    float *d_ssoft; StreamAux *d_saux; // scratch between the two stream-frame decode kernels: [nchan * fcap][272] kept trellis inputs, [nchan * fcap] LICH words / normaliser
    uint8_t *d_lsf_snap, *d_lsf_ver;  // [nchan][nsnap][32] link-setup data as it stood (version 0 = at the start of the call), [nchan][fcap] version per record
    int nsnap;
    int16_t *d_iq_stage[2];           // staging for the _host entry point (double buffered over channel chunks)
    int64_t stage_chunk;
    cudaStream_t copy_stream, aux_stream;   // aux: LSF/packet/BERT frame decode runs beside the stream-frame decode
    cudaEvent_t ev_fork, ev_join;
    cudaEvent_t ev_h2d[2], ev_done[2], ev_piece[M17B_HOST_TAIL_PIECES];
    int afc, bert, last_launches, seam_last;
    void *d_pkt_state;                // [nchan] RxPacketState of m17b_rx_reassemble_packets (app.cuh), allocated on first use
    RxEqState *d_eq;                  // [nchan] equaliser option (m17b_rx_set_equaliser): allocated on first use, nullptr = off
    int eq_on;
    int *d_overflow;                  // sticky flags (m17b_rx_get_overflow): 1 = the symbol seam was given more symbols than the capacity
    // time-sliced pipeline (see rx_pipeline): front end of slice k+1 | timing loop + framer of slice k | frame decode of slice k-1
    int slice_blocks;                 // blocks per slice; 0 = one slice (stages strictly in sequence)
    cudaStream_t s_fe, s_sync, s_dec;
    cudaEvent_t ev_start, ev_fe[M17B_MAX_SLICES], ev_sy[M17B_MAX_SLICES], ev_end;
    int2 *d_frame_rng;                // [M17B_MAX_SLICES][nchan] records completed by each slice
    // channel-group pipeline (see m17b_dsp_rx): the channels are cut into chan_groups contiguous groups, each running its own
    // front end -> sync -> decode -> post chain on its own stream, so that the latency-bound timing loop of one group shares the SMs
    // with the throughput-bound front end / decode of another.  Channels are independent: results do not depend on it.
    int chan_groups;                  // -1 = auto (4 groups from 512 channels up, see rx_groups_for), 0 / 1 = off
    cudaStream_t s_grp[M17B_MAX_GROUPS], s_grp_aux[M17B_MAX_GROUPS];
    cudaEvent_t ev_gfork, ev_gjoin[M17B_MAX_GROUPS], ev_gf[M17B_MAX_GROUPS], ev_gj[M17B_MAX_GROUPS];
    int sync_impl;                    // -1 auto; 0: warp per channel (sync.cuh); 2 / 4: CTA of that many warps per channel (sync_cta.cuh); 33: warp per channel with the taps in shared memory (sync_g.cuh)
    int timing;                       // record cudaEvents around each stage of the next calls (bench only)
    cudaEvent_t ev_stage[M17B_TIMING_RING][5];
    int64_t tcount;                   // calls made since timing was enabled
};

// m17b_rx_reset in ONE launch, a warp per channel (coalesced word stores): the per-channel state, the symbol carry region, counters
// and record counts, the packet reassembly state.  (It used to be a thread-per-channel kernel walking its 2 KB of state + six
// memsets: ~50 us per reset, 3 % of a bench step that restarts its streams every time.)
__global__ void __launch_bounds__(128) k_rx_reset(RxChanState *st, int64_t nchan, float *syms, int64_t sym_pitch, unsigned long long *stats, int32_t *nframes,
                                                  int32_t *nevents, int *overflow, uint32_t *pkt_state, int pkt_words) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nchan) return;
    // zero-initialised statics, then m17_rx_sync_init: m_clk = 1, m_thr = 0, m_index = 10 (m17_rx_sync.cpp:124-127)
    static_assert(sizeof(RxChanState) % 4 == 0, "state is cleared word by word");
    uint32_t *w = (uint32_t *)(st + c);
    for (int i = lane; i < (int)(sizeof(RxChanState) / 4); i += 32) w[i] = 0;
    __syncwarp();
    if (lane == 0) { st[c].clk = 1; st[c].index = 10; nframes[c] = 0; nevents[c] = 0; if (c == 0) *overflow = 0; }
    // only the carry region (the "last 192 symbols of the previous call") is ever read before it is written
    float *sb = syms + c * sym_pitch;
    for (int i = lane; i < M17B_SYM_CARRY; i += 32) sb[i] = 0.0f;
    if (lane < 8) stats[c * 8 + lane] = 0ull;
    if (pkt_state) for (int i = lane; i < pkt_words; i += 32) pkt_state[c * pkt_words + i] = 0u;
}

// Sequential per-channel tail of m17_rx_parse (update_lich / parse_packet / delivery gate).  One WARP per channel, 32 records
// per step, one record header per lane (one strided 16-byte load each).  Only a record that CHANGES per-channel state has to be
// taken in order: a LICH chunk that differs from the cache, an LSF / packet / BERT frame.  So each step alternates between
//   (a) all lanes: which of the records not yet done would change the state as it is now?  (ballot -> first such record j);
//       the records before j only read the state, and their flags follow from it lane-parallel;
//   (b) the warp together on record j: the five new bytes go into the cache, the 30-byte CRC of m_lsf[0] is ONE table look-up
//       per lane (position table d_crcpos, tables.cuh) and an XOR reduction, copy_lich / the snapshot are lane-per-byte copies;
// until the 32 records are done.  While a stream runs the same six chunks repeat and (a) finishes the step at once; on a noisy
// channel every wrongly corrected chunk costs one round of (b) (~100 cycles instead of a 30-step dependent CRC chain on one lane:
// 0.107 -> 0.049 ms on the bench mix).  The CRC verdict of an unchanged cache is carried, never recomputed (a pure function of the 30 bytes).
#define POST_WARPS 1
// For the M17-over-UDP gateway output (net.cuh) the kernel also keeps the history of the validated link-setup data m_lsf[1]
// over the call: snapshot 0 is the cache as the call found it, a new snapshot is taken whenever copy_lich() changes it, and
// every record gets the number of the snapshot that was current when it was parsed.
struct PostWarpSmem { uint4 hdr[32]; uint8_t lsf0[32], lsf1[32]; };
__global__ void __launch_bounds__(POST_WARPS * 32) k_post(m17b_frame_rec *frames, int64_t fcap, const int32_t *__restrict__ nframes, int64_t nchan,
                                                          RxChanState *st, const uint16_t *__restrict__ g_crcpos, unsigned long long *stats,
                                                          uint8_t *__restrict__ lsf_snap, int nsnap, uint8_t *__restrict__ lsf_ver,
                                                          const uint8_t *__restrict__ g_prbs, int bert_on) {
    __shared__ PostWarpSmem sm_all[POST_WARPS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * POST_WARPS + wid;
    if (c >= nchan) return;
    constexpr unsigned FULL = 0xffffffffu;
    PostWarpSmem &sm = sm_all[wid];
    RxChanState *S = st + c;
    sm.lsf0[lane] = S->lsf[0][lane];
    sm.lsf1[lane] = S->lsf[1][lane];
    __syncwarp();
    const unsigned crc_preset = __ldg(g_crcpos + 30 * 256);
    // CRC-16 of 30 bytes, byte `lane` supplied by lane < 30 (m17_crc_array_encode, m17_crc.cpp:26-35); the result is warp-uniform
    auto crc30 = [&](unsigned byte) {
        const unsigned v = lane < 30 ? (unsigned)__ldg(g_crcpos + lane * 256 + byte) : 0u;
        return __reduce_xor_sync(FULL, v) ^ crc_preset;
    };
    // warp-uniform channel state
    bool ok0 = crc30(sm.lsf0[lane]) == 0, ok1 = crc30(sm.lsf1[lane]) == 0;
    uint8_t *snap = lsf_snap + c * (int64_t)nsnap * 32;
    snap[lane] = sm.lsf1[lane];                                               // snapshot 0
    bool dirty = __any_sync(FULL, lane < 30 && sm.lsf0[lane] != sm.lsf1[lane]);
    int ver = 0;
    int packet_idx = S->packet_idx;
    unsigned n_stream = 0, n_gerr = 0, n_deliv = 0, n_lsf = 0;
    const int n = nframes[c];
    m17b_frame_rec *base = frames + c * fcap;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        uint4 h = make_uint4(0, 0, 0, 0);
        if (k < n) h = *(const uint4 *)(base + k);                 // bytes 0..15: sym_off, type, flags, golay_err, nbytes, lich[6], data[0..1]
        sm.hdr[lane] = h;
        __syncwarp();
        const int type = h.y & 0xFF, fl = (h.y >> 8) & 0xFF;
        const bool parsed = (k < n) && (fl & M17B_F_PARSED);
        const bool is_stream = parsed && type == M17B_T_STREAM;
        const int seq = (int)((h.w >> 8) & 0xFF) >> 5;                           // lich[5] >> 5
        const bool chunk = is_stream && seq < 6;                                 // update_lich stores it, m17_rx_parse.cpp:71-85
        const bool other = parsed && (type == M17B_T_LSF || type == M17B_T_PACKET || (type == M17B_T_BERT && bert_on));
        const int ge = is_stream ? (int)((h.y >> 16) & 0xFF) : 0;
        int nf = fl, myver = 0;
        int pos = 0;
        for (;;) {
            // (a) who would change the state as it is now?
            bool differs = false;
            if (chunk && lane >= pos) {
                const uint8_t *q = &sm.lsf0[seq * 5];
                differs = q[0] != (uint8_t)h.z || q[1] != (uint8_t)(h.z >> 8) || q[2] != (uint8_t)(h.z >> 16) || q[3] != (uint8_t)(h.z >> 24) || q[4] != (uint8_t)h.w;
            }
            const bool chg = lane >= pos && (other || (chunk && (differs || (ok0 && dirty))));
            const unsigned mch = __ballot_sync(FULL, chg);
            const int j = mch ? __ffs(mch) - 1 : 32;
            const bool inr = lane >= pos && lane < j;                            // read-only records of this round
            const bool ev = inr && chunk && ok0;                                 // update_lich -> copy_lich (a no-op: the cache is clean) -> parse_lsf
            const bool dl = inr && is_stream && (ok1 || ev);                     // :148-158
            const unsigned ms = __ballot_sync(FULL, inr && is_stream), me = __ballot_sync(FULL, ev), md = __ballot_sync(FULL, dl);
            if (ms) {
                n_stream += __popc(ms); n_lsf += __popc(me); n_deliv += __popc(md);
                n_gerr += __reduce_add_sync(FULL, inr ? (unsigned)ge : 0u);
                if (me) ok1 = true;
            }
            if (inr) { nf = fl | (ev ? M17B_F_LSF_EVENT : 0) | (dl ? M17B_F_DELIVERED : 0); myver = parsed ? ver : 0; }
            if (j == 32) break;
            // (b) record j, the warp together
            const uint4 q = sm.hdr[j];
            const int tj = q.y & 0xFF;
            int flags = (q.y >> 8) & 0xFF;
            if (tj == M17B_T_STREAM) {
                n_stream++;
                n_gerr += (q.y >> 16) & 0xFF;
                const int sj = (int)((q.w >> 8) & 0xFF) >> 5;
                const unsigned nb = lane < 4 ? (q.z >> (8 * lane)) & 0xFFu : q.w & 0xFFu;
                bool changed = false;
                if (lane < 5) { changed = sm.lsf0[sj * 5 + lane] != (uint8_t)nb; sm.lsf0[sj * 5 + lane] = (uint8_t)nb; }
                changed = __any_sync(FULL, changed);
                __syncwarp();
                if (changed) { ok0 = crc30(sm.lsf0[lane]) == 0; dirty = true; }
                if (ok0) {
                    if (dirty) {
                        sm.lsf1[lane] = sm.lsf0[lane];                           // copy_lich (bytes 30, 31 of both are zero)
                        if (ver < nsnap - 1) ver++;
                        snap[ver * 32 + lane] = lane < 30 ? sm.lsf0[lane] : 0;
                        dirty = false;
                    }
                    ok1 = true;
                    flags |= M17B_F_LSF_EVENT; n_lsf++;
                }
                if (ok1) { flags |= M17B_F_DELIVERED; n_deliv++; }              // :148-158
                __syncwarp();
            } else if (tj == M17B_T_LSF) {
                // decode_link_frame checks the CRC of m_packet, not of the decoded bytes (m17_rx_parse.cpp:98);
                // the honest verdict for the decoded LSF is the record's crc field
                if (crc30(lane < 30 ? S->packet[lane] : 0) == 0) { flags |= M17B_F_LSF_EVENT; n_lsf++; }
            } else if (tj == M17B_T_PACKET) {
                // parse_packet (m17_rx_parse.cpp:34-51) including its index bug; the copy is clamped to the buffer
                const m17b_frame_rec *r = base + k0 + j;
                const int last = r->data[25];
                const int eof = last >> 7, fn = (last >> 2) & 0x1F;
                if (eof) {
                    const int room = 800 - packet_idx, mm = fn < room ? fn : room;
                    if (lane < mm) S->packet[packet_idx + lane] = r->data[lane];
                    packet_idx = 0;
                } else {
                    if (lane < 25) S->packet[fn * 25 + lane] = r->data[lane];
                    packet_idx = fn * 25;
                }
                __syncwarp();
            } else if (tj == M17B_T_BERT && bert_on) {
                // the decode_bert_frame the reference left empty (m17_rx_parse.cpp:178-180): the frame's 197 PRBS9 bits go
                // through m17_prbs9_rx_check (m17_prbs9.cpp:40-64), bit by bit, in order
                if (lane == 0) {
                    const m17b_frame_rec *r = base + k0 + j;
                    int idx = S->prbs_idx, state = S->prbs_state;
                    unsigned bad = S->prbs_bad, good = S->prbs_good, eq = S->prbs_eq, dif = S->prbs_dif, nb = 0, ne = 0;
                    for (int b = 0; b < 197; b++) {
                        const unsigned bit = (r->data[b >> 3] >> (7 - (b & 7))) & 1u;
                        const unsigned d = bit ^ g_prbs[idx];
                        if (d) { dif = (dif + 1) & 0xFFFF; eq = 0; } else { eq = (eq + 1) & 0xFFFF; dif = 0; }
                        idx = idx + 1 == 511 ? 0 : idx + 1;
                        if (state == 0) {
                            bad = 0; good = 0;
                            if (eq >= 18) state = 1;
                            if (d) { idx = 0; state = 0; }                              // m17_prbs9_rx_reset
                        } else {
                            nb++; ne += d;
                            if (dif >= 18) state = 0;
                            if (d) bad = (bad + 1) & 0xFFFF;
                            // `if(d == 9) m_rx_good++` (m17_prbs9.cpp:61) can never fire: m_rx_good stays 0
                        }
                    }
                    S->prbs_idx = (uint16_t)idx; S->prbs_state = state;
                    S->prbs_bad = (uint16_t)bad; S->prbs_good = (uint16_t)good; S->prbs_eq = (uint16_t)eq; S->prbs_dif = (uint16_t)dif;
                    S->bert_bits += nb; S->bert_errs += ne;
                }
                __syncwarp();
            }
            if (lane == j) { nf = flags; myver = ver; }
            pos = j + 1;
        }
        if (k < n) {
            if (nf != fl) ((uint8_t *)(base + k))[5] = (uint8_t)nf;
            lsf_ver[c * fcap + k] = (uint8_t)myver;
        }
        __syncwarp();
    }
    S->lsf[0][lane] = sm.lsf0[lane];
    S->lsf[1][lane] = sm.lsf1[lane];
    if (lane == 0) {
        S->packet_idx = packet_idx;
        unsigned long long *q = stats + c * 8;
        q[1] += n_stream; q[2] += n_gerr; q[3] += n_deliv; q[6] += n_lsf;
    }
}

extern "C" int m17b_rx_destroy(m17b_rx *rx) {
    if (!rx) return M17B_E_ARG;
    cudaFree(rx->d_state); cudaFree(rx->d_disc); cudaFree(rx->d_mean); cudaFree(rx->d_syms); cudaFree(rx->d_nsym); cudaFree(rx->d_sym_base);
    cudaFree(rx->d_nframes); cudaFree(rx->d_nevents); cudaFree(rx->d_frames); cudaFree(rx->d_events); cudaFree(rx->d_stats);
    for (int i = 0; i < 2; i++) {
        if (rx->d_iq_stage[i]) cudaFree(rx->d_iq_stage[i]);
        if (rx->ev_h2d[i]) cudaEventDestroy(rx->ev_h2d[i]);
        if (rx->ev_done[i]) cudaEventDestroy(rx->ev_done[i]);
    }
    for (int i = 0; i < M17B_HOST_TAIL_PIECES; i++) if (rx->ev_piece[i]) cudaEventDestroy(rx->ev_piece[i]);
    cudaFree(rx->d_frame_rng); cudaFree(rx->d_lsf_snap); cudaFree(rx->d_lsf_ver); cudaFree(rx->d_overflow); cudaFree(rx->d_eq); cudaFree(rx->d_pkt_state); cudaFree(rx->d_ssoft); cudaFree(rx->d_saux); cudaFree(rx->d_dlist);
    if (rx->s_fe) cudaStreamDestroy(rx->s_fe);
    if (rx->s_sync) cudaStreamDestroy(rx->s_sync);
    if (rx->s_dec) cudaStreamDestroy(rx->s_dec);
    if (rx->ev_start) cudaEventDestroy(rx->ev_start);
    if (rx->ev_end) cudaEventDestroy(rx->ev_end);
    for (int i = 0; i < M17B_MAX_SLICES; i++) { if (rx->ev_fe[i]) cudaEventDestroy(rx->ev_fe[i]); if (rx->ev_sy[i]) cudaEventDestroy(rx->ev_sy[i]); }
    for (int g = 0; g < M17B_MAX_GROUPS; g++) {
        if (rx->s_grp[g]) cudaStreamDestroy(rx->s_grp[g]);
        if (rx->s_grp_aux[g]) cudaStreamDestroy(rx->s_grp_aux[g]);
        if (rx->ev_gjoin[g]) cudaEventDestroy(rx->ev_gjoin[g]);
        if (rx->ev_gf[g]) cudaEventDestroy(rx->ev_gf[g]);
        if (rx->ev_gj[g]) cudaEventDestroy(rx->ev_gj[g]);
    }
    if (rx->ev_gfork) cudaEventDestroy(rx->ev_gfork);
    if (rx->copy_stream) cudaStreamDestroy(rx->copy_stream);
    if (rx->aux_stream) cudaStreamDestroy(rx->aux_stream);
    if (rx->ev_fork) cudaEventDestroy(rx->ev_fork);
    if (rx->ev_join) cudaEventDestroy(rx->ev_join);
    for (int r = 0; r < M17B_TIMING_RING; r++) for (int i = 0; i < 5; i++) if (rx->ev_stage[r][i]) cudaEventDestroy(rx->ev_stage[r][i]);
    free(rx);
    return M17B_OK;
}

__global__ void k_rx_eq_open(RxEqState *e, int64_t nchan) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    eq_init(e[c].e, true, true);
    e[c].mid = 0.0f; e[c].pad[0] = e[c].pad[1] = e[c].pad[2] = 0.0f;
}
extern "C" int m17b_rx_reset(m17b_rx *rx, void *stream) {
    if (!rx) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    cudaStream_t st = as_stream(stream);
    k_rx_reset<<<grid_for(rx->nchan, 4), 128, 0, st>>>(rx->d_state, rx->nchan, rx->d_syms, rx->sym_pitch, rx->d_stats, rx->d_nframes, rx->d_nevents, rx->d_overflow,
                                                       (uint32_t *)rx->d_pkt_state, 832 / 4);                  // sizeof(RxPacketState) = 832, app.cuh
    KERNEL_CHECK();
    if (rx->d_eq) {                                                                                             // eq_open, m17_equalize.cpp:217-224
        k_rx_eq_open<<<grid_for(rx->nchan, 128), 128, 0, st>>>(rx->d_eq, rx->nchan);
        KERNEL_CHECK();
    }
    return M17B_OK;
}

// m17_rx_init / m17_rx_lost (m17_rx_frame.cpp:179-186): reset_sync() -- the sliding sync window reads as zeros -- and
// m_flock = false.  Nothing else: the timing loop, the frame counters and the LICH cache stay as they are.
__global__ void k_rx_framer_reset(RxChanState *st, int64_t nchan) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    for (int i = 0; i < 8; i++) st[c].win[i] = 0.0f;
    st[c].flock = 0;
}
extern "C" int m17b_rx_framer_reset(m17b_rx *rx, void *stream) {
    if (!rx) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    k_rx_framer_reset<<<grid_for(rx->nchan, 128), 128, 0, as_stream(stream)>>>(rx->d_state, rx->nchan);
    KERNEL_CHECK();
    return M17B_OK;
}

extern "C" int m17b_rx_create(m17b_ctx *ctx, int64_t nchan, int64_t max_blocks, m17b_rx **out) {
    if (!ctx || !out || nchan <= 0 || max_blocks <= 0) return M17B_E_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    m17b_rx *rx = (m17b_rx *)calloc(1, sizeof(m17b_rx));
    if (!rx) return M17B_E_NOMEM;
    rx->ctx = ctx; rx->nchan = nchan; rx->max_blocks = max_blocks;
    rx->sync_impl = -1;                // auto: few channels per launch -> 4 warps per channel (latency), many -> 1 warp per channel (fewest instructions)
    if (const char *e = getenv("M17B_SYNC_IMPL")) rx->sync_impl = atoi(e);     // tuning knob: 0 = one warp per channel, 2 / 4 = warps per channel
    rx->sym_pitch = (M17B_SYM_CARRY + max_blocks * M17B_SYM_CAP_PER_BLOCK + 3) & ~(int64_t)3;
    // records: the timing loop yields at most 193 symbols per block, the symbol seam accepts up to 200 per block -- size for the latter
    rx->fcap = (max_blocks * M17B_SYM_CAP_PER_BLOCK + M17B_FRAME_SYMS - 1) / M17B_FRAME_SYMS + 2;
    rx->ecap = 2 * rx->fcap + 4;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&rx->d_state, sizeof(RxChanState) * nchan);
    A((void **)&rx->d_disc, sizeof(float) * nchan * max_blocks * 384);
    A((void **)&rx->d_mean, sizeof(float) * nchan * max_blocks);
    A((void **)&rx->d_syms, sizeof(float) * nchan * rx->sym_pitch);
    A((void **)&rx->d_nsym, sizeof(int32_t) * nchan * max_blocks);
    A((void **)&rx->d_sym_base, sizeof(int32_t) * nchan);
    A((void **)&rx->d_nframes, sizeof(int32_t) * nchan);
    A((void **)&rx->d_nevents, sizeof(int32_t) * nchan);
    A((void **)&rx->d_frames, sizeof(m17b_frame_rec) * nchan * rx->fcap);
    A((void **)&rx->d_events, sizeof(m17b_event_rec) * nchan * rx->ecap);
    A((void **)&rx->d_stats, sizeof(unsigned long long) * nchan * 8);
    A((void **)&rx->d_frame_rng, sizeof(int2) * nchan * M17B_MAX_SLICES);
    rx->nsnap = (int)(max_blocks / 6 + 2);        // a new LSF needs six LICH chunks = six frames
    if (rx->nsnap > 255) rx->nsnap = 255;
    A((void **)&rx->d_lsf_snap, (size_t)nchan * rx->nsnap * 32);
    A((void **)&rx->d_lsf_ver, (size_t)nchan * rx->fcap);
    A((void **)&rx->d_ssoft, (size_t)nchan * rx->fcap * STREAM_NIN * sizeof(float));
    A((void **)&rx->d_saux, (size_t)nchan * rx->fcap * sizeof(StreamAux));
    A((void **)&rx->d_dlist, (size_t)nchan * (rx->fcap + 1) * sizeof(int32_t));
    A((void **)&rx->d_overflow, sizeof(int));
    if (e != cudaSuccess) { m17b_set_cuda_error(e, __FILE__, __LINE__); m17b_rx_destroy(rx); return e == cudaErrorMemoryAllocation ? M17B_E_NOMEM : M17B_E_CUDA; }
#define CREATE_TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { m17b_set_cuda_error(e__, __FILE__, __LINE__); m17b_rx_destroy(rx); return M17B_E_CUDA; } } while (0)
    {   // The LSF / packet decode kernel is launched beside the stream frames' trellis kernel and after the same predecessor; a few
        // long CTAs against thousands of short ones.  CTAs are dispatched kernel by kernel unless a priority says otherwise: at
        // equal priority the trellis kernel's 9216 CTAs went first and the LSF kernel ran after them instead of beside them.
        int lo = 0, hi = 0;
        CREATE_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CREATE_TRY(cudaStreamCreateWithPriority(&rx->aux_stream, cudaStreamNonBlocking, hi));
    }
    CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_fork, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_join, cudaEventDisableTiming));
    {   // the serial chain of the timing loop is the critical path of the pipeline: it gets the highest priority
        int lo = 0, hi = 0;
        CREATE_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CREATE_TRY(cudaStreamCreateWithPriority(&rx->s_sync, cudaStreamNonBlocking, hi));
        CREATE_TRY(cudaStreamCreateWithPriority(&rx->s_dec, cudaStreamNonBlocking, hi < lo ? hi + 1 : lo));
        CREATE_TRY(cudaStreamCreateWithPriority(&rx->s_fe, cudaStreamNonBlocking, lo));
        CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_start, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_end, cudaEventDisableTiming));
        for (int i = 0; i < M17B_MAX_SLICES; i++) {
            CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_fe[i], cudaEventDisableTiming));
            CREATE_TRY(cudaEventCreateWithFlags(&rx->ev_sy[i], cudaEventDisableTiming));
        }
    }
    rx->slice_blocks = 0;          // measured on B200: slicing gives no gain at 1024 channels (the kernels contend for issue slots), see DESIGN.md 4
    if (const char *e2 = getenv("M17B_SLICE_BLOCKS")) rx->slice_blocks = atoi(e2);
    rx->chan_groups = -1;
    if (const char *e3 = getenv("M17B_CHAN_GROUPS")) rx->chan_groups = atoi(e3);
    e = cudaMemset(rx->d_syms, 0, sizeof(float) * nchan * rx->sym_pitch);
    if (e != cudaSuccess) { m17b_set_cuda_error(e, __FILE__, __LINE__); m17b_rx_destroy(rx); return M17B_E_CUDA; }
    int rc = m17b_rx_reset(rx, nullptr);
    if (rc) { m17b_rx_destroy(rx); return rc; }
    CREATE_TRY(cudaStreamSynchronize(nullptr));
#undef CREATE_TRY
    *out = rx;
    return M17B_OK;
}

// Equaliser option: eq_open (m17_equalize.cpp:217-224) once, then eq_train_unknown (:185-213) on every
// (half-symbol, symbol) pair of the matched filter, its output to the framer -- between m17_rx_sync.cpp:77 and m17_rx_frame.cpp:173.
// The reference itself never calls its equaliser; off (the default) is upstream behaviour.  Turning it on (re)opens the equaliser of
// every channel; m17b_rx_reset does the same.  Not available together with AFC.  Runs in the warp-per-channel timing-loop kernel.
extern "C" int m17b_rx_set_equaliser(m17b_rx *rx, int on, void *stream) {
    if (!rx) return M17B_E_ARG;
    if (on && rx->afc) return M17B_E_ARG;
    if (on) {
        CUDA_TRY(cudaSetDevice(rx->ctx->device));
        if (!rx->d_eq) CUDA_TRY(cudaMalloc((void **)&rx->d_eq, sizeof(RxEqState) * rx->nchan));
        k_rx_eq_open<<<grid_for(rx->nchan, 128), 128, 0, as_stream(stream)>>>(rx->d_eq, rx->nchan);
        KERNEL_CHECK();
    }
    rx->eq_on = on != 0;
    return M17B_OK;
}

// radio_set_afc_on / radio_set_afc_off (radio.cpp:146-152).  With AFC on, the NCO step of a block depends on the framer state and
// the discriminator mean of the block before it, so a channel's blocks are serial through the whole chain: m17b_dsp_rx then runs
// the AFC front end of each block (afc.cuh) inside the timing-loop kernel's block loop, one launch per call.
extern "C" int m17b_rx_set_afc(m17b_rx *rx, int on, void *stream) {
    if (!rx) return M17B_E_ARG;
    if (!on && rx->afc) {
        CUDA_TRY(cudaSetDevice(rx->ctx->device));
        k_afc_off<<<grid_for(rx->nchan, 128), 128, 0, as_stream(stream)>>>(rx->d_state, rx->nchan);
        KERNEL_CHECK();
    }
    if (on && rx->eq_on) return M17B_E_ARG;          // the equaliser option and AFC are not combined
    rx->afc = on != 0;
    return M17B_OK;
}

// front end over blocks [t0, t0+Tc) of channels [.., +nc)
static int launch_frontend(const int16_t *d_iq, int64_t nc, int64_t T, int64_t t0, int64_t Tc, RxChanState *state, float *disc, float *mean, cudaStream_t st) {
    k_frontend<<<grid_for(nc * Tc, FE_WARPS * 32), FE_WARPS * 32, 0, st>>>((const uint32_t *)d_iq, nc, T, t0, Tc, state, disc, mean, F32X2_ONE);
    KERNEL_CHECK();
    return M17B_OK;
}

#define STAGE_MARK(i) do { if (rx->timing) CUDA_TRY(cudaEventRecord(rx->ev_stage[rx->tcount % M17B_TIMING_RING][i], st)); } while (0)

// matched filter + timing loop + framer over blocks [t0, t1) of channels [c0, c0+nc)
static int launch_sync(m17b_rx *rx, int64_t c0, int64_t nc, const float *disc, const float *mean, int64_t T, int t0, int t1, int2 *rng, int commit_fe,
                       cudaStream_t st, bool shared_gpu = false) {
    m17b_ctx *ctx = rx->ctx;
#define SYNC_ARGS disc, mean, nc, T, t0, t1, rng, rx->d_state + c0, ctx->d_mf, ctx->d_md, rx->d_syms + c0 * rx->sym_pitch, rx->sym_pitch, rx->d_nsym + c0 * T, \
                  rx->d_sym_base + c0, rx->d_frames + c0 * rx->fcap, rx->fcap, rx->d_nframes + c0, rx->d_events + c0 * rx->ecap, rx->ecap, rx->d_nevents + c0, \
                  rx->d_stats + c0 * 8, commit_fe
    // auto policy (measured on B200): one warp per channel has the fewest instructions and wins from ~512 channels
    // up; below that a channel's serial chain is the whole story and four warps per channel shorten it (only when the launch has
    // the GPU to itself: not inside a channel group).  Beyond one resident wave (8 channels per SM at 168 registers = 1184 on a
    // B200) the tap-pairs-in-shared-memory variant (108 registers, 16 warps per SM) is 15-23 % faster.
    const int impl = rx->sync_impl >= 0 ? rx->sync_impl : (nc <= 256 && !shared_gpu ? 4 : nc <= 8 * 148 ? 0 : 33);
    const f32x2 one = 0x3F8000003F800000ull;                          // (1.0f, 1.0f), passed as data so that ptxas cannot contract the packed adds (sync.cuh, dot2)
    if (rx->eq_on) {
        const size_t smem = sizeof(float) * SY_HIST * SY_WARPS;       // the half-symbol values of a block, per warp
        const unsigned g = grid_for(nc, SY_WARPS);
        if (mean) k_sync_frame<true, false, true><<<g, SY_WARPS * 32, smem, st>>>(SYNC_ARGS, one, nullptr, nullptr, nullptr, rx->d_eq + c0);
        else      k_sync_frame<false, false, true><<<g, SY_WARPS * 32, smem, st>>>(SYNC_ARGS, one, nullptr, nullptr, nullptr, rx->d_eq + c0);
    } else if (impl == 33) {
        const size_t smem = sizeof(SyncGroupSmem) * SY_WARPS;
        const unsigned g = grid_for(nc, SY_WARPS);
        if (mean) k_sync_frame_g<true, 32, true><<<g, SY_WARPS * 32, smem, st>>>(SYNC_ARGS, one);
        else      k_sync_frame_g<false, 32, true><<<g, SY_WARPS * 32, smem, st>>>(SYNC_ARGS, one);
    } else if (impl == 4) {
        if (mean) k_sync_frame_cta<4, true><<<(unsigned)nc, 128, 0, st>>>(SYNC_ARGS);
        else      k_sync_frame_cta<4, false><<<(unsigned)nc, 128, 0, st>>>(SYNC_ARGS);
    } else if (impl == 2) {
        if (mean) k_sync_frame_cta<2, true><<<(unsigned)nc, 64, 0, st>>>(SYNC_ARGS);
        else      k_sync_frame_cta<2, false><<<(unsigned)nc, 64, 0, st>>>(SYNC_ARGS);
    } else {
        const unsigned g = grid_for(nc, SY_WARPS);
        if (mean) k_sync_frame<true><<<g, SY_WARPS * 32, 0, st>>>(SYNC_ARGS, one);
        else      k_sync_frame<false><<<g, SY_WARPS * 32, 0, st>>>(SYNC_ARGS, one);
    }
#undef SYNC_ARGS
    KERNEL_CHECK();
    return M17B_OK;
}

// The chain for channels [c0, c0+nc) over T blocks.  d_iq == NULL: the chain starts at the discriminator seam.
//
// One slice (rx->slice_blocks == 0, short calls, or stage timing on): front end -> sync/framer -> decode -> post in sequence on st.
// Several slices: the call is cut into time slices of slice_blocks blocks.  The front end of slice k+1 (block-parallel, HBM /
// issue bound), the timing loop + framer of slice k (one serial chain per channel: latency-bound, leaves most issue slots
// idle) and the frame decode of slice k-1 (ALU-bound) run on three streams and share the SMs; events carry the only true
// dependencies (FE(k) -> SYNC(k), SYNC(k-1) -> SYNC(k) by stream order, SYNC(k) -> DECODE(k)).  Results are identical
// to the single-slice order: every kernel reads and writes exactly what it would have.
static int rx_pipeline(m17b_rx *rx, int64_t c0, int64_t nc, const int16_t *d_iq, const float *disc_in, int64_t T, cudaStream_t st, bool allow_slices = true,
                       int grp = -1) {
    m17b_ctx *ctx = rx->ctx;
    // the stream / events of the side-by-side LSF/packet decode: the object's own, or the channel group's
    cudaStream_t aux_stream = grp >= 0 ? rx->s_grp_aux[grp] : rx->aux_stream;
    cudaEvent_t ev_fork = grp >= 0 ? rx->ev_gf[grp] : rx->ev_fork, ev_join = grp >= 0 ? rx->ev_gj[grp] : rx->ev_join;
    float *disc_w = rx->d_disc + c0 * T * 384, *mean_w = rx->d_mean + c0 * T;
    const float *disc = d_iq ? disc_w : disc_in, *mean = d_iq ? mean_w : nullptr;
    const int commit_fe = d_iq ? 1 : 0;
    m17b_frame_rec *frames = rx->d_frames + c0 * rx->fcap;
    float *syms = rx->d_syms + c0 * rx->sym_pitch;
    int64_t sb = rx->slice_blocks;
    int nsl = (sb > 0 && !rx->timing && allow_slices) ? (int)((T + sb - 1) / sb) : 1;
    if (nsl > M17B_MAX_SLICES) { nsl = M17B_MAX_SLICES; }
    if (rx->afc && d_iq) {
        // AFC on: the NCO step of a block depends on the framer state and the mean of the block before it, so a channel's blocks
        // are serial through the WHOLE chain -- but only within the channel.  One launch: the warp that owns the channel runs
        // the AFC front end of each block inside the timing loop's block loop (k_sync_frame<true, true>, afc.cuh).
        STAGE_MARK(0);
        STAGE_MARK(1);
        {
            const size_t smem = sizeof(AfcWarpSmem) * SY_WARPS;
            CUDA_TRY(cudaFuncSetAttribute(k_sync_frame<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const f32x2 one = 0x3F8000003F800000ull;
            k_sync_frame<true, true><<<grid_for(nc, SY_WARPS), SY_WARPS * 32, smem, st>>>(
                nullptr, nullptr, nc, T, 0, (int)T, nullptr, rx->d_state + c0, ctx->d_mf, ctx->d_md, rx->d_syms + c0 * rx->sym_pitch, rx->sym_pitch,
                rx->d_nsym + c0 * T, rx->d_sym_base + c0, rx->d_frames + c0 * rx->fcap, rx->fcap, rx->d_nframes + c0, rx->d_events + c0 * rx->ecap, rx->ecap,
                rx->d_nevents + c0, rx->d_stats + c0 * 8, 0, one, (const uint32_t *)d_iq, disc_w, mean_w);
            KERNEL_CHECK();
        }
        STAGE_MARK(2);
        int rc = launch_decode(ctx, syms, rx->sym_pitch, M17B_SYM_CARRY, rx->d_sym_base + c0, frames, rx->fcap, rx->d_nframes + c0, nc, nullptr, rx->d_ssoft + c0 * rx->fcap * STREAM_NIN, rx->d_saux + c0 * rx->fcap, rx->d_dlist + c0 * (rx->fcap + 1), st,
                               aux_stream, ev_fork, ev_join, nullptr, 0, rx->bert);
        if (rc) return rc;
        STAGE_MARK(3);
        k_post<<<grid_for(nc, POST_WARPS), POST_WARPS * 32, 0, st>>>(frames, rx->fcap, rx->d_nframes + c0, nc, rx->d_state + c0, ctx->d_crcpos, rx->d_stats + c0 * 8,
                                                                  rx->d_lsf_snap + c0 * rx->nsnap * 32, rx->nsnap, rx->d_lsf_ver + c0 * rx->fcap,
                                                                  ctx->d_prbs, rx->bert);
        KERNEL_CHECK();
        STAGE_MARK(4);
        rx->last_launches += 4;
        return M17B_OK;
    }
    if (nsl < 2) {
        STAGE_MARK(0);
        if (d_iq) {
            int rcf = launch_frontend(d_iq, nc, T, 0, T, rx->d_state + c0, disc_w, mean_w, st);
            if (rcf) return rcf;
            rx->last_launches += 1;
        }
        STAGE_MARK(1);
        int rc = launch_sync(rx, c0, nc, disc, mean, T, 0, (int)T, nullptr, commit_fe, st, grp >= 0);
        if (rc) return rc;
        STAGE_MARK(2);
        rc = launch_decode(ctx, syms, rx->sym_pitch, M17B_SYM_CARRY, rx->d_sym_base + c0, frames, rx->fcap, rx->d_nframes + c0, nc, nullptr, rx->d_ssoft + c0 * rx->fcap * STREAM_NIN, rx->d_saux + c0 * rx->fcap, rx->d_dlist + c0 * (rx->fcap + 1), st,
                           aux_stream, ev_fork, ev_join, nullptr, 0, rx->bert);
        if (rc) return rc;
        STAGE_MARK(3);
        k_post<<<grid_for(nc, POST_WARPS), POST_WARPS * 32, 0, st>>>(frames, rx->fcap, rx->d_nframes + c0, nc, rx->d_state + c0, ctx->d_crcpos, rx->d_stats + c0 * 8,
                                                                  rx->d_lsf_snap + c0 * rx->nsnap * 32, rx->nsnap, rx->d_lsf_ver + c0 * rx->fcap,
                                                                  ctx->d_prbs, rx->bert);
        KERNEL_CHECK();
        STAGE_MARK(4);
        rx->last_launches += 4;       // sync/framer, two decode kernels, post
        return M17B_OK;
    }
    sb = (T + nsl - 1) / nsl;
    CUDA_TRY(cudaEventRecord(rx->ev_start, st));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_fe, rx->ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_sync, rx->ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_dec, rx->ev_start, 0));
    for (int k = 0; k < nsl; k++) {
        const int64_t t0 = k * sb, t1 = (t0 + sb < T) ? t0 + sb : T;
        if (t0 >= t1) break;
        int2 *rng = rx->d_frame_rng + (int64_t)k * rx->nchan + c0;
        if (d_iq) {
            int rcf = launch_frontend(d_iq, nc, T, t0, t1 - t0, rx->d_state + c0, disc_w, mean_w, rx->s_fe);
            if (rcf) return rcf;
            CUDA_TRY(cudaEventRecord(rx->ev_fe[k], rx->s_fe));
            CUDA_TRY(cudaStreamWaitEvent(rx->s_sync, rx->ev_fe[k], 0));
            rx->last_launches += 1;
        }
        int rc = launch_sync(rx, c0, nc, disc, mean, T, (int)t0, (int)t1, rng, commit_fe, rx->s_sync);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(rx->ev_sy[k], rx->s_sync));
        CUDA_TRY(cudaStreamWaitEvent(rx->s_dec, rx->ev_sy[k], 0));
        const int64_t span = t1 - t0;
        rc = launch_decode(ctx, syms, rx->sym_pitch, M17B_SYM_CARRY, rx->d_sym_base + c0, frames, rx->fcap, rx->d_nframes + c0, nc, nullptr, rx->d_ssoft + c0 * rx->fcap * STREAM_NIN, rx->d_saux + c0 * rx->fcap, rx->d_dlist + c0 * (rx->fcap + 1), rx->s_dec,
                           aux_stream, ev_fork, ev_join, rng, span + span / 64 + 4, rx->bert);
        if (rc) return rc;
        rx->last_launches += 3;
    }
    k_post<<<grid_for(nc, POST_WARPS), POST_WARPS * 32, 0, rx->s_dec>>>(frames, rx->fcap, rx->d_nframes + c0, nc, rx->d_state + c0, ctx->d_crcpos, rx->d_stats + c0 * 8,
                                                                  rx->d_lsf_snap + c0 * rx->nsnap * 32, rx->nsnap, rx->d_lsf_ver + c0 * rx->fcap,
                                                                  ctx->d_prbs, rx->bert);
    KERNEL_CHECK();
    rx->last_launches += 1;
    CUDA_TRY(cudaEventRecord(rx->ev_end, rx->s_dec));
    CUDA_TRY(cudaStreamWaitEvent(st, rx->ev_end, 0));
    return M17B_OK;
}


// The whole batch as chan_groups independent chains on their own streams (fork from / join into st).
static int rx_grouped(m17b_rx *rx, const int16_t *d_iq, const float *d_disc, int64_t T, cudaStream_t st, int G) {
    if (!rx->ev_gfork) {
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_gfork, cudaEventDisableTiming));
        for (int g = 0; g < M17B_MAX_GROUPS; g++) {
            CUDA_TRY(cudaStreamCreateWithPriority(&rx->s_grp[g], cudaStreamNonBlocking, lo));
            CUDA_TRY(cudaStreamCreateWithPriority(&rx->s_grp_aux[g], cudaStreamNonBlocking, hi));   // (see aux_stream)
            CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_gjoin[g], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_gf[g], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_gj[g], cudaEventDisableTiming));
        }
    }
    CUDA_TRY(cudaEventRecord(rx->ev_gfork, st));
    const int64_t per = (rx->nchan + G - 1) / G;
    for (int g = 0; g < G; g++) {
        const int64_t c0 = g * per, nc = (rx->nchan - c0 < per) ? rx->nchan - c0 : per;
        if (nc <= 0) break;
        CUDA_TRY(cudaStreamWaitEvent(rx->s_grp[g], rx->ev_gfork, 0));
        int rc = rx_pipeline(rx, c0, nc, d_iq ? d_iq + c0 * T * 3840 : nullptr, d_disc ? d_disc + c0 * T * 384 : nullptr, T, rx->s_grp[g], false, g);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(rx->ev_gjoin[g], rx->s_grp[g]));
        CUDA_TRY(cudaStreamWaitEvent(st, rx->ev_gjoin[g], 0));
    }
    return M17B_OK;
}
static int rx_groups_for(const m17b_rx *rx) {
    int G = rx->chan_groups;
    // auto: the sync kernel ends with a tail of slow channels on mostly idle SMs, and its serial chains leave most issue slots
    // free; four staggered groups fill both with the neighbours' front end / decode.  Measured with the round-2 kernels
    // (benchmarks/chan_groups.py, profiles/r02f_chan_groups.jsonl): 1024 channels 1.56 -> 1.52 ms, 2048: 3.12 -> 2.83, 4096: 5.72 -> 5.23,
    // 8192 x 100 blocks: 4.73 -> 4.54; below 512 channels nothing is gained.  More than 4 groups exceed the 8 hardware queues.
    if (G < 0) G = rx->nchan >= 512 ? 4 : 1;
    if (G > M17B_MAX_GROUPS) G = M17B_MAX_GROUPS;
    if (G < 2 || rx->timing || rx->nchan < 2 * G) return 1;
    return G;
}

extern "C" int m17b_dsp_rx(m17b_rx *rx, const int16_t *d_iq, int64_t nblocks, void *stream) {
    if (!rx || !d_iq || nblocks <= 0) return M17B_E_ARG;
    if (nblocks > rx->max_blocks) return M17B_E_CAPACITY;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    rx->last_launches = 0; rx->last_blocks = nblocks; rx->seam_last = 0;
    const int G = rx_groups_for(rx);
    int rc = G > 1 ? rx_grouped(rx, d_iq, nullptr, nblocks, as_stream(stream), G) : rx_pipeline(rx, 0, rx->nchan, d_iq, nullptr, nblocks, as_stream(stream));
    if (rx->timing) rx->tcount++;
    return rc;
}

extern "C" int m17b_rx_baseband(m17b_rx *rx, const float *d_disc, int64_t nblocks, void *stream) {
    if (!rx || !d_disc || nblocks <= 0 || ((uintptr_t)d_disc & 15)) return M17B_E_ARG;      // rows are fetched in 16-byte pieces
    if (nblocks > rx->max_blocks) return M17B_E_CAPACITY;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    rx->last_launches = 0; rx->last_blocks = nblocks; rx->seam_last = 1;
    const int G = rx_groups_for(rx);
    int rc = G > 1 ? rx_grouped(rx, nullptr, d_disc, nblocks, as_stream(stream), G) : rx_pipeline(rx, 0, rx->nchan, nullptr, d_disc, nblocks, as_stream(stream));
    if (rx->timing) rx->tcount++;
    return rc;
}

// Symbol seam: m17_rx_symbols (m17_rx_frame.cpp:173-177) on symbols supplied by the caller -- framer, frame decode, post.
// d_syms float [nchan][pitch], d_nsym int32 [nchan] symbols per channel (<= max_blocks * 200); results as for m17b_dsp_rx,
// reported as one "block" (view.nblocks = 1, view.d_nsym[c] = symbols taken).
extern "C" int m17b_rx_symbols(m17b_rx *rx, const float *d_syms, int64_t pitch, const int32_t *d_nsym, void *stream) {
    if (!rx || !d_syms || !d_nsym || pitch < 0) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    m17b_ctx *ctx = rx->ctx;
    cudaStream_t st = as_stream(stream);
    rx->last_launches = 0; rx->last_blocks = 1; rx->seam_last = 2;
    const int64_t cap = rx->sym_pitch - M17B_SYM_CARRY;
    k_framer<<<grid_for(rx->nchan, 4), 128, 0, st>>>(d_syms, pitch, d_nsym, rx->nchan, rx->d_state, rx->d_syms, rx->sym_pitch, cap, rx->d_nsym, rx->d_sym_base,
                                                   rx->d_frames, rx->fcap, rx->d_nframes, rx->d_events, rx->ecap, rx->d_nevents, rx->d_stats, rx->d_overflow);
    KERNEL_CHECK();
    int rc = launch_decode(ctx, rx->d_syms, rx->sym_pitch, M17B_SYM_CARRY, rx->d_sym_base, rx->d_frames, rx->fcap, rx->d_nframes, rx->nchan, nullptr, rx->d_ssoft, rx->d_saux, rx->d_dlist, st,
                           rx->aux_stream, rx->ev_fork, rx->ev_join, nullptr, 0, rx->bert);
    if (rc) return rc;
    k_post<<<grid_for(rx->nchan, POST_WARPS), POST_WARPS * 32, 0, st>>>(rx->d_frames, rx->fcap, rx->d_nframes, rx->nchan, rx->d_state, ctx->d_crcpos, rx->d_stats,
                                                                      rx->d_lsf_snap, rx->nsnap, rx->d_lsf_ver, ctx->d_prbs, rx->bert);
    KERNEL_CHECK();
    rx->last_launches = 4;
    return M17B_OK;
}

// BERT receive.  Upstream decodes nothing for BERT frames (decode_bert_frame is empty); with this switched
// on they are de-punctured / Viterbi-decoded like any other frame and their PRBS9 bits run through m17_prbs9_rx_check.
extern "C" int m17b_rx_set_bert(m17b_rx *rx, int on) {
    if (!rx) return M17B_E_ARG;
    rx->bert = on != 0;
    return M17B_OK;
}
__global__ void k_get_dbg(const RxChanState *st, int64_t nchan, unsigned long long *out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchan) { out[8 * c] = st[c].dbg_cycles; out[8 * c + 1] = st[c].dbg_rounds; for (int i = 0; i < 6; i++) out[8 * c + 2 + i] = st[c].dbg_phase[i]; }
}
// instrumentation: per channel {SM cycles, speculation rounds, cycles per phase x6 (builds with -DM17B_PHASE_CLOCKS)} of the last
// one-warp-per-channel timing-loop launch
extern "C" int m17b_rx_debug_sync(m17b_rx *rx, uint64_t *d_out, void *stream) {
    if (!rx || !d_out) return M17B_E_ARG;
    k_get_dbg<<<grid_for(rx->nchan, 128), 128, 0, as_stream(stream)>>>(rx->d_state, rx->nchan, (unsigned long long *)d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
__global__ void k_get_bert(const RxChanState *st, int64_t nchan, uint32_t *out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    const RxChanState &S = st[c];
    uint32_t *o = out + c * 8;
    o[0] = (uint32_t)S.prbs_state; o[1] = S.prbs_idx; o[2] = S.prbs_bad; o[3] = S.prbs_good; o[4] = S.prbs_eq; o[5] = S.prbs_dif;
    o[6] = S.bert_bits; o[7] = S.bert_errs;
}
extern "C" int m17b_rx_get_bert(m17b_rx *rx, uint32_t *d_out, void *stream) {
    if (!rx || !d_out) return M17B_E_ARG;
    k_get_bert<<<grid_for(rx->nchan, 128), 128, 0, as_stream(stream)>>>(rx->d_state, rx->nchan, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}

extern "C" int m17b_rx_get_view(m17b_rx *rx, m17b_rx_view *v) {
    if (!rx || !v) return M17B_E_ARG;
    v->nchan = rx->nchan; v->nblocks = rx->last_blocks;
    v->d_frames = rx->d_frames; v->frame_cap = rx->fcap; v->d_nframes = rx->d_nframes;
    v->d_syms = rx->d_syms; v->sym_pitch = rx->sym_pitch; v->sym_carry = M17B_SYM_CARRY; v->d_nsym = rx->d_nsym;
    v->d_sym_base = rx->d_sym_base;
    v->d_disc = rx->seam_last ? nullptr : rx->d_disc; v->d_mean = rx->seam_last ? nullptr : rx->d_mean;
    v->d_events = rx->d_events; v->event_cap = rx->ecap; v->d_nevents = rx->d_nevents;
    v->d_stats = (const uint64_t *)rx->d_stats;
    return M17B_OK;
}
// bench instrumentation: when enabled, stage boundaries of m17b_dsp_rx / m17b_rx_baseband are marked with CUDA events
// on the launching stream; m17b_rx_stage_ms returns {front end, sync+framer, decode, post} of the last call (0 if absent).
extern "C" int m17b_rx_set_timing(m17b_rx *rx, int on) {
    if (!rx) return M17B_E_ARG;
    if (on && !rx->ev_stage[0][0])
        for (int r = 0; r < M17B_TIMING_RING; r++) for (int i = 0; i < 5; i++) CUDA_TRY(cudaEventCreate(&rx->ev_stage[r][i]));
    rx->timing = on != 0;
    rx->tcount = 0;
    return M17B_OK;
}
extern "C" int m17b_rx_stage_ms(m17b_rx *rx, int64_t call_index, float *out4) {
    if (!rx || !out4 || !rx->ev_stage[0][0] || call_index < 0 || call_index >= rx->tcount || call_index < rx->tcount - M17B_TIMING_RING) return M17B_E_ARG;
    cudaEvent_t *ev = rx->ev_stage[call_index % M17B_TIMING_RING];
    CUDA_TRY(cudaEventSynchronize(ev[4]));
    for (int i = 0; i < 4; i++) {
        out4[i] = 0;
        if (i == 0 && rx->seam_last) continue;
        CUDA_TRY(cudaEventElapsedTime(&out4[i], ev[i], ev[i + 1]));
    }
    return M17B_OK;
}
// channel groups (0 / 1 = the whole batch as one chain)
extern "C" int m17b_rx_set_chan_groups(m17b_rx *rx, int groups) {
    if (!rx || groups < -1 || groups > M17B_MAX_GROUPS) return M17B_E_ARG;
    rx->chan_groups = groups;
    return M17B_OK;
}
// blocks per pipeline slice (0 = no slicing: the stages run strictly one after the other)
extern "C" int m17b_rx_set_slice_blocks(m17b_rx *rx, int blocks) {
    if (!rx || blocks < 0) return M17B_E_ARG;
    rx->slice_blocks = blocks;
    return M17B_OK;
}
extern "C" int64_t m17b_rx_frame_cap(const m17b_rx *rx) { return rx ? rx->fcap : 0; }
// sticky capacity flags since the last m17b_rx_reset (synchronises the device): bit 0 = m17b_rx_symbols was handed more symbols
// for a channel than max_blocks * 200; the excess was ignored.  (The record / event buffers are sized for the worst case of either
// entry point, so they cannot overflow.)
extern "C" int m17b_rx_get_overflow(m17b_rx *rx, int *h_flags) {
    if (!rx || !h_flags) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h_flags, rx->d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
    return M17B_OK;
}
extern "C" int m17b_rx_last_launches(const m17b_rx *rx) { return rx ? rx->last_launches : 0; }

// The chain for channels [c0, c0+nc) in TIME pieces [bounds[k], bounds[k+1]), piece k starting once ready[k] has
// fired (its samples have arrived): front end, timing loop + framer and frame decode per piece (the kernels' block-range forms,
// results identical to one pass), the per-channel post stage once at the end.
static int rx_pipeline_pieces(m17b_rx *rx, int64_t c0, int64_t nc, const int16_t *d_iq, int64_t T, int npieces, const int64_t *bounds, const cudaEvent_t *ready,
                              cudaStream_t st) {
    m17b_ctx *ctx = rx->ctx;
    float *disc_w = rx->d_disc + c0 * T * 384, *mean_w = rx->d_mean + c0 * T;
    m17b_frame_rec *frames = rx->d_frames + c0 * rx->fcap;
    float *syms = rx->d_syms + c0 * rx->sym_pitch;
    if (npieces > M17B_MAX_SLICES) return M17B_E_ARG;
    // Three streams, as in the time-sliced pipeline: the front end of a piece starts when its samples have landed -- a lane walks a
    // whole 1920-sample block, ~0.27 ms whatever the piece size -- beside the timing loop of the piece before it; the timing loop
    // and the decode of the pieces stay in order.  What remains after the last byte is one front-end latency + the chain over the
    // last piece.
    CUDA_TRY(cudaEventRecord(rx->ev_start, st));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_fe, rx->ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_sync, rx->ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(rx->s_dec, rx->ev_start, 0));
    for (int k = 0; k < npieces; k++) {
        const int64_t t0 = bounds[k], t1 = bounds[k + 1];
        int2 *rng = rx->d_frame_rng + (int64_t)k * rx->nchan + c0;
        CUDA_TRY(cudaStreamWaitEvent(rx->s_fe, ready[k], 0));
        int rc = launch_frontend(d_iq, nc, T, t0, t1 - t0, rx->d_state + c0, disc_w, mean_w, rx->s_fe);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(rx->ev_fe[k], rx->s_fe));
        CUDA_TRY(cudaStreamWaitEvent(rx->s_sync, rx->ev_fe[k], 0));
        rc = launch_sync(rx, c0, nc, disc_w, mean_w, T, (int)t0, (int)t1, rng, 1, rx->s_sync);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(rx->ev_sy[k], rx->s_sync));
        CUDA_TRY(cudaStreamWaitEvent(rx->s_dec, rx->ev_sy[k], 0));
        const int64_t span = t1 - t0;
        rc = launch_decode(ctx, syms, rx->sym_pitch, M17B_SYM_CARRY, rx->d_sym_base + c0, frames, rx->fcap, rx->d_nframes + c0, nc, nullptr, rx->d_ssoft + c0 * rx->fcap * STREAM_NIN, rx->d_saux + c0 * rx->fcap, rx->d_dlist + c0 * (rx->fcap + 1), rx->s_dec,
                           rx->aux_stream, rx->ev_fork, rx->ev_join, rng, span + span / 64 + 4, rx->bert);
        if (rc) return rc;
        rx->last_launches += 4;
    }
    k_post<<<grid_for(nc, POST_WARPS), POST_WARPS * 32, 0, rx->s_dec>>>(frames, rx->fcap, rx->d_nframes + c0, nc, rx->d_state + c0, ctx->d_crcpos, rx->d_stats + c0 * 8,
                                                              rx->d_lsf_snap + c0 * rx->nsnap * 32, rx->nsnap, rx->d_lsf_ver + c0 * rx->fcap,
                                                              ctx->d_prbs, rx->bert);
    KERNEL_CHECK();
    rx->last_launches += 1;
    CUDA_TRY(cudaEventRecord(rx->ev_end, rx->s_dec));
    CUDA_TRY(cudaStreamWaitEvent(st, rx->ev_end, 0));
    return M17B_OK;
}

// End-to-end entry point with host buffers: channels are processed in chunks so the H2D copy of chunk k+1
// overlaps the kernels of chunk k (two staging buffers, a dedicated copy stream); records stream back per chunk.
extern "C" int m17b_dsp_rx_host(m17b_rx *rx, const int16_t *h_iq, int64_t nblocks, m17b_frame_rec *h_frames, int32_t *h_nframes, void *stream) {
    if (!rx || !h_iq || !h_frames || !h_nframes || nblocks <= 0) return M17B_E_ARG;
    if (nblocks > rx->max_blocks) return M17B_E_CAPACITY;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    cudaStream_t st = as_stream(stream);
    const int64_t T = nblocks;
    if (!rx->copy_stream) {
        // chunks of about 128 MiB of IQ per staging buffer (at least 32 channels), all of (nearly) the same size: every chunk costs
        // one full latency of the 250-block timing loop whatever its size, so a ragged little last chunk would add ~1 ms to the
        // tail after the last copy (measured: 34-channel chunks + a 4-channel one 37.3 ms, even chunks 36.5 ms; PCIe alone 35.4)
        int64_t per_chan = rx->max_blocks * 7680;
        int64_t mib = 128;
        if (const char *e = getenv("M17B_HOST_CHUNK_MIB")) mib = atoi(e) > 0 ? atoi(e) : 128;
        int64_t chunk = (mib << 20) / per_chan;
        if (chunk < 32) chunk = 32;
        if (chunk > rx->nchan) chunk = rx->nchan;
        const int64_t nchunks = (rx->nchan + chunk - 1) / chunk;
        chunk = (rx->nchan + nchunks - 1) / nchunks;              // even split: sizes differ by at most one channel
        rx->stage_chunk = chunk;
        CUDA_TRY(cudaStreamCreateWithFlags(&rx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaMalloc((void **)&rx->d_iq_stage[i], (size_t)chunk * per_chan));
            CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_h2d[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < M17B_HOST_TAIL_PIECES; i++) CUDA_TRY(cudaEventCreateWithFlags(&rx->ev_piece[i], cudaEventDisableTiming));
    }
    rx->last_launches = 0; rx->last_blocks = nblocks; rx->seam_last = 0;
    const int64_t nchunks = (rx->nchan + rx->stage_chunk - 1) / rx->stage_chunk;
    int k = 0;
    for (int64_t ci = 0; ci < nchunks; ci++, k ^= 1) {
        const int64_t c0 = ci * rx->nchan / nchunks, nc = (ci + 1) * rx->nchan / nchunks - c0;     // <= stage_chunk
        // the staging buffer may only be overwritten once the kernels that read it two chunks ago are done
        CUDA_TRY(cudaStreamWaitEvent(rx->copy_stream, rx->ev_done[k], 0));
        const int64_t piece_min = 25;
        if (ci == nchunks - 1 && !rx->afc && T >= 4 * piece_min) {
            // The LAST chunk arrives in up to six time pieces (each at least 25 blocks), and every piece is processed while the
            // later ones are still on the bus: what remains after the final byte has landed is the chain over the last piece
            // instead of the timing loop's latency over all T (every channel's blocks are serial), which is the tail of the whole
            // call.  (Two pieces, the second of 25 blocks, run on one stream: 36.4 ms; pieces pipelined on three streams: see DESIGN.md.)
            int np = (int)(T / piece_min);
            if (np > M17B_HOST_TAIL_PIECES) np = M17B_HOST_TAIL_PIECES;
            int64_t bounds[M17B_HOST_TAIL_PIECES + 1];
            cudaEvent_t ready[M17B_HOST_TAIL_PIECES];
            for (int q = 0; q <= np; q++) bounds[q] = q * T / np;
            const size_t pitch = (size_t)T * 7680;
            const char *src = (const char *)(h_iq + c0 * T * 3840);
            char *dst = (char *)rx->d_iq_stage[k];
            for (int q = 0; q < np; q++) {
                CUDA_TRY(cudaMemcpy2DAsync(dst + bounds[q] * 7680, pitch, src + bounds[q] * 7680, pitch, (size_t)(bounds[q + 1] - bounds[q]) * 7680, (size_t)nc,
                                           cudaMemcpyHostToDevice, rx->copy_stream));
                CUDA_TRY(cudaEventRecord(rx->ev_piece[q], rx->copy_stream));
                ready[q] = rx->ev_piece[q];
            }
            int rc = rx_pipeline_pieces(rx, c0, nc, rx->d_iq_stage[k], T, np, bounds, ready, st);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(rx->ev_done[k], st));
            CUDA_TRY(cudaMemcpyAsync(h_frames + c0 * rx->fcap, rx->d_frames + c0 * rx->fcap, sizeof(m17b_frame_rec) * nc * rx->fcap, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_nframes + c0, rx->d_nframes + c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, st));
            continue;
        }
        CUDA_TRY(cudaMemcpyAsync(rx->d_iq_stage[k], h_iq + c0 * T * 3840, (size_t)nc * T * 7680, cudaMemcpyHostToDevice, rx->copy_stream));
        CUDA_TRY(cudaEventRecord(rx->ev_h2d[k], rx->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, rx->ev_h2d[k], 0));
        // note: the per-chunk front-end output lands at the chunk's own offset of d_disc (laid out for T = nblocks)
        // (the channel chunks already overlap copy and compute; no time slicing inside a chunk)
        int rc = rx_pipeline(rx, c0, nc, rx->d_iq_stage[k], nullptr, T, st, false);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(rx->ev_done[k], st));
        CUDA_TRY(cudaMemcpyAsync(h_frames + c0 * rx->fcap, rx->d_frames + c0 * rx->fcap, sizeof(m17b_frame_rec) * nc * rx->fcap, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_nframes + c0, rx->d_nframes + c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return M17B_OK;
}
