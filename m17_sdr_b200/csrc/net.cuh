// net.cuh -- the M17-over-UDP reflector frame as the batched wire format either side of the hot path:
//   "M17 " | stream id (2, big-endian) | LSF bytes 0..27 (dst 6, src 6, TYPE 2, META 14) | FN (2) | payload 16 | CRC-16 = 54 bytes.
// Replaces net_add_magic/_stream_id/_lich/_fn/_payload/_crc and m17_net_new_rx_data (m17_net.cpp:25-74),
// build_lich_to_net / m17_send_stream_frame_to_net (m17_tx_routines.cpp:54-70,298-306) on the way out, and
// m17_parse_m17_data (m17_net.cpp:203-238) / build_lich_from_net (m17_tx_routines.cpp:71-86) on the way in.
// Sockets, the reflector handshake (CONN/ACKN/PING/PONG/DISC) and threads stay with the host application: out of scope.
// One thread per frame; byte work, HBM-trivial (54 B out per 64-B record in).
#pragma once
#include "app.cuh"

__device__ __forceinline__ void net_write_frame(uint8_t *o, uint16_t sid, const uint8_t *lsf28, int have_dst, uint64_t dst, const uint8_t *fn_pld18,
                                                const uint16_t *tab) {
    uint8_t b[54];
    b[0] = 0x4D; b[1] = 0x31; b[2] = 0x37; b[3] = 0x20;                       // net_add_magic
    b[4] = (uint8_t)(sid >> 8); b[5] = (uint8_t)sid;                          // net_add_stream_id
#pragma unroll
    for (int i = 0; i < 28; i++) b[6 + i] = lsf28[i];                         // net_add_lich
    if (have_dst) {
#pragma unroll
        for (int i = 0; i < 6; i++) b[6 + i] = (uint8_t)(dst >> (40 - 8 * i));  // gateway destination (m17_net.cpp:56-61)
    }
#pragma unroll
    for (int i = 0; i < 18; i++) b[34 + i] = fn_pld18[i];                     // net_add_fn + net_add_payload
    uint16_t k = 0xFFFF;
#pragma unroll
    for (int i = 0; i < 52; i++) k = crc16_step(k, b[i], tab);                // net_add_crc
    b[52] = (uint8_t)(k >> 8); b[53] = (uint8_t)k;
    uint16_t *o2 = (uint16_t *)o;                                             // 54-byte frames are 2-byte aligned
#pragma unroll
    for (int i = 0; i < 27; i++) o2[i] = (uint16_t)(b[2 * i] | (b[2 * i + 1] << 8));
}

__global__ void k_net_pack(const uint16_t *__restrict__ sid, const uint8_t *__restrict__ lsf, int64_t lsf_stride, int have_dst, uint64_t dst,
                           const uint16_t *__restrict__ fn, const uint8_t *__restrict__ payload, int64_t n, uint8_t *__restrict__ out,
                           const uint16_t *__restrict__ g_crc) {
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_crc[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint8_t fp[18];
    fp[0] = (uint8_t)(fn[r] >> 8); fp[1] = (uint8_t)fn[r];
    for (int i = 0; i < 16; i++) fp[2 + i] = payload[r * 16 + i];
    net_write_frame(out + r * 54, sid[r], lsf + r * lsf_stride, have_dst, dst, fp, tab);
}

__global__ void k_net_parse(const uint8_t *__restrict__ in, int64_t n, uint8_t *__restrict__ ok, uint16_t *__restrict__ sid, uint8_t *__restrict__ lsf30,
                            uint16_t *__restrict__ fn, uint8_t *__restrict__ payload, const uint16_t *__restrict__ g_crc) {
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_crc[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint16_t *i2 = (const uint16_t *)(in + r * 54);
    uint8_t b[54];
#pragma unroll
    for (int i = 0; i < 27; i++) { const uint16_t v = i2[i]; b[2 * i] = (uint8_t)v; b[2 * i + 1] = (uint8_t)(v >> 8); }
    uint16_t k = 0xFFFF, kl = 0xFFFF;
#pragma unroll
    for (int i = 0; i < 54; i++) k = crc16_step(k, b[i], tab);
    ok[r] = (k == 0) && b[0] == 0x4D && b[1] == 0x31 && b[2] == 0x37 && b[3] == 0x20;   // m17_net_parse_msg :291 + m17_parse_m17_data :204
    sid[r] = (uint16_t)((b[4] << 8) | b[5]);
    fn[r] = (uint16_t)((b[34] << 8) | b[35]);
#pragma unroll
    for (int i = 0; i < 28; i++) { lsf30[r * 30 + i] = b[6 + i]; kl = crc16_step(kl, b[6 + i], tab); }   // build_lich: fresh CRC
    lsf30[r * 30 + 28] = (uint8_t)(kl >> 8); lsf30[r * 30 + 29] = (uint8_t)kl;
    for (int i = 0; i < 16; i++) payload[r * 16 + i] = b[36 + i];
}

// one thread per RX record: delivered stream frames become datagrams, compacted per channel in record order
__global__ void k_net_from_records(const m17b_frame_rec *__restrict__ frames, int64_t fcap, const int32_t *__restrict__ nframes, int64_t nchan,
                                   const uint8_t *__restrict__ lsf_snap, int nsnap, const uint8_t *__restrict__ lsf_ver,
                                   const uint16_t *__restrict__ sid, int have_dst, uint64_t dst, uint8_t *__restrict__ out, int32_t *__restrict__ count,
                                   const uint16_t *__restrict__ g_crc) {
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_crc[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per channel
    if (c >= nchan) return;
    const int n = nframes[c];
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        bool want = false;
        const m17b_frame_rec *r = frames + c * fcap + k;
        if (k < n) want = r->type == M17B_T_STREAM && (r->flags & M17B_F_DELIVERED);   // m17_rx_parse.cpp:148-154
        const unsigned m = __ballot_sync(0xffffffffu, want);
        if (want) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            net_write_frame(out + (c * fcap + slot) * 54, sid[c], lsf_snap + (c * nsnap + lsf_ver[c * fcap + k]) * 32, have_dst, dst, r->data, tab);
        }
        base += __popc(m);
    }
    if (lane == 0) count[c] = base;
}

extern "C" int m17b_net_pack(m17b_ctx *ctx, const uint16_t *d_sid, const uint8_t *d_lsf, int64_t lsf_stride, int have_dst, uint64_t dst,
                             const uint16_t *d_fn, const uint8_t *d_payload, int64_t n, uint8_t *d_out, void *stream) {
    if (!ctx || !d_sid || !d_lsf || !d_fn || !d_payload || !d_out || n < 0 || lsf_stride < 28) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_net_pack<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_sid, d_lsf, lsf_stride, have_dst, dst, d_fn, d_payload, n, d_out, ctx->d_crc);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_net_parse(m17b_ctx *ctx, const uint8_t *d_in, int64_t n, uint8_t *d_ok, uint16_t *d_sid, uint8_t *d_lsf30, uint16_t *d_fn,
                              uint8_t *d_payload, void *stream) {
    if (!ctx || !d_in || !d_ok || !d_sid || !d_lsf30 || !d_fn || !d_payload || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_net_parse<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_in, n, d_ok, d_sid, d_lsf30, d_fn, d_payload, ctx->d_crc);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_rx_net_frames(m17b_rx *rx, const uint16_t *d_sid, int have_dst, uint64_t dst, uint8_t *d_out, int32_t *d_count, void *stream) {
    if (!rx || !d_sid || !d_out || !d_count) return M17B_E_ARG;
    k_net_from_records<<<grid_for(rx->nchan, 4), 128, 0, as_stream(stream)>>>(rx->d_frames, rx->fcap, rx->d_nframes, rx->nchan, rx->d_lsf_snap, rx->nsnap,
                                                                             rx->d_lsf_ver, d_sid, have_dst, dst, d_out, d_count, rx->ctx->d_crc);
    KERNEL_CHECK();
    return M17B_OK;
}

// ================================================================ remaining small m17defines.h entry points, batched
// m17_dsp_demap_symbol (m17_dsp.cpp:35-42): n symbols with their own normaliser -> n x {MSB soft, LSB soft}
__global__ void k_demap_symbols(const float *__restrict__ in, const float *__restrict__ mag, int64_t n, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = demap_soft(in[i], mag[i], false);
    out[2 * i + 1] = demap_soft(in[i], mag[i], true);
}
extern "C" int m17b_demap_symbols(m17b_ctx *ctx, const float *d_in, const float *d_mag, int64_t n, float *d_out, void *stream) {
    if (!ctx || !d_in || !d_mag || !d_out || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_demap_symbols<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(d_in, d_mag, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
// m17_dsp_decimating_filter (m17_dsp.cpp:438-449): out[k] = sum_j in[k*stride + j] * coffs[j], the sum started at 0 and added in
// order (no contraction); n rows of `len` samples (+ flen - 1 of look-ahead, as upstream reads them), in_pitch apart
__global__ void k_decimating_filter(const float *__restrict__ in, int64_t in_pitch, const float *__restrict__ coffs, int stride, int flen, int nout,
                                    int64_t n, float *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nout) return;
    const int64_t r = t / nout; const int k = (int)(t % nout);
    const float *s = in + r * in_pitch + (int64_t)k * stride;
    float sum = 0;
    for (int j = 0; j < flen; j++) sum += s[j] * coffs[j];
    out[t] = sum;
}
extern "C" int m17b_dsp_decimating_filter(m17b_ctx *ctx, const float *d_in, int64_t in_pitch, const float *d_coffs, int stride, int flen, int len, int64_t n,
                                          float *d_out, int *out_len, void *stream) {
    if (!ctx || !d_in || !d_coffs || !d_out || stride < 1 || flen < 1 || len < 0 || n < 0) return M17B_E_ARG;
    const int nout = (len + stride - 1) / stride;                  // for (i = 0; i < len; i += stride)
    if (out_len) *out_len = nout;
    if (n == 0 || nout == 0) return M17B_OK;
    k_decimating_filter<<<grid_for(n * nout, 256), 256, 0, as_stream(stream)>>>(d_in, in_pitch, d_coffs, stride, flen, nout, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
// m17_prbs9_rx_check (m17_prbs9.cpp:40-64) over n independent bit sequences with persistent checker state
// d_state uint32 [n][8]: m_rx_state, m_rx_idx, m_rx_bad, m_rx_good, m_rx_eq_cnt, m_rx_dif_cnt, bits checked in sync, bit errors
// (zeros = m17_prbs9_rx_reset on a fresh process)
__global__ void k_prbs9_rx_check(const uint8_t *__restrict__ bits, int nbits, int64_t n, uint32_t *st, const uint8_t *__restrict__ g_prbs) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t *o = st + r * 8;
    int state = (int)o[0], idx = (int)o[1];
    unsigned bad = o[2], good = o[3], eq = o[4], dif = o[5], nb = o[6], ne = o[7];
    for (int b = 0; b < nbits; b++) {
        const unsigned d = (bits[r * nbits + b] ^ g_prbs[idx]) & 1u;
        if (d) { dif = (dif + 1) & 0xFFFF; eq = 0; } else { eq = (eq + 1) & 0xFFFF; dif = 0; }
        idx = idx + 1 == 511 ? 0 : idx + 1;
        if (state == 0) {
            bad = 0; good = 0;
            if (eq >= 18) state = 1;
            if (d) { idx = 0; state = 0; }
        } else {
            nb++; ne += d;
            if (dif >= 18) state = 0;
            if (d) bad = (bad + 1) & 0xFFFF;
        }
    }
    o[0] = (uint32_t)state; o[1] = (uint32_t)idx; o[2] = bad; o[3] = good; o[4] = eq; o[5] = dif; o[6] = nb; o[7] = ne;
}
extern "C" int m17b_prbs9_rx_check(m17b_ctx *ctx, const uint8_t *d_bits, int nbits, int64_t n, uint32_t *d_state, void *stream) {
    if (!ctx || !d_bits || !d_state || nbits < 0 || n < 0) return M17B_E_ARG;
    if (n == 0 || nbits == 0) return M17B_OK;
    k_prbs9_rx_check<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_bits, nbits, n, d_state, ctx->d_prbs);
    KERNEL_CHECK();
    return M17B_OK;
}
// eq_restart (m17_equalize.cpp:142-145): the U/D factors start over, the tap weights stay
__global__ void k_eq_restart(EqState *st, int64_t nchan) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    EqState &e = st[c];
    for (int j = 0; j < 5; j++) { for (int i = 0; i < j; i++) e.u[i][j] = 0.0f; e.d[j] = 0.1f; }   // eq_k_reset_ud :25-36
}
extern "C" int m17b_eq_restart(m17b_eq *eq, void *stream) {
    if (!eq) return M17B_E_ARG;
    k_eq_restart<<<grid_for(eq->nchan, 128), 128, 0, as_stream(stream)>>>(eq->d_state, eq->nchan);
    KERNEL_CHECK();
    return M17B_OK;
}
