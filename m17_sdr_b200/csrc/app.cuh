// app.cuh -- the application layer just behind the frame decode: packet reassembly with the packet CRC,
// and the GNSS position carried in the LSF META field.
//   * m17b_rx_reassemble_packets: what parse_packet (m17_rx_parse.cpp:34-51) is meant to do.  The reference's version stores
//     non-final chunks at fn*25 but the final chunk at the index of the LAST chunk's start and checks the CRC of a
//     buffer that therefore never holds the whole packet; the per-channel state reproduces that bug for parity (k_post).  Here
//     the 25-byte chunks of non-final frames and the `count` bytes of the EOF frame are concatenated in arrival order and the
//     CRC-16 appended by m17_send_packet_frames (m17_tx_routines.cpp:323-353) is checked over the whole packet.
//   * m17b_gps_decode: gps_decode (gps.cpp:8-27) on the 14 META bytes of an LSF.
#pragma once
#include "tx.cuh"

struct RxPacketState { int32_t len; uint8_t buf[828]; };      // a packet being collected (32 frames x 25 bytes + CRC at most)

// One warp per channel walks the channel's records of the last call in order.  A packet frame's 25 (or `count`) bytes are
// copied by the lanes; on EOF lane 0 runs the CRC over the collected bytes and the packet is published:
//   pkt[c][k] = {offset into bytes[c], length without the CRC, crc_ok}.
__global__ void __launch_bounds__(128) k_reassemble(const m17b_frame_rec *__restrict__ frames, int64_t fcap, const int32_t *__restrict__ nframes, int64_t nchan,
                                                    RxPacketState *st, const uint16_t *__restrict__ g_crc, uint8_t *bytes, int64_t bytes_cap, int32_t *pkt, int max_pkts,
                                                    int32_t *npkt) {
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_crc[i];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * 4 + wid;
    if (c >= nchan) return;
    RxPacketState *S = st + c;
    int len = S->len, np = 0;
    int64_t off = 0;
    const int n = nframes[c];
    for (int k = 0; k < n; k++) {
        const m17b_frame_rec *r = frames + c * fcap + k;
        if (r->type != M17B_T_PACKET || !(r->flags & M17B_F_PARSED)) continue;
        const int meta = r->data[25];
        const int eof = meta >> 7, cnt = eof ? ((meta >> 2) & 0x1F) : 25;
        const int take = cnt > 25 ? 25 : cnt;
        if (len + take <= 828) { if (lane < take) S->buf[len + lane] = r->data[lane]; len += take; }
        else len = 829;                                                       // overlong: dropped at its EOF
        __syncwarp();
        if (eof) {
            int ok = 0;
            if (len >= 2 && len <= 828) {
                if (lane == 0) { uint16_t crc = 0xFFFF; for (int i = 0; i < len; i++) crc = crc16_step(crc, S->buf[i], tab); ok = crc == 0; }
                ok = __shfl_sync(0xffffffffu, ok, 0);
                const int plen = len - 2;
                if (np < max_pkts && off + plen <= bytes_cap) {
                    for (int i = lane; i < plen; i += 32) bytes[c * bytes_cap + off + i] = S->buf[i];
                    if (lane == 0) { int32_t *e = pkt + (c * max_pkts + np) * 3; e[0] = (int32_t)off; e[1] = plen; e[2] = ok; }
                    off += plen;
                    np++;
                }
            } else if (np < max_pkts) {
                if (lane == 0) { int32_t *e = pkt + (c * max_pkts + np) * 3; e[0] = (int32_t)off; e[1] = 0; e[2] = 0; }   // shorter than a CRC, or overlong: no payload
                np++;
            }
            len = 0;
            __syncwarp();
        }
    }
    if (lane == 0) { S->len = len; npkt[c] = np; }
}
extern "C" int m17b_rx_reassemble_packets(m17b_rx *rx, uint8_t *d_bytes, int64_t bytes_cap, int32_t *d_pkt, int max_pkts, int32_t *d_npkt, void *stream) {
    if (!rx || !d_bytes || !d_pkt || !d_npkt || bytes_cap <= 0 || max_pkts <= 0) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(rx->ctx->device));
    if (!rx->d_pkt_state) {
        CUDA_TRY(cudaMalloc((void **)&rx->d_pkt_state, sizeof(RxPacketState) * rx->nchan));
        CUDA_TRY(cudaMemsetAsync(rx->d_pkt_state, 0, sizeof(RxPacketState) * rx->nchan, as_stream(stream)));
    }
    k_reassemble<<<grid_for(rx->nchan, 4), 128, 0, as_stream(stream)>>>(rx->d_frames, rx->fcap, rx->d_nframes, rx->nchan, (RxPacketState *)rx->d_pkt_state, rx->ctx->d_crc,
                                                                     d_bytes, bytes_cap, d_pkt, max_pkts, d_npkt);
    KERNEL_CHECK();
    return M17B_OK;
}

// gps_decode (gps.cpp:8-27): b = the 14 META bytes of an LSF (the function reads one byte past them: the first CRC byte,
// lsf[28], is the low byte of the 48-bit course / speed / object word -- reproduced, the caller passes whole 30-byte LSFs)
__global__ void k_gps_decode(const uint8_t *__restrict__ lsf, int64_t stride, int64_t n, m17b_gps_rec *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *b = lsf + i * stride + 14;
    m17b_gps_rec g;
    g.lat = (double)(int8_t)b[0] + (double)(((uint32_t)b[1] << 8) | b[2]) / 65536.0;
    g.lon = (double)(int16_t)(((uint32_t)b[3] << 8) | b[4]) + (double)(((uint32_t)b[5] << 8) | b[6]) / 65536.0;
    g.alt = (int32_t)(int16_t)((int)(((uint32_t)b[7] << 8) | b[8]) - 1500);
    uint64_t w = 0;
    for (int k = 0; k < 6; k++) w = (w << 8) | b[9 + k];
    g.course = (int32_t)(uint16_t)(w >> 38);
    g.speed = (int32_t)((w >> 28) & 0x3FF);
    g.object = (int32_t)(w & 0xFFFFF);
    out[i] = g;
}
extern "C" int m17b_gps_decode(m17b_ctx *ctx, const uint8_t *d_lsf, int64_t stride, int64_t n, m17b_gps_rec *d_out, void *stream) {
    if (!ctx || !d_lsf || !d_out || n < 0 || stride < 30) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_gps_decode<<<grid_for(n, 128), 128, 0, as_stream(stream)>>>(d_lsf, stride, n, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}
