// tx.cuh -- batched TX: frame formatters (CRC'd LSF / stream / packet / BERT -> 192 dibits) and the 4FSK
// modulator (polyphase RRC x os -> fp32 phase accumulator -> cos/sin -> int16 IQ).
// Replaces m17_fmt_add_* (m17_tx_routines.cpp:24-31,92-117,143-187,201-255) and m17_mod_dibits / mod_filter /
// sub_filter / mod_fsk (m17_modulate.cpp:22-61,79-92).
//
// Formatter mapping: one thread per OUTPUT DIBIT (see k_fmt): interleave + randomise + puncture are index maps and a
// convolutional-code output bit is the parity of a masked 5-bit window, so every dibit is computed independently.
// Modulator: one fused kernel (mod.cuh): FIR by worker warps, the per-channel fp32 phase chain by a scan warp, cos/sin and
// the int16 stores by the workers again, all through shared memory.
#pragma once
#include "rx.cuh"

struct TxChanState {
    float hist[31];      // deviation of the last 31 symbols = m_tx_s (m17_modulate.cpp:8)
    float acc;           // m_acc (m17_modulate.cpp:14)
    uint8_t lich[32];    // m_lich (m17_tx_routines.cpp:13)
    int lich_count, fn, prbs_idx;
};
struct m17b_tx {
    m17b_ctx *ctx;
    int64_t nchan;
    int os;
    float *d_taps;       // 31*os
    float *d_devtab;     // dibit -> deviation (rad/sample), m17_modulate.cpp:9; [4] = blank carrier
    TxChanState *d_state;
    unsigned long long *d_dbg;   // [8] instrumentation of the last m17b_mod_dibits call (m17b_tx_debug_scan)
};
#include "mod.cuh"

// ---------------------------------------------------------------- formatters
// One thread per OUTPUT DIBIT of a frame, a CTA of 192 threads walks frames in batches of FMT_NFB.  Every final bit is
// "parity of (5-bit window of the frame's bit stream AND mask)": a convolutional-code bit is the parity of 3 or 4 of the 5
// register bits (m17_conv.cpp:22-31), a LICH Golay bit is a window with a one-bit mask.  Puncture, interleave and randomise
// are index maps, so a thread's two (window position, mask, randomiser bit) descriptors depend only on its dibit index: they
// are derived once per thread from the constant-memory maps and reused for every frame the CTA formats.  Per frame the CTA
// stages a 48-byte record in shared memory: bytes 0..31 the info bits delayed by 4 (the encoder's zero start / zero tail fall
// out of the padding), bytes 32..43 the four 24-bit Golay words of a stream frame.
#define FMT_NFB 16
#define FMT_PITCH 48
#define FMT_GOLAY_BYTE 32
template <int MODE> __device__ __forceinline__ uint32_t fmt_info_byte(const uint8_t *__restrict__ src, const uint8_t *__restrict__ meta, const uint8_t *__restrict__ prbs,
                                                                      int64_t f, int fn, int prbs0, int m) {
    if (m < 0) return 0;
    if (MODE == 1) return m < 30 ? src[f * 30 + m] : 0;                                     // LSF, 240 bits
    if (MODE == 2) return m == 0 ? (uint32_t)(fn >> 8) & 255u : m == 1 ? (uint32_t)fn & 255u : m < 18 ? src[f * 16 + m - 2] : 0;   // FN(16) + payload(128)
    if (MODE == 3) return m < 25 ? src[f * 25 + m] : m == 25 ? meta[f] : 0;                  // chunk(200) + meta(8)
    uint32_t v = 0;                                                                         // BERT, 197 PRBS9 bits
#pragma unroll
    for (int r = 0; r < 8; r++) { const int q = 8 * m + r; v = (v << 1) | (q < 197 ? (uint32_t)prbs[(prbs0 + q) % 511] : 0u); }
    return v;
}
template <int MODE>
__global__ void __launch_bounds__(192) k_fmt(const uint8_t *__restrict__ src, const uint8_t *__restrict__ meta, int64_t nframes, int64_t F, const TxChanState *__restrict__ st,
                                             uint8_t *__restrict__ dibits, const uint16_t *__restrict__ genc, const uint8_t *__restrict__ prbs) {
    __shared__ __align__(16) uint8_t rec[FMT_NFB][FMT_PITCH];
    const int d = threadIdx.x;
    constexpr uint32_t SYNCW = MODE == 1 ? 0x55F7u : MODE == 2 ? 0xFF5Du : MODE == 3 ? 0x75FFu : 0xDF55u;
    // this thread's two bit descriptors: interleave (out[pi(j)] = in[j], pi an involution), then randomise, dibit = b[i]<<1 | b[i+1]
    int pos[2] = {0, 0}; uint32_t msk[2] = {0, 0}, rnd[2] = {0, 0};
    if (d >= 8) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int i = 2 * (d - 8) + e;
            const int j = c_tx.qpp[i];
            rnd[e] = c_tx.rnd[i];
            if (MODE == 2 && j < 96) { pos[e] = 8 * FMT_GOLAY_BYTE + j; msk[e] = 0x10u; }          // pack_24_to_1 of the Golay words
            else {
                const int p = MODE == 1 ? c_tx.unp1[j] : MODE == 2 ? c_tx.unp2[j - 96] : MODE == 3 ? c_tx.unp3[j] : c_tx.unp2[j];
                pos[e] = p >> 1;                            // trellis step t: window = info bits t-4 .. t
                msk[e] = (p & 1) ? 0x17u : 0x19u;           // G2 = b[t]^b[t-1]^b[t-2]^b[t-4], G1 = b[t]^b[t-3]^b[t-4]  (window MSB = b[t-4])
            }
        }
    }
    const int by0 = pos[0] >> 3, sh0 = 11 - (pos[0] & 7), by1 = pos[1] >> 3, sh1 = 11 - (pos[1] & 7);
    for (int64_t base = (int64_t)blockIdx.x * FMT_NFB; base < nframes; base += (int64_t)gridDim.x * FMT_NFB) {
        __syncthreads();
        for (int idx = d; idx < FMT_NFB * 44; idx += 192) {
            const int fl = idx / 44, n = idx - fl * 44;
            const int64_t f = base + fl;
            if (f >= nframes) continue;
            uint32_t v = 0;
            int fn = 0, prbs0 = 0, lc = 0;
            const TxChanState *S = nullptr;
            if (MODE == 2 || MODE == 4) {
                const int k = (int)(f % F);
                S = st + f / F;
                fn = (S->fn + k) & 0xFFFF;
                lc = (S->lich_count + k) % 6;
                prbs0 = (S->prbs_idx + k * 197) % 511;
            }
            if (n < 32) {
                v = ((fmt_info_byte<MODE>(src, meta, prbs, f, fn, prbs0, n - 1) << 4) | (fmt_info_byte<MODE>(src, meta, prbs, f, fn, prbs0, n) >> 4)) & 255u;
            } else if (MODE == 2) {
                // LICH chunk lc of m_lich, byte 5 = lc << 5 -> four 12-bit words -> m17_golay_encode (m17_tx_routines.cpp:151-164)
                const int w = (n - FMT_GOLAY_BYTE) / 3, kb = (n - FMT_GOLAY_BYTE) - 3 * w;
                uint32_t b[6];
#pragma unroll
                for (int i = 0; i < 5; i++) b[i] = S->lich[lc * 5 + i];
                b[5] = (uint32_t)(lc & 7) << 5;
                const uint32_t dw = w == 0 ? (b[0] << 4) | (b[1] >> 4) : w == 1 ? ((b[1] & 15u) << 8) | b[2] : w == 2 ? (b[3] << 4) | (b[4] >> 4) : ((b[4] & 15u) << 8) | b[5];
                const uint32_t g = (dw << 12) | __ldg(&genc[dw]);
                v = (g >> (16 - 8 * kb)) & 255u;
            }
            rec[fl][n] = (uint8_t)v;
        }
        __syncthreads();
#pragma unroll 4
        for (int fl = 0; fl < FMT_NFB; fl++) {
            const int64_t f = base + fl;
            if (f >= nframes) break;
            uint32_t out;
            if (d < 8) out = (SYNCW >> (14 - 2 * d)) & 3u;                                          // pack_16_to_2
            else {
                const uint8_t *r = rec[fl];
                const uint32_t w0 = ((uint32_t)r[by0] << 8) | r[by0 + 1], w1 = ((uint32_t)r[by1] << 8) | r[by1 + 1];
                const uint32_t b0 = (__popc((w0 >> sh0) & msk[0]) & 1u) ^ rnd[0], b1 = (__popc((w1 >> sh1) & msk[1]) & 1u) ^ rnd[1];
                out = (b0 << 1) | b1;
            }
            dibits[f * 192 + d] = (uint8_t)out;
        }
    }
}
static inline unsigned fmt_grid(int64_t nframes) {
    const int64_t batches = (nframes + FMT_NFB - 1) / FMT_NFB;
    return (unsigned)(batches < 148 * 8 ? batches : 148 * 8);
}
__global__ void k_tx_advance(TxChanState *st, int64_t nchan, int dfn, int dlich, int dprbs) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    st[c].fn = (st[c].fn + dfn) & 0xFFFF;
    st[c].lich_count = (st[c].lich_count + dlich) % 6;
    st[c].prbs_idx = (st[c].prbs_idx + dprbs) % 511;
}
__global__ void k_tx_set_lsf(TxChanState *st, int64_t nchan, const uint8_t *lsf) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nchan * 32) return;
    int64_t c = gid / 32;
    int i = (int)(gid % 32);
    st[c].lich[i] = i < 30 ? lsf[c * 30 + i] : 0;
    if (i == 0) { st[c].lich_count = 0; st[c].fn = 0; }                                              // m17_tx_routines.cpp:98-99
}

extern "C" int m17b_fmt_preamble(uint8_t *d) { if (!d) return M17B_E_ARG; for (int i = 0; i < 192; i += 2) { d[i] = 1; d[i + 1] = 3; } return M17B_OK; }
extern "C" int m17b_fmt_eot(uint8_t *d) { if (!d) return M17B_E_ARG; for (int i = 0; i < 192; i++) d[i] = (i & 7) == 6 ? 3 : 1; return M17B_OK; }
extern "C" int m17b_fmt_link_setup_frame(m17b_ctx *ctx, const uint8_t *d_lsf, int64_t n, uint8_t *d_dibits, void *stream) {
    if (!ctx || !d_lsf || !d_dibits || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_fmt<1><<<fmt_grid(n), 192, 0, as_stream(stream)>>>(d_lsf, nullptr, n, 1, nullptr, d_dibits, ctx->d_genc, ctx->d_prbs);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_fmt_packet_frames(m17b_ctx *ctx, const uint8_t *d_chunk, const uint8_t *d_meta, int64_t n, uint8_t *d_dibits, void *stream) {
    if (!ctx || !d_chunk || !d_meta || !d_dibits || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    k_fmt<3><<<fmt_grid(n), 192, 0, as_stream(stream)>>>(d_chunk, d_meta, n, 1, nullptr, d_dibits, ctx->d_genc, ctx->d_prbs);
    KERNEL_CHECK();
    return M17B_OK;
}
// m17_send_packet_frames (m17_tx_routines.cpp:323-353), batched: append the CRC-16 to each packet, cut it into 25-byte chunks,
// non-final frames carry their index, the final frame EOF + the number of bytes used (25 when the split is exact).
// One thread per packet prepares chunk / meta arrays for the frame formatter; slots past a packet's last frame are marked.
__global__ void k_packet_split(const uint8_t *__restrict__ pk, int64_t stride, const int32_t *__restrict__ len, int64_t n, int max_frames,
                               const uint16_t *__restrict__ g_crc, uint8_t *__restrict__ chunk, uint8_t *__restrict__ meta, int32_t *__restrict__ nframes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = pk + i * stride;
    int L = len[i];
    if (L < 0) L = 0;
    if (L + 2 > 25 * max_frames) L = 25 * max_frames - 2;
    uint16_t k = 0xFFFF;
    for (int b = 0; b < L; b++) k = crc16_step(k, p[b], g_crc);
    const int tot = L + 2, frames = tot / 25, left = tot % 25;
    const int nf = left == 0 ? frames : frames + 1;
    uint8_t *c = chunk + i * (int64_t)max_frames * 25;
    for (int b = 0; b < max_frames * 25; b++) {
        uint8_t v = 0;
        if (b < L) v = p[b]; else if (b == L) v = (uint8_t)(k >> 8); else if (b == L + 1) v = (uint8_t)k;
        c[b] = v;                                                  // m17_fmt_add_packet zero-pads the last chunk (:206-207)
    }
    for (int f = 0; f < max_frames; f++) {
        uint8_t m = 0xFF;                                          // slot not used by this packet
        if (f < nf - 1) m = (uint8_t)(f << 2);                     // eof = 0, frame number
        else if (f == nf - 1) m = (uint8_t)(0x80 | ((left == 0 ? 25 : left) << 2));
        meta[i * max_frames + f] = m;
    }
    nframes[i] = nf;
}
__global__ void k_packet_blank(const uint8_t *__restrict__ meta, int64_t nslots, uint8_t *__restrict__ dibits) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nslots * 192) return;
    if (meta[t / 192] == 0xFF) dibits[t] = 4;                       // unused slot: blank carrier symbols (m17_mod_carrier)
}
extern "C" int m17b_send_packet_frames(m17b_ctx *ctx, const uint8_t *d_packets, int64_t stride, const int32_t *d_len, int64_t n, int max_frames,
                                       uint8_t *d_dibits, int32_t *d_nframes, void *stream) {
    if (!ctx || !d_packets || !d_len || !d_dibits || !d_nframes || n < 0 || max_frames < 1 || max_frames > 32 || stride < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    cudaStream_t st = as_stream(stream);
    uint8_t *chunk, *meta;
    CUDA_TRY(cudaMallocAsync((void **)&chunk, (size_t)n * max_frames * 25, st));
    CUDA_TRY(cudaMallocAsync((void **)&meta, (size_t)n * max_frames, st));
    k_packet_split<<<grid_for(n, 128), 128, 0, st>>>(d_packets, stride, d_len, n, max_frames, ctx->d_crc, chunk, meta, d_nframes);
    k_fmt<3><<<fmt_grid(n * max_frames), 192, 0, st>>>(chunk, meta, n * max_frames, 1, nullptr, d_dibits, ctx->d_genc, ctx->d_prbs);
    k_packet_blank<<<grid_for(n * max_frames * 192, 256), 256, 0, st>>>(meta, n * max_frames, d_dibits);
    KERNEL_CHECK();
    CUDA_TRY(cudaFreeAsync(chunk, st));
    CUDA_TRY(cudaFreeAsync(meta, st));
    return M17B_OK;
}
extern "C" int m17b_fmt_stream_frames(m17b_tx *tx, const uint8_t *d_payload, int64_t F, uint8_t *d_dibits, void *stream) {
    if (!tx || !d_payload || !d_dibits || F < 0) return M17B_E_ARG;
    if (F == 0) return M17B_OK;
    cudaStream_t st = as_stream(stream);
    k_fmt<2><<<fmt_grid(tx->nchan * F), 192, 0, st>>>(d_payload, nullptr, tx->nchan * F, F, tx->d_state, d_dibits, tx->ctx->d_genc, tx->ctx->d_prbs);
    KERNEL_CHECK();
    k_tx_advance<<<grid_for(tx->nchan, 128), 128, 0, st>>>(tx->d_state, tx->nchan, (int)(F & 0xFFFF), (int)(F % 6), 0);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_fmt_bert_frames(m17b_tx *tx, int64_t F, uint8_t *d_dibits, void *stream) {
    if (!tx || !d_dibits || F < 0) return M17B_E_ARG;
    if (F == 0) return M17B_OK;
    cudaStream_t st = as_stream(stream);
    k_fmt<4><<<fmt_grid(tx->nchan * F), 192, 0, st>>>(nullptr, nullptr, tx->nchan * F, F, tx->d_state, d_dibits, tx->ctx->d_genc, tx->ctx->d_prbs);
    KERNEL_CHECK();
    k_tx_advance<<<grid_for(tx->nchan, 128), 128, 0, st>>>(tx->d_state, tx->nchan, 0, 0, (int)((F * 197) % 511));
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_tx_set_lsf(m17b_tx *tx, const uint8_t *d_lsf, void *stream) {
    if (!tx || !d_lsf) return M17B_E_ARG;
    k_tx_set_lsf<<<grid_for(tx->nchan * 32, 256), 256, 0, as_stream(stream)>>>(tx->d_state, tx->nchan, d_lsf);
    KERNEL_CHECK();
    return M17B_OK;
}

extern "C" int m17b_tx_destroy(m17b_tx *tx) {
    if (!tx) return M17B_E_ARG;
    cudaFree(tx->d_taps); cudaFree(tx->d_devtab); cudaFree(tx->d_state); cudaFree(tx->d_dbg);
    free(tx);
    return M17B_OK;
}
extern "C" int m17b_tx_reset(m17b_tx *tx, void *stream) {
    if (!tx) return M17B_E_ARG;
    CUDA_TRY(cudaMemsetAsync(tx->d_state, 0, sizeof(TxChanState) * tx->nchan, as_stream(stream)));
    return M17B_OK;
}
extern "C" int m17b_tx_create(m17b_ctx *ctx, int64_t nchan, int os, m17b_tx **out) {
    if (!ctx || !out || nchan <= 0 || os <= 0 || os > 160) return M17B_E_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    m17b_tx *tx = (m17b_tx *)calloc(1, sizeof(m17b_tx));
    if (!tx) return M17B_E_NOMEM;
    tx->ctx = ctx; tx->nchan = nchan; tx->os = os;
    // m17_mod_init (m17_modulate.cpp:65-76): RRC alpha 0.5, 31*os taps, os samples/symbol, taps sum to 10
    const int nt = 31 * os;
    float *h = (float *)malloc(sizeof(float) * nt);
    m17b_build_rrc_filter(h, 0.5f, nt, os);
    m17b_set_filter_gain(h, 10, 1, nt);
    int rc = upload(&tx->d_taps, h, nt);
    free(h);
    if (rc) { free(tx); return rc; }
    const float dev[5] = {(float)(M_PI / 30.0), (float)(M_PI / 10.0), (float)(-M_PI / 30), (float)(-M_PI / 10.0), 0.0f};
    rc = upload(&tx->d_devtab, dev, 5);
    if (rc) { cudaFree(tx->d_taps); free(tx); return rc; }
    CUDA_TRY(cudaMalloc((void **)&tx->d_state, sizeof(TxChanState) * nchan));
    CUDA_TRY(cudaMemset(tx->d_state, 0, sizeof(TxChanState) * nchan));
    CUDA_TRY(cudaMalloc((void **)&tx->d_dbg, 64));
    CUDA_TRY(cudaMemset(tx->d_dbg, 0, 64));
    *out = tx;
    return M17B_OK;
}
extern "C" int m17b_mod_dibits(m17b_tx *tx, const uint8_t *d_syms, int64_t nsym, int16_t *d_iq, float *d_freq, void *stream) {
    if (!tx || !d_syms || !d_iq || nsym < 0) return M17B_E_ARG;
    if (nsym == 0) return M17B_OK;
    CUDA_TRY(cudaSetDevice(tx->ctx->device));
    const ModGeom g = mod_geometry(tx->nchan, tx->os, nsym);
    const size_t smem = mod_smem_floats(g) * sizeof(float);
    const unsigned grid = (unsigned)((tx->nchan + g.G - 1) / g.G);
    CUDA_TRY(cudaMemsetAsync(tx->d_dbg, 0, 64, as_stream(stream)));
    if (tx->os == 10) {
        CUDA_TRY(cudaFuncSetAttribute(k_mod_fused<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mod_fused<10><<<grid, TXM_THREADS, smem, as_stream(stream)>>>(d_syms, g, tx->d_taps, tx->d_devtab, tx->d_state, tx->nchan, d_iq, d_freq, tx->d_dbg);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(k_mod_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mod_fused<0><<<grid, TXM_THREADS, smem, as_stream(stream)>>>(d_syms, g, tx->d_taps, tx->d_devtab, tx->d_state, tx->nchan, d_iq, d_freq, tx->d_dbg);
    }
    KERNEL_CHECK();
    return M17B_OK;
}

extern "C" int m17b_tx_debug_scan(m17b_tx *tx, uint64_t *h_out8) {
    if (!tx || !h_out8) return M17B_E_ARG;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h_out8, tx->d_dbg, 64, cudaMemcpyDeviceToHost));
    return M17B_OK;
}

// ---------------------------------------------------------------- equaliser (m17_equalize.cpp), thread per channel
// (EqState and the training step live in eq.cuh, shared with the equaliser option of the live chain)
struct m17b_eq { m17b_ctx *ctx; int64_t nchan; EqState *d_state; };
__global__ void k_eq_reset(EqState *st, int64_t nchan, int full) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    eq_init(st[c], full != 0, true);
}
__global__ void k_eq_train(EqState *st, int64_t nchan, const float *in, const float *train, int64_t nsym, float *out) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchan) return;
    EqState e = st[c];
    const float *x2 = in + c * nsym * 2;
    for (int64_t n = 0; n < nsym; n++)
        out[c * nsym + n] = eq_step(e, x2[2 * n], x2[2 * n + 1], train != nullptr, train ? train[c * nsym + n] : 0.0f);
    st[c] = e;
}
extern "C" int m17b_eq_create(m17b_ctx *ctx, int64_t nchan, m17b_eq **out) {
    if (!ctx || !out || nchan <= 0) return M17B_E_ARG;
    m17b_eq *eq = (m17b_eq *)calloc(1, sizeof(m17b_eq));
    if (!eq) return M17B_E_NOMEM;
    eq->ctx = ctx; eq->nchan = nchan;
    CUDA_TRY(cudaMalloc((void **)&eq->d_state, sizeof(EqState) * nchan));
    k_eq_reset<<<grid_for(nchan, 128), 128>>>(eq->d_state, nchan, 1);
    KERNEL_CHECK();
    *out = eq;
    return M17B_OK;
}
extern "C" int m17b_eq_destroy(m17b_eq *eq) { if (!eq) return M17B_E_ARG; cudaFree(eq->d_state); free(eq); return M17B_OK; }
extern "C" int m17b_eq_reset(m17b_eq *eq, void *stream) {
    if (!eq) return M17B_E_ARG;
    k_eq_reset<<<grid_for(eq->nchan, 128), 128, 0, as_stream(stream)>>>(eq->d_state, eq->nchan, 0);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_eq_train(m17b_eq *eq, const float *d_in, const float *d_train, int64_t nsym, float *d_out, void *stream) {
    if (!eq || !d_in || !d_out || nsym < 0) return M17B_E_ARG;
    if (nsym == 0) return M17B_OK;
    k_eq_train<<<grid_for(eq->nchan, 64), 64, 0, as_stream(stream)>>>(eq->d_state, eq->nchan, d_in, d_train, nsym, d_out);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- synthetic channel (bench / demo input only)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void k_synth(int16_t *iq, int64_t nchan, int64_t nsamp, const float *sigma, const float *f0, uint64_t seed) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nchan * nsamp) return;
    const int64_t c = gid / nsamp, n = gid % nsamp;
    short2 v = ((short2 *)iq)[gid];
    double re = v.x, im = v.y;
    const float f = f0 ? f0[c] : 0.0f;
    if (f != 0.0f) {
        double sn, cs;
        sincospi(2.0 * (double)f * (double)n, &sn, &cs);
        const double r2 = re * cs - im * sn, i2 = re * sn + im * cs;
        re = r2; im = i2;
    }
    const float sg = sigma ? sigma[c] : 0.0f;
    if (sg > 0.0f) {
        const uint64_t r = mix64(seed ^ mix64((uint64_t)gid));
        const float u1 = ((uint32_t)(r >> 40) + 1) * (1.0f / 16777216.0f);      // (0,1]
        const float u2 = (uint32_t)(r & 0xFFFFFF) * (1.0f / 16777216.0f);
        const float mag = sg * sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        re += mag * cs; im += mag * sn;
    }
    int ri = (int)rint(fmin(fmax(re, -32768.0), 32767.0)), ii = (int)rint(fmin(fmax(im, -32768.0), 32767.0));
    if (ri == 0 && ii == 0) ri = 1;                                             // dsp_limit divides by |z|
    ((short2 *)iq)[gid] = make_short2((short)ri, (short)ii);
}
extern "C" int m17b_synth_channel(m17b_ctx *ctx, int16_t *d_iq, int64_t nchan, int64_t nsamp, const float *d_sigma, const float *d_f0,
                                  uint64_t seed, void *stream) {
    if (!ctx || !d_iq || nchan <= 0 || nsamp <= 0) return M17B_E_ARG;
    k_synth<<<grid_for(nchan * nsamp, 256), 256, 0, as_stream(stream)>>>(d_iq, nchan, nsamp, d_sigma, d_f0, seed);
    KERNEL_CHECK();
    return M17B_OK;
}
