// decode.cuh -- K=5 rate-1/2 soft Viterbi and the fused per-frame decode:
//   192 symbols -> demap -> de-randomise -> de-interleave -> de-puncture (one constant-memory gather map)
//   -> Golay(24,12) x4 (stream) -> Viterbi -> byte pack -> CRC-16 -> 64-byte record.
// Replaces m17_rx_parse / decode_link_frame / decode_stream_frame / decode_packet_frame
// (m17_rx_parse.cpp:86-226), m17_viterbi_decode (m17_conv.cpp:73-113,148-168).
//
// Mapping: ONE THREAD PER FRAME.  The 16 path metrics live in registers (no shuffles, 16 independent
// butterflies per step give the ILP), the 16 survivor decisions of a step are one uint16 in shared memory
// laid out [step][thread] (conflict-free), and the traceback is a per-thread pointer chase through that
// column.  Frames of a warp are staged into shared memory with coalesced row loads (pitch 193 floats, so
// the per-thread column reads that follow hit 32 different banks).  The gather map index is uniform across
// the warp, so constant-memory reads broadcast.
#pragma once
#include "fec.cuh"

// Encoder output (G1<<1|G2) for 5-bit register r: G1 = 1+D^3+D^4 -> bits {4,1,0}; G2 = 1+D+D^2+D^4 -> bits {4,3,2,0}
__host__ __device__ constexpr int conv_sym(int r) {
    return ((((r >> 4) ^ (r >> 1) ^ r) & 1) << 1) | (((r >> 4) ^ (r >> 3) ^ (r >> 2) ^ r) & 1);
}

template <int X> __device__ __forceinline__ float pick4(float m0, float m1, float m2, float m3) {
    return X == 0 ? m0 : X == 1 ? m1 : X == 2 ? m2 : m3;
}

// One trellis step (m17_conv.cpp:73-113).  New state v is reached from w = (2v)&15 (even) and y = w+1 (odd);
// the branch symbol is the encoder output for register (v>>3)<<4 | predecessor.  Strict '>' keeps the even
// predecessor, ties go to the odd one.  Returns the 16 decisions (bit v set = odd predecessor chosen).
template <int V> struct Acs {
    __device__ __forceinline__ static void run(const float (&acm)[16], float (&tm)[16], float m0, float m1, float m2, float m3, unsigned &dec) {
        constexpr int w = (2 * V) & 15, y = w + 1, hi = (V >> 3) << 4;
        constexpr int x = conv_sym(hi | w), z = conv_sym(hi | y);
        float a = acm[w] + pick4<x>(m0, m1, m2, m3);
        float b = acm[y] + pick4<z>(m0, m1, m2, m3);
        bool even = a > b;
        tm[V] = even ? a : b;
        dec |= even ? 0u : (1u << V);
        Acs<V + 1>::run(acm, tm, m0, m1, m2, m3, dec);
    }
};
template <> struct Acs<16> {
    __device__ __forceinline__ static void run(const float (&)[16], float (&)[16], float, float, float, float, unsigned &) {}
};
__device__ __forceinline__ unsigned viterbi_step(float (&acm)[16], float s1, float s2) {
    // branch metrics: correlation of (+-s1, +-s2) with the expected pair, each a single rounded add
    float n1 = -s1, n2 = -s2;
    float m0 = n1 + n2, m1 = n1 + s2, m2 = s1 + n2, m3 = s1 + s2;
    float tm[16];
    unsigned dec = 0;
    Acs<0>::run(acm, tm, m0, m1, m2, m3, dec);
#pragma unroll
    for (int v = 0; v < 16; v++) acm[v] = tm[v];
    return dec;
}
__device__ __forceinline__ void viterbi_init(float (&acm)[16]) {
#pragma unroll
    for (int v = 0; v < 16; v++) acm[v] = 0.0f;
    acm[0] = 1.0f;    // m17_conv.cpp:153
}
// traceback step: state at step t+1 is s; returns predecessor (the state at step t)
__device__ __forceinline__ unsigned trace_prev(unsigned s, unsigned dec) { return ((s << 1) & 15u) | ((dec >> s) & 1u); }

// ---------------------------------------------------------------- stand-alone batched Viterbi
// d_soft [n][len] -> d_bits [n][len/2].  CTA = NT frames; soft values are staged in chunks of CH steps.
template <int NT, int CH>
__global__ void __launch_bounds__(NT) k_viterbi(const float *soft, int len, int64_t n, uint8_t *bits) {
    extern __shared__ unsigned char smem_raw[];
    const int steps = len / 2;
    uint16_t *dec = (uint16_t *)smem_raw;                                  // [steps][NT]
    float *chunk = (float *)(smem_raw + (((size_t)steps * NT * 2 + 15) & ~(size_t)15));   // [NT][2*CH+1]
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int64_t f0 = (int64_t)blockIdx.x * NT;
    const int64_t f = f0 + tid;
    float acm[16];
    viterbi_init(acm);
    for (int t0 = 0; t0 < steps; t0 += CH) {
        const int nst = min(CH, steps - t0), nval = 2 * nst;
        __syncwarp();
        for (int r = 0; r < 32; r++) {                                      // each warp stages its own 32 rows
            int64_t fr = f0 + wbase + r;
            if (fr < n)
                for (int k = lane; k < nval; k += 32) chunk[(wbase + r) * (2 * CH + 1) + k] = soft[fr * len + 2 * t0 + k];
        }
        __syncwarp();
        if (f < n) {
            const float *row = &chunk[tid * (2 * CH + 1)];
            for (int t = 0; t < nst; t++) dec[(t0 + t) * NT + tid] = (uint16_t)viterbi_step(acm, row[2 * t], row[2 * t + 1]);
        }
    }
    if (f >= n) return;
    unsigned s = 0;
    uint8_t *o = bits + f * steps;
    for (int t = steps - 1; t >= 0; t--) {
        s = trace_prev(s, dec[t * NT + tid]);
        o[t] = (uint8_t)((s >> 3) & 1);                                     // m17_conv.cpp:164-165
    }
}
extern "C" int m17b_viterbi_decode(m17b_ctx *ctx, const float *d_soft, int len, int64_t n, uint8_t *d_bits, void *stream) {
    if (!ctx || !d_soft || !d_bits || len <= 0 || (len & 1) || len > 1024 || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    constexpr int NT = 64, CH = 32;
    size_t smem = (((size_t)(len / 2) * NT * 2 + 15) & ~(size_t)15) + (size_t)NT * (2 * CH + 1) * 4;
    CUDA_TRY(cudaFuncSetAttribute(k_viterbi<NT, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_viterbi<NT, CH><<<grid_for(n, NT), NT, smem, as_stream(stream)>>>(d_soft, len, n, d_bits);
    KERNEL_CHECK();
    return M17B_OK;
}

// ---------------------------------------------------------------- fused frame decode
struct FrameLsf    { static constexpr int STEPS = 244, NBYTES = 30; __device__ static const uint16_t *map() { return c_maps.p1; } };
struct FrameStream { static constexpr int STEPS = 148, NBYTES = 18; __device__ static const uint16_t *map() { return c_maps.p2; } };
struct FramePacket { static constexpr int STEPS = 210, NBYTES = 26; __device__ static const uint16_t *map() { return c_maps.p3; } };

__device__ __forceinline__ float gather_soft(const float *row, float cor, unsigned e) {
    if (e == MAP_ERASE) return 0.0f;                                        // m17_puncture.cpp:52,63,75
    float v = demap_soft(row[e & 0xFFu], cor, (e & MAP_LSB) != 0);
    return (e & MAP_NEG) ? -v : v;                                          // m17_correlate.cpp:29
}

// Viterbi + traceback + pack for one frame; row = this thread's 192 symbols in smem (reused as byte scratch
// afterwards), dec = this thread's decision column.  Writes NBYTES decoded bytes to obytes (smem).
template <class F, int NT>
__device__ __forceinline__ void decode_conv(const float *row, float cor, uint16_t *dec, int tid, uint8_t *obytes) {
    const uint16_t *map = F::map();
    float acm[16];
    viterbi_init(acm);
    for (int t = 0; t < F::STEPS; t++) {
        float s1 = gather_soft(row, cor, map[2 * t]);
        float s2 = gather_soft(row, cor, map[2 * t + 1]);
        dec[t * NT + tid] = (uint16_t)viterbi_step(acm, s1, s2);
    }
    // traceback from state 0; out[t] = MSB of the state at step t = input bit t-1.  Callers discard out[0] and pack
    // out[1..8*NBYTES] MSB first (pack_1_to_8(&bits[1],...), m17_rx_parse.cpp:97,142,171).
    unsigned s = 0, acc = 0;
    for (int t = F::STEPS - 1; t >= 1; t--) {
        s = trace_prev(s, dec[t * NT + tid]);
        if (t <= 8 * F::NBYTES) {
            acc |= ((s >> 3) & 1u) << ((8 - t) & 7);                        // t = 8j+8 is bit 0 of byte j ... t = 8j+1 is bit 7
            if ((t & 7) == 1) { obytes[(t - 1) >> 3] = (uint8_t)acc; acc = 0; }
        }
    }
}

// frames: records pre-filled with sym_off/type/flags by the framer (or by k_parse_init); the symbols of record r
// of channel c start at syms[c*sym_pitch + sym_carry + (rec.sym_off - sym_base[c])].
template <int NT>
__global__ void __launch_bounds__(NT) k_decode_frames(const float *__restrict__ syms, int64_t sym_pitch, int sym_carry,
                                                      const int32_t *__restrict__ sym_base, m17b_frame_rec *frames, int64_t fcap,
                                                      const int32_t *__restrict__ nframes, int tiles_per_chan, float *soft_out,
                                                      const uint16_t *__restrict__ g_crc, const uint16_t *__restrict__ genc,
                                                      const uint16_t *__restrict__ gerr) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int PITCH = 193;
    float *rows = (float *)smem_raw;                                        // [NT][193]
    uint16_t *dec = (uint16_t *)(smem_raw + (size_t)NT * PITCH * 4);        // [244][NT]
    __shared__ uint16_t crc_tab[256];
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    for (int i = tid; i < 256; i += NT) crc_tab[i] = g_crc[i];
    __syncthreads();
    const int64_t c = blockIdx.x / tiles_per_chan;
    const int tile = blockIdx.x % tiles_per_chan;
    const int nfr = min((int64_t)nframes[c], fcap);
    const int slot = tile * NT + tid;
    m17b_frame_rec *rec = frames + c * fcap + slot;
    int type = -1, flags = 0;
    int64_t src = 0;
    if (slot < nfr) {
        uint2 hd = *(const uint2 *)rec;
        type = (hd.y & 0xFF);
        flags = (hd.y >> 8) & 0xFF;
        src = c * sym_pitch + sym_carry + ((int32_t)hd.x - (sym_base ? sym_base[c] : 0));
    }
    const bool work = (slot < nfr) && (flags & M17B_F_PARSED) && type >= 1 && type <= 4;
    // stage: each warp loads the rows of its own 32 frames (coalesced 128-byte requests)
    for (int r = 0; r < 32; r++) {
        const bool w_r = __shfl_sync(0xffffffffu, (int)work, r) != 0;
        const long long s_r = __shfl_sync(0xffffffffu, (long long)src, r);
        if (w_r) {
            float *dst = &rows[(wbase + r) * PITCH];
#pragma unroll
            for (int k = 0; k < 6; k++) dst[lane + 32 * k] = __ldg(&syms[s_r + lane + 32 * k]);
        }
    }
    __syncwarp();
    if (!work) return;
    float *row = &rows[tid * PITCH];
    float hdr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) hdr[i] = row[i];
    const float cor = demap_cor(hdr);
    if (soft_out) {
        float *so = soft_out + (c * fcap + slot) * 368;
        for (int k = 0; k < 184; k++) { so[2 * k] = demap_soft(row[8 + k], cor, false); so[2 * k + 1] = demap_soft(row[8 + k], cor, true); }
    }
    uint32_t golay_e = 0, nbytes = 0;
    uint32_t lw01 = 0, lw23 = 0;
    uint8_t *ob = (uint8_t *)&rows[tid * PITCH] ;                           // byte scratch AFTER the ACS pass (row no longer needed)
    uint8_t dbytes[32];
    if (type == M17B_T_STREAM) {
        // 4 x hard_decode_24_bits (m17_bit_utils.cpp:180-187: bit = soft >= 0) + m_17_golay_decode
        uint32_t w[4];
        for (int q = 0; q < 4; q++) {
            uint32_t word = 0;
            for (int b = 0; b < 24; b++) word = (word << 1) | (gather_soft(row, cor, c_maps.lich[24 * q + b]) >= 0 ? 1u : 0u);
            golay_e += (uint32_t)golay_decode_word(word, genc, gerr, &w[q]);
        }
        lw01 = (w[0] << 12) | w[1];                                          // pack_12_to_8_x4x6, m17_bit_utils.cpp:152-172
        lw23 = (w[2] << 12) | w[3];
        decode_conv<FrameStream, NT>(row, cor, dec, tid, ob);
        nbytes = FrameStream::NBYTES;
    } else if (type == M17B_T_LSF) {
        decode_conv<FrameLsf, NT>(row, cor, dec, tid, ob);
        nbytes = FrameLsf::NBYTES;
    } else if (type == M17B_T_PACKET) {
        decode_conv<FramePacket, NT>(row, cor, dec, tid, ob);
        nbytes = FramePacket::NBYTES;
    }
    uint16_t crc = 0;
    if (nbytes) {
        crc = 0xFFFF;
        for (uint32_t i = 0; i < nbytes; i++) crc = crc16_step(crc, ob[i], crc_tab);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) dbytes[i] = (i < (int)nbytes) ? ob[i] : 0;
    // assemble bytes 4..47 and 52..55 of the record (0..3, 46..51 belong to the framer)
    uint8_t *rb = (uint8_t *)rec;
    if (type == M17B_T_PACKET && (dbytes[25] & 0x80)) flags |= M17B_F_PKT_EOF;
    rb[5] = (uint8_t)flags;
    rb[6] = (uint8_t)golay_e;
    rb[7] = (uint8_t)nbytes;
    rb[8] = (uint8_t)(lw01 >> 16); rb[9] = (uint8_t)(lw01 >> 8); rb[10] = (uint8_t)lw01;
    rb[11] = (uint8_t)(lw23 >> 16); rb[12] = (uint8_t)(lw23 >> 8); rb[13] = (uint8_t)lw23;
#pragma unroll
    for (int i = 0; i < 30; i++) rb[14 + i] = dbytes[i];
    *(uint16_t *)(rb + 44) = crc;
    rec->cor = cor;
}

// stand-alone m17_rx_parse for n independent frames: initialise records, then run the fused decode
__global__ void k_parse_init(const uint8_t *type, int64_t n, m17b_frame_rec *rec) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4 *p = (uint4 *)(rec + r);
    p[0] = make_uint4((uint32_t)(r * 192), (uint32_t)type[r] | (M17B_F_PARSED << 8), 0, 0);
    p[1] = z; p[2] = z; p[3] = z;
}
__global__ void k_set_i32(int32_t *p, int32_t v) { *p = v; }

template <int NT> static size_t decode_smem() { return (size_t)NT * 193 * 4 + (size_t)244 * NT * 2; }
#define DECODE_NT 64

static int launch_decode(m17b_ctx *ctx, const float *syms, int64_t sym_pitch, int sym_carry, const int32_t *sym_base,
                         m17b_frame_rec *frames, int64_t fcap, const int32_t *nframes, int64_t nchan, float *soft_out, cudaStream_t st) {
    const int tiles = (int)((fcap + DECODE_NT - 1) / DECODE_NT);
    const size_t smem = decode_smem<DECODE_NT>();
    static bool attr_set = false;
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(k_decode_frames<DECODE_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
    if ((int64_t)tiles * nchan > 0x7fffffffLL) return M17B_E_ARG;
    k_decode_frames<DECODE_NT><<<(unsigned)(tiles * nchan), DECODE_NT, smem, st>>>(syms, sym_pitch, sym_carry, sym_base, frames, fcap, nframes, tiles,
                                                                                 soft_out, ctx->d_crc, ctx->d_genc, ctx->d_gerr);
    KERNEL_CHECK();
    return M17B_OK;
}

extern "C" int m17b_rx_parse_frames(m17b_ctx *ctx, const float *d_sym, const uint8_t *d_type, int64_t n, m17b_frame_rec *d_rec, float *d_soft, void *stream) {
    if (!ctx || !d_sym || !d_type || !d_rec || n < 0 || n > 0x7fffffffLL / 192) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    cudaStream_t st = as_stream(stream);
    int32_t *d_n;
    CUDA_TRY(cudaMallocAsync((void **)&d_n, sizeof(int32_t), st));
    k_set_i32<<<1, 1, 0, st>>>(d_n, (int32_t)n);
    k_parse_init<<<grid_for(n, 256), 256, 0, st>>>(d_type, n, d_rec);
    KERNEL_CHECK();
    int rc = launch_decode(ctx, d_sym, 0, 0, nullptr, d_rec, n, d_n, 1, d_soft, st);
    CUDA_TRY(cudaFreeAsync(d_n, st));
    return rc;
}

// ---------------------------------------------------------------- config-4 microbenchmark: punctured soft frames in, bytes out
// d_soft [n][NIN] already de-randomised/de-interleaved (i.e. so[] of m17_rx_parse.cpp:91-95); de-puncture + Viterbi + pack.
template <class F, int PAT, int NIN, int NT>
__global__ void __launch_bounds__(NT) k_viterbi_punct(const float *__restrict__ soft, int64_t n, uint8_t *__restrict__ bytes) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int PITCH = NIN + 1;
    float *rows = (float *)smem_raw;                                        // [NT][NIN+1]
    uint16_t *dec = (uint16_t *)(smem_raw + (((size_t)NT * PITCH * 4 + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int64_t f0 = (int64_t)blockIdx.x * NT, f = f0 + tid;
    for (int r = 0; r < 32; r++) {
        int64_t fr = f0 + wbase + r;
        if (fr < n) for (int k = lane; k < NIN; k += 32) rows[(wbase + r) * PITCH + k] = __ldg(&soft[fr * NIN + k]);
    }
    __syncwarp();
    if (f >= n) return;
    const float *row = &rows[tid * PITCH];
    float acm[16];
    viterbi_init(acm);
    int k = 0;
    for (int t = 0; t < F::STEPS; t++) {
        float s1 = d_punct_keeps(PAT, 2 * t) ? row[k++] : 0.0f;
        float s2 = d_punct_keeps(PAT, 2 * t + 1) ? row[k++] : 0.0f;
        dec[t * NT + tid] = (uint16_t)viterbi_step(acm, s1, s2);
    }
    unsigned s = 0, acc = 0;
    uint8_t *o = bytes + f * F::NBYTES;
    for (int t = F::STEPS - 1; t >= 1; t--) {
        s = trace_prev(s, dec[t * NT + tid]);
        if (t <= 8 * F::NBYTES) {
            acc |= ((s >> 3) & 1u) << ((8 - t) & 7);
            if ((t & 7) == 1) { o[(t - 1) >> 3] = (uint8_t)acc; acc = 0; }
        }
    }
}
template <class F, int PAT, int NIN> static int launch_vp(const float *d_soft, int64_t n, uint8_t *d_bytes, cudaStream_t st) {
    constexpr int NT = 64;
    size_t smem = (((size_t)NT * (NIN + 1) * 4 + 15) & ~(size_t)15) + (size_t)F::STEPS * NT * 2;
    CUDA_TRY(cudaFuncSetAttribute(k_viterbi_punct<F, PAT, NIN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_viterbi_punct<F, PAT, NIN, NT><<<grid_for(n, NT), NT, smem, st>>>(d_soft, n, d_bytes);
    KERNEL_CHECK();
    return M17B_OK;
}
extern "C" int m17b_viterbi_punctured(m17b_ctx *ctx, int pattern, const float *d_soft, int64_t n, uint8_t *d_bytes, void *stream) {
    if (!ctx || !d_soft || !d_bytes || n < 0) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    if (pattern == 1) return launch_vp<FrameLsf, 1, 368>(d_soft, n, d_bytes, as_stream(stream));
    if (pattern == 2) return launch_vp<FrameStream, 2, 272>(d_soft, n, d_bytes, as_stream(stream));
    if (pattern == 3) return launch_vp<FramePacket, 3, 368>(d_soft, n, d_bytes, as_stream(stream));
    return M17B_E_ARG;
}
