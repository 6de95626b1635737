// decode.cuh -- K=5 rate-1/2 soft Viterbi and the per-frame decode:
//   192 symbols -> demap -> de-randomise -> de-interleave -> de-puncture (one gather map)
//   -> Golay(24,12) x4 (stream) -> Viterbi -> byte pack -> CRC-16 -> 64-byte record.
// Replaces m17_rx_parse / decode_link_frame / decode_stream_frame / decode_packet_frame
// (m17_rx_parse.cpp:86-226), m17_viterbi_decode (m17_conv.cpp:73-113,148-168).
//
// The trellis pass is ONE THREAD PER FRAME everywhere: the 16 path metrics live in registers (no shuffles, 16 independent
// butterflies per step give the ILP, packed adds), the 16 survivor decisions of a step are one uint16 in per-thread local memory
// (interleaved by the hardware, so the accesses coalesce), the traceback is a per-thread pointer chase.
// Three kernels share it:
//   k_stream_gather + k_stream_acs  stream frames (nearly all frames): a warp per 8 record slots produces the 272 kept trellis
//                                   inputs in trellis order + the LICH / Golay words, and lists the frames of the other types;
//                                   the trellis kernel stages 16 steps of 32 frames at a time (8.4 KB of shared memory per warp);
//   k_decode_frames                 LSF / packet / BERT frames from that list, whole 192-symbol rows staged (pitch 193 floats, so the
//                                   per-thread column reads hit 32 banks), gather map read warp-uniformly from constant memory;
//   k_viterbi / k_viterbi_punct     the stand-alone batched decoders (m17b_viterbi_decode, m17b_viterbi_punctured).
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
// one butterfly: a = acm[w]+met[x], b = acm[y]+met[z]; keep a iff a > b (ties and NaN go to the odd predecessor, exactly
// like 'if(tempa>tempb)' in the BF macro); the decision bit is OR-ed in under the predicate (no select + shift + or).
__device__ __forceinline__ float acs_select(float a, float b, unsigned &dec, unsigned bit) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %2, %3;\n\tselp.f32 %0, %2, %3, p;\n\t@!p or.b32 %1, %1, %4;\n\t}"
        : "=f"(r), "+r"(dec) : "f"(a), "f"(b), "r"(bit));
    return r;
}
// Packed form (32 scalar adds per step become 16 FADD2: -19 % instructions in the ACS loop, frame decode 0.43 -> 0.41 ms): the
// two candidates of new state V are (from[w] + met[x], from[w+1] + met[z]) with w = (2V) & 15 -- one FADD2 on the register pair (from[w], from[w+1]) and one of only four distinct metric pairs (met[0],met[3]), (met[2],met[1]), (met[1],met[2]),
// (met[3],met[0]) (conv_sym over the 16 states).  Each half is the same IEEE add as the scalar form.
template <int V> struct AcsP {
    __device__ __forceinline__ static void run(const f32x2 (&ap)[8], float (&tm)[16], const f32x2 (&mp)[4], unsigned (&dec)[4]) {
        constexpr int w = (2 * V) & 15, hi = (V >> 3) << 4;
        constexpr int x = conv_sym(hi | w), z = conv_sym(hi | (w + 1));
        static_assert(x + z == 3, "metric pairs are (k, 3 - k)");
        float a, b;
        unpack2(add2(ap[w >> 1], mp[x]), a, b);
        tm[V] = acs_select(a, b, dec[V & 3], 1u << V);
        AcsP<V + 1>::run(ap, tm, mp, dec);
    }
};
template <> struct AcsP<16> {
    __device__ __forceinline__ static void run(const f32x2 (&)[8], float (&)[16], const f32x2 (&)[4], unsigned (&)[4]) {}
};
// one trellis step from metrics `from` into `to` (ping-pong, so no register copies); returns the 16 decisions
__device__ __forceinline__ unsigned viterbi_step_pp(const float (&from)[16], float (&to)[16], float s1, float s2) {
    // branch metrics: correlation of (+-s1, +-s2) with the expected pair, each a single rounded add
    const float n1 = -s1, n2 = -s2;
    const float m0 = n1 + n2, m1 = n1 + s2, m2 = s1 + n2, m3 = s1 + s2;
    unsigned dec[4] = {0, 0, 0, 0};
    const f32x2 mp[4] = {pack2(m0, m3), pack2(m1, m2), pack2(m2, m1), pack2(m3, m0)};      // mp[x] = (met[x], met[3 - x])
    f32x2 ap[8];
#pragma unroll
    for (int j = 0; j < 8; j++) ap[j] = pack2(from[2 * j], from[2 * j + 1]);
    AcsP<0>::run(ap, to, mp, dec);
    return (dec[0] | dec[1]) | (dec[2] | dec[3]);
}
__device__ __forceinline__ unsigned viterbi_step(float (&acm)[16], float s1, float s2) {
    float tm[16];
    const unsigned dec = viterbi_step_pp(acm, tm, s1, s2);
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
struct FrameBert   { static constexpr int STEPS = 201, NBYTES = 25; __device__ static const uint16_t *map() { return c_maps.bert; } };

// The trellis pass over a warp's 32 rows of kept inputs (row r at base + r * NIN, rows without their bit in rowmask are not
// touched), survivors into dec; lane_base = base + lane.
template <class F, int PAT, int NIN>
__device__ __forceinline__ void viterbi_chunked(const float *lane_base, unsigned rowmask, float (&tile)[2][32][33], int lane, uint16_t (&dec)[F::STEPS]) {
    constexpr int NCH = (F::STEPS + VP_CHUNK - 1) / VP_CHUNK;
    static_assert(F::STEPS % 2 == 0 && VP_CHUNK % 2 == 0, "two steps per iteration");
    auto stage = [&](int c) {                                       // chunk c: inputs [koff[c], koff[c + 1]) of every row
        if (c < NCH) {
            const int k0 = c_punct.koff[PAT - 1][c], cnt = min((int)c_punct.koff[PAT - 1][c + 1], NIN) - k0;
            if (lane < cnt) {
                const float *src = lane_base + k0;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[c & 1][0][lane]);
#pragma unroll 8
                for (int r = 0; r < 32; r++)
                    if ((rowmask >> r) & 1u) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + r * 33 * 4), "l"(src + (int64_t)r * NIN));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    stage(0);
    stage(1);
    float ma[16], mb[16];
    viterbi_init(ma);
    for (int c = 0; c < NCH; c++) {
        asm volatile("cp.async.wait_group 1;");                      // chunk c has landed (chunk c + 1 may still be in flight)
        __syncwarp();
        const float *row = tile[c & 1][lane];
        const int t0 = c * VP_CHUNK, t1 = min(t0 + VP_CHUNK, F::STEPS);
        int k = 0;
        for (int t = t0; t < t1; t += 2) {
            const unsigned k0 = c_punct.keep[PAT - 1][t], k1 = c_punct.keep[PAT - 1][t + 1];      // warp-uniform
            const float s1 = (k0 & 1) ? row[k++] : 0.0f;
            const float s2 = (k0 & 2) ? row[k++] : 0.0f;
            const float s3 = (k1 & 1) ? row[k++] : 0.0f;
            const float s4 = (k1 & 2) ? row[k++] : 0.0f;
            dec[t] = (uint16_t)viterbi_step_pp(ma, mb, s1, s2);
            dec[t + 1] = (uint16_t)viterbi_step_pp(mb, ma, s3, s4);
        }
        __syncwarp();                                                // every lane is done with the tile: it may be refilled
        stage(c + 2);
    }
}
// traceback from state 0 and byte pack (see decode_conv)
template <class F>
__device__ __forceinline__ void viterbi_traceback(const uint16_t (&dec)[F::STEPS], uint8_t *obytes) {
    unsigned s = 0;
    for (int t = F::STEPS - 1; t > 8 * F::NBYTES; t--) s = trace_prev(s, dec[t]);
    // (fully unrolled over the bytes: 900 of the trellis kernel's 1800 SASS instructions, run once per frame.  One byte per trip --
    //  784 instructions, 64 registers, the output bytes in local memory -- was tried for the instruction cache's sake: no change
    //  in the step, 1.82 -> 1.76 G frames/s on the stand-alone kernel.)
    for (int j = F::NBYTES - 1; j >= 0; j--) {
        unsigned d[8], acc = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) d[b] = dec[8 * j + 8 - b];
#pragma unroll
        for (int b = 0; b < 8; b++) { s = trace_prev(s, d[b]); acc |= ((s >> 3) & 1u) << b; }
        obytes[j] = (uint8_t)acc;
    }
}

// soft value for one gather-map entry; the row already holds m = sym * cor (m17_dsp.cpp:38) for the payload symbols
__device__ __forceinline__ float gather_soft(const float *row, unsigned e) {
    if (e == MAP_ERASE) return 0.0f;                                        // m17_puncture.cpp:52,63,75
    const float m = row[e & 0xFFu];
    const float v = (e & MAP_LSB) ? demap_lsb(m) : -m;                        // m17_dsp.cpp:40-41
    return (e & MAP_NEG) ? -v : v;                                          // m17_correlate.cpp:29
}
// hard_decode_24_bits (m17_bit_utils.cpp:180-187) only asks whether the soft value is >= 0, and that needs no arithmetic: the LSB
// value |m| - 0.6666 is never zero (0.6666 is not a float) and is >= 0 exactly when |m| > 0.6666f rounded down, i.e. |m| >=
// the float just above 0.6666; the MSB value -m is >= 0 when m <= 0 (-0.0 >= 0 holds).  De-randomising flips the sign, and
// with it '>=' into '<=' (only the MSB value can be zero).  NaN compares false in the reference in every case, here too.
__device__ __forceinline__ bool gather_hard(const float *row, unsigned e) {
    const float m = row[e & 0xFFu];
    constexpr float c_up = 0.66660005f;                                      // smallest float above 0.6666 (0.6666f is below it)
    static_assert((double)c_up > 0.6666 && (double)c_up - 0.6666 < 5.9e-8 && (double)0.6666f < 0.6666, "c_up is the float just above 0.6666");
    if (e & MAP_LSB) return (e & MAP_NEG) ? fabsf(m) < c_up : fabsf(m) >= c_up;
    return (e & MAP_NEG) ? m >= 0.0f : m <= 0.0f;
}

// Viterbi + traceback + pack for one frame; row = this thread's scaled symbols in smem (reused as byte scratch
// afterwards), dec = this thread's decision column.  Writes NBYTES decoded bytes to obytes (smem).
template <class F, int NT>
__device__ __forceinline__ void decode_conv(const float *row, uint16_t *dec_smem, int tid, uint8_t *obytes) {
#ifndef M17B_DEC_SMEM
    // survivor words in per-thread LOCAL memory (interleaved per thread by the hardware, so the accesses coalesce) instead of
    // shared memory: frees 9.5 / 15.6 KB of shared memory per warp for occupancy
    uint16_t decl[F::STEPS];
    uint16_t *dec = decl;
#define DIDX(t) (t)
#else
    uint16_t *dec = dec_smem;
#define DIDX(t) ((t) * NT + tid)
#endif
    const uint16_t *map = F::map();
    float ma[16], mb[16];
    viterbi_init(ma);
    // software pipeline: the four soft values of the NEXT two steps are gathered (constant-memory map -> smem -> demap)
    // while the 32 butterflies of the current two steps issue.  An odd step count (BERT: 201) ends with one single step.
    constexpr int PAIRS = F::STEPS & ~1;
    uint2 e = *(const uint2 *)map;                                          // four uint16 entries, warp-uniform
    float s1 = gather_soft(row, e.x & 0xFFFFu), s2 = gather_soft(row, e.x >> 16);
    float s3 = gather_soft(row, e.y & 0xFFFFu), s4 = gather_soft(row, e.y >> 16);
    for (int t = 0; t < PAIRS; t += 2) {
        float n1 = 0.f, n2 = 0.f, n3 = 0.f, n4 = 0.f;
        if (t + 2 < PAIRS) {
            e = *(const uint2 *)(map + 2 * t + 4);
            n1 = gather_soft(row, e.x & 0xFFFFu); n2 = gather_soft(row, e.x >> 16);
            n3 = gather_soft(row, e.y & 0xFFFFu); n4 = gather_soft(row, e.y >> 16);
        } else if (F::STEPS & 1) {
            const unsigned ee = *(const unsigned *)(map + 2 * t + 4);
            n1 = gather_soft(row, ee & 0xFFFFu); n2 = gather_soft(row, ee >> 16);
        }
        dec[DIDX(t)] = (uint16_t)viterbi_step_pp(ma, mb, s1, s2);
        dec[DIDX(t + 1)] = (uint16_t)viterbi_step_pp(mb, ma, s3, s4);
        s1 = n1; s2 = n2; s3 = n3; s4 = n4;
    }
    if (F::STEPS & 1) dec[DIDX(F::STEPS - 1)] = (uint16_t)viterbi_step_pp(ma, mb, s1, s2);
    // traceback from state 0; out[t] = MSB of the state at step t = input bit t-1.  Callers discard out[0] and pack
    // out[1..8*NBYTES] MSB first (pack_1_to_8(&bits[1],...), m17_rx_parse.cpp:97,142,171).
    // The survivor words do not depend on the state, so they are fetched eight steps at a time; only the 3-op state
    // update is serial.
    unsigned s = 0;
    static_assert(F::STEPS - 1 >= 8 * F::NBYTES && (F::STEPS - 1 - 8 * F::NBYTES) < 8, "tail shorter than a byte");
    for (int t = F::STEPS - 1; t > 8 * F::NBYTES; t--) s = trace_prev(s, dec[DIDX(t)]);   // tail bits, discarded
    for (int j = F::NBYTES - 1; j >= 0; j--) {
        unsigned d[8], acc = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) d[b] = dec[DIDX(8 * j + 8 - b)];                             // t = 8j+8 ... 8j+1
#pragma unroll
        for (int b = 0; b < 8; b++) { s = trace_prev(s, d[b]); acc |= ((s >> 3) & 1u) << b; }   // t = 8j+8 is bit 0 ... t = 8j+1 is bit 7
        obytes[j] = (uint8_t)acc;
    }
#undef DIDX
}

// frames: records pre-filled with sym_off/type/flags by the framer (or by k_parse_init); the symbols of record r
// of channel c start at syms[c*sym_pitch + sym_carry + (rec.sym_off - sym_base[c])].
// This kernel handles the LSF / packet / BERT frames (a handful per channel and call); stream frames go through k_stream_gather +
// k_stream_acs below.  Its frames come as a compact list (dlist[0] = count, dlist[1 + i] = channel * fcap + slot, appended by
// k_stream_gather, which sees every record header anyway): thread i of the grid takes list entry i, so the warps are full.  (The
// first form mapped a CTA to 32 consecutive slots of one channel: one or two active lanes per warp -- 65 536 channels x 25 blocks
// spent more time here, one LSF per channel, than in the front end.)  CTAs beyond the list exit at once.
template <int NT>
__global__ void __launch_bounds__(NT) k_decode_frames(const float *__restrict__ syms, int64_t sym_pitch, int sym_carry,
                                                      const int32_t *__restrict__ sym_base, m17b_frame_rec *frames, int64_t fcap,
                                                      const int32_t *__restrict__ dlist, float *soft_out,
                                                      const uint16_t *__restrict__ g_crc, const uint16_t *__restrict__ genc,
                                                      const uint16_t *__restrict__ gerr, int bert_on) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int PITCH = 193;
    float *rows = (float *)smem_raw;                                        // [NT][193]
    uint16_t *dec = (uint16_t *)(smem_raw + (size_t)NT * PITCH * 4);        // [148 or 244][NT]
    __shared__ uint16_t crc_tab[256];
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int64_t nlist = dlist[0];
    if ((int64_t)blockIdx.x * NT >= nlist) return;
    for (int i = tid; i < 256; i += NT) crc_tab[i] = g_crc[i];
    // a small grid walks the list (the host does not know its length: a grid sized for the worst case was 9216 CTAs of 24.7 KB
    // shared memory each to dispatch, nearly all empty, ahead of the trellis kernel launched beside this one)
    for (int64_t b0 = (int64_t)blockIdx.x * NT; b0 < nlist; b0 += (int64_t)gridDim.x * NT) {
    const int64_t li = b0 + tid;
    const bool work = li < nlist;
    const int64_t fi = work ? (int64_t)dlist[1 + li] : 0;               // channel * fcap + slot (every entry is a parsed LSF / packet / BERT record)
    const int64_t c = fi / fcap;
    const int slot = (int)(fi - c * fcap);
    m17b_frame_rec *rec = frames + fi;
    int type = -1, flags = 0;
    uint32_t w0 = 0;
    int64_t src = 0;
    if (work) {
        const uint2 hd = *(const uint2 *)rec;
        w0 = hd.x;
        type = (hd.y & 0xFF);
        flags = (hd.y >> 8) & 0xFF;
        src = c * sym_pitch + sym_carry + (int32_t)(hd.x - (sym_base ? (uint32_t)sym_base[c] : 0u));
    }
    // stage: each warp loads the rows of its own 32 frames (coalesced 128-byte requests)
    // (asynchronous copies: all 32 x 6 row pieces are in flight together -- with plain loads every row's stores waited for that
    //  row's loads, 12 % of the kernel's stall samples)
    for (int r = 0; r < 32; r++) {
        const bool w_r = __shfl_sync(0xffffffffu, (int)work, r) != 0;
        const long long s_r = __shfl_sync(0xffffffffu, (long long)src, r);
        if (w_r) {
            float *dst = &rows[(wbase + r) * PITCH];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst + lane + 32 * k);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(syms + s_r + lane + 32 * k));
            }
        }
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    if (work) {
    float *row = &rows[tid * PITCH];
    float hdr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) hdr[i] = row[i];
    const float cor = demap_cor(hdr);
    for (int k = 8; k < 192; k++) row[k] = row[k] * cor;                    // m = in * mag  (m17_dsp.cpp:38), once per symbol
    if (soft_out) {
        float *so = soft_out + (c * fcap + slot) * 368;
        for (int k = 0; k < 184; k++) {
            const float m = row[8 + k];
            so[2 * k] = -m;
            so[2 * k + 1] = demap_lsb(m);
        }
    }
    const uint32_t golay_e = 0, lw01 = 0, lw23 = 0;
    uint32_t nbytes = 0;
    uint8_t *ob = (uint8_t *)row;                                           // byte scratch AFTER the ACS pass (row no longer needed)
    if (type == M17B_T_LSF) {
        decode_conv<FrameLsf, NT>(row, dec, tid, ob);
        nbytes = FrameLsf::NBYTES;
    } else if (type == M17B_T_PACKET) {
        decode_conv<FramePacket, NT>(row, dec, tid, ob);
        nbytes = FramePacket::NBYTES;
    } else if (type == M17B_T_BERT && bert_on) {
        // decode_bert_frame is empty upstream (m17_rx_parse.cpp:178-180: demap only); with the BERT receive extension on
        // (m17b_rx_set_bert): inverse of m17_fmt_add_bert_frame (m17_tx_routines.cpp:226-238): 197 PRBS9 bits + 3 of the 4 tail bits -> 25 bytes
        decode_conv<FrameBert, NT>(row, dec, tid, ob);
        nbytes = FrameBert::NBYTES;
    }
    uint32_t crc = 0;
    if (nbytes) {
        uint16_t k = 0xFFFF;
        for (uint32_t i = 0; i < nbytes; i++) k = crc16_step(k, ob[i], crc_tab);
        crc = k;
        for (uint32_t i = nbytes; i < 32; i++) ob[i] = 0;
    } else {
        for (int i = 0; i < 32; i++) ob[i] = 0;
    }
    const uint16_t *oh = (const uint16_t *)ob;                              // data[] as 15 half-words
    if (type == M17B_T_PACKET && (ob[25] & 0x80)) flags |= M17B_F_PKT_EOF;
    // assemble the 64-byte record in registers: words 0/1 (sym_off, type) and the votes/errors/variance fields come
    // from the framer, everything else from this kernel
    uint32_t *rw = (uint32_t *)rec;
    const uint32_t w11_old = rw[11], w12 = rw[12];
    uint4 q0, q1, q2, q3;
    q0.x = w0;
    q0.y = (uint32_t)type | ((uint32_t)flags << 8) | (golay_e << 16) | (nbytes << 24);
    q0.z = ((lw01 >> 16) & 0xFF) | (((lw01 >> 8) & 0xFF) << 8) | ((lw01 & 0xFF) << 16) | (((lw23 >> 16) & 0xFF) << 24);
    q0.w = ((lw23 >> 8) & 0xFF) | ((lw23 & 0xFF) << 8) | ((uint32_t)oh[0] << 16);
    q1.x = oh[1] | ((uint32_t)oh[2] << 16);   q1.y = oh[3] | ((uint32_t)oh[4] << 16);
    q1.z = oh[5] | ((uint32_t)oh[6] << 16);   q1.w = oh[7] | ((uint32_t)oh[8] << 16);
    q2.x = oh[9] | ((uint32_t)oh[10] << 16);  q2.y = oh[11] | ((uint32_t)oh[12] << 16);
    q2.z = oh[13] | ((uint32_t)oh[14] << 16); q2.w = crc | (w11_old & 0xFFFF0000u);
    q3.x = w12; q3.y = __float_as_uint(cor); q3.z = 0; q3.w = 0;
    uint4 *r4 = (uint4 *)rec;
    r4[0] = q0; r4[1] = q1; r4[2] = q2; r4[3] = q3;
    }
    __syncthreads();                                                        // the rows are restaged by the next trip
    }
}

// ---------------------------------------------------------------- stream frames: the decode in two kernels
// Stream frames are nearly all the frames there are, and the one-kernel form above holds a frame's 192 symbols in shared memory for
// the whole trellis pass (they are read in interleaver order): 24.7 KB per warp, 9 warps per SM, an issue rate of 45 %.  So the
// two halves run as separate kernels.  k_stream_gather, one warp per frame: demap, de-randomise, de-interleave, de-puncture into
// the 272 kept trellis inputs IN TRELLIS ORDER (1088 bytes per frame, written and read back through L2), and the LICH: 96 hard
// bits by ballot, 4 x Golay(24,12).  k_stream_acs, one thread per frame: the chunk-staged trellis pass of the punctured Viterbi
// kernel below (8.4 KB of shared memory per warp), traceback, CRC, record.
struct StreamAux { uint32_t lw01, lw23, golay_e; float cor; };
#define SG_WARPS 4
#define SG_FPW   8            // record slots per warp: the per-warp set-up (channel, map entries) is paid once per 8 frames
__global__ void __launch_bounds__(SG_WARPS * 32) k_stream_gather(const float *__restrict__ syms, int64_t sym_pitch, int sym_carry, const int32_t *__restrict__ sym_base,
                                                                 const m17b_frame_rec *__restrict__ frames, int64_t fcap, const int32_t *__restrict__ nframes,
                                                                 const int2 *__restrict__ frame_rng, int slots_per_chan, int64_t nchan,
                                                                 const uint16_t *__restrict__ smap, const uint16_t *__restrict__ genc, const uint16_t *__restrict__ gerr,
                                                                 float *__restrict__ ssoft, StreamAux *__restrict__ saux, float *soft_out, int32_t *dlist) {
    // (the first form of this kernel took one frame per warp and staged the map in shared memory per CTA: ~600 warp instructions
    //  per frame at an issue rate of 90 %, 0.16 ms for 245 000 frames; the map entries a lane uses are the same for every frame)
    __shared__ float soft2_s[SG_WARPS][368];                     // the frame's soft values in natural order: (-m, |m| - 0.6666) per payload symbol
    constexpr unsigned FULL = 0xffffffffu;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned item = blockIdx.x * SG_WARPS + wid;           // (channel, group of SG_FPW slots); 32-bit: the launch checks the grid
    const unsigned groups = (unsigned)slots_per_chan / SG_FPW;
    const int64_t c = item / groups;
    if (c >= nchan) return;
    int lo = 0, nfr;
    if (frame_rng) { const int2 r = frame_rng[c]; lo = r.x; nfr = r.y; }
    else nfr = (int)min((int64_t)nframes[c], fcap);
    int slot = lo + (int)(item - (unsigned)c * groups) * SG_FPW;
    const int send = min(slot + SG_FPW, nfr);
    if (slot >= send) return;
    // The lane's entries of the stream-frame map (d_smap: index 2 * (symbol - 8) + lsb into the soft values, bit 15 = de-randomiser
    // sign) as byte offsets and sign masks: trellis inputs lane, lane + 32, ... and LICH bits lane, lane + 32, lane + 64.
    // (e & 0x8000) << 16 is the sign bit: m17_de_correlate_1 negates the soft value where the randomiser bit is set (m17_correlate.cpp:29)
    unsigned moff[9], msgn[9], loff[3], lsgn[3];
#pragma unroll
    for (int j = 0; j < 9; j++) {
        const int idx = 32 * j + lane;
        const unsigned e = idx < STREAM_NIN ? (unsigned)__ldg(smap + idx) : 0u;
        moff[j] = e & 0x1FFu; msgn[j] = (e & 0x8000u) << 16;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const unsigned e = __ldg(smap + STREAM_NIN + 32 * j + lane);
        loff[j] = e & 0x1FFu; lsgn[j] = (e & 0x8000u) << 16;
    }
    const float *sbase = syms + c * sym_pitch + sym_carry;
    const uint32_t base_g = sym_base ? (uint32_t)sym_base[c] : 0u;
    const m17b_frame_rec *frec = frames + c * fcap;
    float *soft2 = soft2_s[wid];
    float2 *s2 = (float2 *)soft2;
    // The group's record headers in one load (lane i: slot + i), the symbols of the next stream frame fetched while the current
    // one is worked on: a warp's frames are a serial chain otherwise (header -> symbols -> normaliser -> picks -> Golay tables).
    uint2 hdl = make_uint2(0, 0);
    if (lane < SG_FPW && slot + lane < send) hdl = *(const uint2 *)(frec + slot + lane);
    const bool parsed = (hdl.y >> 8) & M17B_F_PARSED;
    const int htype = hdl.y & 0xFF;
    const bool mine = htype == M17B_T_STREAM && parsed;
    // the other frame types go on k_decode_frames' list (dlist[0] = count): one atomic per warp that has any
    const bool other = parsed && (htype == M17B_T_LSF || htype == M17B_T_PACKET || htype == M17B_T_BERT);
    const unsigned om = __ballot_sync(FULL, other);
    if (om) {
        int base = 0;
        if (lane == 0) base = atomicAdd(dlist, __popc(om));
        base = __shfl_sync(FULL, base, 0);
        if (other) dlist[1 + base + __popc(om & ((1u << lane) - 1u))] = (int32_t)(c * fcap + slot + lane);
    }
    unsigned todo = __ballot_sync(FULL, mine);                                  // bit i: slot + i holds a parsed stream frame
    if (!todo) return;
    float v[6], vn[6];
    auto fetch = [&](int i, float (&dst)[6]) {
        const uint32_t off = __shfl_sync(FULL, hdl.x, i);
        const float *src = sbase + (int32_t)(off - base_g);
#pragma unroll
        for (int k = 0; k < 6; k++) dst[k] = src[lane + 32 * k];
    };
    fetch(__ffs(todo) - 1, v);
    // the frame whose Golay decode is still in flight: received word (lanes 0..3), its parity look-up, record index, normaliser
    bool pend = false;
    uint32_t pword = 0, pe1 = 0;
    int64_t pfidx = 0;
    float pcor = 0.0f;
    auto lich_finish = [&](uint32_t e, uint32_t wd, int64_t fi, float cr) {      // e = gerr[syndrome]: error weight << 12 | data error pattern
        const uint32_t w = lane < 4 ? ((wd >> 12) & 0xFFF) ^ (e & 0xFFF) : 0u;
        int ge = lane < 4 ? (int)(e >> 12) : 0;
        const uint32_t w0 = __shfl_sync(FULL, w, 0), w1 = __shfl_sync(FULL, w, 1), w2 = __shfl_sync(FULL, w, 2), w3 = __shfl_sync(FULL, w, 3);
        ge += __shfl_xor_sync(FULL, ge, 1);
        ge += __shfl_xor_sync(FULL, ge, 2);
        if (lane == 0) {
            StreamAux a;
            a.lw01 = (w0 << 12) | w1;                                            // pack_12_to_8_x4x6, m17_bit_utils.cpp:152-172
            a.lw23 = (w2 << 12) | w3;
            a.golay_e = (uint32_t)ge;
            a.cor = cr;
            *(uint4 *)(saux + fi) = *(const uint4 *)&a;
        }
    };
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        if (todo) fetch(__ffs(todo) - 1, vn);
        // demap normaliser from the 8 sync symbols, summed in order (m17_dsp.cpp:35-37)
        float sa = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; q++) sa += fabsf(__shfl_sync(FULL, v[0], q));
        const float cor = 8.0f / sa;
        // both soft values of every payload symbol once (m = in * mag, m17_dsp.cpp:38-41); the maps only pick and sign them
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int idx = lane + 32 * k - 8;
            if (idx >= 0) { const float m = v[k] * cor; s2[idx] = make_float2(-m, demap_lsb(m)); }
        }
        __syncwarp();
        // 4 x hard_decode_24_bits (m17_bit_utils.cpp:180-187: bit = soft >= 0) + m_17_golay_decode: ballot bit i = LICH bit i, and
        // the words are MSB first, so the reversed ballots read as one 96-bit big-endian string.  (Before the picks: the two
        // dependent table look-ups of the Golay decoder are then in flight while the 272 trellis inputs are written.)
        const unsigned B0 = __brev(__ballot_sync(FULL, __uint_as_float(__float_as_uint(soft2[loff[0]]) ^ lsgn[0]) >= 0.0f));
        const unsigned B1 = __brev(__ballot_sync(FULL, __uint_as_float(__float_as_uint(soft2[loff[1]]) ^ lsgn[1]) >= 0.0f));
        const unsigned B2 = __brev(__ballot_sync(FULL, __uint_as_float(__float_as_uint(soft2[loff[2]]) ^ lsgn[2]) >= 0.0f));
        const uint32_t word = lane == 0 ? B0 >> 8 : lane == 1 ? ((B0 & 0xFFu) << 16) | (B1 >> 16) : lane == 2 ? ((B1 & 0xFFFFu) << 8) | (B2 >> 24) : B2 & 0xFFFFFFu;
        // m_17_golay_decode (m17_golay.cpp:103-116) is two dependent table look-ups; they are spread over two frames so that
        // nothing waits for them: this frame's parity look-up and the previous frame's syndrome look-up are issued here, the
        // previous frame's LICH record is completed after this frame's picks.
        uint32_t e2 = 0, e1 = 0;
        if (lane < 4) {
            if (pend) e2 = __ldg(&gerr[(pword & 0xFFF) ^ pe1]);
            e1 = __ldg(&genc[(word >> 12) & 0xFFF]);
        }
        const int64_t fidx = c * fcap + slot + i;
        float *o = ssoft + fidx * STREAM_NIN + lane;
#pragma unroll
        for (int j = 0; j < 8; j++) o[32 * j] = __uint_as_float(__float_as_uint(soft2[moff[j]]) ^ msgn[j]);
        if (lane < STREAM_NIN - 256) o[256] = __uint_as_float(__float_as_uint(soft2[moff[8]]) ^ msgn[8]);
        if (soft_out) {
            float *so = soft_out + fidx * 368;
            for (int k = lane; k < 368; k += 32) so[k] = soft2[k];
        }
        if (pend) lich_finish(e2, pword, pfidx, pcor);
        pend = true; pword = word; pe1 = e1; pfidx = fidx; pcor = cor;
        __syncwarp();                                                            // soft2 is rewritten by the next frame
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = vn[k];
    }
    if (pend) {
        uint32_t e2 = 0;
        if (lane < 4) e2 = __ldg(&gerr[(pword & 0xFFF) ^ pe1]);
        lich_finish(e2, pword, pfidx, pcor);
    }
}
__global__ void __launch_bounds__(32) k_stream_acs(const float *__restrict__ ssoft, const StreamAux *__restrict__ saux, m17b_frame_rec *frames, int64_t fcap,
                                                   const int32_t *__restrict__ nframes, const int2 *__restrict__ frame_rng, int tiles_per_chan,
                                                   const uint16_t *__restrict__ g_crc) {
    __shared__ float tile[2][32][33];
    __shared__ uint16_t crc_tab[256];
    const int lane = threadIdx.x;
    const int64_t c = blockIdx.x / tiles_per_chan;
    const int tl = blockIdx.x % tiles_per_chan;
    int lo = 0, nfr;
    if (frame_rng) { const int2 r = frame_rng[c]; lo = r.x; nfr = r.y; }
    else nfr = (int)min((int64_t)nframes[c], fcap);
    const int slot = lo + tl * 32 + lane;
    m17b_frame_rec *rec = frames + c * fcap + slot;
    int type = -1, flags = 0;
    uint32_t w0 = 0;
    if (slot < nfr) { const uint2 hd = *(const uint2 *)rec; w0 = hd.x; type = hd.y & 0xFF; flags = (hd.y >> 8) & 0xFF; }
    const bool work = slot < nfr && (flags & M17B_F_PARSED) && type == M17B_T_STREAM;
    const unsigned wmask = __ballot_sync(0xffffffffu, work);
    if (!wmask) return;
    for (int i = lane; i < 256; i += 32) crc_tab[i] = g_crc[i];
    uint16_t dec[FrameStream::STEPS];
    viterbi_chunked<FrameStream, 2, STREAM_NIN>(ssoft + (c * fcap + lo + tl * 32) * STREAM_NIN + lane, wmask, tile, lane, dec);
    if (!work) return;
    __align__(4) uint8_t ob[32];
    viterbi_traceback<FrameStream>(dec, ob);
    constexpr uint32_t nbytes = FrameStream::NBYTES;
    uint16_t k = 0xFFFF;
    for (uint32_t i = 0; i < nbytes; i++) k = crc16_step(k, ob[i], crc_tab);
    const uint32_t crc = k;
    for (uint32_t i = nbytes; i < 32; i++) ob[i] = 0;
    const uint16_t *oh = (const uint16_t *)ob;                              // data[] as 15 half-words
    const uint4 ax = *(const uint4 *)(saux + c * fcap + slot);
    const uint32_t lw01 = ax.x, lw23 = ax.y, golay_e = ax.z;
    // the 64-byte record: words 0/1 (sym_off, type) and the votes/errors/variance fields come from the framer, the LICH, its
    // Golay error count and the demap normaliser from k_stream_gather
    uint32_t *rw = (uint32_t *)rec;
    const uint32_t w11_old = rw[11], w12 = rw[12];
    uint4 q0, q1, q2, q3;
    q0.x = w0;
    q0.y = (uint32_t)type | ((uint32_t)flags << 8) | (golay_e << 16) | (nbytes << 24);
    q0.z = ((lw01 >> 16) & 0xFF) | (((lw01 >> 8) & 0xFF) << 8) | ((lw01 & 0xFF) << 16) | (((lw23 >> 16) & 0xFF) << 24);
    q0.w = ((lw23 >> 8) & 0xFF) | ((lw23 & 0xFF) << 8) | ((uint32_t)oh[0] << 16);
    q1.x = oh[1] | ((uint32_t)oh[2] << 16);   q1.y = oh[3] | ((uint32_t)oh[4] << 16);
    q1.z = oh[5] | ((uint32_t)oh[6] << 16);   q1.w = oh[7] | ((uint32_t)oh[8] << 16);
    q2.x = oh[9] | ((uint32_t)oh[10] << 16);  q2.y = oh[11] | ((uint32_t)oh[12] << 16);
    q2.z = oh[13] | ((uint32_t)oh[14] << 16); q2.w = crc | (w11_old & 0xFFFF0000u);
    q3.x = w12; q3.y = ax.w; q3.z = 0; q3.w = 0;
    uint4 *r4 = (uint4 *)rec;
    r4[0] = q0; r4[1] = q1; r4[2] = q2; r4[3] = q3;
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

#define DECODE_NT 32
static size_t decode_smem() { return (size_t)DECODE_NT * 193 * 4; }

// frame_rng / max_frames: decode only the records [rng.x, rng.y) of each channel (at most max_frames of them).
// ssoft [nchan * fcap][STREAM_NIN] floats and saux [nchan * fcap] are the scratch between the two stream-frame kernels.
static int launch_decode(m17b_ctx *ctx, const float *syms, int64_t sym_pitch, int sym_carry, const int32_t *sym_base,
                         m17b_frame_rec *frames, int64_t fcap, const int32_t *nframes, int64_t nchan, float *soft_out, float *ssoft, StreamAux *saux, int32_t *dlist, cudaStream_t st,
                         cudaStream_t aux = nullptr, cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr,
                         const int2 *frame_rng = nullptr, int64_t max_frames = 0, int bert_on = 0) {
    const int tiles = (int)(((frame_rng ? max_frames : fcap) + DECODE_NT - 1) / DECODE_NT);
    // (set on every call: the attribute is per device, a process may drive several)
    CUDA_TRY(cudaFuncSetAttribute(k_decode_frames<DECODE_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem()));
    if ((int64_t)tiles * nchan * 32 > 0x7fffffffLL || nchan * fcap > 0x7fffffffLL) return M17B_E_ARG;          // (channel, slot) items and list entries are 32-bit
    const unsigned grid = (unsigned)(tiles * nchan);
    // dlist [1 + nchan * fcap]: the gather kernel walks every record header, decodes the stream frames' inputs and lists the others
    CUDA_TRY(cudaMemsetAsync(dlist, 0, sizeof(int32_t), st));
    k_stream_gather<<<(grid * (32 / SG_FPW) + SG_WARPS - 1) / SG_WARPS, SG_WARPS * 32, 0, st>>>(syms, sym_pitch, sym_carry, sym_base, frames, fcap, nframes, frame_rng, tiles * 32, nchan,
                                                                     ctx->d_smap, ctx->d_genc, ctx->d_gerr, ssoft, saux, soft_out, dlist);
    // the kernels touch disjoint records: run the (rare, long) LSF / packet one beside the stream frames' trellis pass
    cudaStream_t st2 = st;
    if (aux) { CUDA_TRY(cudaEventRecord(ev_fork, st)); CUDA_TRY(cudaStreamWaitEvent(aux, ev_fork, 0)); st2 = aux; }
    const unsigned dgrid = grid < 4u * 148u ? grid : 4u * 148u;
    k_decode_frames<DECODE_NT><<<dgrid, DECODE_NT, decode_smem(), st2>>>(syms, sym_pitch, sym_carry, sym_base, frames, fcap, dlist, soft_out,
                                                                       ctx->d_crc, ctx->d_genc, ctx->d_gerr, bert_on);
    k_stream_acs<<<grid, 32, 0, st>>>(ssoft, saux, frames, fcap, nframes, frame_rng, tiles, ctx->d_crc);
    KERNEL_CHECK();
    if (aux) { CUDA_TRY(cudaEventRecord(ev_join, aux)); CUDA_TRY(cudaStreamWaitEvent(st, ev_join, 0)); }
    return M17B_OK;
}

extern "C" int m17b_rx_parse_frames(m17b_ctx *ctx, const float *d_sym, const uint8_t *d_type, int64_t n, m17b_frame_rec *d_rec, float *d_soft, void *stream) {
    if (!ctx || !d_sym || !d_type || !d_rec || n < 0 || n > 0x7fffffffLL / 192) return M17B_E_ARG;
    if (n == 0) return M17B_OK;
    cudaStream_t st = as_stream(stream);
    int32_t *d_n;
    float *d_ssoft;
    StreamAux *d_saux;
    CUDA_TRY(cudaMallocAsync((void **)&d_n, sizeof(int32_t), st));
    CUDA_TRY(cudaMallocAsync((void **)&d_ssoft, (size_t)n * STREAM_NIN * sizeof(float), st));
    CUDA_TRY(cudaMallocAsync((void **)&d_saux, (size_t)n * sizeof(StreamAux), st));
    int32_t *d_dlist;
    CUDA_TRY(cudaMallocAsync((void **)&d_dlist, (size_t)(n + 1) * sizeof(int32_t), st));
    k_set_i32<<<1, 1, 0, st>>>(d_n, (int32_t)n);
    k_parse_init<<<grid_for(n, 256), 256, 0, st>>>(d_type, n, d_rec);
    KERNEL_CHECK();
    int rc = launch_decode(ctx, d_sym, 0, 0, nullptr, d_rec, n, d_n, 1, d_soft, d_ssoft, d_saux, d_dlist, st);
    CUDA_TRY(cudaFreeAsync(d_dlist, st));
    CUDA_TRY(cudaFreeAsync(d_n, st));
    CUDA_TRY(cudaFreeAsync(d_ssoft, st));
    CUDA_TRY(cudaFreeAsync(d_saux, st));
    return rc;
}

// ---------------------------------------------------------------- config-4 microbenchmark: punctured soft frames in, bytes out
// d_soft [n][NIN] already de-randomised/de-interleaved (i.e. so[] of m17_rx_parse.cpp:91-95); de-puncture + Viterbi + pack.
// One thread per frame as in the fused decoder, one warp per CTA.  What bounds this kernel is how many frames an SM can hold:
// a frame's inputs are consumed in order, so only the next VP_CHUNK trellis steps' worth (at most 32 floats per frame) is staged
// at a time -- two 4 KB tiles per warp, filled one chunk ahead with cp.async (coalesced 128-byte row pieces, pitch 33 so the
// per-thread column reads hit 32 banks) -- and the survivor words go to per-thread local memory (interleaved by the hardware,
// so the accesses coalesce; written once, read once in the traceback).  8.4 KB of shared memory per warp instead of 44 KB.
template <class F, int PAT, int NIN>
__global__ void __launch_bounds__(32) k_viterbi_punct(const float *__restrict__ soft, int64_t n, uint8_t *__restrict__ bytes) {
    __shared__ float tile[2][32][33];
    const int lane = threadIdx.x;
    const int64_t f0 = (int64_t)blockIdx.x * 32, f = f0 + lane;
    const int nrows = (int)min((int64_t)32, n - f0);
    uint16_t dec[F::STEPS];
    viterbi_chunked<F, PAT, NIN>(soft + f0 * NIN + lane, nrows == 32 ? 0xffffffffu : (1u << nrows) - 1u, tile, lane, dec);
    if (f >= n) return;
    viterbi_traceback<F>(dec, bytes + f * F::NBYTES);
}
template <class F, int PAT, int NIN> static int launch_vp(const float *d_soft, int64_t n, uint8_t *d_bytes, cudaStream_t st) {
    k_viterbi_punct<F, PAT, NIN><<<grid_for(n, 32), 32, 0, st>>>(d_soft, n, d_bytes);
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
