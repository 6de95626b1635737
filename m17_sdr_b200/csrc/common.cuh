// common.cuh -- shared declarations for libm17b200 (single translation unit: m17b200.cu includes every part).
//
// Arithmetic contract: the whole TU is compiled with -fmad=false (no FMA contraction), IEEE sqrt/div
// (nvcc defaults, no fast-math), no flush-to-zero.  The reference is a generic x86-64 -O3 build
// (makefile:6: no -mfma, no -ffast-math), so with the same operand types and the same evaluation
// order the fp32 results are bit-identical.  m17b_ctx_create() runs a self-test that fails loudly if
// the library was built with contraction enabled.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "m17b200.h"

#define M17B_NF 40          // polyphase branches      m17_rx_sync.cpp:3
#define M17B_FN 31          // taps per branch         m17_rx_sync.cpp:4
#define M17B_SYM_CARRY 192  // symbols of the previous call kept in front of the stream buffer

void m17b_set_cuda_error(cudaError_t e, const char *file, int line);

#define CUDA_TRY(x)                                                              \
    do {                                                                         \
        cudaError_t e__ = (x);                                                   \
        if (e__ != cudaSuccess) { m17b_set_cuda_error(e__, __FILE__, __LINE__); return M17B_E_CUDA; } \
    } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

// ---------------------------------------------------------------- soft-bit gather maps (constant memory)
// One uint16 per CODED position of a frame type (or per LICH bit): where in the 192-symbol frame the soft
// value comes from after de-randomise (m17_correlate.cpp:27-31), de-interleave (m17_interleave.cpp:8-12)
// and de-puncture (m17_puncture.cpp:47-79) are folded together.
//   bits 0..7 : frame symbol index 8..191        bit 8 : 1 = LSB soft bit (|m|-0.6666), 0 = MSB (-m)
//   bit 9     : 1 = negate (randomiser bit set)  0xFFFF: punctured position -> 0.0f erasure
#define STREAM_NIN 272       // kept (unpunctured) trellis inputs of a stream frame: 296 coded bits less the 24 P2 removes
#define MAP_ERASE 0xFFFFu
#define MAP_LSB   0x100u
#define MAP_NEG   0x200u

struct __align__(16) GatherMaps {
    uint16_t p1[488];    // LSF    : 244 steps
    uint16_t p2[296];    // stream : 148 steps (type-3 bits 96..367)
    uint16_t p3[420];    // packet : 210 steps
    uint16_t lich[96];   // stream : 4 x 24 Golay bits
    uint16_t bert[402];  // BERT   : 201 steps (197 PRBS9 bits + 4 tail); the 369th kept bit is never sent -> erasure
};
// TX-side maps: for final (interleaved, randomised) bit i of a frame -> which type-3 bit feeds it
// (QPP is an involution) and, per type-3 bit, which coded (pre-puncture) position it is.
#define VP_CHUNK 16          // trellis steps per staged chunk of the stand-alone punctured Viterbi kernel (at most 32 kept inputs)
struct __align__(16) PunctSteps {
    uint8_t keep[3][244];      // per trellis step: bit0/bit1 = first/second coded bit survives P1/P2/P3
    uint16_t koff[3][17];      // kept inputs before trellis step VP_CHUNK * j (prefix of popcount(keep)), j = 0 .. ceil(steps / VP_CHUNK)
};
struct __align__(16) TxMaps {
    uint16_t qpp[368];   // pi(i) = (45 i + 92 i^2) mod 368          m17_interleave.cpp:5
    uint8_t  rnd[368];   // randomiser bits, MSB first               m17_correlate.cpp:35-42
    uint16_t unp1[368];  // kept index -> coded position, P1         m17_puncture.cpp:4-6
    uint16_t unp2[368];  // P2 (first 368 kept positions)            m17_puncture.cpp:8
    uint16_t unp3[368];  // P3                                       m17_puncture.cpp:10
};

// per-GPU context
struct m17b_ctx {
    int device;
    uint16_t *d_crc;     // [256]   CRC-16/M17 byte table            m17_crc.cpp:8-24
    uint16_t *d_crcpos;  // [30][256] + [1]  CRC of a 30-byte LSF as an XOR of per-position contributions (k_post): entry [i][b] = the byte
                         //          table entry of b carried through the 29 - i zero-byte steps that follow it, [30][0] = the 0xFFFF preset carried through 30
    uint16_t *d_genc;    // [4096]  Golay parity table               m17_golay.cpp:31-40
    uint16_t *d_gerr;    // [4096]  Golay syndrome table             m17_golay.cpp:49-72
    float    *d_mf;      // [40][31] matched-filter bank             m17_rx_sync.cpp:13
    float    *d_md;      // [40][31] derivative bank                 m17_rx_sync.cpp:14
    uint8_t  *d_prbs;    // [511]   PRBS9 sequence                   m17_prbs9.cpp:16-26
    uint16_t *d_smap;    // [272 + 96] stream frame: gather-map entries of the kept trellis inputs in order, then the LICH bits (decode.cuh)
    float     h_mf[M17B_NF * M17B_FN], h_md[M17B_NF * M17B_FN];
};

// ---------------------------------------------------------------- per-channel persistent RX state
// Everything the reference keeps in file statics for one channel (SURVEY 8b "persistent state").
struct RxChanState {
    // AFC: NCO phase (dsp_nco_mixer's static acc, m17_dsp.cpp:391) and loop state m_afc_delta (radio.cpp:9); afc.cuh
    double nco_acc;
    float afc_delta;
    int   afc_pad;
    // front end: dsp_arctan_disc2 statics (m17_dsp.cpp:195-196)
    float z0re, z0im, z1re, z1im;
    float nz0re, nz0im, nz1re, nz1im;   // written by the front-end kernel, committed by the sync kernel (no read/write race
                                        // between the block-parallel items of one channel)
    int   disc_count;
    // timing loop (m17_rx_sync.cpp:7-12,78)
    int   clk, thr, index;
    float sum, dif;
    float tail[30];          // last 30 (mean-removed) discriminator samples = m_buff[1..30]
    // framer (m17_rx_frame.cpp:14-18,104)
    int   flock, fclk, ferr;
    float win[8];            // m_sync sliding window (zeros after reset_sync)
    float head[8];           // first 8 symbols of the frame being collected (m_f_sym[0..7])
    int   frame_start;       // stream index of m_f_sym[0]
    int   sym_total;         // symbols emitted so far (stream index of the next symbol)
    int   prev_n;            // symbols written by the previous call (for the carry copy)
    // parser (m17_rx_parse.cpp:5-7)
    int   packet_idx;
    uint8_t lsf[2][32];      // m_lsf[2][30] padded
    uint8_t packet[800];     // m_packet
    // BERT receive: PRBS9 checker statics m_rx_idx, m_rx_state, m_rx_bad, m_rx_good, m_rx_eq_cnt, m_rx_dif_cnt
    // (m17_prbs9.cpp:7-12) and two running totals (bits checked while in sync, bit errors among them)
    uint16_t prbs_idx, prbs_bad, prbs_good, prbs_eq, prbs_dif, prbs_pad;
    int      prbs_state;
    uint32_t bert_bits, bert_errs;
    // instrumentation: SM cycles the timing-loop kernel spent on this channel in its last launch, rounds it ran
    unsigned long long dbg_cycles;
    uint32_t dbg_rounds, dbg_pad;
    unsigned long long dbg_phase[6];   // cycles in: staging, timing loop, emission, framer, carry, (spare)
};

// ---------------------------------------------------------------- packed fp32 pairs (sm_100 FMUL2 / FADD2 / FFMA2)
// Each half is an independent IEEE round-to-nearest operation, so results are bit-identical to the scalar forms; they
// only halve the instruction count.  NOTE: ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 even
// with --fmad=false, so a packed product must never feed a packed add directly -- products are consumed by scalar adds.
typedef unsigned long long f32x2;
#define F32X2_ONE 0x3F8000003F800000ull      // (1.0f, 1.0f), passed to kernels as an ARGUMENT (see the note above)
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint16_t crc16_step(uint16_t crc, uint8_t byte, const uint16_t *tab) {
    // m17_crc.cpp:30-33
    return (uint16_t)((crc << 8) ^ tab[((crc >> 8) ^ byte) & 0xFF]);
}

// demap one soft bit from a frame's symbols (m17_dsp.cpp:35-42): m = sym*cor; MSB soft = -m;
// LSB soft = (float)(fabs(m) - 0.6666) with the subtraction in double, as the double literal forces.
// reference formulation of the LSB soft value, kept for the exhaustive self-test
__device__ __forceinline__ float demap_lsb_ieee(float m) { return __double2float_rn((double)fabsf(m) - 0.6666); }
// The same value in five fp32 adds (no F2F / DADD: B200's fp64 and conversion pipes made this one expression 30 % of the frame
// decoder's stall samples).  0.6666 = c_hi + c_lo + 1.1e-16 with c_hi, c_lo floats; s = a - c_hi rounds, e = (-c_hi - s) + a
// recovers what the rounding dropped wherever it matters, and s + (e - c_lo) rounds once more.  That this equals
// (float)((double)a - 0.6666) -- two roundings of the exact difference -- for EVERY finite a >= 0 is established by enumeration
// (m17b_selftest_demap on the GPU, all 2^32 bit patterns of m; benchmarks/demap_lsb_enum.py on the CPU).  a = Inf would turn
// into NaN in the error term; from 2^25 up the value is a itself, which the guard returns.
__device__ __forceinline__ float demap_lsb(float m) {
    const float a = fabsf(m);
    constexpr float c_hi = 0.6666f;
    constexpr float c_lo = (float)(0.6666 - (double)0.6666f);
    const float s = a - c_hi;
    const float e = (-c_hi - s) + a;
    const float v = s + (e - c_lo);
    return a >= 33554432.0f ? a : v;
}
__device__ __forceinline__ float demap_soft(float sym, float cor, bool lsb) {
    float m = sym * cor;
    return lsb ? demap_lsb(m) : -m;
}
