/*
 * m17b200.h -- C ABI of libm17b200.so: the batched, B200-native (sm_100a) replacement for the
 * m17gismo baseband hot path of G4GUO/m17_sdr.
 *
 * The reference has no FFI layer: its boundary is the set of free C++ functions declared in
 * m17gismo/m17defines.h and linked statically.  Every entry point below is the batched
 * form of one of those functions and cites the reference definition it replaces (paths relative to
 * /root/reference/m17gismo).  The header-only C++ shim include/m17gismo_b200.hpp re-creates the
 * original names/signatures at batch = 1 on top of this ABI.
 *
 * Conventions
 *   - plain C, no torch/CUDA types: streams are passed as void* (a cudaStream_t; NULL = default stream);
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers; all buffers are caller-owned;
 *   - every function returns 0 on success or a negative M17B_E_* code (the reference returns void or a
 *     length and has no error channel; lengths are implied by the arguments here);
 *   - functions only enqueue work on the stream unless their name ends in _host;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with M17B_E_CUDA.
 *   - a context / rx / tx object is thread-compatible, not thread-safe (the reference runs the whole
 *     path on one thread, m17_tx_rx.cpp:238-257).
 */
#ifndef M17B200_H
#define M17B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M17B_VERSION 100

#define M17B_OK            0
#define M17B_E_ARG        -1   /* bad argument                                   */
#define M17B_E_CUDA       -2   /* CUDA runtime error (see m17b_last_cuda_error)  */
#define M17B_E_NOMEM      -3
#define M17B_E_UNSUPPORTED -4  /* feature not available in this build             */
#define M17B_E_CAPACITY   -5   /* nblocks exceeds the capacity given at create   */

#define M17B_BLOCK_SAMPLES 1920   /* N_SAMPLES         m17defines.h:17 */
#define M17B_DISC_PER_BLOCK 384   /* N_SAMPLES/5       m17_dsp.cpp:463 */
#define M17B_FRAME_SYMS     192   /* FRAME_SYM_LENGTH  m17defines.h:66 */
#define M17B_SOFT_BITS      368
#define M17B_SYM_CAP_PER_BLOCK 200

/* sync / frame types, m17_rx_frame.cpp:5-12 */
enum { M17B_T_PREAMBLE = 0, M17B_T_LSF = 1, M17B_T_STREAM = 2, M17B_T_PACKET = 3, M17B_T_BERT = 4, M17B_T_EOT = 5 };

/* record flags */
#define M17B_F_SYNC_OK    0x01 /* m17_locked_sync_check() passed            m17_rx_frame.cpp:144 */
#define M17B_F_PARSED     0x02 /* m17_rx_parse() semantics applied          m17_rx_frame.cpp:145,153 */
#define M17B_F_LOS        0x04 /* framer dropped lock on this frame         m17_rx_frame.cpp:138-150 */
#define M17B_F_DELIVERED  0x08 /* stream payload passed upward              m17_rx_parse.cpp:148-158 */
#define M17B_F_LSF_EVENT  0x10 /* parse_lsf() would have run                m17_rx_parse.cpp:82,99 */
#define M17B_F_PKT_EOF    0x20 /* packet frame with EOF bit                 m17_rx_parse.cpp:173 */

/*
 * One record per completed 192-symbol frame.  Replaces the reference's synchronous up-calls
 * (m17_aos/m17_los, m17_db_*, gui_*, m17_txrx_spkr_audio, m17_net_new_rx_data:
 * m17_rx_parse.cpp:20-32,128,145,153-157; m17_rx_frame.cpp:139,149,169).
 */
typedef struct {
    int32_t  sym_off;      /* index of the frame's first symbol in the channel's emitted symbol stream: a counter modulo 2^32
                              since the last m17b_rx_reset (it wraps after 10.3 days of continuous symbols at 4800/s; all kernels
                              only ever use differences of indices, consumers should compare (int32_t)(a - b) likewise)   */
    uint8_t  type;         /* M17B_T_*  (m17_sync_check winner)                                        */
    uint8_t  flags;        /* M17B_F_*                                                                 */
    uint8_t  golay_err;    /* stream frames: sum of the 4 Golay error counts                            */
    uint8_t  nbytes;       /* bytes in data[]: 30 LSF / 18 stream (FN+16) / 26 packet / 0               */
    uint8_t  lich[6];      /* stream frames: corrected LICH chunk                                       */
    uint8_t  data[30];     /* Viterbi output, MSB first, tail discarded                                 */
    uint16_t crc;          /* CRC-16/M17 over data[0..nbytes): 0 = valid for LSF frames                 */
    uint8_t  votes;        /* sync-word sign mismatches                                                 */
    uint8_t  frame_errors; /* consecutive bad sync words so far                                         */
    float    variance;     /* sync-word magnitude spread                                                */
    float    cor;          /* demap normaliser                                                          */
    uint8_t  rsvd[8];
} m17b_frame_rec;          /* 64 bytes */

#define M17B_EV_AOS 1
#define M17B_EV_LOS 2
typedef struct { int32_t sym_idx; int32_t kind; } m17b_event_rec;

typedef struct m17b_ctx m17b_ctx;   /* per-GPU: lookup tables, filter banks (replaces the init chain main.cpp:108-126) */
typedef struct m17b_rx  m17b_rx;    /* per-batch RX: per-channel state blobs + stage buffers                          */
typedef struct m17b_tx  m17b_tx;    /* per-batch TX: per-channel modulator/formatter state                            */
typedef struct m17b_eq  m17b_eq;    /* per-batch equaliser state                                                      */

int         m17b_version(void);
const char *m17b_error_string(int code);
const char *m17b_last_cuda_error(void);

/* m17_prbs9_init, m17_crc_init, m17_init_conv, m17_init_de_correlate, m17_dsp_init, m17_fmt_init,
   m17_golay_init, m17_rx_sync_init, m17_mod_init  (main.cpp:108-126) */
int m17b_ctx_create(int device, m17b_ctx **out);
int m17b_ctx_destroy(m17b_ctx *ctx);
/* filter design, host side, double math: m17_dsp_build_rrc_filter (m17_dsp.cpp:295-315),
   m17_dsp_set_filter_gain (m17_dsp.cpp:420-429) */
int m17b_build_rrc_filter(float *h_taps, float rolloff, int ntaps, int samples_per_symbol);
int m17b_set_filter_gain(float *h_taps, float gain, int stride, int ntaps);
/* copies of the RX polyphase banks m_mf/m_md [40][31] (m17_rx_sync.cpp:12-14,101-123) */
int m17b_get_sync_taps(const m17b_ctx *ctx, float *h_mf, float *h_md);

/* ------------------------------------------------------------------ batched primitives (device buffers) */
/* m17_crc_array_encode (m17_crc.cpp:26-35): n arrays of len bytes, stride bytes apart -> n CRCs */
int m17b_crc_array_encode(m17b_ctx *ctx, const uint8_t *d_in, int64_t stride, int len, int64_t n, uint16_t *d_crc, void *stream);
/* m17_golay_encode (m17_golay.cpp:94-102) */
int m17b_golay_encode(m17b_ctx *ctx, const uint16_t *d_data, int64_t n, uint32_t *d_words, void *stream);
/* m_17_golay_decode (m17_golay.cpp:103-116): corrected data + error count 0..4 */
int m17b_golay_decode(m17b_ctx *ctx, const uint32_t *d_words, int64_t n, uint16_t *d_data, uint8_t *d_err, void *stream);
/* m17_conv_encode_8 (m17_conv.cpp:53-71): n x nbytes -> n x 2*(8*nbytes+4) bit-bytes */
int m17b_conv_encode_8(m17b_ctx *ctx, const uint8_t *d_in, int nbytes, int64_t n, uint8_t *d_out, void *stream);
/* m17_conv_encode_1 (m17_conv.cpp:33-49): n x nbits -> n x 2*(nbits+4) */
int m17b_conv_encode_1(m17b_ctx *ctx, const uint8_t *d_in, int nbits, int64_t n, uint8_t *d_out, void *stream);
/* m17_viterbi_decode (m17_conv.cpp:148-168): n x len soft values -> n x len/2 bit-bytes (out[0]=0 quirk kept) */
int m17b_viterbi_decode(m17b_ctx *ctx, const float *d_soft, int len, int64_t n, uint8_t *d_bits, void *stream);
/* m17_punc_p1/p2/p3 (m17_puncture.cpp:12-41): returns kept length through *out_len (host int) */
int m17b_punc(m17b_ctx *ctx, int pattern, const uint8_t *d_in, int len, int64_t n, uint8_t *d_out, int *out_len, void *stream);
/* m17_de_punc_p1/p2/p3 (m17_puncture.cpp:47-79): len = OUTPUT length; in_len = kept length */
int m17b_de_punc(m17b_ctx *ctx, int pattern, const float *d_in, int in_len, int len, int64_t n, float *d_out, void *stream);
/* m17_interleave / m17_de_interleave (m17_interleave.cpp:3-12), n frames of 368 */
int m17b_interleave(m17b_ctx *ctx, const uint8_t *d_in, int64_t n, uint8_t *d_out, void *stream);
int m17b_de_interleave(m17b_ctx *ctx, const float *d_in, int64_t n, float *d_out, void *stream);
/* m17_de_correlate_8 / _1(uint8) / _1(float) (m17_correlate.cpp:11-31); in-place allowed */
int m17b_de_correlate_8(m17b_ctx *ctx, uint8_t *d_io, int len, int64_t n, void *stream);
int m17b_de_correlate_1_u8(m17b_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, int len, int64_t n, void *stream);
int m17b_de_correlate_1_f32(m17b_ctx *ctx, const float *d_in, float *d_out, int len, int64_t n, void *stream);
/* m17_dsp_demap_frame (m17_dsp.cpp:82-95): n x 192 symbols -> n x 368 soft bits */
int m17b_demap_frame(m17b_ctx *ctx, const float *d_sym, int64_t n, float *d_soft, void *stream);
/* m17_dsp_demap_symbol (m17_dsp.cpp:35-42): n symbols, each with its own normaliser -> d_out [n][2] = {-m, |m| - 0.6666} */
int m17b_demap_symbols(m17b_ctx *ctx, const float *d_in, const float *d_mag, int64_t n, float *d_out, void *stream);
/* m17_dsp_decimating_filter (m17_dsp.cpp:438-449): n rows in_pitch floats apart, each len samples (+ flen-1 of look-ahead),
   out[k] = sum_j in[k*stride+j]*coffs[j]; d_out [n][ceil(len/stride)]; the output length is returned through *out_len */
int m17b_dsp_decimating_filter(m17b_ctx *ctx, const float *d_in, int64_t in_pitch, const float *d_coffs, int stride, int flen, int len, int64_t n,
                               float *d_out, int *out_len, void *stream);
/* m17_prbs9_rx_check (m17_prbs9.cpp:40-64) on n bit sequences (d_bits [n][nbits], one bit per byte) with persistent checker
   state d_state uint32 [n][8] (layout of m17b_rx_get_bert; all zero = m17_prbs9_rx_reset in a fresh process) */
int m17b_prbs9_rx_check(m17b_ctx *ctx, const uint8_t *d_bits, int nbits, int64_t n, uint32_t *d_state, void *stream);
/* m17_sync_check (m17_rx_frame.cpp:47-81): n x 8 symbols -> type, votes, variance */
int m17b_sync_check(m17b_ctx *ctx, const float *d_vec, int64_t n, uint8_t *d_type, uint8_t *d_votes, float *d_var, void *stream);
/* m17_prbs9_tx_load (m17_prbs9.cpp:27-32): n sequences of len bits starting at phase start[i] (NULL = 0) */
int m17b_prbs9_tx_load(m17b_ctx *ctx, const int32_t *d_start, int len, int64_t n, uint8_t *d_out, void *stream);
/* m17_rx_parse for n independent frames (m17_rx_parse.cpp:185-226) WITHOUT cross-frame LICH state:
   d_sym [n][192], d_type [n]; fills type-dependent record fields; d_soft (optional) [n][368] demapped bits */
int m17b_rx_parse_frames(m17b_ctx *ctx, const float *d_sym, const uint8_t *d_type, int64_t n, m17b_frame_rec *d_rec, float *d_soft, void *stream);
/* config-4 microbenchmark path: n punctured soft frames (pattern 1/2/3: 368/272/368 values) ->
   depuncture + Viterbi + pack -> n x {30,18,26} bytes */
int m17b_viterbi_punctured(m17b_ctx *ctx, int pattern, const float *d_soft, int64_t n, uint8_t *d_bytes, void *stream);

/* ------------------------------------------------------------------ RX chain */
/* nchan channels, at most max_blocks 40-ms blocks per call */
int m17b_rx_create(m17b_ctx *ctx, int64_t nchan, int64_t max_blocks, m17b_rx **out);
int m17b_rx_destroy(m17b_rx *rx);
/* m17_rx_init + m17_rx_sync_init initial state for every channel (m17_rx_frame.cpp:183-186, m17_rx_sync.cpp:124-127) */
int m17b_rx_reset(m17b_rx *rx, void *stream);
/* m17_rx_init / m17_rx_lost alone (m17_rx_frame.cpp:179-186): reset_sync() + m_flock = false for every channel; the timing loop,
   the discriminator history, the LICH cache and the counters are left as they are */
int m17b_rx_framer_reset(m17b_rx *rx, void *stream);
/* radio_set_afc_on / radio_set_afc_off (radio.cpp:146-152): with AFC on, m17b_dsp_rx runs dsp_nco_mixer + radio_afc
   (m17_dsp.cpp:390-408, radio.cpp:196-208); the loop is closed through the framer, so a channel's blocks are serial through
   the whole chain: the front end of each block then runs inside the timing-loop kernel (one launch per call).  Off (the
   reference default) uses the block-parallel front end.  No effect on m17b_rx_baseband. */
int m17b_rx_set_afc(m17b_rx *rx, int on, void *stream);
/* Equaliser option.  The reference ships a 5-tap T/2 RLS equaliser (m17_equalize.cpp) that nothing calls.  With the
   option on, every symbol of the timing loop is paired with the matched filter's output half a symbol earlier (the same polyphase
   branch, rx_sync_filter m17_rx_sync.cpp:25-31 one sample before), the pairs go through eq_train_unknown (m17_equalize.cpp:185-213)
   and the framer (m17_rx_symbols, m17_rx_frame.cpp:173) sees the equaliser's output; eq_open (:217-224) runs when the option is turned
   on and on m17b_rx_reset.  Off (the default) is upstream behaviour.  Not available together with AFC (M17B_E_ARG).  As upstream's
   code stands the option degrades reception (its decision levels assume unit-scale symbols): it exists to evaluate the
   reference's equaliser in the chain, bit-exact against the same wiring of the reference's own functions. */
int m17b_rx_set_equaliser(m17b_rx *rx, int on, void *stream);
/* BERT receive.  The reference sends BERT frames (m17_fmt_add_bert_frame) but its decode_bert_frame is empty
   (m17_rx_parse.cpp:178-180) and m17_prbs9_rx_check (m17_prbs9.cpp:40-64) is never called.  With on != 0, BERT frames are
   de-punctured (P2, 402 coded bits), Viterbi-decoded (201 steps) into data[0..25) and their 197 PRBS9 bits go through the
   reference's checker.  Off (the default) reproduces upstream: BERT records carry no data. */
int m17b_rx_set_bert(m17b_rx *rx, int on);
/* checker state per channel, d_out uint32 [nchan][8]: m_rx_state, m_rx_idx, m_rx_bad, m_rx_good, m_rx_eq_cnt, m_rx_dif_cnt
   (m17_prbs9.cpp:7-12), then two running totals: bits checked while in sync, bit errors among them */
int m17b_rx_get_bert(m17b_rx *rx, uint32_t *d_out, void *stream);
/* m17_dsp_rx (m17_dsp.cpp:461-476) for nchan channels x nblocks blocks: d_iq = int16 [nchan][nblocks*1920][2] */
int m17b_dsp_rx(m17b_rx *rx, const int16_t *d_iq, int64_t nblocks, void *stream);
/* baseband seam (m17_test.cpp:49-51): m17_rx_sync_samples + m17_rx_symbols on d_disc = float [nchan][nblocks*384], 16-byte aligned */
int m17b_rx_baseband(m17b_rx *rx, const float *d_disc, int64_t nblocks, void *stream);
/* symbol seam: m17_rx_symbols / m17_rx_sym (m17_rx_frame.cpp:126-177) on symbols supplied by the caller (another demodulator, an
   equaliser in front of the framer): framer FSM, frame decode, LICH / packet post stage.  d_syms float [nchan][pitch],
   d_nsym int32 [nchan] = symbols of each channel (<= max_blocks*200).  Results as for m17b_dsp_rx, reported as one block. */
int m17b_rx_symbols(m17b_rx *rx, const float *d_syms, int64_t pitch, const int32_t *d_nsym, void *stream);
/* device views of the last call's results (valid until the next call on this rx) */
typedef struct {
    int64_t nchan, nblocks;
    const m17b_frame_rec *d_frames; int64_t frame_cap; const int32_t *d_nframes;   /* [nchan][frame_cap], [nchan] */
    const float *d_syms; int64_t sym_pitch, sym_carry; const int32_t *d_nsym;      /* [nchan][sym_pitch] (new symbols start at sym_carry), [nchan][nblocks] */
    const int32_t *d_sym_base;                                                     /* [nchan] stream index of d_syms[c][sym_carry] */
    const float *d_disc; const float *d_mean;                                      /* [nchan][nblocks][384] raw, [nchan][nblocks] (NULL on the baseband seam) */
    const m17b_event_rec *d_events; int64_t event_cap; const int32_t *d_nevents;   /* [nchan][event_cap], [nchan] */
    const uint64_t *d_stats;                                                       /* [nchan][8]: frames, stream frames, golay errs, delivered, aos, los, lsf events, symbols */
} m17b_rx_view;
int m17b_rx_get_view(m17b_rx *rx, m17b_rx_view *out);
/* end-to-end form with HOST buffers (pinned or pageable): H2D of the IQ, the chain, D2H of records.
   h_frames [nchan][frame_cap] (frame_cap from m17b_rx_frame_cap), h_nframes [nchan]; synchronises the stream. */
int64_t m17b_rx_frame_cap(const m17b_rx *rx);
/* sticky capacity flags since the last m17b_rx_reset (synchronises the device): bit 0 = m17b_rx_symbols was handed more than
   max_blocks*200 symbols for some channel and ignored the excess.  Record and event buffers are sized for the worst case of
   every entry point (frame_cap = ceil(max_blocks*200/192) + 2) and cannot overflow. */
int m17b_rx_get_overflow(m17b_rx *rx, int *h_flags);
int m17b_dsp_rx_host(m17b_rx *rx, const int16_t *h_iq, int64_t nblocks, m17b_frame_rec *h_frames, int32_t *h_nframes, void *stream);
/* bench instrumentation: mark stage boundaries with CUDA events on the launching stream; stage_ms returns the device
   time of {front end, matched filter+timing+framer, frame decode, LICH/packet post} of device call number call_index
   (0-based, counted since timing was switched on; the last 64 calls are kept).  Not recorded by the _host entry point. */
int m17b_rx_set_timing(m17b_rx *rx, int on);
int m17b_rx_stage_ms(m17b_rx *rx, int64_t call_index, float *h_out4);
/* Time-sliced pipelining of m17b_dsp_rx / m17b_rx_baseband: the call is cut into slices of `blocks` 40-ms blocks and the
   front end of slice k+1, the timing loop + framer of slice k and the frame decode of slice k-1 run concurrently on
   internal streams (results are identical).  0 = no slicing (stages strictly in sequence), the default. */
int m17b_rx_set_slice_blocks(m17b_rx *rx, int blocks);
/* Scheduling knob: run the batch as `groups` (0..8) contiguous channel groups, each a complete chain on its own stream, so
   that one group's latency-bound timing loop shares the SMs with another group's front end / decode.  Channels are
   independent (m17_dsp_rx keeps all state per receiver), so results do not depend on it.  0 / 1 = one chain; -1 = automatic
   (the default: 4 groups from 512 channels up, one chain below). */
int m17b_rx_set_chan_groups(m17b_rx *rx, int groups);
/* instrumentation: d_out uint64 [nchan][8] = {SM cycles, speculation rounds, cycles in staging / timing loop / emission / framer /
   carry (only in builds with -DM17B_PHASE_CLOCKS), spare} the one-warp-per-channel timing-loop kernel spent on each channel in its
   last launch (the kernel's time is that of its slowest channel) */
int m17b_rx_debug_sync(m17b_rx *rx, uint64_t *d_out, void *stream);
/* number of kernels the last m17b_dsp_rx / m17b_rx_baseband call launched */
int m17b_rx_last_launches(const m17b_rx *rx);

/* exhaustive check that the front end's fast limiter arithmetic equals the reference formulation (double product,
   sqrtf, 1.0/m) bit-for-bit on raw IQ words [first, first+count) of the 2^32 possible int16 pairs; returns the number
   of mismatching samples (must be 0) and optionally the first dump_cap offending raw words.  The full range takes a few
   seconds on a B200. */
int m17b_selftest_frontend(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream);
/* the same for the limiter's normaliser by itself (dsp_limit, m17_dsp.cpp:412-419: m = sqrtf(re*re + im*im), 1.0 / m): the fast
   form's m and g against IEEE sqrt and division on the float bit patterns [first, first+count) of s = re*re + im*im, whatever
   re and im were (the AFC front end limits mixer outputs, not int16 grid points); dump = {bits(s), m_fast, m_ieee, g_fast, g_ieee} */
int m17b_selftest_limiter(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream);

/* exhaustive check that the frame decoder's fp32 form of the LSB soft value, (float)(fabs(m) - 0.6666) with the subtraction in
   double (m17_dsp_demap_frame / _symbol, m17_dsp.cpp:40-41,91), and its sign-only hard decision (hard_decode_24_bits,
   m17_bit_utils.cpp:180-187) equal the double formulation on the float bit patterns [first, first+count) of m = sym * cor;
   returns the number of mismatches (must be 0) and optionally the first dump_cap offenders as {bits(m), fast, reference}. */
int m17b_selftest_demap(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream);

/* ------------------------------------------------------------------ TX chain */
/* oversample: radio_get_oversample() (radio.cpp:211-219), 10 or 80; m17_mod_init (m17_modulate.cpp:65-76) */
int m17b_tx_create(m17b_ctx *ctx, int64_t nchan, int oversample, m17b_tx **out);
int m17b_tx_destroy(m17b_tx *tx);
int m17b_tx_reset(m17b_tx *tx, void *stream);
/* build_lich (m17_tx_routines.cpp:37-53): d_lsf [nchan][30] becomes each channel's m_lich; counters reset (:98-99) */
int m17b_tx_set_lsf(m17b_tx *tx, const uint8_t *d_lsf, void *stream);
/* m17_fmt_add_tx_preamble / m17_fmt_add_eot (m17_tx_routines.cpp:24-31,242-255): [192] dibits, host helper */
int m17b_fmt_preamble(uint8_t *h_dibits);
int m17b_fmt_eot(uint8_t *h_dibits);
/* m17_fmt_add_link_setup_frame (m17_tx_routines.cpp:92-117, encoded with adequately sized buffers: SURVEY D1):
   d_lsf [n][30] -> d_dibits [n][192] */
int m17b_fmt_link_setup_frame(m17b_ctx *ctx, const uint8_t *d_lsf, int64_t n, uint8_t *d_dibits, void *stream);
/* m17_fmt_add_stream_frame (m17_tx_routines.cpp:143-187) for F consecutive frames per channel:
   d_payload [nchan][F][16] -> d_dibits [nchan][F][192]; advances each channel's m_lich_count / m_fn */
int m17b_fmt_stream_frames(m17b_tx *tx, const uint8_t *d_payload, int64_t F, uint8_t *d_dibits, void *stream);
/* m17_fmt_add_packet (m17_tx_routines.cpp:201-222, safe buffers: D2): d_chunk [n][25], d_meta [n] (= eof<<7 | nf<<2) */
int m17b_fmt_packet_frames(m17b_ctx *ctx, const uint8_t *d_chunk, const uint8_t *d_meta, int64_t n, uint8_t *d_dibits, void *stream);
/* m17_send_packet_frames (m17_tx_routines.cpp:323-353) up to the dibits: n packets of d_len[i] bytes (stride apart) get their
   CRC-16 appended and are cut into 25-byte chunks (non-final frames: frame number; final frame: EOF + bytes used, 25 when the
   split is exact) -> d_dibits [n][max_frames][192], d_nframes [n]; slots past a packet's last frame hold blank-carrier
   symbols (4).  max_frames <= 32 (the frame counter is 5 bits); packets longer than 25*max_frames-2 bytes are truncated. */
int m17b_send_packet_frames(m17b_ctx *ctx, const uint8_t *d_packets, int64_t stride, const int32_t *d_len, int64_t n, int max_frames,
                            uint8_t *d_dibits, int32_t *d_nframes, void *stream);
/* m17_fmt_add_bert_frame (m17_tx_routines.cpp:226-238, intended behaviour: D5): F frames per channel */
int m17b_fmt_bert_frames(m17b_tx *tx, int64_t F, uint8_t *d_dibits, void *stream);
/* m17_mod_dibits / m17_mod_carrier (m17_modulate.cpp:42-61,79-92): d_syms [nchan][nsym] (0..3 dibit, 4 = carrier)
   -> d_iq int16 [nchan][nsym*os][2]; d_freq (optional) float [nchan][nsym*os] = the filtered deviation m_sum */
int m17b_mod_dibits(m17b_tx *tx, const uint8_t *d_syms, int64_t nsym, int16_t *d_iq, float *d_freq, void *stream);
/* instrumentation of the last m17b_mod_dibits call, SM cycles summed over the CTAs: h_out8 = {phase-scan warp: waiting for the
   FIR, walking its rows; chunks walked; CTAs; first worker warp: symbol fill + worker barrier, FIR, waiting for the scan,
   cos/sin + stores}.  Synchronises the device. */
int m17b_tx_debug_scan(m17b_tx *tx, uint64_t *h_out8);
/* exhaustive check that the modulator's fp32 form of mod_fsk's per-symbol phase wrap (m17_modulate.cpp:33-37: divide by 2 pi in
   double, modf, multiply back, every store rounding to float) equals the double formulation bit-for-bit on the float bit
   patterns [first, first+count) (Inf/NaN skipped); returns the number of mismatches (must be 0) and optionally the first
   dump_cap offenders as {input bits, fast result, reference result}.  All 2^32 patterns take well under a second on a B200. */
int m17b_selftest_tx_wrap(m17b_ctx *ctx, uint64_t first, uint64_t count, uint64_t *h_mismatches, uint32_t *h_dump, int dump_cap, void *stream);

/* ------------------------------------------------------------------ Pluto front-end decimator */
typedef struct m17b_dec m17b_dec;   /* per-batch decimator state: the 31-sample history m_rx_buff carries (radio.cpp:15,167) */
/* m17_dsp_build_lpf_filter (m17_dsp.cpp:347-360) and m17_dsp_float_to_short (m17_dsp.cpp:382-386), host side */
int m17b_build_lpf_filter(float *h_taps, float bw, int ntaps);
int m17b_float_to_short(const float *h_in, int16_t *h_out, int len);
/* build_pluto_rx_dec_filter (radio.cpp:44-51) + zeroed m_rx_buff */
int m17b_dec_create(m17b_ctx *ctx, int64_t nchan, m17b_dec **out);
int m17b_dec_destroy(m17b_dec *dec);
int m17b_dec_reset(m17b_dec *dec, void *stream);
int m17b_dec_get_taps(const m17b_dec *dec, int16_t *h_taps31);
/* radio_receive_samples, Pluto branch: rx_decimate_filter / sub_filter (radio.cpp:18-40,157-177) for nchan channels.
   d_in int16 [nchan][8*nout][2] at 384 kS/s -> d_out int16 [nchan][nout][2] at 48 kS/s (nout a multiple of 4; the
   reference produces 240 per 1920-sample chunk); the 31-sample filter history carries over from call to call */
int m17b_dec_run(m17b_dec *dec, const int16_t *d_in, int64_t nout, int16_t *d_out, void *stream);

/* ------------------------------------------------------------------ wideband channeliser (second half) */
/* The Pluto decimator (radio.cpp:18-40) generalised from one channel to a raster: one int16 IQ capture at 1.2 MS/s
   (25 x 48 kS/s) -> 96 channels spaced 12.5 kHz at 48 kS/s, in the layout m17b_dsp_rx reads.  Integer arithmetic throughout
   (polyphase int16 FIR with int32 accumulators, fixed-point 96-point DFT, >> 15), stated in plain C by the oracle
   (m17o_chan_run), which is pinned at M = 1, D = 8 against radio.cpp itself.  taps_per_branch = 4, 8, 12 or 16 (filter length
   96 x that; 12 gives > 60 dB at the adjacent channel). */
typedef struct m17b_chan m17b_chan;
#define M17B_CHAN_M 96
#define M17B_CHAN_D 25
int m17b_chan_create(m17b_ctx *ctx, int64_t ncaptures, int taps_per_branch, m17b_chan **out);
int m17b_chan_destroy(m17b_chan *ch);
int m17b_chan_reset(m17b_chan *ch, void *stream);
int m17b_chan_get_taps(const m17b_chan *ch, int16_t *h_taps, int *len);
/* d_in int16 [ncaptures][25*nout][2] -> d_out int16 [ncaptures*96][out_pitch][2] (out_pitch >= nout samples per row); row
   capture*96 + k is the channel k * 12.5 kHz above the capture's centre (k >= 48: below it).  Filter history and window phase
   carry over from call to call. */
int m17b_chan_run(m17b_chan *ch, const int16_t *d_in, int64_t nout, int16_t *d_out, int64_t out_pitch, void *stream);

/* ------------------------------------------------------------------ M17-over-UDP reflector frame */
#define M17B_NET_FRAME_BYTES 54
/* net_add_magic/_stream_id/_lich/_fn/_payload/_crc (m17_net.cpp:25-49), build_lich_to_net + m17_send_stream_frame_to_net
   (m17_tx_routines.cpp:54-70,298-306): n frames.  d_lsf: LSF bytes (>= 28 used) lsf_stride apart; have_dst != 0 replaces the
   destination call with dst (the gateway's "<reflector> <module>" address, m17_net.cpp:56-61).  d_out [n][54] */
int m17b_net_pack(m17b_ctx *ctx, const uint16_t *d_sid, const uint8_t *d_lsf, int64_t lsf_stride, int have_dst, uint64_t dst,
                  const uint16_t *d_fn, const uint8_t *d_payload, int64_t n, uint8_t *d_out, void *stream);
/* m17_parse_m17_data (m17_net.cpp:203-238) + build_lich_from_net (m17_tx_routines.cpp:71-86): d_in [n][54] -> d_ok [n]
   (magic and CRC over all 54 bytes), stream id, the 30-byte LSF the TX side uses (bytes 6..33 + fresh CRC), FN, payload [n][16] */
int m17b_net_parse(m17b_ctx *ctx, const uint8_t *d_in, int64_t n, uint8_t *d_ok, uint16_t *d_sid, uint8_t *d_lsf30, uint16_t *d_fn,
                   uint8_t *d_payload, void *stream);
/* m17_net_new_rx_data for the last m17b_dsp_rx / m17b_rx_baseband call (m17_rx_parse.cpp:148-154): every DELIVERED stream
   record of channel c becomes one datagram with the link-setup data that was valid when it was parsed, stream id d_sid[c].
   d_out [nchan][frame_cap][54] (datagrams of a channel packed from slot 0 in record order), d_count [nchan] */
int m17b_rx_net_frames(m17b_rx *rx, const uint16_t *d_sid, int have_dst, uint64_t dst, uint8_t *d_out, int32_t *d_count, void *stream);

/* ------------------------------------------------------------------ application layer behind the frame decode */
/* Packet reassembly as parse_packet (m17_rx_parse.cpp:34-51) is meant to work (the reference's own indexing cannot validate a
   multi-frame packet; that behaviour is reproduced in the per-channel state for parity).  Walks the records of the
   last m17b_dsp_rx / m17b_rx_baseband / m17b_rx_symbols call in order, per channel: 25 bytes of every non-final packet frame,
   `count` bytes of the EOF frame, then the CRC-16 appended by m17_send_packet_frames (m17_tx_routines.cpp:323-353) is checked
   over the whole packet.  A packet may span calls (the partial packet is carried; m17b_rx_reset clears it).
   d_bytes [nchan][bytes_cap]: payloads (CRC stripped) back to back; d_pkt int32 [nchan][max_pkts][3] = {offset, length, crc_ok};
   d_npkt [nchan].  Packets that do not fit bytes_cap / max_pkts are dropped. */
int m17b_rx_reassemble_packets(m17b_rx *rx, uint8_t *d_bytes, int64_t bytes_cap, int32_t *d_pkt, int max_pkts, int32_t *d_npkt, void *stream);
/* gps_decode (gps.cpp:8-27) on the META field of n link-setup frames (d_lsf: 30-byte LSFs `stride` bytes apart; the reference
   reads 15 bytes from META, i.e. one byte into the CRC -- reproduced) */
typedef struct { double lat, lon; int32_t alt, course, speed, object; } m17b_gps_rec;
int m17b_gps_decode(m17b_ctx *ctx, const uint8_t *d_lsf, int64_t stride, int64_t n, m17b_gps_rec *d_out, void *stream);

/* ------------------------------------------------------------------ equaliser (m17_equalize.cpp) */
int m17b_eq_create(m17b_ctx *ctx, int64_t nchan, m17b_eq **out);   /* eq_open  :217-224 */
int m17b_eq_destroy(m17b_eq *eq);
int m17b_eq_reset(m17b_eq *eq, void *stream);                       /* eq_reset :137-141 */
int m17b_eq_restart(m17b_eq *eq, void *stream);                     /* eq_restart :142-145 (U/D factors only) */
/* eq_train_known / eq_train_unknown (:163-213) for nsym symbols per channel: d_in [nchan][nsym][2],
   d_train [nchan][nsym] or NULL (decision-directed) -> d_out [nchan][nsym] */
int m17b_eq_train(m17b_eq *eq, const float *d_in, const float *d_train, int64_t nsym, float *d_out, void *stream);

/* ------------------------------------------------------------------ synthetic channel (bench input only) */
/* in-place AWGN on int16 IQ + carrier rotation; per-channel sigma (LSB units) and f0 (cycles/sample);
   counter-based RNG, deterministic in (seed, channel, sample); never emits an exact (0,0) sample */
int m17b_synth_channel(m17b_ctx *ctx, int16_t *d_iq, int64_t nchan, int64_t nsamp, const float *d_sigma, const float *d_f0,
                       uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* M17B200_H */
