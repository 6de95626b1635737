/*
 * m17_oracle.h -- CPU restatement ("port") of the m17gismo baseband hot path, in plain C with
 * explicit per-channel state.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  It is the checker, never
 * the product: the product (m17_sdr_b200/) has no CPU path at all.
 *
 * Parity status: PINNED.  Every function is checked (tests/test_oracle_vs_ref.py, run wherever
 * oracle/_ref/libm17ref.so exists) against the reference's own objects built unmodified from
 * /root/reference/m17gismo by oracle/ref/Makefile, and (tests/test_oracle_golden.py, runs
 * anywhere) against golden vectors in tests/golden/ generated from those objects by
 * tests/golden/make_golden.py.  The reference itself ships no tests or vectors.
 *
 * Each function cites the reference file:line it follows (paths relative to m17gismo/).
 */
#ifndef M17_ORACLE_H
#define M17_ORACLE_H
#include <stdint.h>
#include "m17_records.h"

#ifdef __cplusplus
extern "C" {
#endif

#define M17O_NF 40            /* polyphase branches   m17_rx_sync.cpp:3 */
#define M17O_FN 31            /* taps per branch      m17_rx_sync.cpp:4 */
#define M17O_BLOCK 1920       /* N_SAMPLES            m17defines.h:17   */
#define M17O_FRAME_SYMS 192   /* FRAME_SYM_LENGTH     m17defines.h:66   */

void m17o_init(void);                                      /* main.cpp:108-126 (table builds) */

/* ---- filter design */
void m17o_rrc_design(float *taps, float rolloff, int ntaps, int sps);        /* m17_dsp.cpp:295-315 */
void m17o_set_gain(float *taps, float gain, int stride, int ntaps);          /* m17_dsp.cpp:420-429 */
void m17o_get_sync_taps(float *mf, float *md);                               /* [40][31] each, m17_rx_sync.cpp:101-123 */

/* ---- FEC / bit-domain primitives */
uint16_t m17o_crc(const uint8_t *in, int len);                               /* m17_crc.cpp:26-35 */
uint32_t m17o_golay_encode(uint16_t data);                                   /* m17_golay.cpp:94-102 */
int      m17o_golay_decode(uint32_t word, uint16_t *odata);                  /* m17_golay.cpp:103-116 */
void     m17o_golay_errtab(uint16_t *out4096);
int  m17o_conv_encode_8(const uint8_t *in, uint8_t *out, int len);           /* m17_conv.cpp:53-71 */
int  m17o_conv_encode_1(const uint8_t *in, uint8_t *out, int len);           /* m17_conv.cpp:33-49 */
int  m17o_viterbi(const float *in, uint8_t *out, int len);                   /* m17_conv.cpp:73-113,148-168 */
int  m17o_punc(int p, const uint8_t *in, uint8_t *out, int len);             /* m17_puncture.cpp:12-41 */
int  m17o_depunc(int p, const float *in, float *out, int len);               /* m17_puncture.cpp:47-79 */
void m17o_interleave(const uint8_t *in, uint8_t *out, int len);              /* m17_interleave.cpp:3-7 */
void m17o_deinterleave(const float *in, float *out, int len);                /* m17_interleave.cpp:8-12 */
void m17o_derand_bytes(uint8_t *io, int len);                                /* m17_correlate.cpp:11-15 */
void m17o_derand_bits(const uint8_t *in, uint8_t *out, int len);             /* m17_correlate.cpp:16-20 */
void m17o_derand_soft(const float *in, float *out, int len);                 /* m17_correlate.cpp:27-31 */
void m17o_demap_frame(const float *sym192, float *soft368);                  /* m17_dsp.cpp:35-42,82-95 */
void m17o_gps_decode(const uint8_t *meta15, double *latlon2, int32_t *alt_course_speed_object);   /* gps.cpp:8-27 */
void m17o_demap_symbol(float in, float mag, float *out2);                    /* m17_dsp.cpp:35-42 */
int  m17o_decimating_filter(const float *in, float *out, const float *coffs, int stride, int flen, int len);   /* m17_dsp.cpp:438-449 */
uint32_t m17o_hard24(const float *in);                                       /* m17_bit_utils.cpp:180-187 */
void m17o_sync_check(const float *v8, int *type, int *votes, float *var);    /* m17_rx_frame.cpp:22-81 */
void m17o_prbs9_seq(uint8_t *out511);                                        /* m17_prbs9.cpp:16-26 */
uint64_t m17o_encode_call(const char *call9);                                /* m17_bit_utils.cpp:191-209 */
void     m17o_decode_call(uint64_t w, char *out10);                          /* m17_bit_utils.cpp:210-226 */

/* ---- equaliser (m17_equalize.cpp; dead code in the reference, standalone parity) */
typedef struct { float c[5], g[5], u[5][5], d[5], E, q, y, fbr, samples[5]; } m17o_eq;
void  m17o_eq_open(m17o_eq *e);                                              /* :217-224 */
void  m17o_eq_reset(m17o_eq *e);                                             /* :137-141 */
float m17o_eq_train_known(m17o_eq *e, const float *in2, float train);        /* :163-180 */
float m17o_eq_train_unknown(m17o_eq *e, const float *in2);                   /* :185-213 */

/* ---- TX */
typedef struct {
    int     os;                /* radio_get_oversample(): 10 Lime / 80 Pluto  radio.cpp:211-219 */
    float  *taps;              /* 31*os RRC taps, gain 10   m17_modulate.cpp:72-73 */
    float   s[31];             /* m_tx_s    m17_modulate.cpp:8  */
    float   acc;               /* m_acc     m17_modulate.cpp:14 */
    uint8_t lich[30];          /* m_lich    m17_tx_routines.cpp:13 */
    int     lich_count;        /* m_lich_count :14 */
    int     fn;                /* m_fn :15 */
    int     prbs_idx;          /* m_tx_idx  m17_prbs9.cpp:7 */
} m17o_tx;
m17o_tx *m17o_tx_new(int os);
void     m17o_tx_free(m17o_tx *t);
int  m17o_build_lsf(uint64_t dst, uint64_t src, uint16_t typeword, const uint8_t *meta14, uint8_t *out30); /* m17_tx_routines.cpp:37-53 */
int  m17o_fmt_preamble(uint8_t *dibits);                                     /* m17_tx_routines.cpp:24-31 */
int  m17o_fmt_eot(uint8_t *dibits);                                          /* :242-255 */
int  m17o_fmt_lsf(const uint8_t *lsf30, uint8_t *dibits);                    /* :92-117 with safe buffers (D1) */
int  m17o_fmt_stream(m17o_tx *t, const uint8_t *payload16, uint8_t *dibits); /* :143-187 */
int  m17o_fmt_packet(const uint8_t *chunk, int len, int eof, int nf, uint8_t *dibits); /* :201-222 safe (D2) */
int  m17o_fmt_bert(m17o_tx *t, uint8_t *dibits);                             /* :226-238 intended (D5) */
/* modulate nsym symbols (value 0..3 dibit, 4 = blank carrier); writes nsym*os IQ pairs (and the
   frequency samples m_sum when freq != NULL); returns samples written.  m17_modulate.cpp:22-61,79-92 */
long m17o_mod(m17o_tx *t, const uint8_t *syms, long nsym, int16_t *iq, float *freq);

/* ---- M17-over-UDP reflector frame, 54 bytes */
void m17o_net_pack(uint16_t sid, const uint8_t *lsf28, int have_dst, uint64_t dst, uint16_t fn, const uint8_t *pld16, uint8_t *out54);   /* m17_net.cpp:25-74 */
int  m17o_net_parse(const uint8_t *b54, uint16_t *sid, uint8_t *lsf30, uint16_t *fn, uint8_t *pld16);   /* m17_net.cpp:203-238, m17_tx_routines.cpp:71-86 */

/* ---- Pluto front-end decimator: int16 31-tap symmetric low-pass, decimate by 8, 384 kS/s -> 48 kS/s */
typedef struct { int16_t taps[31]; int16_t hist[31][2]; } m17o_dec;
void m17o_lpf_design(float *taps, float bw, int ntaps);                      /* m17_dsp.cpp:347-360 */
void m17o_dec_open(m17o_dec *d);                                             /* build_pluto_rx_dec_filter, radio.cpp:44-51 */
void m17o_dec_run(m17o_dec *d, const int16_t *in, long nout, int16_t *out);  /* radio.cpp:18-40,157-177: in [8*nout][2] -> out [nout][2] */

/* ---- RX */
typedef struct m17o_rx m17o_rx;
m17o_rx *m17o_rx_new(void);
void     m17o_rx_free(m17o_rx *r);
void     m17o_rx_set_afc(m17o_rx *r, int on);
void     m17o_rx_set_bert(m17o_rx *r, int on);                               /* BERT receive extension (decode_bert_frame is empty upstream) */
void     m17o_rx_set_eq(m17o_rx *r, int on);                                 /* equaliser option (eq_open + eq_train_unknown on T/2 pairs; no call site upstream; pinned by oracle/ref/eq_shim.cpp) */
void     m17o_rx_get_bert(const m17o_rx *r, uint32_t *out8);                 /* state, idx, bad, good, eq, dif (m17_prbs9.cpp:7-12), bits, errs */
void     m17o_prbs9_rx_check(m17o_rx *r, uint8_t bit);                       /* m17_prbs9.cpp:40-64 */
void     m17o_set_bert_out(uint32_t *p);
void     m17o_rx_trace(m17o_rx *r, float *disc, int32_t *nsym, float *syms, long symcap,
                       m17_frame_rec *frames, long fcap, float *soft, m17_event_rec *events, long ecap);
void     m17o_rx_counts(const m17o_rx *r, int64_t *out4);     /* blocks, syms, frames, events */
void     m17o_dsp_rx(m17o_rx *r, const int16_t *iq, int nsamp);              /* m17_dsp.cpp:461-476 */
void     m17o_rx_baseband(m17o_rx *r, const float *disc, int n);             /* m17_test.cpp:49-51 seam */
int      m17o_sync_samples(m17o_rx *r, const float *in, float *out, int len);/* m17_rx_sync.cpp:77-99 */
void     m17o_frontend(m17o_rx *r, const int16_t *iq, int nsamp, float *disc, int *ndisc, float *mean);

/* batch driver, same contract as ref_rx_run() in oracle/ref/ref_shim.cpp but threads, not forks */
int m17o_rx_run(const void *in, int seam, long C, long T, int nthreads,
                float *disc, int32_t *nsym, float *syms, long symcap,
                m17_frame_rec *frames, long fcap, float *soft,
                m17_event_rec *events, long ecap, int64_t *counts);
/* timing driver: returns seconds for the slowest worker */
double m17o_rx_time(const int16_t *iq, long C, long T, int nthreads);

#ifdef __cplusplus
}
#endif
/* wideband channeliser: the Pluto decimator (radio.cpp:18-40) generalised to M channels out of one capture (see m17_oracle.c) */
typedef struct { int M, D, L, lg2, has3; int16_t *h; int32_t *tw2, *rot; int32_t s3; int16_t *hist; long long n_done; } m17o_chan;
m17o_chan *m17o_chan_open(int M, int D, int L, const int16_t *taps);
void m17o_chan_free(m17o_chan *c);
void m17o_chan_run(m17o_chan *c, const int16_t *in, long nout, int16_t *out, long out_pitch);   /* in [D*nout][2] -> out [M][out_pitch][2] */

#endif
