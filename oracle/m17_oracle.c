/*
 * m17_oracle.c -- CPU restatement of the m17gismo baseband hot path.  See m17_oracle.h.
 * TEST INFRASTRUCTURE ONLY (the checker; never measured as the product, never shipped).
 *
 * Arithmetic notes (the reference is C++: fabs/sqrt/cos/sin of a float resolve to the float
 * overloads; products with double literals are done in double and rounded once on store):
 * every expression below keeps the reference's operand types and evaluation order, and the
 * file must be compiled WITHOUT -ffast-math and without FMA contraction (-ffp-contract=off),
 * like the reference's generic x86-64 -O3 build (makefile:6).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "m17_oracle.h"

/* ================================================================== tables */
static uint16_t t_crc[256];                 /* m17_crc.cpp:8-24  */
static uint16_t t_genc[4096];               /* m17_golay.cpp:31-40 */
static uint16_t t_gerr[4096];               /* m17_golay.cpp:49-72 */
static uint8_t  t_rand[368];                /* m17_correlate.cpp:35-42 */
static uint8_t  t_prbs[511];                /* m17_prbs9.cpp:16-26 */
static uint8_t  t_conv[32][2];              /* m17_conv.cpp:24-29 */
static float    t_mf[M17O_NF][M17O_FN];     /* m17_rx_sync.cpp:13 */
static float    t_md[M17O_NF][M17O_FN];     /* m17_rx_sync.cpp:14 */
static int      t_ready = 0;

/* M17 randomiser sequence, 46 bytes (protocol constant; m17_correlate.cpp:3-7) */
static const uint8_t k_rand_bytes[46] = {
    0xD6,0xB5,0xE2,0x30,0x82,0xFF,0x84,0x62,0xBA,0x4E,0x96,0x90,0xD8,0x98,0xDD,0x5D,0x0C,0xC8,0x52,0x43,0x91,0x1D,0xF8,
    0x6E,0x68,0x2F,0x35,0xDA,0x14,0xEA,0xCD,0x76,0x19,0x8D,0xD5,0x80,0xD1,0x33,0x87,0x13,0x57,0x18,0x2D,0x29,0x78,0xC3 };
/* Golay(24,12) parity generator rows (protocol constant; m17_golay.cpp:11) */
static const uint16_t k_golay_rows[12] = { 0xC75,0x63B,0xF68,0x7B4,0x3DA,0xD99,0x6CD,0x367,0xDC6,0xA97,0x93E,0x8EB };
/* sync-word templates as +-1 symbol signs (m17_rx_frame.cpp:5-12): preamble, LSF 0x55F7, stream 0xFF5D,
   packet 0x75FF, BERT 0xDF55, EOT 0x555D.  Stored as bitmasks: bit i set = template[i] is -1. */
static const uint8_t k_sync_neg[6] = { 0xAA, 0xB0, 0x4F, 0xF2, 0x0D, 0x40 };

static float sync_tpl(int t, int i) { return (k_sync_neg[t] >> i) & 1 ? -1.0f : 1.0f; }

/* puncture keep-pattern: P1 period 61 (zeros where i%4==2), P2 period 12 (last dropped),
   P3 period 8 (last dropped).  m17_puncture.cpp:4-10 */
static int punc_keep(int p, int i) {
    if (p == 1) { int k = i % 61; return (k % 4) != 2; }
    if (p == 2) return (i % 12) != 11;
    return (i % 8) != 7;
}
static int qpp(int i) { return (i * 45 + 92 * i * i) % 368; }   /* m17_interleave.cpp:5,10 */

static int popcount24(uint32_t w) { int n = 0; for (int i = 0; i < 24; i++) { n += w & 1; w >>= 1; } return n; }

void m17o_rrc_design(float *taps, float rolloff, int ntaps, int sps) {
    /* m17_dsp.cpp:295-315; note the integer division centring the taps */
    double B = (rolloff + 0.0001);
    double t = -(ntaps - 1) / 2;
    double Ts = sps;
    for (int i = 0; i < ntaps; i++) {
        double a = 2.0 * B / (M_PI * sqrt(Ts));
        double b = cos((1.0 + B) * M_PI * t / Ts);
        double c;
        if (t == 0) c = (1.0 - B) * M_PI / (4 * B);
        else        c = sin((1.0 - B) * M_PI * t / Ts) / (4.0 * B * t / Ts);
        double d = (1.0 - (4.0 * B * t / Ts) * (4.0 * B * t / Ts));
        taps[i] = (float)(a * (b + c) / d);
        t = t + 1.0;
    }
}
void m17o_set_gain(float *taps, float gain, int stride, int ntaps) {
    /* m17_dsp.cpp:420-429: float running sum, float divide */
    float sum = 0;
    for (int i = 0; i < ntaps; i++) sum += taps[i * stride];
    gain = gain / sum;
    for (int i = 0; i < ntaps; i++) taps[i * stride] = taps[i * stride] * gain;
}

void m17o_init(void) {
    if (t_ready) return;
    /* CRC-16 poly 0x5935, MSB first (m17_crc.cpp:4-24) */
    for (int i = 0; i < 256; i++) {
        uint16_t x = (uint16_t)(i << 8);
        for (int n = 0; n < 8; n++) x = (x & 0x8000) ? (uint16_t)((x << 1) ^ 0x5935) : (uint16_t)(x << 1);
        t_crc[i] = x;
    }
    /* Golay parity table (m17_golay.cpp:31-40) */
    for (int i = 0; i < 4096; i++) {
        uint16_t p = 0;
        for (int n = 0; n < 12; n++) if (i & (0x800 >> n)) p ^= k_golay_rows[n];
        t_genc[i] = p;
    }
    /* Golay syndrome table (m17_golay.cpp:49-72): ascending scan of all 2^24 words, weight<=4,
       last writer wins; the 0x400 pre-fill of entries 0..0xFFE is reproduced */
    for (int i = 0; i < 0xFFF; i++) t_gerr[i] = 0x400;
    t_gerr[0xFFF] = 0;
    for (uint32_t w = 0; w < 0x1000000; w++) {
        int bits = popcount24(w);
        if (bits < 5) {
            uint16_t data = (uint16_t)(w >> 12), parity = (uint16_t)(w & 0xFFF);
            t_gerr[parity ^ t_genc[data]] = (uint16_t)((bits << 12) | data);
        }
    }
    /* randomiser bits, MSB first (m17_correlate.cpp:35-42) */
    for (int i = 0; i < 368; i++) t_rand[i] = (k_rand_bytes[i >> 3] >> (7 - (i & 7))) & 1;
    /* PRBS9 x^9+x^5+1 seed 1 (m17_prbs9.cpp:16-26) */
    { uint16_t sr = 1; int n = 0; do { uint8_t b = ((sr >> 8) ^ (sr >> 4)) & 1; sr = ((sr << 1) | b) & 0x1FF; t_prbs[n++] = b; } while (sr != 1 && n < 511); }
    /* conv encoder output table G1=1+D^3+D^4, G2=1+D+D^2+D^4 on the 5-bit register (m17_conv.cpp:24-29) */
    for (int i = 0; i < 32; i++) {
        t_conv[i][0] = ((i >> 4) ^ (i >> 1) ^ i) & 1;
        t_conv[i][1] = ((i >> 4) ^ (i >> 3) ^ (i >> 2) ^ i) & 1;
    }
    /* matched / derivative polyphase banks (m17_rx_sync.cpp:101-123) */
    {
        enum { N = M17O_NF * M17O_FN };
        static float mf[N], md[N];
        m17o_rrc_design(mf, 0.5f, N, M17O_NF * 2);
        for (int i = 0; i < N; i++) md[i] = mf[(i + 1) % N] - mf[(i + N - 1) % N];
        for (int i = 0; i < M17O_NF; i++)
            for (int j = 0; j < M17O_FN; j++) { t_mf[i][j] = mf[i + j * M17O_NF]; t_md[i][j] = md[i + j * M17O_NF]; }
        for (int i = 0; i < M17O_NF; i++) m17o_set_gain(t_mf[i], 1.0f, 1, M17O_FN);
    }
    t_ready = 1;
}
void m17o_get_sync_taps(float *mf, float *md) { m17o_init(); memcpy(mf, t_mf, sizeof(t_mf)); memcpy(md, t_md, sizeof(t_md)); }
void m17o_golay_errtab(uint16_t *out) { m17o_init(); memcpy(out, t_gerr, sizeof(t_gerr)); }
void m17o_prbs9_seq(uint8_t *out) { m17o_init(); memcpy(out, t_prbs, 511); }

/* ================================================================== bit-domain primitives */
uint16_t m17o_crc(const uint8_t *in, int len) {
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < len; i++) crc = (uint16_t)((crc << 8) ^ t_crc[((crc >> 8) ^ in[i]) & 0xFF]);
    return crc;
}
uint32_t m17o_golay_encode(uint16_t data) { return ((uint32_t)data << 12) | t_genc[data & 0xFFF]; }
int m17o_golay_decode(uint32_t word, uint16_t *odata) {
    uint16_t data = (word >> 12) & 0xFFF, parity = word & 0xFFF;
    uint16_t e = t_gerr[parity ^ t_genc[data]];
    *odata = data ^ (e & 0xFFF);
    return (e & 0xF000) >> 12;
}
static int conv_step(uint8_t *sr, int bit, uint8_t *out, int idx) {
    if (bit) *sr |= 0x10;
    out[idx++] = t_conv[*sr][0];
    out[idx++] = t_conv[*sr][1];
    *sr >>= 1;
    return idx;
}
int m17o_conv_encode_8(const uint8_t *in, uint8_t *out, int len) {
    int idx = 0; uint8_t sr = 0;
    for (int i = 0; i < len; i++) for (int n = 7; n >= 0; n--) idx = conv_step(&sr, (in[i] >> n) & 1, out, idx);
    for (int i = 0; i < 4; i++) idx = conv_step(&sr, 0, out, idx);
    return idx;
}
int m17o_conv_encode_1(const uint8_t *in, uint8_t *out, int len) {
    int idx = 0; uint8_t sr = 0;
    for (int i = 0; i < len; i++) idx = conv_step(&sr, in[i] != 0, out, idx);
    for (int i = 0; i < 4; i++) idx = conv_step(&sr, 0, out, idx);
    return idx;
}
/*
 * Soft Viterbi, K=5 r=1/2 (m17_conv.cpp:73-113,148-168).  Max-correlation metric, no normalisation,
 * acm[0]=1.0 start bias.  New state v has predecessors w=(2v)&15 and y=w+1; the branch symbol of
 * predecessor p entering v is the encoder output for register (v>>3)<<4 | p (the butterfly list at
 * :93-108 is exactly this).  Strict '>' keeps w, ties go to the ODD predecessor y.  Traceback from
 * state 0; out[i] is the MSB of the state reached at step i, i.e. input bit i-1 (:160-166).
 */
int m17o_viterbi(const float *in, uint8_t *out, int len) {
    float acm[16]; int steps = 0;
    uint8_t (*path)[16] = (uint8_t (*)[16])malloc((size_t)(len / 2 + 1) * 16);
    for (int i = 0; i < 16; i++) acm[i] = 0.0f;
    acm[0] = 1.0f;
    for (int i = 0; i < len; i += 2) {
        float m1 = in[i], m2 = in[i + 1], metric[4], tm[16];
        float m1X = m1, m0X = -m1, mX1 = m2, mX0 = -m2;
        metric[0] = (m0X + mX0); metric[1] = (m0X + mX1); metric[2] = (m1X + mX0); metric[3] = (m1X + mX1);
        for (int v = 0; v < 16; v++) {
            int w = (2 * v) & 15, y = w + 1, hi = (v >> 3) << 4;
            int x = (t_conv[hi | w][0] << 1) | t_conv[hi | w][1];
            int z = (t_conv[hi | y][0] << 1) | t_conv[hi | y][1];
            float a = acm[w] + metric[x], b = acm[y] + metric[z];
            if (a > b) { tm[v] = a; path[steps][v] = (uint8_t)w; } else { tm[v] = b; path[steps][v] = (uint8_t)y; }
        }
        for (int v = 0; v < 16; v++) acm[v] = tm[v];
        steps++;
    }
    uint8_t state = 0;
    for (int i = steps - 1; i >= 0; i--) { state = path[i][state]; out[i] = (state & 0x08) ? 1 : 0; }
    free(path);
    return steps;
}
int m17o_punc(int p, const uint8_t *in, uint8_t *out, int len) {
    int idx = 0;
    for (int i = 0; i < len; i++) if (punc_keep(p, i)) out[idx++] = in[i];   /* in-place safe: idx <= i */
    return idx;
}
int m17o_depunc(int p, const float *in, float *out, int len) {
    int idx = 0;
    for (int i = 0; i < len; i++) out[i] = punc_keep(p, i) ? in[idx++] : 0.0f;
    return len;
}
void m17o_interleave(const uint8_t *in, uint8_t *out, int len) { for (int i = 0; i < len; i++) out[qpp(i)] = in[i]; }
void m17o_deinterleave(const float *in, float *out, int len) { for (int i = 0; i < len; i++) out[qpp(i)] = in[i]; }
void m17o_derand_bytes(uint8_t *io, int len) { for (int i = 0; i < len; i++) io[i] ^= k_rand_bytes[i % 46]; }
void m17o_derand_bits(const uint8_t *in, uint8_t *out, int len) { for (int i = 0; i < len; i++) out[i] = (in[i] ^ t_rand[i % 368]) & 1; }
void m17o_derand_soft(const float *in, float *out, int len) { for (int i = 0; i < len; i++) out[i] = t_rand[i % 368] ? -in[i] : in[i]; }

/* gps_decode (gps.cpp:8-27): b points at the 14 META bytes of an LSF; 15 bytes are read (the last one is the first CRC byte).
   out = {lat, lon} doubles, then alt, course, speed, object as int32 */
void m17o_gps_decode(const uint8_t *b, double *latlon, int32_t *out4) {
    latlon[0] = (double)(int8_t)b[0] + (double)(uint16_t)((b[1] << 8) | b[2]) / 65536.0;
    latlon[1] = (double)(int16_t)((b[3] << 8) | b[4]) + (double)(uint16_t)((b[5] << 8) | b[6]) / 65536.0;
    out4[0] = (int16_t)(((b[7] << 8) | b[8]) - 1500);
    uint64_t w = 0;
    for (int k = 0; k < 6; k++) w = (w << 8) | b[9 + k];
    out4[1] = (uint16_t)(w >> 38);
    out4[2] = (int32_t)((w >> 28) & 0x3FF);
    out4[3] = (int32_t)(w & 0xFFFFF);
}
/* m17_dsp_demap_symbol (m17_dsp.cpp:35-42): one symbol with the caller's normaliser */
void m17o_demap_symbol(float in, float mag, float *out) {
    float m = in * mag;
    out[0] = -m;
    out[1] = (float)(fabs(m) - 0.6666);
}
/* m17_dsp_decimating_filter (m17_dsp.cpp:438-449): sum starts at 0 and every product is added in tap order */
int m17o_decimating_filter(const float *in, float *out, const float *coffs, int stride, int flen, int len) {
    int idx = 0;
    for (int i = 0; i < len; i += stride) {
        float sum = 0;
        for (int j = 0; j < flen; j++) sum += in[i + j] * coffs[j];
        out[idx++] = sum;
    }
    return idx;
}
void m17o_demap_frame(const float *s, float *out) {
    /* m17_dsp.cpp:82-95 and :35-42.  cor is a double quotient rounded to float; the LSB soft value is a
       double subtraction (0.6666 is a double literal) rounded to float.  Positive means bit 1. */
    float sum = 0;
    for (int i = 0; i < 8; i++) sum += fabsf(s[i]);
    float cor = (float)(8.0 / sum);
    for (int i = 8, o = 0; i < M17O_FRAME_SYMS; i++, o += 2) {
        float m = s[i] * cor;
        out[o] = -m;
        out[o + 1] = (float)(fabsf(m) - 0.6666);
    }
}
uint32_t m17o_hard24(const float *in) {
    uint32_t w = 0;
    for (int i = 0; i < 24; i++) { w <<= 1; w |= in[i] >= 0 ? 1 : 0; }
    return w;
}
static float sync_variance(const float *in) {
    /* find_variance, m17_rx_frame.cpp:22-43 (note the else: a new max never updates min) */
    float mmin = fabsf(in[0]), mmax = mmin, v;
    for (int i = 1; i < 8; i++) { v = fabsf(in[i]); if (v > mmax) mmax = v; else if (v < mmin) mmin = v; }
    v = (mmax - mmin) / mmax;
    if (v != v) v = 1.0f;
    return v;
}
void m17o_sync_check(const float *v, int *type, int *votes, float *var) {
    /* m17_rx_frame.cpp:47-81: sequential 8-term sums, argmax with strict '>' from (0, type 0) */
    float sums[6];
    for (int t = 0; t < 6; t++) sums[t] = v[0] * sync_tpl(t, 0);
    for (int i = 1; i < 8; i++) for (int t = 0; t < 6; t++) sums[t] += v[i] * sync_tpl(t, i);
    *var = sync_variance(v);
    float mmax = 0; int nmax = 0;
    for (int t = 0; t < 6; t++) if (sums[t] > mmax) { mmax = sums[t]; nmax = t; }
    *type = nmax;
    int n = 0;
    for (int i = 0; i < 8; i++) if (v[i] * sync_tpl(nmax, i) < 0) n++;
    *votes = n;
}
static int sync_accept(int type, int votes, float var, int locked) {
    /* m17_rx_frame.cpp:82-103: the variance is compared against double literals */
    if (votes > (locked ? 1 : 0)) return 0;
    if (type >= 1 && type <= 4) return (double)var < (locked ? 0.5 : 0.3);
    return 0;
}
uint64_t m17o_encode_call(const char *call) {
    uint64_t w = 0;
    for (int i = 8; i >= 0; i--) {
        char c = call[i]; w *= 40;
        if (c >= 'A' && c <= 'Z') w += c - 'A' + 1;
        else if (c >= '0' && c <= '9') w += c - '0' + 27;
        else if (c == '-') w += 37; else if (c == '/') w += 38; else if (c == '.') w += 39;
    }
    return w;
}
void m17o_decode_call(uint64_t w, char *out) {
    if (w == 0xFFFFFFFFFFFFull) { strcpy(out, "BROADCAST"); return; }
    for (int i = 0; i < 9; i++) {
        int c = (int)(w % 40); w /= 40;
        out[i] = c == 0 ? ' ' : c <= 26 ? (char)('A' + c - 1) : c <= 36 ? (char)('0' + c - 27) : c == 37 ? '-' : c == 38 ? '/' : '.';
    }
    out[9] = 0;
}

/* ================================================================== equaliser (m17_equalize.cpp) */
#define KN 5
static void eq_reset_ud(m17o_eq *e) { for (int j = 0; j < KN; j++) { for (int i = 0; i < j; i++) e->u[i][j] = 0.0f; e->d[j] = 0.1f; } }
void m17o_eq_reset(m17o_eq *e) { eq_reset_ud(e); for (int i = 0; i < KN; i++) e->c[i] = 0.0f; }
void m17o_eq_open(m17o_eq *e) { memset(e, 0, sizeof(*e)); e->q = 0.08f; e->E = 0.01f; m17o_eq_reset(e); }
static void eq_gain(m17o_eq *e, const float *x) {
    /* m17_equalize.cpp:40-100, eqs 6.2-6.22 of the UD-factorised Kalman update */
    float f[KN], h[KN], a[KN];
    f[0] = x[0];
    for (int j = 1; j < KN; j++) { f[j] = e->u[0][j] * x[0] + x[j]; for (int i = 1; i < j; i++) f[j] += e->u[i][j] * x[i]; }
    for (int j = 0; j < KN; j++) e->g[j] = e->d[j] * f[j];
    a[0] = e->E + e->g[0] * f[0];
    for (int j = 1; j < KN; j++) a[j] = a[j - 1] + e->g[j] * f[j];
    float hq = 1 + e->q, ht = a[KN - 1] * e->q;
    e->y = (float)1.0 / (a[0] + ht);
    e->d[0] = e->d[0] * hq * (e->E + ht) * e->y;
    for (int j = 1; j < KN; j++) {
        float B = a[j - 1] + ht;
        h[j] = -f[j] * e->y;
        e->y = (float)1.0 / (a[j] + ht);
        e->d[j] = e->d[j] * hq * B * e->y;
        for (int i = 0; i < j; i++) { float B0 = e->u[i][j]; e->u[i][j] = B0 + h[j] * e->g[i]; e->g[i] += e->g[j] * B0; }
    }
}
static float eq_step(m17o_eq *e, const float *in2, int known, float train) {
    for (int i = 0; i < KN - 2; i++) e->samples[i] = e->samples[i + 2];      /* :152-159 */
    e->samples[KN - 2] = in2[0]; e->samples[KN - 1] = in2[1];
    float sym = e->samples[0] * e->c[0];                                     /* :122-135 */
    for (int i = 1; i < KN; i++) sym += e->samples[i] * e->c[i];
    if (!known) {                                                            /* :195-205 */
        if (sym > 0) train = (sym >= 0.66) ? 1.0f : 0.333f; else train = (sym <= -0.66) ? -1.0f : -0.333f;
    }
    float err = train - sym;
    eq_gain(e, e->samples);                                                  /* :105-121 */
    err *= e->y;
    for (int i = 0; i < KN; i++) e->c[i] += err * e->g[i];
    e->fbr = train;
    return sym;
}
float m17o_eq_train_known(m17o_eq *e, const float *in2, float train) { return eq_step(e, in2, 1, train); }
float m17o_eq_train_unknown(m17o_eq *e, const float *in2) { return eq_step(e, in2, 0, 0.0f); }

/* ================================================================== TX */
m17o_tx *m17o_tx_new(int os) {
    m17o_init();
    m17o_tx *t = (m17o_tx *)calloc(1, sizeof(*t));
    t->os = os;
    t->taps = (float *)malloc(sizeof(float) * 31 * os);
    m17o_rrc_design(t->taps, 0.5f, 31 * os, os);          /* m17_modulate.cpp:72 */
    m17o_set_gain(t->taps, 10, 1, 31 * os);               /* :73 */
    return t;
}
void m17o_tx_free(m17o_tx *t) { if (t) { free(t->taps); free(t); } }

int m17o_build_lsf(uint64_t dst, uint64_t src, uint16_t tw, const uint8_t *meta, uint8_t *o) {
    for (int i = 0; i < 6; i++) { o[i] = (uint8_t)(dst >> (40 - 8 * i)); o[6 + i] = (uint8_t)(src >> (40 - 8 * i)); }
    o[12] = (uint8_t)(tw >> 8); o[13] = (uint8_t)tw;
    memcpy(o + 14, meta, 14);
    uint16_t crc = m17o_crc(o, 28);
    o[28] = (uint8_t)(crc >> 8); o[29] = (uint8_t)crc;
    return 30;
}
static int put_sync(uint16_t w, uint8_t *d) { for (int i = 0; i < 8; i++) d[i] = (w >> (14 - 2 * i)) & 3; return 8; }
static int bits_to_dibits(const uint8_t *b, uint8_t *d, int nbits) { int n = 0; for (int i = 0; i < nbits; i += 2) d[n++] = (uint8_t)((b[i] << 1) | b[i + 1]); return n; }
int m17o_fmt_preamble(uint8_t *d) { for (int i = 0; i < 192; i += 2) { d[i] = 1; d[i + 1] = 3; } return 192; }
int m17o_fmt_eot(uint8_t *d) { for (int i = 0; i < 192; i++) d[i] = (i % 8) == 6 ? 3 : 1; return 192; }
static int finish_frame(uint16_t sync, uint8_t *a, uint8_t *b, uint8_t *dibits) {
    /* interleave -> randomise -> sync + dibits; a holds 368 type-3 bits, b is scratch */
    m17o_interleave(a, b, 368);
    m17o_derand_bits(b, a, 368);
    int n = put_sync(sync, dibits);
    return n + bits_to_dibits(a, dibits + n, 368);
}
int m17o_fmt_lsf(const uint8_t *lsf30, uint8_t *dibits) {
    uint8_t a[512], b[512];
    int len = m17o_conv_encode_8(lsf30, a, 30);          /* 488 */
    len = m17o_punc(1, a, b, len);                       /* 368 */
    memcpy(a, b, 368);
    return finish_frame(0x55F7, a, b, dibits);
}
int m17o_fmt_stream(m17o_tx *t, const uint8_t *payload, uint8_t *dibits) {
    uint8_t tmp[24], a[512], b[512];
    /* LICH chunk + counter, 4 x Golay (m17_tx_routines.cpp:151-164) */
    memcpy(tmp, &t->lich[t->lich_count * 5], 5);
    tmp[5] = (uint8_t)((t->lich_count & 7) << 5);
    t->lich_count = (t->lich_count + 1) % 6;
    uint16_t dw[4];
    dw[0] = (uint16_t)((tmp[0] << 4) | (tmp[1] >> 4)); dw[1] = (uint16_t)(((tmp[1] & 15) << 8) | tmp[2]);
    dw[2] = (uint16_t)((tmp[3] << 4) | (tmp[4] >> 4)); dw[3] = (uint16_t)(((tmp[4] & 15) << 8) | tmp[5]);
    int len = 0;
    for (int i = 0; i < 4; i++) { uint32_t w = m17o_golay_encode(dw[i]); for (int k = 23; k >= 0; k--) a[len++] = (w >> k) & 1; }
    /* FN + payload, conv, P2 (:166-176) */
    tmp[0] = (uint8_t)(t->fn >> 8); tmp[1] = (uint8_t)t->fn;
    t->fn = (t->fn + 1) & 0xFFFF;
    memcpy(tmp + 2, payload, 16);
    int n = m17o_conv_encode_8(tmp, b, 18);              /* 296 */
    n = m17o_punc(2, b, a + 96, n);                      /* 272 */
    return finish_frame(0xFF5D, a, b, dibits);
}
int m17o_fmt_packet(const uint8_t *chunk, int len, int eof, int nf, uint8_t *dibits) {
    uint8_t tmp[26], a[512], b[512];
    if (len > 25) return 0;
    memset(tmp, 0, 26); memcpy(tmp, chunk, len);
    tmp[25] = (uint8_t)((eof ? 0x80 : 0) | (nf << 2));
    m17o_conv_encode_8(tmp, a, 26);                      /* 424 */
    m17o_punc(3, a, b, 420);                             /* 368 */
    memcpy(a, b, 368);
    return finish_frame(0x75FF, a, b, dibits);
}
int m17o_fmt_bert(m17o_tx *t, uint8_t *dibits) {
    uint8_t a[512], b[512];
    for (int i = 0; i < 197; i++) { a[i] = t_prbs[t->prbs_idx]; t->prbs_idx = (t->prbs_idx + 1) % 511; }
    int n = m17o_conv_encode_1(a, b, 197);               /* 402 */
    m17o_punc(2, b, a, n);                               /* 369, first 368 used */
    return finish_frame(0xDF55, a, b, dibits);
}
long m17o_mod(m17o_tx *t, const uint8_t *syms, long nsym, int16_t *iq, float *freq) {
    /* dibit -> deviation in rad/sample (m17_modulate.cpp:9): the LUT entries are double quotients
       rounded to float on initialisation */
    const float lu[5] = { (float)(M_PI / 30.0), (float)(M_PI / 10.0), (float)(-M_PI / 30), (float)(-M_PI / 10.0), 0.0f };
    const int os = t->os;
    long n = 0;
    for (long k = 0; k < nsym; k++) {
        for (int i = 0; i < 30; i++) t->s[i] = t->s[i + 1];              /* :51-55 */
        t->s[30] = lu[syms[k]];
        for (int i = 0, ph = os - 1; i < os; i++, ph--) {                /* :57-59 and :42-48 */
            const float *c = &t->taps[ph];
            float sum = t->s[0] * c[0];
            for (int j = 1; j < 31; j++) sum += t->s[j] * c[j * os];
            if (freq) freq[n] = sum;
            t->acc += sum;                                               /* :23-27 float accumulator */
            iq[2 * n]     = (int16_t)(cosf(t->acc) * 0x3FFF);
            iq[2 * n + 1] = (int16_t)(sinf(t->acc) * 0x3FFF);
            n++;
        }
        /* wrap once per symbol through double modf, rounding to float at every store (:33-37) */
        double ip;
        t->acc = (float)(t->acc / (2.0 * M_PI));
        t->acc = (float)modf(t->acc, &ip);
        t->acc = (float)(t->acc * 2.0 * M_PI);
    }
    return n;
}

/* ================================================================== M17-over-UDP reflector frame */
/* m17_net_new_rx_data / net_add_* (m17_net.cpp:25-74), m17_send_stream_frame_to_net (m17_tx_routines.cpp:298-306):
   "M17 " | stream id (2, BE) | LSF bytes 0..27 (dst6 src6 type2 meta14) | FN (2, BE) | payload 16 | CRC-16 over the first 52.
   have_dst: the RX gateway path overwrites the destination call with the reflector's (m17_net.cpp:56-61) */
void m17o_net_pack(uint16_t sid, const uint8_t *lsf28, int have_dst, uint64_t dst, uint16_t fn, const uint8_t *pld16, uint8_t *o) {
    o[0] = 0x4D; o[1] = 0x31; o[2] = 0x37; o[3] = 0x20;
    o[4] = (uint8_t)(sid >> 8); o[5] = (uint8_t)sid;
    memcpy(o + 6, lsf28, 28);
    if (have_dst) for (int i = 0; i < 6; i++) o[6 + i] = (uint8_t)(dst >> (40 - 8 * i));
    o[34] = (uint8_t)(fn >> 8); o[35] = (uint8_t)fn;
    memcpy(o + 36, pld16, 16);
    uint16_t crc = m17o_crc(o, 52);
    o[52] = (uint8_t)(crc >> 8); o[53] = (uint8_t)crc;
}
/* m17_parse_m17_data (m17_net.cpp:203-238) + build_lich_from_net (m17_tx_routines.cpp:71-86): accept iff the CRC over all 54
   bytes is 0; the TX side's LSF is bytes 6..33 (the TYPE word survives m17_upack_type/m17_pack_type unchanged) + a fresh CRC */
int m17o_net_parse(const uint8_t *b, uint16_t *sid, uint8_t *lsf30, uint16_t *fn, uint8_t *pld16) {
    *sid = (uint16_t)((b[4] << 8) | b[5]);
    memcpy(lsf30, b + 6, 28);
    uint16_t crc = m17o_crc(lsf30, 28);
    lsf30[28] = (uint8_t)(crc >> 8); lsf30[29] = (uint8_t)crc;
    *fn = (uint16_t)((b[34] << 8) | b[35]);
    memcpy(pld16, b + 36, 16);
    return m17o_crc(b, 54) == 0;
}

/* ================================================================== Pluto /8 decimator (radio.cpp:18-51,157-177) */
void m17o_lpf_design(float *taps, float bw, int ntaps) {
    /* m17_dsp.cpp:347-360: rectangular-window sinc, double math, first tap time by integer division */
    double B = bw, t = -(ntaps - 1) / 2;
    for (int i = 0; i < ntaps; i++) {
        double a = (t == 0) ? 2.0 * B : 2.0 * B * sin(M_PI * t * B) / (M_PI * t * B);
        taps[i] = (float)a;
        t = t + 1.0;
    }
}
void m17o_dec_open(m17o_dec *d) {
    float f[31];
    memset(d, 0, sizeof(*d));
    m17o_lpf_design(f, 0.125f, 31);
    m17o_set_gain(f, 0.9, 1, 31);                                /* radio.cpp:48: gain literal 0.9 (double) narrowed to the float parameter */
    for (int i = 0; i < 31; i++) d->taps[i] = (int16_t)(f[i] * 0x7FFF);     /* m17_dsp_float_to_short, m17_dsp.cpp:382-386 */
}
void m17o_dec_run(m17o_dec *d, const int16_t *in, long nout, int16_t *out) {
    /* streaming form of rx_decimate_filter over m_rx_buff: output k uses inputs 8k-31 .. 8k-1 (the 31-sample history
       carried at the head of the buffer, radio.cpp:167), symmetric taps folded, int32 accumulate, arithmetic >> 15 */
    for (long k = 0; k < nout; k++) {
        int32_t re, im, w[31][2];
        for (int j = 0; j < 31; j++) {
            long n = 8 * k - 31 + j;
            if (n >= 0) { w[j][0] = in[2 * n]; w[j][1] = in[2 * n + 1]; }
            else        { w[j][0] = d->hist[31 + n][0]; w[j][1] = d->hist[31 + n][1]; }
        }
        re = w[15][0] * d->taps[15];
        im = w[15][1] * d->taps[15];
        for (int i = 0; i < 15; i++) { re += d->taps[i] * (w[i][0] + w[30 - i][0]); im += d->taps[i] * (w[i][1] + w[30 - i][1]); }
        out[2 * k] = (int16_t)(re >> 15);
        out[2 * k + 1] = (int16_t)(im >> 15);
    }
    /* new history = last 31 inputs */
    int16_t nh[31][2];
    for (int j = 0; j < 31; j++) {
        long n = 8 * nout - 31 + j;
        if (n >= 0) { nh[j][0] = in[2 * n]; nh[j][1] = in[2 * n + 1]; } else { nh[j][0] = d->hist[31 + n][0]; nh[j][1] = d->hist[31 + n][1]; }
    }
    memcpy(d->hist, nh, sizeof(nh));
}

/* ================================================================== RX */
struct m17o_rx {
    /* front end */
    int   disc_count; float z0re, z0im, z1re, z1im;      /* m17_dsp.cpp:195-196 */
    double nco_acc;                                      /* :391 */
    int   afc_on; float afc_delta;                       /* radio.cpp:8-10 */
    /* timing loop */
    float buff[M17O_FN]; int clk, thr, index; float sum, dif;   /* m17_rx_sync.cpp:7-12,78 */
    /* equaliser option (m17_equalize.cpp has no call site upstream): the matched filter's output half a
       symbol before each symbol instant, and the equaliser's statics */
    int   eq_on; float mid; m17o_eq eq;
    /* framer */
    float fsym[192], win[8]; int flock, fclk, ferr;      /* m17_rx_frame.cpp:14-18,104 */
    /* parser */
    uint8_t lsf[2][30], packet[800]; int packet_idx;     /* m17_rx_parse.cpp:5-7 */
    int in_frame; uint32_t g_errors, n_frames;           /* m17defines.h:104-107 */
    /* BERT receive (extension: the reference's decode_bert_frame is empty, m17_rx_parse.cpp:178-180): PRBS9 checker statics
       m_rx_idx, m_rx_state, m_rx_bad, m_rx_good, m_rx_eq_cnt, m_rx_dif_cnt (m17_prbs9.cpp:7-12) + two running totals */
    int bert_on; uint16_t prbs_idx; int prbs_state; uint16_t prbs_bad, prbs_good, prbs_eq, prbs_dif;
    uint32_t bert_bits, bert_errs;
    /* trace */
    float *t_disc; int32_t *t_nsym; float *t_syms; long symcap;
    m17_frame_rec *t_frames; long fcap; float *t_soft; m17_event_rec *t_events; long ecap;
    long n_blocks, n_syms, n_rec, n_events;
    long cur_sym;
};
m17o_rx *m17o_rx_new(void) {
    m17o_init();
    m17o_rx *r = (m17o_rx *)calloc(1, sizeof(*r));
    r->clk = 1; r->index = 10;                           /* m17_rx_sync.cpp:124-127 */
    return r;
}
void m17o_rx_free(m17o_rx *r) { free(r); }
void m17o_rx_set_afc(m17o_rx *r, int on) { r->afc_on = on; }
void m17o_rx_set_bert(m17o_rx *r, int on) { r->bert_on = on; }
void m17o_rx_set_eq(m17o_rx *r, int on) { r->eq_on = on; if (on) m17o_eq_open(&r->eq); }
void m17o_rx_get_bert(const m17o_rx *r, uint32_t *o) {
    o[0] = (uint32_t)r->prbs_state; o[1] = r->prbs_idx; o[2] = r->prbs_bad; o[3] = r->prbs_good; o[4] = r->prbs_eq; o[5] = r->prbs_dif;
    o[6] = r->bert_bits; o[7] = r->bert_errs;
}
/* m17_prbs9_rx_check (m17_prbs9.cpp:40-64), literally: hunting restarts the local sequence on every mismatch, 18 matches in a
   row give sync, 18 mismatches in a row lose it; `if(d == 9) m_rx_good++` can never fire (d is 0 or 1) and is kept as is */
void m17o_prbs9_rx_check(m17o_rx *r, uint8_t bit) {
    uint8_t d = bit ^ t_prbs[r->prbs_idx];
    if (d) { r->prbs_dif++; r->prbs_eq = 0; } else { r->prbs_eq++; r->prbs_dif = 0; }
    r->prbs_idx = (uint16_t)((r->prbs_idx + 1) % 511);
    if (r->prbs_state == 0) {
        r->prbs_bad = 0; r->prbs_good = 0;
        if (r->prbs_eq >= 18) r->prbs_state = 1;
        if (d) { r->prbs_idx = 0; r->prbs_state = 0; }           /* m17_prbs9_rx_reset */
    } else {
        r->bert_bits++; if (d) r->bert_errs++;                   /* running totals (ours) */
        if (r->prbs_dif >= 18) r->prbs_state = 0;
        if (d) r->prbs_bad++;
        if (d == 9) r->prbs_good++;
    }
}
void m17o_rx_trace(m17o_rx *r, float *disc, int32_t *nsym, float *syms, long symcap, m17_frame_rec *frames, long fcap,
                   float *soft, m17_event_rec *events, long ecap) {
    r->t_disc = disc; r->t_nsym = nsym; r->t_syms = syms; r->symcap = symcap; r->t_frames = frames; r->fcap = fcap;
    r->t_soft = soft; r->t_events = events; r->ecap = ecap;
}
void m17o_rx_counts(const m17o_rx *r, int64_t *o) { o[0] = r->n_blocks; o[1] = r->n_syms; o[2] = r->n_rec; o[3] = r->n_events; }

static void rx_event(m17o_rx *r, int kind) {
    if (r->t_events && r->n_events < r->ecap) { r->t_events[r->n_events].sym_idx = (int32_t)r->cur_sym; r->t_events[r->n_events].kind = kind; }
    r->n_events++;
}
static m17_frame_rec *rx_new_rec(m17o_rx *r, m17_frame_rec *scratch) {
    m17_frame_rec *f = (r->t_frames && r->n_rec < r->fcap) ? &r->t_frames[r->n_rec] : scratch;
    memset(f, 0, sizeof(*f));
    return f;
}

/* ---- front end: int16 IQ -> limiter -> discriminator /5 -> block-mean removal */
void m17o_frontend(m17o_rx *r, const int16_t *iq, int nsamp, float *disc, int *ndisc, float *mean) {
    float offset = 0; int idx = 0;
    float delta = 0;
    /* radio_get_afc_status / radio_get_afc_delta (radio.cpp:201-208, m17_dsp.cpp:468) */
    if (r->afc_on) { if (r->in_frame) delta = r->afc_delta; else { r->afc_delta = 0; delta = 0; } }
    for (int i = 0; i < nsamp; i++) {
        /* dsp_short_to_float (m17_dsp.cpp:136-141): int16 * 0.00003 in double, rounded to float */
        float re = (float)(iq[2 * i] * 0.00003), im = (float)(iq[2 * i + 1] * 0.00003);
        if (r->afc_on) {                                             /* dsp_nco_mixer :390-399 */
            float c = (float)cos(r->nco_acc), s = (float)sin(r->nco_acc);
            r->nco_acc += delta;
            float nre = (re * c) - (im * s), nim = (re * s) + (im * c);
            re = nre; im = nim;
        }
        /* dsp_limit (:412-419): sqrtf of the float sum of squares; gain = 1.0/m in double, rounded */
        float m = sqrtf(re * re + im * im);
        float g = (float)(1.0 / m);
        re *= g; im *= g;
        /* dsp_arctan_disc2 (:194-222) */
        float a = r->z0im * (re - r->z1re);
        float b = r->z0re * (im - r->z1im);
        float u = b - a;
        r->z1re = r->z0re; r->z1im = r->z0im; r->z0re = re; r->z0im = im;
        r->disc_count = (r->disc_count + 1) % 5;
        if (r->disc_count == 0) disc[idx++] = u * 0.5f;
        offset += u * 0.5f;
    }
    if (r->afc_on) {                                                 /* :401-407 */
        double ip;
        r->nco_acc = r->nco_acc / (2.0 * M_PI);
        r->nco_acc = modf(r->nco_acc, &ip);
        r->nco_acc = r->nco_acc * 2.0 * M_PI;
        if (r->nco_acc != r->nco_acc) r->nco_acc = 0;
    }
    offset = offset / nsamp;
    if (r->afc_on && r->in_frame) r->afc_delta -= offset * 0.1;      /* radio_afc, radio.cpp:196-200 */
    for (int i = 0; i < idx; i++) disc[i] = disc[i] - offset;
    *ndisc = idx; *mean = offset;
}

/* ---- matched filter + symbol-timing loop (m17_rx_sync.cpp:25-99) */
static float dot31(const float *in, const float *c) { float s = in[0] * c[0]; for (int i = 1; i < M17O_FN; i++) s += in[i] * c[i]; return s; }
/* `mid` (optional): the T/2-spaced companion of every symbol for the equaliser option -- the SAME matched filter branch
   (rx_sync_filter(m_buff, m_mf[m_index], FN), m17_rx_sync.cpp:25-31,84) evaluated on the sample before the symbol instant,
   i.e. on the "else" sample of the loop below after m17_sync_adjust has run, so that both halves of a pair use one branch.
   A symbol inserted by a forward bit slip (:56-58) gets mid = 0 like the symbol itself. */
static int sync_samples_mid(m17o_rx *r, const float *in, float *out, float *mid, int len) {
    int m_idx = 0;                        /* may legally reach -1 on a backward slip at block start:
                                             the next symbol is then written to out[-1], i.e. lost */
    const int thresh = r->flock ? 80 : 10;   /* m17_rx_lock() cannot change inside a block (:91-94) */
    for (int i = 0; i < len; i++) {
        for (int k = 0; k < M17O_FN - 1; k++) r->buff[k] = r->buff[k + 1];
        r->buff[M17O_FN - 1] = in[i];
        r->clk = (r->clk + 1) % 2;
        if (r->clk) {
            r->sum = dot31(r->buff, t_mf[r->index]);
            r->dif = dot31(r->buff, t_md[r->index]);
            if (m_idx >= 0) { out[m_idx] = r->sum; if (mid) mid[m_idx] = r->mid; }
            m_idx++;
        } else {
            float dif = r->dif;                                  /* sync_update :38-42 */
            if (r->sum < 0) dif = -dif;
            if (dif > 0) r->thr++;
            if (dif < 0) r->thr--;
            if (r->thr > thresh) {                               /* m17_sync_adjust :45-72 */
                r->index = (r->index + 1) % M17O_NF; r->thr = 0;
                if (r->index == 0) { r->clk = 1; if (m_idx >= 0) { out[m_idx] = 0; if (mid) mid[m_idx] = 0; } m_idx++; }
            }
            if (r->thr < -thresh) {
                r->thr = 0; r->index = (r->index + M17O_NF - 1) % M17O_NF;
                if (r->index == M17O_NF - 1) { r->clk = 1; m_idx--; }
            }
            if (mid) r->mid = dot31(r->buff, t_mf[r->index]);
        }
    }
    return m_idx < 0 ? 0 : m_idx;
}
int m17o_sync_samples(m17o_rx *r, const float *in, float *out, int len) { return sync_samples_mid(r, in, out, 0, len); }

/* ---- frame decode (m17_rx_parse.cpp:86-226) */
static void pack_bits(const uint8_t *bits, uint8_t *out, int nbits) {
    for (int i = 0; i < nbits; i += 8) { uint8_t b = 0; for (int k = 0; k < 8; k++) b = (uint8_t)((b << 1) | bits[i + k]); out[i >> 3] = b; }
}
static void rx_parse_lsf_event(m17_frame_rec *f) { f->flags |= M17R_F_LSF_EVENT; }   /* parse_lsf :52-70 */
static void rx_parse(m17o_rx *r, const float *s, int type, m17_frame_rec *f) {
    float sb[368], so[368], dp[488]; uint8_t bits[256];
    if (type < 1 || type > 4) return;                            /* preamble / EOT: frame-id only (:194-197,218-221) */
    m17o_demap_frame(s, sb);
    { float sum = 0; for (int i = 0; i < 8; i++) sum += fabsf(s[i]); f->cor = (float)(8.0 / sum); }
    if (r->t_soft && r->n_rec < r->fcap) memcpy(&r->t_soft[r->n_rec * 368], sb, sizeof(sb));
    if (type == 4 && !r->bert_on) return;                        /* decode_bert_frame is empty (:178-180) */
    m17o_derand_soft(sb, sb, 368);
    m17o_deinterleave(sb, so, 368);
    if (type == 4) {
        /* what decode_bert_frame was meant to be (inverse of m17_fmt_add_bert_frame, m17_tx_routines.cpp:226-238): de-puncture P2
           to 402 coded bits (the 369th kept bit was never sent: an erasure), Viterbi over 201 steps, 197 PRBS9 bits to the checker */
        float sp[369];
        memcpy(sp, so, sizeof(so)); sp[368] = 0.0f;
        m17o_depunc(2, sp, dp, 402);
        m17o_viterbi(dp, bits, 402);
        pack_bits(bits + 1, f->data, 200); f->nbytes = 25;
        for (int i = 1; i <= 197; i++) m17o_prbs9_rx_check(r, bits[i]);
        f->crc = m17o_crc(f->data, f->nbytes);
        return;
    }
    if (type == 1) {                                             /* decode_link_frame :86-101 */
        m17o_depunc(1, so, dp, 488);
        m17o_viterbi(dp, bits, 488);
        pack_bits(bits + 1, f->data, 240); f->nbytes = 30;
        if (m17o_crc(r->packet, 30) == 0) rx_parse_lsf_event(f);  /* D3: checks m_packet, not the new bytes */
    } else if (type == 2) {                                      /* decode_stream_frame :105-160 */
        uint16_t w[4]; int e = 0;
        for (int k = 0; k < 4; k++) e += m17o_golay_decode(m17o_hard24(&so[24 * k]), &w[k]);
        r->g_errors += (uint32_t)e; r->n_frames++;               /* m17_db_golay_errors, m17_dbase.cpp:79-82 */
        f->golay_err = (uint8_t)e;
        uint32_t w01 = ((uint32_t)w[0] << 12) | w[1], w23 = ((uint32_t)w[2] << 12) | w[3];   /* pack_12_to_8_x4x6 */
        f->lich[0] = (uint8_t)(w01 >> 16); f->lich[1] = (uint8_t)(w01 >> 8); f->lich[2] = (uint8_t)w01;
        f->lich[3] = (uint8_t)(w23 >> 16); f->lich[4] = (uint8_t)(w23 >> 8); f->lich[5] = (uint8_t)w23;
        int seq = f->lich[5] >> 5;                               /* update_lich :71-85 */
        if (seq < 6) {
            memcpy(&r->lsf[0][seq * 5], f->lich, 5);
            if (m17o_crc(r->lsf[0], 30) == 0) { memcpy(r->lsf[1], r->lsf[0], 30); rx_parse_lsf_event(f); }
        }
        m17o_depunc(2, &so[96], dp, 296);
        m17o_viterbi(dp, bits, 296);
        pack_bits(bits + 1, f->data, 144); f->nbytes = 18;
        if (m17o_crc(r->lsf[1], 30) == 0) f->flags |= M17R_F_DELIVERED;   /* :148-158 */
    } else {                                                     /* decode_packet_frame :161-177 */
        m17o_depunc(3, so, dp, 420);
        m17o_viterbi(dp, bits, 420);
        pack_bits(bits + 1, f->data, 208); f->nbytes = 26;
        int eof = f->data[25] >> 7, fn = (f->data[25] >> 2) & 0x1F;
        if (eof) {                                               /* parse_packet :34-51, incl. defect D4 */
            f->flags |= M17R_F_PKT_EOF;
            int room = 800 - r->packet_idx; int n = fn < room ? fn : room;   /* the reference can overrun here; we clamp */
            if (n > 0) memcpy(&r->packet[r->packet_idx], f->data, (size_t)n);
            r->packet_idx = 0;
        } else {
            memcpy(&r->packet[fn * 25], f->data, 25);
            r->packet_idx = fn * 25;
        }
    }
    f->crc = m17o_crc(f->data, f->nbytes);
}

/* ---- framer (m17_rx_frame.cpp:126-172) */
static void rx_los(m17o_rx *r) { r->in_frame = 0; rx_event(r, M17R_EV_LOS); }            /* m17_dbase.cpp:68-74 */
static void rx_sym(m17o_rx *r, float sym) {
    int type, votes; float var; m17_frame_rec scratch;
    if (r->flock) {
        r->fsym[r->fclk] = sym;
        r->fclk = (r->fclk + 1) % 192;
        if (r->fclk == 0) {
            m17o_sync_check(r->fsym, &type, &votes, &var);
            m17_frame_rec *f = rx_new_rec(r, &scratch);
            f->sym_off = (int32_t)(r->cur_sym - 191); f->type = (uint8_t)type; f->votes = (uint8_t)votes; f->variance = var;
            int ok = sync_accept(type, votes, var, 1);
            if (ok) f->flags |= M17R_F_SYNC_OK;
            if (type == 5) {                                     /* EOT :137-140 */
                r->flock = 0; memset(r->win, 0, sizeof(r->win));
                f->flags |= M17R_F_LOS; f->frame_errors = (uint8_t)r->ferr;
                rx_los(r);
            } else if (ok) {
                f->flags |= M17R_F_PARSED;
                rx_parse(r, r->fsym, type, f);
                r->ferr = 0; f->frame_errors = 0;
            } else {
                r->ferr++; f->frame_errors = (uint8_t)r->ferr;
                if (r->ferr > 5) {                               /* N_FERROR :122,147-150 */
                    r->flock = 0; memset(r->win, 0, sizeof(r->win));
                    f->flags |= M17R_F_LOS;
                    rx_los(r);
                } else { f->flags |= M17R_F_PARSED; rx_parse(r, r->fsym, type, f); }
            }
            r->n_rec++;
        }
    } else {
        for (int i = 0; i < 7; i++) r->win[i] = r->win[i + 1];
        r->win[7] = sym;
        m17o_sync_check(r->win, &type, &votes, &var);
        if (sync_accept(type, votes, var, 0)) {
            memcpy(r->fsym, r->win, sizeof(r->win));
            r->fclk = 8; r->ferr = 0; r->flock = 1;
            r->g_errors = 0; r->n_frames = 0; r->in_frame = 1;   /* m17_aos, m17_dbase.cpp:60-67 */
            rx_event(r, M17R_EV_AOS);
        }
    }
}
static void rx_symbols(m17o_rx *r, const float *sym, int n) {
    for (int i = 0; i < n; i++) {
        r->cur_sym = r->n_syms;
        if (r->t_syms && r->n_syms < r->symcap) r->t_syms[r->n_syms] = sym[i];
        rx_sym(r, sym[i]);
        r->n_syms++;
    }
}
void m17o_rx_baseband(m17o_rx *r, const float *disc, int n) {
    float tmp[960];
    if (r->t_disc && n <= 384) memcpy(&r->t_disc[r->n_blocks * 384], disc, sizeof(float) * (size_t)n);
    int k;
    if (r->eq_on) {
        /* equaliser option: eq_train_unknown (m17_equalize.cpp:185-213) on every (half-symbol, symbol) pair, its output is the
           symbol the framer sees (between m17_rx_sync.cpp:77 and m17_rx_frame.cpp:173) */
        float mid[960];
        k = sync_samples_mid(r, disc, tmp + 1, mid + 1, n);
        for (int q = 0; q < k; q++) { float in2[2] = { mid[1 + q], tmp[1 + q] }; tmp[1 + q] = m17o_eq_train_unknown(&r->eq, in2); }
    } else
    k = m17o_sync_samples(r, disc, tmp + 1, n);     /* +1: room for the D6 out[-1] write */
    if (r->t_nsym) r->t_nsym[r->n_blocks] = k;
    rx_symbols(r, tmp + 1, k);
    r->n_blocks++;
}
void m17o_dsp_rx(m17o_rx *r, const int16_t *iq, int nsamp) {
    float disc[M17O_BLOCK]; int nd; float mean;
    m17o_frontend(r, iq, nsamp, disc, &nd, &mean);
    m17o_rx_baseband(r, disc, nd);
}

/* ================================================================== batch drivers */
typedef struct {
    const void *in; int seam; long C, T; int nthr, tid;
    float *disc; int32_t *nsym; float *syms; long symcap; m17_frame_rec *frames; long fcap; float *soft;
    m17_event_rec *events; long ecap; int64_t *counts; double secs;
} job_t;
static uint32_t *g_bert_out = 0;                                /* optional [C][8] sink for the BERT checker state of a batch run */
void m17o_set_bert_out(uint32_t *p) { g_bert_out = p; }
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void *job_main(void *p) {
    job_t *j = (job_t *)p;
    double t0 = now_s();
    for (long c = j->tid; c < j->C; c += j->nthr) {
        m17o_rx *r = m17o_rx_new();
        m17o_rx_trace(r, j->disc ? j->disc + c * j->T * 384 : 0, j->nsym ? j->nsym + c * j->T : 0,
                      j->syms ? j->syms + c * j->symcap : 0, j->symcap, j->frames ? j->frames + c * j->fcap : 0, j->fcap,
                      j->soft ? j->soft + c * j->fcap * 368 : 0, j->events ? j->events + c * j->ecap : 0, j->ecap);
        if (j->seam & 32) m17o_rx_set_bert(r, 1);         /* seam flag 32: BERT receive extension on */
        if (j->seam & 64) m17o_rx_set_eq(r, 1);           /* seam flag 64: equaliser option on */
        if (j->seam & 16) m17o_rx_set_afc(r, 1);          /* seam flag 16: AFC on (radio_set_afc_on, radio.cpp:146-148) */
        if ((j->seam & 15) == 0) { const int16_t *iq = (const int16_t *)j->in + c * j->T * 3840; for (long t = 0; t < j->T; t++) m17o_dsp_rx(r, iq + t * 3840, 1920); }
        else { const float *d = (const float *)j->in + c * j->T * 384; for (long t = 0; t < j->T; t++) m17o_rx_baseband(r, d + t * 384, 384); }
        if (j->counts) m17o_rx_counts(r, j->counts + c * 4);
        if (g_bert_out) m17o_rx_get_bert(r, g_bert_out + c * 8);
        m17o_rx_free(r);
    }
    j->secs = now_s() - t0;
    return 0;
}
static double run_jobs(job_t *proto, int nthr) {
    if (nthr < 1) nthr = 1;
    if (nthr > 256) nthr = 256;
    job_t jobs[256]; pthread_t th[256]; double mx = 0;
    m17o_init();
    for (int i = 0; i < nthr; i++) { jobs[i] = *proto; jobs[i].nthr = nthr; jobs[i].tid = i; pthread_create(&th[i], 0, job_main, &jobs[i]); }
    for (int i = 0; i < nthr; i++) { pthread_join(th[i], 0); if (jobs[i].secs > mx) mx = jobs[i].secs; }
    return mx;
}
int m17o_rx_run(const void *in, int seam, long C, long T, int nthreads, float *disc, int32_t *nsym, float *syms, long symcap,
                m17_frame_rec *frames, long fcap, float *soft, m17_event_rec *events, long ecap, int64_t *counts) {
    job_t j; memset(&j, 0, sizeof(j));
    j.in = in; j.seam = seam; j.C = C; j.T = T; j.disc = disc; j.nsym = nsym; j.syms = syms; j.symcap = symcap;
    j.frames = frames; j.fcap = fcap; j.soft = soft; j.events = events; j.ecap = ecap; j.counts = counts;
    run_jobs(&j, nthreads);
    return 0;
}
double m17o_rx_time(const int16_t *iq, long C, long T, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.in = iq; j.seam = 0; j.C = C; j.T = T;
    return run_jobs(&j, nthreads);
}

/* ================================================================== wideband channeliser (second half)
   The reference's Pluto path low-pass filters and decimates ONE 384 kS/s channel to 48 kS/s (sub_filter / rx_decimate_filter,
   radio.cpp:18-40: int16 taps, int32 accumulate, >> 15).  Generalisation to M channels spaced Fs / M out of one capture at
   Fs = D * 48 kS/s: the same integer FIR, folded into M polyphase branches, followed by an M-point DFT in fixed point.
     window of output n   : x[nD - L + i], i = 0..L-1          (the reference's window for M = 1, D = 8, L = 31)
     fold                 : z[p] = sum_q h[p + qM] * x[nD - L + p + qM]                      (int32, wraps like the reference)
     DFT                  : Z[k] = sum_p z[p] e^{-j 2 pi k p / M}, M = 2^b or 3 * 2^b: prime-factor split 3 x 2^b (no twiddles
                            between the parts), radix-2 decimation in time; twiddles are Q30 integers (+-1 exact), a product is
                            (int64 a * w) >> 30 per real multiply, so the twiddles 1 and -j act exactly
     phase of the window  : Y[k] = Z[k] * e^{-j 2 pi k (nD - L) / M}   (Q30 table, index 0 = the identity)
     output               : y[k][n] = (int16)(Y[k] >> 15)                                   (the reference's scaling)
   For M = 1 every step but the fold and the shift is the identity: the output IS radio.cpp's decimator (pinned in
   tests/test_oracle_vs_ref.py with the reference's own taps).  TEST INFRASTRUCTURE ONLY. */
static int32_t chq_mul(int32_t a, int32_t w) { return (int32_t)(((int64_t)a * (int64_t)w) >> 30); }
static int32_t chq_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t chq_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t chq_tw(double v) { double s = v * 1073741824.0; s = s < 0 ? s - 0.5 : s + 0.5; return (int32_t)s; }    /* Q30: 1.0 and -1.0 are exact */
static void chq_cmul(int32_t ar, int32_t ai, int32_t wr, int32_t wi, int32_t *or_, int32_t *oi) {
    *or_ = chq_sub(chq_mul(ar, wr), chq_mul(ai, wi));
    *oi = chq_add(chq_mul(ar, wi), chq_mul(ai, wr));
}
m17o_chan *m17o_chan_open(int M, int D, int L, const int16_t *taps) {
    int b = 0, m2 = M, has3 = 0;
    if (M >= 3 && M % 3 == 0) { has3 = 1; m2 = M / 3; }
    while ((1 << b) < m2) b++;
    if ((1 << b) != m2 || M < 1 || D < 1 || L < 1) return NULL;
    m17o_chan *c = (m17o_chan *)calloc(1, sizeof(*c));
    c->M = M; c->D = D; c->L = L; c->lg2 = b; c->has3 = has3;
    c->h = (int16_t *)malloc(sizeof(int16_t) * L); memcpy(c->h, taps, sizeof(int16_t) * L);
    c->hist = (int16_t *)calloc((size_t)2 * L, sizeof(int16_t));
    c->tw2 = (int32_t *)calloc((size_t)2 * (m2 > 1 ? m2 / 2 : 1), sizeof(int32_t));
    for (int j = 0; j < m2 / 2; j++) { c->tw2[2 * j] = chq_tw(cos(2.0 * M_PI * j / m2)); c->tw2[2 * j + 1] = chq_tw(-sin(2.0 * M_PI * j / m2)); }
    c->rot = (int32_t *)calloc((size_t)2 * M, sizeof(int32_t));
    for (int r = 0; r < M; r++) { c->rot[2 * r] = chq_tw(cos(2.0 * M_PI * r / M)); c->rot[2 * r + 1] = chq_tw(-sin(2.0 * M_PI * r / M)); }
    c->s3 = chq_tw(sqrt(3.0) / 2.0);
    c->n_done = 0;
    return c;
}
void m17o_chan_free(m17o_chan *c) { if (c) { free(c->h); free(c->hist); free(c->tw2); free(c->rot); free(c); } }
static int chq_bitrev(int v, int bits) { int r = 0; for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i); return r; }
/* in-place 2^b-point DFT of (re, im), input in natural order */
static void chq_fft2(const m17o_chan *c, int32_t *re, int32_t *im) {
    const int N = 1 << c->lg2;
    int32_t tr[1024], ti[1024];
    for (int i = 0; i < N; i++) { tr[i] = re[chq_bitrev(i, c->lg2)]; ti[i] = im[chq_bitrev(i, c->lg2)]; }
    for (int m = 2; m <= N; m <<= 1) {
        for (int k = 0; k < N; k += m)
            for (int j = 0; j < m / 2; j++) {
                const int ti_ = j * (N / m);                       /* twiddle index of W_N^(j N / m) */
                int32_t vr = tr[k + j + m / 2], vi = ti[k + j + m / 2], xr, xi;
                chq_cmul(vr, vi, c->tw2[2 * ti_], c->tw2[2 * ti_ + 1], &xr, &xi);   /* (W = 1 and W = -j are exact in Q30) */
                const int32_t ur = tr[k + j], ui = ti[k + j];
                tr[k + j] = chq_add(ur, xr); ti[k + j] = chq_add(ui, xi);
                tr[k + j + m / 2] = chq_sub(ur, xr); ti[k + j + m / 2] = chq_sub(ui, xi);
            }
    }
    for (int i = 0; i < N; i++) { re[i] = tr[i]; im[i] = ti[i]; }
}
/* in [D * nout][2] -> out [M][out_pitch][2] (channel k = bin k: centre frequency k Fs / M, k >= M/2 negative) */
void m17o_chan_run(m17o_chan *c, const int16_t *in, long nout, int16_t *out, long out_pitch) {
    const int M = c->M, D = c->D, L = c->L, N2 = 1 << c->lg2;
    const long nin = (long)D * nout;
    int16_t *x = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)(L + nin));      /* history then the new samples */
    memcpy(x, c->hist, sizeof(int16_t) * 2 * L);
    memcpy(x + 2 * L, in, sizeof(int16_t) * 2 * nin);
    int32_t *zr = (int32_t *)malloc(sizeof(int32_t) * M), *zi = (int32_t *)malloc(sizeof(int32_t) * M);
    int32_t *Zr = (int32_t *)malloc(sizeof(int32_t) * M), *Zi = (int32_t *)malloc(sizeof(int32_t) * M);
    int32_t br[1024], bi[1024];
    const int c1 = c->has3 ? N2 * ((N2 % 3 == 1) ? 1 : 2) : 0;                 /* N2 * (N2^-1 mod 3) */
    int inv3 = 1; while (c->has3 && (3 * inv3) % N2 != 1 % N2) inv3++;
    const int c2 = c->has3 ? 3 * inv3 : 1;                                     /* 3 * (3^-1 mod N2) */
    for (long n = 0; n < nout; n++) {
        /* output n of this call: window x[nD - L .. nD - 1] (the reference's sub_filter(&in[i*8]) over a buffer that starts with
           the 31 carried samples, radio.cpp:36-38,167) */
        const int16_t *w = x + 2 * (n * D);
        for (int p = 0; p < M; p++) {
            uint32_t sr = 0, si = 0;
            for (int i = p; i < L; i += M) { sr += (uint32_t)((int32_t)c->h[i] * (int32_t)w[2 * i]); si += (uint32_t)((int32_t)c->h[i] * (int32_t)w[2 * i + 1]); }
            zr[p] = (int32_t)sr; zi[p] = (int32_t)si;
        }
        if (!c->has3) {
            if (N2 > 1) chq_fft2(c, zr, zi);
            for (int k = 0; k < M; k++) { Zr[k] = zr[k]; Zi[k] = zi[k]; }
        } else {
            for (int k1 = 0; k1 < 3; k1++) {
                for (int n2 = 0; n2 < N2; n2++) {
                    const int p0 = (3 * n2) % M, p1 = (N2 + 3 * n2) % M, p2 = (2 * N2 + 3 * n2) % M;
                    if (k1 == 0) { br[n2] = chq_add(chq_add(zr[p0], zr[p1]), zr[p2]); bi[n2] = chq_add(chq_add(zi[p0], zi[p1]), zi[p2]); }
                    else {
                        /* W3 = -1/2 - j sqrt(3)/2: z0 - (z1 + z2)/2 -+ j sqrt(3)/2 (z1 - z2) */
                        const int32_t t1r = chq_add(zr[p1], zr[p2]), t1i = chq_add(zi[p1], zi[p2]);
                        const int32_t t2r = chq_sub(zr[p1], zr[p2]), t2i = chq_sub(zi[p1], zi[p2]);
                        const int32_t m1r = chq_sub(zr[p0], t1r >> 1), m1i = chq_sub(zi[p0], t1i >> 1);
                        const int32_t m2r = chq_mul(t2r, c->s3), m2i = chq_mul(t2i, c->s3);
                        if (k1 == 1) { br[n2] = chq_add(m1r, m2i); bi[n2] = chq_sub(m1i, m2r); }      /* m1 - j m2 */
                        else         { br[n2] = chq_sub(m1r, m2i); bi[n2] = chq_add(m1i, m2r); }      /* m1 + j m2 */
                    }
                }
                if (N2 > 1) chq_fft2(c, br, bi);
                for (int k2 = 0; k2 < N2; k2++) { const int k = (c1 * k1 + c2 * k2) % M; Zr[k] = br[k2]; Zi[k] = bi[k2]; }
            }
        }
        const long long i0 = (long long)(c->n_done + n) * D - L;             /* absolute index of the window's first sample */
        for (int k = 0; k < M; k++) {
            long long r = ((long long)k * (i0 % M)) % M; if (r < 0) r += M;
            int32_t yr, yi;
            chq_cmul(Zr[k], Zi[k], c->rot[2 * r], c->rot[2 * r + 1], &yr, &yi);     /* r = 0: exactly the identity (Q30) */
            out[2 * ((long)k * out_pitch + n)] = (int16_t)(yr >> 15);
            out[2 * ((long)k * out_pitch + n) + 1] = (int16_t)(yi >> 15);
        }
    }
    memcpy(c->hist, x + 2 * nin, sizeof(int16_t) * 2 * L);
    c->n_done += nout;
    free(x); free(zr); free(zi); free(Zr); free(Zi);
}
