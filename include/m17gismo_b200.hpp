// m17gismo_b200.hpp -- header-only C++ shim: the ORIGINAL m17gismo function names and signatures
// (m17defines.h:219-244,259-330,349-375,398-409) at batch = 1, implemented on the C ABI of m17b200.h.
//
// A maintainer of the reference removes the hot-path .cpp files from the build (m17_dsp, m17_rx_sync, m17_rx_frame,
// m17_rx_parse, m17_conv, m17_puncture, m17_interleave, m17_correlate, m17_golay, m17_crc, m17_prbs9, m17_equalize,
// m17_modulate, the m17_fmt_*/m17_send_* half of m17_tx_routines), includes this header in ONE translation unit with
// M17GISMO_B200_IMPLEMENTATION defined, and links libm17b200.so.  All work is done on the GPU; calls are synchronous.
// Batch = 1 is for drop-in compatibility and tests -- throughput comes from calling the batched ABI directly.
//
// Differences a caller can observe (all consequences of returning records instead of up-calls):
//  * RX results arrive through m17b_shim_callbacks (frame records + AOS/LOS events) right after m17_dsp_rx /
//    m17_rx_symbols returns; the reference's gui_*/m17_db_*/m17_net_new_rx_data up-calls are the caller's to make.
//  * m17_rx_sync_samples runs the framer too (the kernel is fused); m17_rx_symbols on the symbols it returned then only
//    dispatches the records.  Any other symbols (m17_rx_symbols / m17_rx_sym on a caller's own stream) go through the
//    stand-alone framer (m17b_rx_symbols).
//  * blocks must be whole: m17_dsp_rx takes 1920 IQ samples, m17_rx_sync_samples 384 samples (as m17_dsp_rx feeds it).
//  * TX IQ is handed to m17b_shim_callbacks::transmit in 1920-sample blocks (radio_transmit_samples, radio.cpp:178).
#ifndef M17GISMO_B200_HPP
#define M17GISMO_B200_HPP
#include <stdint.h>
#include "m17b200.h"

typedef uint16_t uint12_t;   // m17defines.h:20-22
typedef uint32_t uint24_t;
typedef uint64_t uint48_t;
typedef struct { float re, im; } fcmplx;      // m17defines.h:125-128
typedef struct { int16_t re, im; } scmplx;    // m17defines.h:130-133
typedef struct { uint8_t p_s, dt, et, est, can, reserved; } M17Type;   // m17defines.h:34-41

struct m17b_shim_callbacks {
    void (*frame)(const m17b_frame_rec *rec, void *user);     // every completed frame, in order
    void (*event)(const m17b_event_rec *ev, void *user);      // AOS / LOS (m17_aos / m17_los)
    void (*transmit)(const scmplx *iq, uint32_t n, void *user);   // radio_transmit_samples
    void *user;
};
void m17b_shim_set_callbacks(const m17b_shim_callbacks *cb);
void m17b_shim_set_oversample(int os);                        // radio_get_oversample(): 10 (default) or 80, before m17_mod_init
int  m17b_shim_last_error(void);

// init chain (main.cpp:108-126)
void m17_prbs9_init(void); void m17_crc_init(void); void m17_init_conv(void); void m17_init_de_correlate(void);
void m17_dsp_init(void); void m17_fmt_init(void); void m17_golay_init(void); void m17_rx_sync_init(void); void m17_mod_init(void);
// filter design
void m17_dsp_build_rrc_filter(float *filter, float rolloff, int ntaps, int samples_per_symbol);
void m17_dsp_set_filter_gain(float *filter, float gain, int stride, int ntaps);
// RX
void m17_dsp_rx(scmplx *in, int len);
int  m17_rx_sync_samples(float *in, float *out, int len);
void m17_rx_symbols(float *sym, int len);
void m17_rx_sym(float sym);                                            // m17_rx_frame.cpp:126
void m17_rx_init(void); void m17_rx_lost(void); bool m17_rx_lock(void);
void radio_set_afc_on(void); void radio_set_afc_off(void); bool radio_get_afc_status(void);   // radio.cpp:146-155
void m17_dsp_demap_frame(float *in, float *out);
// FEC primitives
int m17_punc_p1(uint8_t *in, uint8_t *out, int len); int m17_punc_p2(uint8_t *in, uint8_t *out, int len); int m17_punc_p3(uint8_t *in, uint8_t *out, int len);
int m17_de_punc_p1(float *in, float *out, int len); int m17_de_punc_p2(float *in, float *out, int len); int m17_de_punc_p3(float *in, float *out, int len);
void m17_interleave(uint8_t *in, uint8_t *out, int len);
void m17_de_interleave(float *in, float *out, int len);
void m17_de_correlate_8(uint8_t *in, int len);
void m17_de_correlate_1(uint8_t *in, uint8_t *out, int len);
void m17_de_correlate_1(float *in, float *out, int len);
int m17_conv_encode_1(uint8_t *in, uint8_t *out, int len);
int m17_conv_encode_8(uint8_t *in, uint8_t *out, int len);
int m17_viterbi_decode(float *in, uint8_t *out, int len);
uint24_t m17_golay_encode(uint12_t data);
int m_17_golay_decode(uint24_t word, uint12_t &odata);
uint16_t m17_crc_array_encode(uint8_t *in, int len);
void m17_prbs9_tx_load(uint8_t *out, int len); void m17_prbs9_tx_reset(void);
void m17_prbs9_rx_reset(void); int m17_prbs9_rx_check(uint8_t bit);          // m17_prbs9.cpp:36-64
const uint32_t *m17b_shim_prbs9_state(void);   // {m_rx_state, m_rx_idx, m_rx_bad, m_rx_good, m_rx_eq_cnt, m_rx_dif_cnt, bits, errors}: file statics upstream
void m17_dsp_demap_symbol(float in, float mag, float *out);
void m17_dsp_build_lpf_filter(float *filter, float bw, int ntaps);
void m17_dsp_float_to_short(float *in, int16_t *out, int len);
int  m17_dsp_decimating_filter(float *in, float *out, float *coffs, int stride, int flen, int len);
void m17_rx_parse(float *s, uint8_t type);     // one frame, record handed to the frame callback (no LICH state: see m17_dsp_rx)
// equaliser
void eq_open(void); void eq_reset(void); void eq_restart(void); float eq_train_known(float *in, float train); float eq_train_unknown(float *in);
// TX
uint16_t m17_pack_type(M17Type type);
M17Type m17_upack_type(uint16_t word);                                // m17_bit_utils.cpp:245-254
uint48_t m17_encode_call(const char *call);                           // m17_bit_utils.cpp:191-208 (call padded to 9 characters)
char *m17_decode_call(uint48_t word, char *call);                     // m17_bit_utils.cpp:209-226
void m17_mod_dibits(uint8_t *dibits, int len); void m17_mod_carrier(void);
void m17_send_preamble(void);
void m17_send_link_setup_frame(uint48_t dest, uint48_t src, M17Type type, uint8_t *meta);
void m17_send_stream_frame(uint8_t *payload);
void m17_send_bert_frame(void);
void m17_send_packet_frames(uint8_t *packet, int len);     // writes the CRC into packet[len], packet[len+1] like the reference
void m17_send_eot(void);
void m17_send_carrier(void);

#ifdef M17GISMO_B200_IMPLEMENTATION
#include <cuda_runtime.h>
#include <string.h>
#include <vector>

namespace m17b_shim {
struct State {
    m17b_ctx *ctx = nullptr; m17b_rx *rx = nullptr; m17b_tx *tx = nullptr; m17b_eq *eq = nullptr;
    void *dA = nullptr, *dB = nullptr, *dC = nullptr; size_t cap = 0;
    m17b_shim_callbacks cb = {nullptr, nullptr, nullptr, nullptr};
    int os = 10, err = 0, prbs_idx = 0;
    uint32_t prbs_rx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool lock = false;
    int from_sync = 0;                         // symbols of the last m17_rx_sync_samples call that the fused kernel has already framed
    std::vector<scmplx> txbuf;                 // m_tx_samples: flushed every 1920 samples (m17_modulate.cpp:30-33)
    std::vector<m17b_frame_rec> pend_fr; std::vector<m17b_event_rec> pend_ev;
    int chk(int rc) { if (rc) err = rc; return rc; }
    bool ensure() {
        if (ctx) return true;
        if (chk(m17b_ctx_create(0, &ctx))) return false;
        if (chk(m17b_rx_create(ctx, 1, 1, &rx))) return false;
        if (chk(m17b_eq_create(ctx, 1, &eq))) return false;
        return true;
    }
    bool ensure_tx() { if (!ensure()) return false; if (!tx && chk(m17b_tx_create(ctx, 1, os, &tx))) return false; return true; }
    bool afc = false;
    bool scratch(size_t bytes) {
        if (bytes <= cap) return true;
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
        cap = bytes * 2 + 4096;
        return cudaMalloc(&dA, cap) == cudaSuccess && cudaMalloc(&dB, cap) == cudaSuccess && cudaMalloc(&dC, cap) == cudaSuccess;
    }
    void up(void *d, const void *h, size_t n) { cudaMemcpy(d, h, n, cudaMemcpyHostToDevice); }
    void down(void *h, const void *d, size_t n) { cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost); }
    void collect() {                            // pull the last call's records/events to the host
        m17b_rx_view v;
        if (chk(m17b_rx_get_view(rx, &v))) return;
        int32_t nf = 0, ne = 0;
        down(&nf, v.d_nframes, 4); down(&ne, v.d_nevents, 4);
        size_t f0 = pend_fr.size(), e0 = pend_ev.size();
        pend_fr.resize(f0 + nf); pend_ev.resize(e0 + ne);
        if (nf) down(&pend_fr[f0], v.d_frames, sizeof(m17b_frame_rec) * nf);
        if (ne) down(&pend_ev[e0], v.d_events, sizeof(m17b_event_rec) * ne);
    }
    void dispatch() {
        // events and frames are merged in stream order: an event belongs before the first frame that ends after it
        // (stream indices are counters modulo 2^32 -- they wrap after 10.3 days at 4800 symbols/s -- so positions are compared
        //  through their wrapped difference, never directly)
        size_t e = 0;
        auto rel = [](const m17b_event_rec &ev, const m17b_frame_rec &f) { return (int32_t)((uint32_t)ev.sym_idx - ((uint32_t)f.sym_off + 191u)); };
        for (auto &f : pend_fr) {
            while (e < pend_ev.size() && rel(pend_ev[e], f) <= 0 && !(pend_ev[e].kind == M17B_EV_LOS && rel(pend_ev[e], f) == 0)) fire(pend_ev[e++]);
            if (cb.frame) cb.frame(&f, cb.user);
            while (e < pend_ev.size() && pend_ev[e].kind == M17B_EV_LOS && rel(pend_ev[e], f) == 0) fire(pend_ev[e++]);
        }
        while (e < pend_ev.size()) fire(pend_ev[e++]);
        pend_fr.clear(); pend_ev.clear();
    }
    void fire(const m17b_event_rec &ev) { lock = ev.kind == M17B_EV_AOS; if (cb.event) cb.event(&ev, cb.user); }
    void tx_syms(const uint8_t *syms, int n) {
        if (!ensure_tx() || !scratch((size_t)n * os * 4)) return;
        up(dA, syms, n);
        if (chk(m17b_mod_dibits(tx, (const uint8_t *)dA, n, (int16_t *)dB, nullptr, nullptr))) return;
        size_t o = txbuf.size();
        txbuf.resize(o + (size_t)n * os);
        down(&txbuf[o], dB, (size_t)n * os * 4);
        size_t done = 0;
        while (txbuf.size() - done >= 1920) { if (cb.transmit) cb.transmit(&txbuf[done], 1920, cb.user); done += 1920; }
        txbuf.erase(txbuf.begin(), txbuf.begin() + done);
    }
};
inline State &S() { static State s; return s; }
template <class Tin, class Tout, class F> inline void unary(const Tin *in, size_t nin, Tout *out, size_t nout, F f) {
    State &s = S();
    if (!s.ensure() || !s.scratch(nin * sizeof(Tin) > nout * sizeof(Tout) ? nin * sizeof(Tin) : nout * sizeof(Tout))) return;
    s.up(s.dA, in, nin * sizeof(Tin));
    if (s.chk(f(s, (const Tin *)s.dA, (Tout *)s.dB))) return;
    s.down(out, s.dB, nout * sizeof(Tout));
}
}  // namespace m17b_shim

void m17b_shim_set_callbacks(const m17b_shim_callbacks *cb) { m17b_shim::S().cb = cb ? *cb : m17b_shim_callbacks{nullptr, nullptr, nullptr, nullptr}; }
void m17b_shim_set_oversample(int os) { m17b_shim::S().os = os; }
int  m17b_shim_last_error(void) { return m17b_shim::S().err; }

void m17_prbs9_init(void) { m17b_shim::S().ensure(); m17b_shim::S().prbs_idx = 0; }
void m17_crc_init(void) { m17b_shim::S().ensure(); }
void m17_init_conv(void) { m17b_shim::S().ensure(); }
void m17_init_de_correlate(void) { m17b_shim::S().ensure(); }
void m17_dsp_init(void) { m17b_shim::S().ensure(); }
void m17_fmt_init(void) { m17b_shim::S().ensure(); }
void m17_golay_init(void) { m17b_shim::S().ensure(); }
void m17_rx_sync_init(void) { auto &s = m17b_shim::S(); if (s.ensure()) { s.chk(m17b_rx_reset(s.rx, nullptr)); s.lock = false; } }
void m17_mod_init(void) { m17b_shim::S().ensure_tx(); }
// m17_rx_init / m17_rx_lost (m17_rx_frame.cpp:179-186): reset_sync() + m_flock = false, nothing else
void m17_rx_init(void) { auto &s = m17b_shim::S(); if (s.ensure()) { s.chk(m17b_rx_framer_reset(s.rx, nullptr)); s.lock = false; } }
void m17_rx_lost(void) { m17_rx_init(); }
bool m17_rx_lock(void) { return m17b_shim::S().lock; }
void radio_set_afc_on(void) { auto &s = m17b_shim::S(); if (s.ensure() && !s.chk(m17b_rx_set_afc(s.rx, 1, nullptr))) s.afc = true; }
void radio_set_afc_off(void) { auto &s = m17b_shim::S(); if (s.ensure() && !s.chk(m17b_rx_set_afc(s.rx, 0, nullptr))) s.afc = false; }
bool radio_get_afc_status(void) { return m17b_shim::S().afc; }

void m17_dsp_build_rrc_filter(float *f, float rolloff, int ntaps, int sps) { m17b_build_rrc_filter(f, rolloff, ntaps, sps); }
void m17_dsp_set_filter_gain(float *f, float gain, int stride, int ntaps) { m17b_set_filter_gain(f, gain, stride, ntaps); }

void m17_dsp_rx(scmplx *in, int len) {
    auto &s = m17b_shim::S();
    if (len != M17B_BLOCK_SAMPLES) { s.err = M17B_E_ARG; return; }
    if (!s.ensure() || !s.scratch((size_t)len * 4)) return;
    s.up(s.dA, in, (size_t)len * 4);
    if (s.chk(m17b_dsp_rx(s.rx, (const int16_t *)s.dA, 1, nullptr))) return;
    s.collect();
    s.dispatch();
}
int m17_rx_sync_samples(float *in, float *out, int len) {
    auto &s = m17b_shim::S();
    if (len != M17B_DISC_PER_BLOCK) { s.err = M17B_E_ARG; return 0; }
    if (!s.ensure() || !s.scratch((size_t)len * 4)) return 0;
    s.up(s.dA, in, (size_t)len * 4);
    if (s.chk(m17b_rx_baseband(s.rx, (const float *)s.dA, 1, nullptr))) return 0;
    m17b_rx_view v;
    if (s.chk(m17b_rx_get_view(s.rx, &v))) return 0;
    int32_t n = 0;
    s.down(&n, v.d_nsym, 4);
    s.down(out, v.d_syms + v.sym_carry, sizeof(float) * n);
    s.collect();
    s.from_sync = n;
    return n;
}
void m17_rx_symbols(float *sym, int len) {
    // symbols that m17_rx_sync_samples just returned were framed by the fused kernel already: only the records are dispatched
    // (m17_dsp_rx's own sequence, m17_dsp.cpp:470-473); any other symbols go through the framer on their own (m17b_rx_symbols)
    auto &s = m17b_shim::S();
    if (s.from_sync > 0 && len == s.from_sync) { s.from_sync = 0; s.dispatch(); return; }
    s.from_sync = 0;
    if (len <= 0 || !s.ensure()) return;
    for (int o = 0; o < len; o += 200) {                                 // the batch-1 receiver holds one block: 200 symbols per call
        const int32_t n = len - o < 200 ? len - o : 200;
        if (!s.scratch(1024)) return;
        s.up(s.dA, sym + o, sizeof(float) * n);
        s.up(s.dC, &n, 4);
        if (s.chk(m17b_rx_symbols(s.rx, (const float *)s.dA, 200, (const int32_t *)s.dC, nullptr))) return;
        s.collect();
    }
    s.dispatch();
}
void m17_rx_sym(float sym) { m17_rx_symbols(&sym, 1); }
void m17_dsp_demap_frame(float *in, float *out) {
    m17b_shim::unary<float, float>(in, 192, out, 368, [](m17b_shim::State &s, const float *a, float *b) { return m17b_demap_frame(s.ctx, a, 1, b, nullptr); });
}

static inline int m17b_shim_punc(int p, uint8_t *in, uint8_t *out, int len) {
    int kept = 0;
    m17b_shim::unary<uint8_t, uint8_t>(in, len, out, len, [&](m17b_shim::State &s, const uint8_t *a, uint8_t *b) { return m17b_punc(s.ctx, p, a, len, 1, b, &kept, nullptr); });
    return kept;                                   // in == out is fine: data is staged on the device (m17_tx_routines.cpp:176)
}
int m17_punc_p1(uint8_t *in, uint8_t *out, int len) { return m17b_shim_punc(1, in, out, len); }
int m17_punc_p2(uint8_t *in, uint8_t *out, int len) { return m17b_shim_punc(2, in, out, len); }
int m17_punc_p3(uint8_t *in, uint8_t *out, int len) { return m17b_shim_punc(3, in, out, len); }
static inline int m17b_shim_depunc(int p, float *in, float *out, int len) {
    // number of kept inputs the pattern consumes for `len` outputs
    int kept = 0;
    for (int i = 0; i < len; i++) kept += (p == 1) ? (((i % 61) & 3) != 2) : (p == 2) ? ((i % 12) != 11) : ((i % 8) != 7);
    m17b_shim::unary<float, float>(in, kept, out, len, [&](m17b_shim::State &s, const float *a, float *b) { return m17b_de_punc(s.ctx, p, a, kept, len, 1, b, nullptr); });
    return len;
}
int m17_de_punc_p1(float *in, float *out, int len) { return m17b_shim_depunc(1, in, out, len); }
int m17_de_punc_p2(float *in, float *out, int len) { return m17b_shim_depunc(2, in, out, len); }
int m17_de_punc_p3(float *in, float *out, int len) { return m17b_shim_depunc(3, in, out, len); }
void m17_interleave(uint8_t *in, uint8_t *out, int len) {
    if (len != 368) { m17b_shim::S().err = M17B_E_ARG; return; }
    m17b_shim::unary<uint8_t, uint8_t>(in, 368, out, 368, [](m17b_shim::State &s, const uint8_t *a, uint8_t *b) { return m17b_interleave(s.ctx, a, 1, b, nullptr); });
}
void m17_de_interleave(float *in, float *out, int len) {
    if (len != 368) { m17b_shim::S().err = M17B_E_ARG; return; }
    m17b_shim::unary<float, float>(in, 368, out, 368, [](m17b_shim::State &s, const float *a, float *b) { return m17b_de_interleave(s.ctx, a, 1, b, nullptr); });
}
void m17_de_correlate_8(uint8_t *in, int len) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(len)) return;
    s.up(s.dA, in, len);
    if (s.chk(m17b_de_correlate_8(s.ctx, (uint8_t *)s.dA, len, 1, nullptr))) return;
    s.down(in, s.dA, len);
}
void m17_de_correlate_1(uint8_t *in, uint8_t *out, int len) {
    m17b_shim::unary<uint8_t, uint8_t>(in, len, out, len, [&](m17b_shim::State &s, const uint8_t *a, uint8_t *b) { return m17b_de_correlate_1_u8(s.ctx, a, b, len, 1, nullptr); });
}
void m17_de_correlate_1(float *in, float *out, int len) {
    m17b_shim::unary<float, float>(in, len, out, len, [&](m17b_shim::State &s, const float *a, float *b) { return m17b_de_correlate_1_f32(s.ctx, a, b, len, 1, nullptr); });
}
int m17_conv_encode_1(uint8_t *in, uint8_t *out, int len) {
    m17b_shim::unary<uint8_t, uint8_t>(in, len, out, 2 * (len + 4), [&](m17b_shim::State &s, const uint8_t *a, uint8_t *b) { return m17b_conv_encode_1(s.ctx, a, len, 1, b, nullptr); });
    return 2 * (len + 4);
}
int m17_conv_encode_8(uint8_t *in, uint8_t *out, int len) {
    m17b_shim::unary<uint8_t, uint8_t>(in, len, out, 2 * (8 * len + 4), [&](m17b_shim::State &s, const uint8_t *a, uint8_t *b) { return m17b_conv_encode_8(s.ctx, a, len, 1, b, nullptr); });
    return 2 * (8 * len + 4);
}
int m17_viterbi_decode(float *in, uint8_t *out, int len) {
    m17b_shim::unary<float, uint8_t>(in, len, out, len / 2, [&](m17b_shim::State &s, const float *a, uint8_t *b) { return m17b_viterbi_decode(s.ctx, a, len, 1, b, nullptr); });
    return len / 2;
}
uint24_t m17_golay_encode(uint12_t data) {
    uint32_t w = 0; uint16_t d = data & 0xFFF;
    m17b_shim::unary<uint16_t, uint32_t>(&d, 1, &w, 1, [](m17b_shim::State &s, const uint16_t *a, uint32_t *b) { return m17b_golay_encode(s.ctx, a, 1, b, nullptr); });
    return w;
}
int m_17_golay_decode(uint24_t word, uint12_t &odata) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(16)) return 4;
    s.up(s.dA, &word, 4);
    if (s.chk(m17b_golay_decode(s.ctx, (const uint32_t *)s.dA, 1, (uint16_t *)s.dB, (uint8_t *)s.dC, nullptr))) return 4;
    uint16_t d = 0; uint8_t e = 0;
    s.down(&d, s.dB, 2); s.down(&e, s.dC, 1);
    odata = d;
    return e;
}
uint16_t m17_crc_array_encode(uint8_t *in, int len) {
    uint16_t c = 0xFFFF;
    if (len <= 0) return c;
    m17b_shim::unary<uint8_t, uint16_t>(in, len, &c, 1, [&](m17b_shim::State &s, const uint8_t *a, uint16_t *b) { return m17b_crc_array_encode(s.ctx, a, len, len, 1, b, nullptr); });
    return c;
}
void m17_prbs9_tx_reset(void) { m17b_shim::S().prbs_idx = 0; }
void m17_prbs9_tx_load(uint8_t *out, int len) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(len + 16)) return;
    int32_t st = s.prbs_idx;
    s.up(s.dA, &st, 4);
    if (s.chk(m17b_prbs9_tx_load(s.ctx, (const int32_t *)s.dA, len, 1, (uint8_t *)s.dB, nullptr))) return;
    s.down(out, s.dB, len);
    s.prbs_idx = (s.prbs_idx + len) % 511;
}
void eq_open(void) { auto &s = m17b_shim::S(); if (s.ensure()) { m17b_eq_destroy(s.eq); s.eq = nullptr; s.chk(m17b_eq_create(s.ctx, 1, &s.eq)); } }
void eq_reset(void) { auto &s = m17b_shim::S(); if (s.ensure()) s.chk(m17b_eq_reset(s.eq, nullptr)); }
void eq_restart(void) { auto &s = m17b_shim::S(); if (s.ensure()) s.chk(m17b_eq_restart(s.eq, nullptr)); }
void m17_prbs9_rx_reset(void) { auto &s = m17b_shim::S(); s.prbs_rx[0] = 0; s.prbs_rx[1] = 0; }      // m_rx_idx = 0, NO_SYNC (:36-39)
int m17_prbs9_rx_check(uint8_t bit) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(64)) return 0;
    s.up(s.dA, &bit, 1);
    s.up(s.dC, s.prbs_rx, 32);
    if (s.chk(m17b_prbs9_rx_check(s.ctx, (const uint8_t *)s.dA, 1, 1, (uint32_t *)s.dC, nullptr))) return 0;
    s.down(s.prbs_rx, s.dC, 32);
    return 0;                                                                                          // as upstream (:63)
}
const uint32_t *m17b_shim_prbs9_state(void) { return m17b_shim::S().prbs_rx; }
void m17_dsp_demap_symbol(float in, float mag, float *out) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(16)) return;
    s.up(s.dA, &in, 4); s.up(s.dC, &mag, 4);
    if (s.chk(m17b_demap_symbols(s.ctx, (const float *)s.dA, (const float *)s.dC, 1, (float *)s.dB, nullptr))) return;
    s.down(out, s.dB, 8);
}
void m17_dsp_build_lpf_filter(float *filter, float bw, int ntaps) { m17b_build_lpf_filter(filter, bw, ntaps); }
void m17_dsp_float_to_short(float *in, int16_t *out, int len) { m17b_float_to_short(in, out, len); }
int m17_dsp_decimating_filter(float *in, float *out, float *coffs, int stride, int flen, int len) {
    auto &s = m17b_shim::S();
    const size_t nin = (size_t)len + flen;
    if (len <= 0 || stride <= 0 || flen <= 0 || !s.ensure() || !s.scratch(nin * 4)) return 0;
    int nout = 0;
    s.up(s.dA, in, ((size_t)len + flen - 1) * 4);                 // upstream reads flen-1 samples past len
    s.up(s.dC, coffs, (size_t)flen * 4);
    if (s.chk(m17b_dsp_decimating_filter(s.ctx, (const float *)s.dA, (int64_t)nin, (const float *)s.dC, stride, flen, len, 1, (float *)s.dB, &nout, nullptr))) return 0;
    s.down(out, s.dB, (size_t)nout * 4);
    return nout;
}
void m17_rx_parse(float *sym, uint8_t type) {
    auto &s = m17b_shim::S();
    if (!s.ensure() || !s.scratch(192 * 4)) return;
    s.up(s.dA, sym, 192 * 4); s.up(s.dC, &type, 1);
    if (s.chk(m17b_rx_parse_frames(s.ctx, (const float *)s.dA, (const uint8_t *)s.dC, 1, (m17b_frame_rec *)s.dB, nullptr, nullptr))) return;
    m17b_frame_rec r;
    s.down(&r, s.dB, sizeof(r));
    if (s.cb.frame) s.cb.frame(&r, s.cb.user);
}
static inline float m17b_shim_eq(float *in, const float *train) {
    auto &s = m17b_shim::S();
    float y = 0;
    if (!s.ensure() || !s.scratch(16)) return y;
    s.up(s.dA, in, 8);
    if (train) s.up(s.dC, train, 4);
    if (s.chk(m17b_eq_train(s.eq, (const float *)s.dA, train ? (const float *)s.dC : nullptr, 1, (float *)s.dB, nullptr))) return y;
    s.down(&y, s.dB, 4);
    return y;
}
float eq_train_known(float *in, float train) { return m17b_shim_eq(in, &train); }
float eq_train_unknown(float *in) { return m17b_shim_eq(in, nullptr); }

uint16_t m17_pack_type(M17Type t) {           // m17_bit_utils.cpp:230-244 (host-side helper)
    uint16_t w = t.reserved; w <<= 4; w |= t.can; w <<= 2; w |= t.est; w <<= 2; w |= t.et; w <<= 2; w |= t.dt; w <<= 1; w |= t.p_s;
    return w;
}
M17Type m17_upack_type(uint16_t w) {          // m17_bit_utils.cpp:245-254
    M17Type t;
    t.reserved = (w >> 11) & 0x1F; t.can = (w >> 7) & 0xF; t.est = (w >> 5) & 0x3; t.et = (w >> 3) & 0x3; t.dt = (w >> 1) & 0x3; t.p_s = w & 0x1;
    return t;
}
uint48_t m17_encode_call(const char *call) {  // m17_bit_utils.cpp:191-208: base 40, last character most significant; ' ' and anything else = 0
    uint48_t word = 0;
    for (int i = 8; i >= 0; i--) {
        const char c = call[i];
        word *= 40;
        if (c >= 'A' && c <= 'Z') word += c - 'A' + 1;
        else if (c >= '0' && c <= '9') word += c - '0' + 27;
        else if (c == '-') word += 37;
        else if (c == '/') word += 38;
        else if (c == '.') word += 39;
    }
    return word;
}
char *m17_decode_call(uint48_t word, char *call) {   // m17_bit_utils.cpp:209-226
    if (word == 0xFFFFFFFFFFFFull) { memcpy(call, "BROADCAST", 10); return call; }
    for (int i = 0; i < 9; i++) {
        const int c = (int)(word % 40);
        call[i] = c == 0 ? ' ' : c <= 26 ? (char)('A' + c - 1) : c <= 36 ? (char)('0' + c - 27) : c == 37 ? '-' : c == 38 ? '/' : '.';
        word /= 40;
    }
    call[9] = 0;
    return call;
}
void m17_mod_dibits(uint8_t *dibits, int len) { m17b_shim::S().tx_syms(dibits, len); }
void m17_mod_carrier(void) { uint8_t c[192]; memset(c, 4, sizeof(c)); m17b_shim::S().tx_syms(c, 192); }
void m17_send_carrier(void) { m17_mod_carrier(); }
void m17_send_preamble(void) { uint8_t d[192]; m17b_fmt_preamble(d); m17_mod_dibits(d, 192); }
void m17_send_eot(void) { uint8_t d[192]; m17b_fmt_eot(d); m17_mod_dibits(d, 192); }
void m17_send_link_setup_frame(uint48_t dest, uint48_t src, M17Type type, uint8_t *meta) {
    auto &s = m17b_shim::S();
    if (!s.ensure_tx() || !s.scratch(256)) return;
    uint8_t lsf[30];                               // build_lich, m17_tx_routines.cpp:37-53
    for (int i = 0; i < 6; i++) { lsf[i] = (uint8_t)(dest >> (40 - 8 * i)); lsf[6 + i] = (uint8_t)(src >> (40 - 8 * i)); }
    uint16_t tw = m17_pack_type(type);
    lsf[12] = (uint8_t)(tw >> 8); lsf[13] = (uint8_t)tw;
    memcpy(lsf + 14, meta, 14);
    uint16_t crc = m17_crc_array_encode(lsf, 28);
    lsf[28] = (uint8_t)(crc >> 8); lsf[29] = (uint8_t)crc;
    s.up(s.dA, lsf, 30);
    if (s.chk(m17b_tx_set_lsf(s.tx, (const uint8_t *)s.dA, nullptr))) return;     // also m_lich_count = m_fn = 0 (:98-99)
    if (s.chk(m17b_fmt_link_setup_frame(s.ctx, (const uint8_t *)s.dA, 1, (uint8_t *)s.dB, nullptr))) return;
    uint8_t d[192];
    s.down(d, s.dB, 192);
    m17_mod_dibits(d, 192);
}
void m17_send_stream_frame(uint8_t *payload) {
    auto &s = m17b_shim::S();
    if (!s.ensure_tx() || !s.scratch(256)) return;
    s.up(s.dA, payload, 16);
    if (s.chk(m17b_fmt_stream_frames(s.tx, (const uint8_t *)s.dA, 1, (uint8_t *)s.dB, nullptr))) return;
    uint8_t d[192];
    s.down(d, s.dB, 192);
    m17_mod_dibits(d, 192);
}
void m17_send_packet_frames(uint8_t *packet, int len) {      // m17_tx_routines.cpp:323-353
    auto &s = m17b_shim::S();
    if (len < 0 || len > 798) { s.err = M17B_E_ARG; return; }
    if (!s.ensure_tx() || !s.scratch(32 * 192 + 1024)) return;
    uint16_t crc = m17_crc_array_encode(packet, len);
    packet[len] = (uint8_t)(crc >> 8); packet[len + 1] = (uint8_t)crc;      // the caller's buffer must be 2 bytes longer (:325)
    int32_t l32 = len, nf = 0;
    s.up(s.dA, packet, (size_t)len);
    s.up((uint8_t *)s.dC, &l32, 4);
    if (s.chk(m17b_send_packet_frames(s.ctx, (const uint8_t *)s.dA, 0, (const int32_t *)s.dC, 1, 32, (uint8_t *)s.dB, (int32_t *)s.dC + 1, nullptr))) return;
    s.down(&nf, (int32_t *)s.dC + 1, 4);
    std::vector<uint8_t> d((size_t)nf * 192);
    s.down(d.data(), s.dB, d.size());
    for (int f = 0; f < nf; f++) m17_mod_dibits(&d[(size_t)f * 192], 192);
}
void m17_send_bert_frame(void) {
    auto &s = m17b_shim::S();
    if (!s.ensure_tx() || !s.scratch(256)) return;
    if (s.chk(m17b_fmt_bert_frames(s.tx, 1, (uint8_t *)s.dB, nullptr))) return;
    uint8_t d[192];
    s.down(d, s.dB, 192);
    m17_mod_dibits(d, 192);
}
#endif  // M17GISMO_B200_IMPLEMENTATION
#endif  // M17GISMO_B200_HPP
