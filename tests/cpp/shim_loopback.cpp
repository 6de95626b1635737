// shim_loopback.cpp -- the reference's own call sequence, unchanged names, running on the GPU through
// include/m17gismo_b200.hpp: init chain (main.cpp:108-126), one "over" sent with m17_send_* (m17_tx_rx.cpp:95-115),
// the transmitted IQ looped straight back into m17_dsp_rx block by block (what m17_test.cpp:42-52 does at baseband),
// plus the known-answer values of SURVEY.md section 4 on the primitive entry points.
#define M17GISMO_B200_IMPLEMENTATION
#include "m17gismo_b200.hpp"
#include <stdio.h>
#include <string.h>
#include <vector>

static std::vector<scmplx> g_iq;
static std::vector<m17b_frame_rec> g_frames;
static int g_aos = 0, g_los = 0, g_fail = 0;
static void on_tx(const scmplx *iq, uint32_t n, void *) { g_iq.insert(g_iq.end(), iq, iq + n); }
static void on_frame(const m17b_frame_rec *r, void *) { g_frames.push_back(*r); }
static void on_event(const m17b_event_rec *e, void *) { if (e->kind == M17B_EV_AOS) g_aos++; else g_los++; }
#define EXPECT(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); g_fail++; } } while (0)

int main() {
    m17b_shim_callbacks cb = {on_frame, on_event, on_tx, nullptr};
    m17b_shim_set_callbacks(&cb);
    m17_prbs9_init(); m17_crc_init(); m17_init_conv(); m17_init_de_correlate(); m17_dsp_init(); m17_fmt_init();
    m17_golay_init(); m17_rx_sync_init(); m17_mod_init();
    if (m17b_shim_last_error()) { printf("init failed: %d (%s)\n", m17b_shim_last_error(), m17b_last_cuda_error()); return 2; }

    // ---- known answers
    uint8_t msg[] = "123456789";
    EXPECT(m17_crc_array_encode(msg, 9) == 0x772B);
    uint8_t a1[] = "A";
    EXPECT(m17_crc_array_encode(a1, 1) == 0x206E);
    EXPECT(m17_golay_encode(0xABC) == 0xABC23C && m17_golay_encode(0x001) == 0x0018EB);
    uint12_t od = 0;
    EXPECT(m_17_golay_decode(0xABC23C ^ 0x111000, od) == 3 && od == 0xABC);
    EXPECT(m_17_golay_decode(0xABC23C ^ 0x00F000, od) == 4 && od == 0x0F3);
    uint8_t bits8[8] = {1, 0, 1, 1, 0, 0, 1, 0}, coded[64];
    EXPECT(m17_conv_encode_1(bits8, coded, 8) == 24);
    const char *kat = "110110001111101001101100";
    for (int i = 0; i < 24; i++) EXPECT(coded[i] == kat[i] - '0');
    float zeros[296] = {0}; uint8_t vb[148];
    EXPECT(m17_viterbi_decode(zeros, vb, 296) == 148);
    int ones = 0; for (int i = 0; i < 148; i++) ones += vb[i];
    EXPECT(ones == 144 && vb[0] == 0 && vb[1] == 1 && vb[147] == 0);
    uint8_t perm[368], back[368], src[368];
    for (int i = 0; i < 368; i++) src[i] = (uint8_t)(i * 7);
    m17_interleave(src, perm, 368); m17_interleave(perm, back, 368);
    for (int i = 0; i < 368; i++) EXPECT(back[i] == src[i]);                 // QPP is an involution
    uint8_t prbs[16]; m17_prbs9_tx_reset(); m17_prbs9_tx_load(prbs, 16);
    const char *pk = "0000100011000010";
    for (int i = 0; i < 16; i++) EXPECT(prbs[i] == pk[i] - '0');

    // ---- the smaller entry points of m17defines.h
    float so2[2];
    m17_dsp_demap_symbol(0.5f, 2.0f, so2);                                   // m = 1: MSB soft -1, LSB soft 1 - 0.6666
    EXPECT(so2[0] == -1.0f && so2[1] == (float)(1.0 - 0.6666));
    float lpf[31]; int16_t lpq[31];
    m17_dsp_build_lpf_filter(lpf, 0.125f, 31);
    EXPECT(lpf[15] == 0.25f && lpf[14] == lpf[16] && lpf[0] == lpf[30]);
    m17_dsp_float_to_short(lpf, lpq, 31);
    EXPECT(lpq[15] == (int16_t)(0.25f * 0x7FFF));
    float din[40], dco[4] = {1.0f, 2.0f, 3.0f, 4.0f}, dout[16];
    for (int i = 0; i < 40; i++) din[i] = (float)i;
    EXPECT(m17_dsp_decimating_filter(din, dout, dco, 4, 4, 32) == 8);
    for (int k = 0; k < 8; k++) EXPECT(dout[k] == (float)(4 * k) + 2.0f * (4 * k + 1) + 3.0f * (4 * k + 2) + 4.0f * (4 * k + 3));
    uint8_t pr[64]; m17_prbs9_tx_reset(); m17_prbs9_tx_load(pr, 64); m17_prbs9_rx_reset();
    for (int i = 0; i < 64; i++) m17_prbs9_rx_check(pr[i]);
    EXPECT(m17b_shim_prbs9_state()[0] == 1 && m17b_shim_prbs9_state()[1] == 64 && m17b_shim_prbs9_state()[7] == 0);   // in sync, no errors
    m17_prbs9_tx_reset();
    eq_open(); eq_restart(); eq_reset();
    EXPECT(m17b_shim_last_error() == 0);
    // host-side helpers of m17_bit_utils.cpp:191-254
    char call[10];
    EXPECT(m17_encode_call("G4GUO    ") == 0x0000025EA29Full);               // the source address used below
    EXPECT(!strcmp(m17_decode_call(m17_encode_call("AB1CD/P-."), call), "AB1CD/P-."));
    EXPECT(!strcmp(m17_decode_call(0xFFFFFFFFFFFFull, call), "BROADCAST"));
    M17Type t0 = {1, 2, 1, 3, 9, 17};
    M17Type t1 = m17_upack_type(m17_pack_type(t0));
    EXPECT(t1.p_s == 1 && t1.dt == 2 && t1.et == 1 && t1.est == 3 && t1.can == 9 && t1.reserved == 17);

    // ---- one over, looped back
    const int F = 14;
    uint8_t meta[14] = {0}, payload[F][16];
    M17Type ty = {1, 2, 0, 0, 0, 0};                                         // stream, voice -> TYPE 0x0005
    EXPECT(m17_pack_type(ty) == 0x0005);
    m17_send_carrier(); m17_send_preamble(); m17_send_preamble();
    m17_send_link_setup_frame(0xFFFFFFFFFFFFull, 0x0000025EA29Full, ty, meta);
    for (int f = 0; f < F; f++) { for (int i = 0; i < 16; i++) payload[f][i] = (uint8_t)(f * 16 + i); m17_send_stream_frame(payload[f]); }
    m17_send_eot(); m17_send_carrier(); m17_send_carrier();
    EXPECT(g_iq.size() == (size_t)(F + 7) * 1920);
    for (size_t b = 0; b + 1920 <= g_iq.size(); b += 1920) m17_dsp_rx(&g_iq[b], 1920);
    int stream = 0, delivered = 0, exact = 0, lsf_ok = 0;
    for (auto &r : g_frames) {
        if (r.type == M17B_T_LSF && r.crc == 0) lsf_ok++;
        if (r.type != M17B_T_STREAM) continue;
        stream++;
        if (!(r.flags & M17B_F_DELIVERED)) continue;
        delivered++;
        int fn = (r.data[0] << 8) | r.data[1];
        if (fn < F && !memcmp(r.data + 2, payload[fn], 16)) exact++;
    }
    EXPECT(g_aos == 1 && g_los == 1 && !m17_rx_lock());
    EXPECT(lsf_ok == 1 && stream == F && delivered >= F - 6 && exact == delivered);
    // ---- m17_rx_lost() only clears the sync window and the lock flag (m17_rx_frame.cpp:179-186); m17_rx_symbols / m17_rx_sym on
    //      symbols that did not come from m17_rx_sync_samples run the stand-alone framer: feed the over's own symbol stream again
    {
        std::vector<float> sy;
        std::vector<float> disc(384), outsym(400);
        size_t nf0 = g_frames.size(); int aos0 = g_aos;
        // a clean sync word + payload at the symbol seam, then 184 symbols
        const float sw[8] = {1, 1, 1, 1, -1, -1, 1, -1};
        for (int f = 0; f < 3; f++) { for (int k = 0; k < 8; k++) sy.push_back(sw[k]); for (int k = 0; k < 184; k++) sy.push_back(((k * 7 + f) & 1) ? 0.333f : -1.0f); }
        m17_rx_lost();
        EXPECT(!m17_rx_lock());
        m17_rx_symbols(sy.data(), 192 + 100);                                // ragged pieces through the symbol seam
        for (int k = 292; k < 300; k++) m17_rx_sym(sy[k]);
        m17_rx_symbols(sy.data() + 300, (int)sy.size() - 300);
        EXPECT(g_aos == aos0 + 1 && m17_rx_lock());
        // (the word above is the LSF sync word 0x55F7 = +3 +3 +3 +3 -3 -3 +3 -3; the third frame completes on the last symbol)
        EXPECT(g_frames.size() == nf0 + 3 && g_frames[nf0].type == M17B_T_LSF && g_frames.back().type == M17B_T_LSF);
        if (g_frames.size() != nf0 + 3) printf("symbol seam: %zu new frames, first type %d sym_off %d\n", g_frames.size() - nf0, g_frames.size() > nf0 ? g_frames[nf0].type : -1, g_frames.size() > nf0 ? g_frames[nf0].sym_off : -1);
        m17_rx_lost();
        EXPECT(m17b_shim_last_error() == 0);
    }
    printf("{\"shim_loopback\": \"%s\", \"frames\": %zu, \"stream\": %d, \"delivered\": %d, \"exact\": %d, \"aos\": %d, \"los\": %d, \"err\": %d}\n",
           g_fail ? "FAIL" : "PASS", g_frames.size(), stream, delivered, exact, g_aos, g_los, m17b_shim_last_error());
    return g_fail || m17b_shim_last_error() ? 1 : 0;
}
