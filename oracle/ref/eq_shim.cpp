/* oracle/ref/eq_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Pins the oracle's equaliser probe (m17_oracle.c, seam flag 64; SURVEY 8f
 * rank 3) to the reference's own functions.  m17_equalize.cpp has no call site upstream, and the T/2-spaced input it wants is the
 * matched filter's output between symbol instants, which m17_rx_sync_samples never computes.  This translation unit INCLUDES
 * m17_rx_sync.cpp in place (compile-time inclusion from $(REF), nothing is copied into this repository) to reach its file statics,
 * feeds the unmodified m17_rx_sync_samples ONE sample per call, and after every vote sample evaluates the reference's own
 * rx_sync_filter(m_buff, m_mf[m_index], FN) -- the branch the next symbol will use.  eq_train_unknown (m17_equalize.cpp:185-213)
 * then runs on every (half-symbol, symbol) pair and its output goes to m17_rx_symbols (m17_rx_frame.cpp:173).
 * Built into oracle/_ref/libm17ref_eq.so together with ref_shim.cpp (trace taps) and the other reference objects. */
#include "m17_rx_sync.cpp"

static float g_mid = 0.0f;
extern "C" void refe_open(void) { eq_open(); g_mid = 0.0f; }
/* one 384-sample block of 2-sps discriminator samples; returns the number of symbols handed to the framer */
extern "C" int refe_block(float *in, int len) {
    float out[964], mid[964], one[4];
    float *o = out + 2, *m = mid + 2;
    int idx = 0;                                   /* m17_rx_sync_samples' own m_idx over a whole block (may reach -1) */
    for (int i = 0; i < len; i++) {
        const bool symbol_sample = ((m_clk + 1) % 2) == 1;
        const int r = m17_rx_sync_samples(&in[i], one, 1);
        if (symbol_sample) {                       /* r == 1: out[0] = sum */
            if (idx >= 0) { o[idx] = one[0]; m[idx] = g_mid; }
            idx++;
        } else {
            if (r == 1) { if (idx >= 0) { o[idx] = 0.0f; m[idx] = 0.0f; } idx++; }      /* forward bit slip: a zero symbol (:56-58) */
            if (r == -1) idx--;                                                          /* backward bit slip (:68-70) */
            g_mid = rx_sync_filter(m_buff, m_mf[m_index], FN);
        }
    }
    const int n = idx < 0 ? 0 : idx;
    for (int q = 0; q < n; q++) { float in2[2] = { m[q], o[q] }; o[q] = eq_train_unknown(in2); }
    m17_rx_symbols(o, n);
    return n;
}
