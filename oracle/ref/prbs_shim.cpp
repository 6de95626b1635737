/* oracle/ref/prbs_shim.cpp -- TEST INFRASTRUCTURE ONLY.  The reference's PRBS9 receive checker keeps its verdict in file statics
 * with no accessor (m17_prbs9.cpp:7-12), so this translation unit INCLUDES the reference source in place (compile-time inclusion
 * from $(REF), nothing is copied into this repository) and exports the statics.  Built into oracle/_ref/libm17ref_prbs.so. */
#include "m17_prbs9.cpp"
extern "C" void refp_init(void) { m17_prbs9_init(); }
extern "C" void refp_check(const uint8_t *bits, long n) { for (long i = 0; i < n; i++) m17_prbs9_rx_check(bits[i]); }
extern "C" void refp_state(unsigned *o) { o[0] = (unsigned)m_rx_state; o[1] = m_rx_idx; o[2] = m_rx_bad; o[3] = m_rx_good; o[4] = m_rx_eq_cnt; o[5] = m_rx_dif_cnt; }
extern "C" void refp_tx(uint8_t *out, int len) { m17_prbs9_tx_load(out, len); }
