"""CPU probe (oracle only, no GPU): what happens when m17_equalize.cpp is wired into the live receive chain literally
(SURVEY 8f rank 3).  The reference has no call site for eq_* (m17_equalize.cpp:163-213); the oracle's seam flag 64 inserts
eq_open() + eq_train_unknown() on (half-symbol, symbol) pairs of the matched filter between m17_rx_sync_samples and m17_rx_symbols
(m17_rx_sync.cpp:77 -> m17_rx_frame.cpp:173; the half-symbol output is the same polyphase branch one sample earlier).
Prints, per channel class, stream frames delivered with the exact payload with the equaliser off / on, and the symbol scale.
usage: python tests/eq_live_chain_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))  # the oracle is test infrastructure: this probe lives under tests/
import m17_oracles as O, signals

P = O.Port()
classes = [None, 30, 26, 22]
X, pl = signals.stream_channels(P, 16, 40, 5, ebn0=[classes[c % 4] for c in range(16)], f0_max=500.0)
Cn, T = X.shape[0], X.shape[1] // O.BLOCK


def run(flags):
    o = O._alloc_rx_out(Cn, T, True, False, np.zeros)
    P.L.m17o_rx_run(O._p(X), flags, Cn, T, 8, O._p(o.disc), O._p(o.nsym), O._p(o.syms), o.symcap, O._p(o.frames), o.fcap,
                    O._p(o.soft), O._p(o.events), o.ecap, O._p(o.counts))
    return o


def exact(o, c):
    n = int(o.counts[c, 2])
    f = o.frames[c, :n]
    sent = {bytes(p) for p in pl[c]}
    return sum(1 for r in f if (r["flags"] & 0x08) and bytes(r["data"][2:18]) in sent)


off, on = run(0), run(64)
for k, e in enumerate(classes):
    ch = [c for c in range(Cn) if c % 4 == k]
    print("Eb/N0(IQ) %s dB: exact payloads of %d sent  off %d  on %d" % ("inf" if e is None else e, 40 * len(ch),
          sum(exact(off, c) for c in ch), sum(exact(on, c) for c in ch)))
s_off = off.syms[0, 2000:4000]
s_on = on.syms[0, 2000:4000]
print("symbol scale (clean channel, symbols 2000..4000): off max|s| %.4f   on max|s| %.3f, median|s| %.3f" % (
      np.abs(s_off).max(), np.abs(s_on).max(), np.median(np.abs(s_on))))
print("decision levels of eq_train_unknown: +-1.0 / +-0.333 (m17_equalize.cpp:195-205); outer symbol level of the chain: %.4f" % np.percentile(np.abs(s_off), 95))
