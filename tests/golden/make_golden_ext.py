#!/usr/bin/env python
"""Generate tests/golden/m17_golden_ext.npz: outputs of the UNMODIFIED reference for the pieces added around the hot path
(AFC loop, Pluto /8 decimator, M17-over-UDP frames, PRBS9 receive checker).  Build container only (needs oracle/_ref):
    python tests/golden/make_golden_ext.py
Everything stored is an output of the reference's own code on the stored, seeded inputs."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from m17_oracles import ORACLE_DIR, Port, Ref, RefRadio, _p  # noqa: E402


def main():
    P, R, RR = Port(), Ref(), RefRadio()
    rng = np.random.default_rng(20261019)
    g = {}
    # ---- AFC on: the golden IQ capture of m17_golden.npz through the reference with radio_set_afc_on()
    G = np.load(os.path.join(HERE, "m17_golden.npz"))
    X = np.ascontiguousarray(G["rx_iq"])
    o = R.rx_run(X, seam=0, afc=True)
    ns = int(o.counts[:, 1].max()); nf = int(o.counts[:, 2].max()); ne = int(o.counts[:, 3].max())
    g["afc_nsym"], g["afc_counts"] = np.array(o.nsym), np.array(o.counts)
    g["afc_syms"] = np.array(o.syms[:, :ns])
    g["afc_frames"] = np.array(o.frames[:, :max(nf, 1)]).view(np.uint8).reshape(X.shape[0], -1, 64)
    g["afc_events"] = np.array(o.events[:, :max(ne, 1)]).view(np.int32).reshape(X.shape[0], -1, 2)
    g["afc_disc_c1"] = np.array(o.disc[1])
    # ---- Pluto /8 decimator: two 1920-sample output blocks, random + extremes
    D = rng.integers(-32768, 32768, (2, 2 * 8 * 1920, 2)).astype(np.int16)
    D[1, :4000] = 32767; D[1, 4000:8000] = -32768
    g["dec_in"], g["dec_out"] = D, RR.pluto_run(D)
    # ---- M17-over-UDP frames
    n = 12
    lsf = np.stack([P.build_lsf(int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 16)),
                                rng.integers(0, 256, 14, dtype=np.uint8)) for _ in range(n)])
    sid = rng.integers(0, 1 << 16, n).astype(np.uint16); fn = rng.integers(0, 1 << 16, n).astype(np.uint16)
    pld = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    g["udp_lsf"], g["udp_sid"], g["udp_fn"], g["udp_pld"] = lsf, sid, fn, pld
    g["udp_frames"] = np.stack([RR.net_rx_data(int(sid[i]), lsf[i], int(fn[i]), pld[i]) for i in range(n)])   # dst = encode_call(" ")
    g["udp_lich"] = np.stack([RR.lich_from_net(f) for f in g["udp_frames"]])
    # ---- PRBS9 receive checker
    seq = np.tile(P.prbs9(), 5)
    bits = seq[:2400].copy(); bits[rng.integers(200, 2400, 25)] ^= 1; bits[1200:1240] ^= 1
    L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libm17ref_prbs.so"))
    L.refp_init(); L.refp_check(_p(bits), C.c_long(len(bits)))
    st = np.zeros(6, np.uint32); L.refp_state(_p(st))
    g["prbs_bits"], g["prbs_state"] = bits, st
    # ---- m17_dsp_demap_symbol / m17_dsp_decimating_filter (m17_dsp.cpp:35-42,438-449)
    sy = rng.normal(0, 1, 500).astype(np.float32); mg = rng.uniform(0.2, 4, 500).astype(np.float32)
    g["dsym_in"], g["dsym_mag"], g["dsym_out"] = sy, mg, R.demap_symbol(sy, mg)
    x = rng.normal(0, 1, (3, 230)).astype(np.float32); cf = rng.normal(0, 0.2, 31).astype(np.float32)
    g["dfil_in"], g["dfil_coffs"] = x, cf
    g["dfil_out"] = np.stack([R.decimating_filter(r, cf, 5, 200) for r in x])
    # ---- gps_decode (gps.cpp:8-27) on random LSFs
    gl = rng.integers(0, 256, (40, 30), dtype=np.uint8)
    g["gps_lsf"] = gl
    g["gps_out"] = np.array([R.gps_decode(l) for l in gl], np.float64)
    out = os.path.join(HERE, "m17_golden_ext.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes; afc frames", o.counts[:, 2])


if __name__ == "__main__":
    main()
