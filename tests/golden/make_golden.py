#!/usr/bin/env python
"""Generate tests/golden/m17_golden.npz from the UNMODIFIED reference objects (oracle/_ref/libm17ref.so).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle`):
    python tests/golden/make_golden.py
The fixture travels with the repo; the GPU box and the CPU test tier read it and never touch /root/reference.
Everything in the file is an OUTPUT OF THE REFERENCE'S OWN CODE on seeded inputs (the inputs are stored too).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from m17_oracles import Port, Ref, add_iq_noise, delay_iq, rotate_iq  # noqa: E402
import signals  # noqa: E402


def main():
    R = Ref()
    P = Port()          # only used as a signal generator for RX inputs (its TX is itself pinned below against R)
    rng = np.random.default_rng(20261018)
    g = {}
    # ---- known-answer values
    g["crc_in_lens"] = np.array([0, 1, 9, 256, 30], np.int32)
    crc_msgs = [b"", b"A", b"123456789", bytes(range(256)), bytes(rng.integers(0, 256, 30, dtype=np.uint8))]
    g["crc_msg4"] = np.frombuffer(crc_msgs[4], np.uint8)
    g["crc_out"] = np.array([R.crc(m) for m in crc_msgs], np.uint16)
    g["golay_enc"] = np.array([R.golay_encode(d) for d in range(4096)], np.uint32)
    g["golay_errtab"] = R.golay_errtab()
    w = rng.integers(0, 1 << 24, 512).astype(np.uint32)
    g["golay_dec_in"] = w
    dec = [R.golay_decode(int(x)) for x in w]
    g["golay_dec_data"] = np.array([d[0] for d in dec], np.uint16)
    g["golay_dec_err"] = np.array([d[1] for d in dec], np.uint8)
    for nb in (18, 26, 30):
        d = rng.integers(0, 256, (4, nb), dtype=np.uint8)
        g[f"conv8_in_{nb}"] = d
        g[f"conv8_out_{nb}"] = np.stack([R.conv_encode_8(r) for r in d])
    b = rng.integers(0, 2, (2, 197), dtype=np.uint8)
    g["conv1_in"] = b
    g["conv1_out"] = np.stack([R.conv_encode_1(r) for r in b])
    for ln in (296, 420, 488):
        s = rng.normal(0, 1, (12, ln)).astype(np.float32)
        s[0] = 0
        s[1] = np.round(s[1] * 4) / 4
        g[f"vit_in_{ln}"] = s
        g[f"vit_out_{ln}"] = np.stack([R.viterbi(r) for r in s])
    for p, ln in ((1, 488), (2, 296), (3, 420)):
        bits = rng.integers(0, 2, ln, dtype=np.uint8)
        g[f"punc_in_{p}"] = bits
        g[f"punc_out_{p}"] = R.punc(p, bits)
        s = rng.normal(0, 1, len(g[f"punc_out_{p}"])).astype(np.float32)
        g[f"depunc_in_{p}"] = s
        g[f"depunc_out_{p}"] = R.depunc(p, s, ln)
    bits = rng.integers(0, 2, 368, dtype=np.uint8)
    soft = rng.normal(0, 1, 368).astype(np.float32)
    g["bits368"], g["soft368"] = bits, soft
    g["interleave_out"] = R.interleave(bits)
    g["deinterleave_out"] = R.deinterleave(soft)
    g["derand_bits_out"] = R.derand_bits(bits)
    g["derand_soft_out"] = R.derand_soft(soft)
    g["derand_bytes_out"] = R.derand_bytes(np.zeros(46, np.uint8))
    sy = (rng.normal(0, 1, (6, 192)) * rng.uniform(0.2, 2, (6, 1))).astype(np.float32)
    g["demap_in"] = sy
    g["demap_out"] = np.stack([R.demap_frame(r) for r in sy])
    v = rng.normal(0, 1, (96, 8)).astype(np.float32)
    tpl = np.array([[1, 1, 1, 1, -1, -1, 1, -1], [-1, -1, -1, -1, 1, 1, -1, 1], [1, -1, 1, 1, -1, -1, -1, -1], [-1, 1, -1, -1, 1, 1, 1, 1],
                    [1, 1, 1, 1, 1, 1, -1, 1], [1, -1, 1, -1, 1, -1, 1, -1]], np.float32)
    v[:64] = tpl[rng.integers(0, 6, 64)] * (1 + 0.2 * rng.normal(0, 1, (64, 8))).astype(np.float32)
    v[3] = 0
    sc = [R.sync_check(r) for r in v]
    g["sync_in"] = v
    g["sync_type"] = np.array([s[0] for s in sc], np.uint8)
    g["sync_votes"] = np.array([s[1] for s in sc], np.uint8)
    g["sync_var"] = np.array([s[2] for s in sc], np.float32)
    g["rrc_1240_80"] = R.rrc(0.5, 1240, 80)
    g["rrc_310_10"] = R.set_gain(R.rrc(0.5, 310, 10), 10, 1, 310)
    g["rrc_62_2"] = R.set_gain(R.rrc(0.5, 62, 2), 1.0, 1, 62)
    g["prbs9"] = R.prbs9(511)
    calls = ["G4GUO    ", "G4GUO/P  ", "AB1CD-9 .", "M17      "]
    g["call_enc"] = np.array([R.encode_call(c) for c in calls], np.uint64)
    # ---- equaliser (fresh process-global state: this script is the only user)
    lv = np.array([1.0, 1 / 3, -1 / 3, -1.0], np.float32)
    tr = lv[rng.integers(0, 4, 300)]
    x = np.convolve(np.repeat(tr, 2).astype(np.float64), [0.15, 0.8, 0.25, -0.1])[:600] + rng.normal(0, 0.02, 600)
    pairs = x.reshape(300, 2).astype(np.float32)
    R.L.ref_eq_open()
    import ctypes as C
    y = np.zeros(300, np.float32)
    for i in range(300):
        pp = np.ascontiguousarray(pairs[i])
        y[i] = R.L.ref_eq_train_known(pp.ctypes.data_as(C.c_void_p), float(tr[i])) if i < 120 else R.L.ref_eq_train_unknown(pp.ctypes.data_as(C.c_void_p))
    g["eq_pairs"], g["eq_train"], g["eq_out"] = pairs, tr, y
    # ---- TX: frame formatters and the modulator
    lsf = R.build_lsf(0xFFFFFFFFFFFF, R.encode_call("G4GUO    "), 0x0005)
    g["lsf"] = lsf
    g["dibits_lsf"] = R.fmt_lsf(lsf)
    pl = rng.integers(0, 256, (8, 16), dtype=np.uint8)
    pl[0] = np.arange(16)
    g["stream_payload"] = pl
    g["dibits_stream"] = R.fmt_stream_frames(lsf, pl)
    chunk = bytes(range(1, 0x3B, 3))
    g["packet_chunk"] = np.frombuffer(chunk, np.uint8)
    g["dibits_packet"] = R.fmt_packet(chunk, 1, 22)
    g["dibits_bert"] = R.fmt_bert(3)
    g["dibits_preamble"], g["dibits_eot"] = R.fmt_preamble(), R.fmt_eot()
    iq, dib = R.tx_stream_run(lsf[None], pl[None, :2], lead=1, npre=1, tail=1, nproc=1)
    g["tx_iq"] = iq[0]
    g["tx_dibits"] = dib[0]
    # ---- RX chain from int16 IQ: clean+delay, 24 dB + offset, 21 dB + offset, noise only
    F = 8
    plr = rng.integers(0, 256, (3, F, 16), dtype=np.uint8)
    iqs, _ = R.tx_stream_run(np.repeat(lsf[None], 3, 0), plr, lead=1, npre=2, tail=2, nproc=3)
    T = iqs.shape[1] // 1920 + 1
    X = np.zeros((4, T * 1920, 2), np.int16)
    X[0] = delay_iq(iqs[0], 777, T * 1920)
    X[1] = add_iq_noise(rotate_iq(delay_iq(iqs[1], 1234, T * 1920), 812.5), 24.0, rng)
    X[2] = add_iq_noise(rotate_iq(delay_iq(iqs[2], 45, T * 1920), -640.0), 21.0, rng)
    X[3] = add_iq_noise(delay_iq(iqs[0][:0], 0, T * 1920), 3.0, rng)
    o = R.rx_run(X, seam=0, want_soft=False)
    g["rx_iq"], g["rx_payload"] = X, plr
    g["rx_nsym"], g["rx_counts"] = np.array(o.nsym), np.array(o.counts)
    ns = int(o.counts[:, 1].max()); nf = int(o.counts[:, 2].max()); ne = int(o.counts[:, 3].max())
    g["rx_syms"] = np.array(o.syms[:, :ns])
    g["rx_frames"] = np.array(o.frames[:, :max(nf, 1)]).view(np.uint8).reshape(4, -1, 64)
    g["rx_events"] = np.array(o.events[:, :max(ne, 1)]).view(np.int32).reshape(4, -1, 2)
    g["rx_disc_c0"] = np.array(o.disc[0])
    # ---- baseband seam, Eb/N0 12 / 8 / 4 dB (the seam where the 0..12 dB sweep is meaningful)
    D, plb = signals.baseband_channels(P, 3, 8, 77, [12.0, 8.0, 4.0])
    ob = R.rx_run(D, seam=1)
    g["bb_disc"], g["bb_payload"] = D, plb
    g["bb_nsym"], g["bb_counts"] = np.array(ob.nsym), np.array(ob.counts)
    ns = int(ob.counts[:, 1].max()); nf = int(ob.counts[:, 2].max()); ne = int(ob.counts[:, 3].max())
    g["bb_syms"] = np.array(ob.syms[:, :ns])
    g["bb_frames"] = np.array(ob.frames[:, :max(nf, 1)]).view(np.uint8).reshape(3, -1, 64)
    g["bb_events"] = np.array(ob.events[:, :max(ne, 1)]).view(np.int32).reshape(3, -1, 2)
    out = os.path.join(HERE, "m17_golden.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes;", "rx frames per channel", o.counts[:, 2], "bb frames", ob.counts[:, 2])


if __name__ == "__main__":
    main()
