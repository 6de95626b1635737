"""GPU parity checks of libm17b200 against the CPU oracle (oracle/libm17oracle.so).

Used two ways: imported by the pytest -m gpu tests (each check raises AssertionError on a mismatch), and run
directly (`python tests/gpu_check.py`) for a verbose report while bringing kernels up on the GPU box.
Every call goes through the C ABI via m17_sdr_b200.api; the oracle is only the checker.
"""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from m17_oracles import F_DELIVERED, REC_DTYPE, Port, lsf_for  # noqa: E402
import signals  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits_eq(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def first_diff(a, b):
    a = np.asarray(a).reshape(-1); b = np.asarray(b).reshape(-1)
    n = min(len(a), len(b))
    d = np.nonzero(a[:n].view(np.uint32 if a.dtype == np.float32 else a.dtype) != b[:n].view(np.uint32 if b.dtype == np.float32 else b.dtype))[0]
    return None if len(d) == 0 else (int(d[0]), a[d[0]], b[d[0]], len(d))


# ---------------------------------------------------------------------------------------------- primitives
def check_primitives(ctx, P, seed=11, n=64):
    rng = np.random.default_rng(seed)
    # CRC
    for ln in (1, 2, 9, 30, 52, 256):
        d = rng.integers(0, 256, (n, ln), dtype=np.uint8)
        got = ctx.m17_crc_array_encode(dev(d)).cpu().numpy()
        exp = np.array([P.crc(bytes(r)) for r in d], np.uint16)
        assert np.array_equal(got, exp), ("crc", ln)
    # Golay
    d12 = np.arange(4096, dtype=np.uint16)
    enc = ctx.m17_golay_encode(dev(d12)).cpu().numpy().astype(np.uint32)
    assert np.array_equal(enc, np.array([P.golay_encode(x) for x in d12], np.uint32)), "golay encode"
    w = rng.integers(0, 1 << 24, 20000).astype(np.int32)
    gd, ge = ctx.m_17_golay_decode(dev(w))
    exp = [P.golay_decode(int(x)) for x in w]
    assert np.array_equal(gd.cpu().numpy(), np.array([e[0] for e in exp], np.uint16)), "golay decode data"
    assert np.array_equal(ge.cpu().numpy(), np.array([e[1] for e in exp], np.uint8)), "golay decode errs"
    # conv encoders
    for nb in (18, 26, 30):
        d = rng.integers(0, 256, (n, nb), dtype=np.uint8)
        got = ctx.m17_conv_encode_8(dev(d)).cpu().numpy()
        assert np.array_equal(got, np.stack([P.conv_encode_8(r) for r in d])), ("conv8", nb)
    b = rng.integers(0, 2, (n, 197), dtype=np.uint8)
    assert np.array_equal(ctx.m17_conv_encode_1(dev(b)).cpu().numpy(), np.stack([P.conv_encode_1(r) for r in b])), "conv1"
    # puncture / depuncture
    for p, ln in ((1, 488), (2, 296), (3, 420), (2, 402)):
        b = rng.integers(0, 2, (n, ln), dtype=np.uint8)
        got = ctx.m17_punc(p, dev(b)).cpu().numpy()
        exp = np.stack([P.punc(p, r) for r in b])
        assert np.array_equal(got, exp), ("punc", p, ln)
        s = rng.normal(0, 1, (n, exp.shape[1])).astype(np.float32)
        got = ctx.m17_de_punc(p, dev(s), ln).cpu().numpy()
        assert bits_eq(got, np.stack([P.depunc(p, r, ln) for r in s])), ("depunc", p, ln)
    # interleave, randomiser
    b = rng.integers(0, 2, (n, 368), dtype=np.uint8)
    s = rng.normal(0, 1, (n, 368)).astype(np.float32)
    assert np.array_equal(ctx.m17_interleave(dev(b)).cpu().numpy(), np.stack([P.interleave(r) for r in b])), "interleave"
    assert bits_eq(ctx.m17_de_interleave(dev(s)).cpu().numpy(), np.stack([P.deinterleave(r) for r in s])), "deinterleave"
    assert np.array_equal(ctx.m17_de_correlate_1(dev(b)).cpu().numpy(), np.stack([P.derand_bits(r) for r in b])), "derand u8"
    assert bits_eq(ctx.m17_de_correlate_1(dev(s)).cpu().numpy(), np.stack([P.derand_soft(r) for r in s])), "derand f32"
    by = rng.integers(0, 256, (n, 54), dtype=np.uint8)
    assert np.array_equal(ctx.m17_de_correlate_8(dev(by)).cpu().numpy(), np.stack([P.derand_bytes(r) for r in by])), "derand bytes"
    # demap, sync check
    sy = (rng.normal(0, 1, (n, 192)) * rng.uniform(0.1, 3, (n, 1))).astype(np.float32)
    assert bits_eq(ctx.m17_dsp_demap_frame(dev(sy)).cpu().numpy(), np.stack([P.demap_frame(r) for r in sy])), "demap"
    v = rng.normal(0, 1, (4000, 8)).astype(np.float32)
    tpl = np.array([[1, 1, 1, 1, -1, -1, 1, -1], [-1, -1, -1, -1, 1, 1, -1, 1], [1, -1, 1, 1, -1, -1, -1, -1], [1, 1, 1, 1, 1, 1, -1, 1]], np.float32)
    v[:2000] = tpl[rng.integers(0, 4, 2000)] * (1 + 0.15 * rng.normal(0, 1, (2000, 8))).astype(np.float32)
    v[5] = 0
    ty, vo, va = ctx.m17_sync_check(dev(v))
    exp = [P.sync_check(r) for r in v]
    assert np.array_equal(ty.cpu().numpy(), np.array([e[0] for e in exp], np.uint8)), "sync type"
    assert np.array_equal(vo.cpu().numpy(), np.array([e[1] for e in exp], np.uint8)), "sync votes"
    assert bits_eq(va.cpu().numpy(), np.array([e[2] for e in exp], np.float32)), "sync variance"
    # m17_dsp_demap_symbol, m17_dsp_decimating_filter (m17_dsp.cpp:35-42,438-449): oracle on random inputs, and the reference's
    # own outputs from the fixture
    sy1 = rng.normal(0, 1, 500).astype(np.float32); mg = rng.uniform(0.2, 4, 500).astype(np.float32)
    assert bits_eq(ctx.m17_dsp_demap_symbol(dev(sy1), dev(mg)).cpu().numpy(), P.demap_symbol(sy1, mg)), "demap symbol"
    xin = rng.normal(0, 1, (3, 200 + 30)).astype(np.float32); cf = rng.normal(0, 0.2, 31).astype(np.float32)
    exp = np.stack([P.decimating_filter(r, cf, 5, 200) for r in xin])
    assert bits_eq(ctx.m17_dsp_decimating_filter(dev(xin), dev(cf), 5, 200).cpu().numpy(), exp), "decimating filter"
    GX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "m17_golden_ext.npz"))
    assert bits_eq(ctx.m17_dsp_demap_symbol(dev(GX["dsym_in"]), dev(GX["dsym_mag"])).cpu().numpy(), GX["dsym_out"]), "demap symbol (reference fixture)"
    assert bits_eq(ctx.m17_dsp_decimating_filter(dev(GX["dfil_in"]), dev(GX["dfil_coffs"]), 5, 200).cpu().numpy(), GX["dfil_out"]), "decimating filter (reference fixture)"
    # gps_decode (gps.cpp:8-27) on the META field: oracle on random LSFs and the reference's outputs from the fixture
    gl = rng.integers(0, 256, (200, 30), dtype=np.uint8)
    gg = ctx.gps_decode(dev(gl))
    for i in range(len(gl)):
        assert tuple(gg[i].tolist()) == P.gps_decode(gl[i]), ("gps_decode", i)
    gg = ctx.gps_decode(dev(GX["gps_lsf"]))
    assert np.array_equal(np.array([tuple(r.tolist()) for r in gg], np.float64), GX["gps_out"]), "gps_decode (reference fixture)"
    seq3 = np.tile(P.prbs9(), 3)
    pb = np.stack([seq3[:900], np.roll(seq3, 5)[:900], rng.integers(0, 2, 900).astype(np.uint8)]).astype(np.uint8)
    pb[0, [300, 301, 640]] ^= 1
    st = ctx.m17_prbs9_rx_check(dev(pb[:, :400]))
    st = ctx.m17_prbs9_rx_check(dev(pb[:, 400:]), st).cpu().numpy().view(np.uint32)          # state carried between calls
    assert np.array_equal(st, np.stack([P.prbs_check(b) for b in pb])), "prbs9 rx check"
    # PRBS9
    seq = P.prbs9()
    got = ctx.m17_prbs9_tx_load(3, 600).cpu().numpy()
    assert np.array_equal(got[0], np.concatenate([seq, seq])[:600]), "prbs9"
    # filter design (host side of the library)
    for args in ((0.5, 1240, 80), (0.5, 310, 10), (0.5, 62, 2), (0.5, 2480, 80)):
        assert bits_eq(ctx.m17_dsp_build_rrc_filter(*args), P.rrc(*args)), ("rrc", args)
    mf, md = ctx.sync_taps()
    pmf, pmd = P.sync_taps()
    assert bits_eq(mf, pmf) and bits_eq(md, pmd), "sync taps"
    return "primitives ok"


def check_viterbi(ctx, P, seed=12, n=300):
    rng = np.random.default_rng(seed)
    for ln in (296, 420, 488, 24, 2):
        s = rng.normal(0, 1, (n, ln)).astype(np.float32)
        s[0] = 0                                   # pure tie-break frame
        s[1] = np.round(s[1] * 4) / 4              # many exact ties
        s[2, ::3] = 0
        got = ctx.m17_viterbi_decode(dev(s)).cpu().numpy()
        exp = np.stack([P.viterbi(r) for r in s])
        assert np.array_equal(got, exp), ("viterbi", ln, first_diff(got, exp))
    # punctured + packed form on encoded noisy frames
    for p, nb, full in ((1, 30, 488), (2, 18, 296), (3, 26, 420)):
        data = rng.integers(0, 256, (n, nb), dtype=np.uint8)
        soft = []
        exp = []
        for r in data:
            cod = P.conv_encode_8(r)[:full]
            x = (2.0 * P.punc(p, cod) - 1.0) + rng.normal(0, 0.8, len(P.punc(p, cod)))
            x = (np.round(np.clip(x, -2, 2) * 32) / 32).astype(np.float32)
            soft.append(x)
            bits = P.viterbi(P.depunc(p, x, full))
            exp.append(np.packbits(bits[1:1 + 8 * nb]))
        got = ctx.viterbi_punctured(p, dev(np.stack(soft))).cpu().numpy()
        assert np.array_equal(got, np.stack(exp)), ("viterbi_punctured", p, first_diff(got, np.stack(exp)))
    return "viterbi ok"


def oracle_parse_frame(P, sym, ty):
    """m17_rx_parse for one frame, composed from oracle primitives (no LICH state)."""
    r = np.zeros((), REC_DTYPE)
    r["type"] = ty
    r["flags"] = 2
    if ty < 1 or ty > 4:
        return r, None
    sb = P.demap_frame(sym)
    acc = np.float32(0)
    for k in range(8):
        acc = np.float32(acc + np.abs(np.float32(sym[k])))
    r["cor"] = np.float32(8.0 / np.float64(acc))
    if ty == 4:
        return r, sb
    so = P.deinterleave(P.derand_soft(sb))
    if ty == 1:
        bits = P.viterbi(P.depunc(1, so, 488)); nb = 30
    elif ty == 2:
        e = 0; w = []
        for k in range(4):
            d, ee = P.golay_decode(P.hard24(so[24 * k:24 * k + 24])); w.append(d); e += ee
        r["golay_err"] = e
        w01, w23 = (w[0] << 12) | w[1], (w[2] << 12) | w[3]
        r["lich"] = [(w01 >> 16) & 255, (w01 >> 8) & 255, w01 & 255, (w23 >> 16) & 255, (w23 >> 8) & 255, w23 & 255]
        bits = P.viterbi(P.depunc(2, so[96:], 296)); nb = 18
    else:
        bits = P.viterbi(P.depunc(3, so, 420)); nb = 26
    data = np.packbits(bits[1:1 + 8 * nb])
    r["data"][:nb] = data
    r["nbytes"] = nb
    r["crc"] = P.crc(bytes(data))
    if ty == 3 and data[25] & 0x80:
        r["flags"] |= 0x20
    return r, sb


def check_parse_frames(ctx, P, seed=13, n=200):
    rng = np.random.default_rng(seed)
    lsf = lsf_for(P)
    frames = [P.fmt_lsf(lsf)] + list(P.fmt_stream_frames(lsf, rng.integers(0, 256, (8, 16), dtype=np.uint8))) + \
             [P.fmt_packet(bytes(rng.integers(0, 256, 25, dtype=np.uint8)), 0, 3), P.fmt_packet(b"hello", 1, 5)] + list(P.fmt_bert(2))
    lu = np.array([1 / 3, 1.0, -1 / 3, -1.0], np.float32)
    sym = np.zeros((n, 192), np.float32)
    types = np.zeros(n, np.uint8)
    for i in range(n):
        d = frames[i % len(frames)]
        s = lu[d] * np.float32(rng.uniform(0.3, 2.0)) + rng.normal(0, 0.25 if i % 3 else 0.02, 192).astype(np.float32)
        sym[i] = s
        types[i] = P.sync_check(s[:8])[0] if i % 11 else rng.integers(0, 6)
    rec, soft = ctx.m17_rx_parse(dev(sym), dev(types), want_soft=True)
    rec = rec.cpu().numpy().view(REC_DTYPE).reshape(n)
    soft = soft.cpu().numpy()
    for i in range(n):
        e, sb = oracle_parse_frame(P, sym[i], int(types[i]))
        for name in ("type", "golay_err", "nbytes", "lich", "data", "crc"):
            assert np.array_equal(rec[i][name], e[name]), ("parse", i, int(types[i]), name, rec[i][name], e[name])
        assert (rec[i]["flags"] & 0x22) == (e["flags"] & 0x22), ("parse flags", i, rec[i]["flags"], e["flags"])
        if sb is not None:
            assert bits_eq(rec[i]["cor"], e["cor"]), ("cor", i, rec[i]["cor"], e["cor"])
            assert bits_eq(soft[i], sb), ("soft bits", i, first_diff(soft[i], sb))
    return "parse_frames ok"


# ---------------------------------------------------------------------------------------------- RX chain
def compare_chain(res, o, seam, Cn, verbose=False):
    """res = Rx.results() of the CUDA chain, o = oracle rx_run output."""
    msgs = []
    T = o.nsym.shape[1]
    if seam == 0:
        disc = res["disc_raw"] - res["mean"][:, :, None]
        if not bits_eq(disc, o.disc):
            msgs.append(("disc", first_diff(disc, o.disc)))
    if not np.array_equal(res["nsym"], o.nsym):
        msgs.append(("nsym", first_diff(res["nsym"], o.nsym)))
    for c in range(Cn):
        ns = int(o.counts[c, 1]); nf = int(o.counts[c, 2]); ne = int(o.counts[c, 3])
        if int(res["nsym"][c].sum()) != ns:
            msgs.append(("sym count", c, int(res["nsym"][c].sum()), ns)); continue
        if not bits_eq(res["syms"][c, :ns], o.syms[c, :ns]):
            msgs.append(("syms", c, first_diff(res["syms"][c, :ns], o.syms[c, :ns])))
        if int(res["nframes"][c]) != nf:
            msgs.append(("nframes", c, int(res["nframes"][c]), nf)); continue
        fa, fb = res["frames"][c, :nf], o.frames[c, :nf]
        for name in REC_DTYPE.names:
            if name == "rsvd":
                continue
            if not bits_eq(fa[name], fb[name]):
                bad = [k for k in range(nf) if not bits_eq(fa[name][k], fb[name][k])]
                msgs.append(("rec." + name, c, bad[:5], fa[name][bad[0]], fb[name][bad[0]], int(fb["type"][bad[0]])))
        if int(res["nevents"][c]) != ne or not np.array_equal(res["events"][c, :ne], o.events[c, :ne]):
            msgs.append(("events", c, res["events"][c, :ne], o.events[c, :ne]))
    if verbose:
        for m in msgs[:20]:
            print("   MISMATCH", m)
    assert not msgs, msgs[:4]


def check_decimator(ctx, P, seed=41):
    """Pluto /8 decimator (radio.cpp:18-51,157-177): random and extreme int16 input, ragged call sizes with the history
    carried between calls, bit-exact against the oracle; then decimator -> m17_dsp_rx on an 8x oversampled M17 signal."""
    import m17_sdr_b200 as m
    rng = np.random.default_rng(seed)
    Cn, nout = 5, 3 * 1920
    X = rng.integers(-32768, 32768, (Cn, 8 * nout, 2)).astype(np.int16)
    X[1] = 32767
    X[2, ::2] = -32768
    X[3] = -32768
    exp = P.dec_run(X)
    dec = m.Decimator(ctx, Cn)
    assert np.array_equal(dec.taps(), P.dec_taps()), "decimator taps"
    got = dec.radio_receive_samples(dev(X)).cpu().numpy()
    assert np.array_equal(got, exp), ("decimate one call", first_diff(got.view(np.uint32), exp.view(np.uint32)))
    dec.reset()
    o = 0
    for n in (4, 236, 1920, 512, 516, 2572):
        got = dec.radio_receive_samples(dev(X[:, 8 * o:8 * (o + n)])).cpu().numpy()
        assert np.array_equal(got, exp[:, o:o + n]), ("decimate split", o, n)
        o += n
    assert o == nout
    dec.close()
    # a Pluto-rate capture: TX at oversample 80 (384 kS/s), decimator, then the RX chain; everything against the oracle
    lsf = lsf_for(P)
    pl = rng.integers(0, 256, (2, 6, 16), dtype=np.uint8)
    chans = []
    for c in range(2):
        iq, _, _ = P.tx_stream_over(lsf, pl[c], os_=80)
        chans.append(iq)
    n = max(len(x) for x in chans)
    T = n // (8 * 1920) + 2
    Y = np.zeros((2, T * 8 * 1920, 2), np.int16)
    for c in range(2):
        d = int(rng.integers(0, 8 * 1920))
        Y[c, d:d + len(chans[c])] = chans[c][: Y.shape[1] - d]
        Y[c, :d] = chans[c][0]; Y[c, d + len(chans[c]):] = chans[c][-1]
    Y = signals.add_iq_noise(Y, 30.0, rng)
    dec = m.Decimator(ctx, 2)
    x48 = dec.radio_receive_samples(dev(Y))
    e48 = P.dec_run(Y)
    assert np.array_equal(x48.cpu().numpy(), e48), "decimate M17 capture"
    dec.close()
    z = (e48[..., 0] == 0) & (e48[..., 1] == 0)       # the zero filter history gives (0,0) at the very start: the limiter divides by |z|
    e48[z, 0] = 1
    res = run_chain(ctx, e48, 0)
    ora = P.rx_run(e48, seam=0)
    compare_chain(res, ora, 0, 2)
    ndel = int(sum(((res["frames"][c, :res["nframes"][c]]["flags"] & F_DELIVERED) != 0).sum() for c in range(2)))
    return "decimator ok (%d outputs x %d ch exact, split calls exact, 384 kS/s capture -> %d frames, %d delivered)" % (nout, Cn, int(res["nframes"].sum()), ndel)


def check_udp_frames(ctx, P, seed=43):
    """M17-over-UDP frames: pack / parse primitives against the oracle, and the gateway output of a decoded
    RX call -- one datagram per delivered stream frame carrying the LSF recovered from the LICH -- against datagrams the oracle
    builds from the oracle's own records; finally datagrams -> TX formatter -> RX again (gateway loop)."""
    import m17_sdr_b200 as m
    rng = np.random.default_rng(seed)
    n = 300
    lsf = np.stack([P.build_lsf(int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 16)),
                                rng.integers(0, 256, 14, dtype=np.uint8)) for _ in range(n)])
    sid = rng.integers(0, 1 << 16, n).astype(np.uint16); fn = rng.integers(0, 1 << 16, n).astype(np.uint16)
    pld = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    dst = P.encode_call("M17-M17 C")
    for d in (None, dst):
        got = ctx.net_pack(dev(sid), dev(lsf), dev(fn), dev(pld), dst=d).cpu().numpy()
        exp = np.stack([P.net_pack(int(sid[i]), lsf[i], int(fn[i]), pld[i], dst=d) for i in range(n)])
        assert np.array_equal(got, exp), ("net_pack", d)
    bad = exp.copy()
    bad[::3, rng.integers(0, 54)] ^= 0x10
    ok, s2, l2, f2, p2 = [t.cpu().numpy() for t in ctx.net_parse(dev(bad))]
    for i in range(n):
        eok, esid, elsf, efn, epld = P.net_parse(bad[i])
        assert bool(ok[i]) == eok and s2[i] == esid and f2[i] == efn and np.array_equal(l2[i], elsf) and np.array_equal(p2[i], epld), ("net_parse", i)
    # gateway output of an RX call: two transmissions with different LSFs on every channel, so the LSF cache changes mid-call
    Cn, F = 6, 9
    X = []
    plA, plB = rng.integers(0, 256, (Cn, F, 16), dtype=np.uint8), rng.integers(0, 256, (Cn, F, 16), dtype=np.uint8)
    for c in range(Cn):
        a, _, _ = P.tx_stream_over(lsf_for(P, src="G4GUO    "), plA[c])
        b, _, _ = P.tx_stream_over(lsf_for(P, src="AB%dCD    " % c, typeword=0x0085), plB[c])
        X.append(np.concatenate([a, b]))
    T = max(len(x) for x in X) // 1920 + 2
    Y = np.zeros((Cn, T * 1920, 2), np.int16)
    for c in range(Cn):
        Y[c, :len(X[c])] = X[c]; Y[c, len(X[c]):] = X[c][-1]
    Y = signals.add_iq_noise(Y, 28.0, rng)
    rx = m.Rx(ctx, Cn, T)
    rx.m17_dsp_rx(dev(Y))
    res = rx.results()
    sidc = (np.arange(Cn) * 257 + 5).astype(np.uint16)
    out, cnt = rx.m17_net_new_rx_data(dev(sidc), dst=dst)
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    o = P.rx_run(Y, seam=0)
    ntot = 0
    for c in range(Cn):
        # oracle: walk the oracle's records, keeping m_lsf[1] as update_lich would (m17_rx_parse.cpp:71-85)
        lsf0, lsf1, exp = np.zeros(30, np.uint8), np.zeros(30, np.uint8), []
        for r in o.frames[c, : int(o.counts[c, 2])]:
            if r["type"] != 2 or not (r["flags"] & 2):
                continue
            seq = int(r["lich"][5]) >> 5
            if seq < 6:
                lsf0[5 * seq:5 * seq + 5] = r["lich"][:5]
                if P.crc(bytes(lsf0)) == 0:
                    lsf1 = lsf0.copy()
            if r["flags"] & F_DELIVERED:
                exp.append(P.net_pack(int(sidc[c]), lsf1, (int(r["data"][0]) << 8) | int(r["data"][1]), r["data"][2:18], dst=dst))
        assert cnt[c] == len(exp), ("datagram count", c, cnt[c], len(exp))
        assert np.array_equal(out[c, :cnt[c]], np.stack(exp)), ("datagrams", c)
        srcs = {bytes(e[12:18]) for e in exp}
        assert len(srcs) == 2, "both transmissions' LSFs must appear"
        ntot += len(exp)
    rx.close()
    return "udp frames ok (%d packed/parsed exact, %d gateway datagrams over %d channels with an LSF change mid-call)" % (n, ntot, Cn)


def check_rx_symbols(ctx, P, seed=53):
    """Symbol seam (m17_rx_symbols on caller-supplied symbols): the oracle's own symbol streams of a noisy baseband run go
    through framer + decode + post in one call and in ragged pieces; records and events must equal the oracle's."""
    import m17_sdr_b200 as m
    D, _ = signals.baseband_channels(P, 7, 12, seed, [None, 12, 9, 7, 5, 3, 1])
    o = P.rx_run(D, seam=1)
    Cn = D.shape[0]
    ns = o.counts[:, 1].astype(np.int64)
    T = D.shape[1] // 384
    rx = m.Rx(ctx, Cn, T)
    for pieces in (1, 3):
        rx.reset()
        frames, events = [[] for _ in range(Cn)], [[] for _ in range(Cn)]
        cuts = [np.round(np.linspace(0, n, pieces + 1) + (np.arange(pieces + 1) % 2) * 5 * (pieces > 1)).astype(np.int64).clip(0, n) for n in ns]
        for k in range(pieces):
            cnt = np.array([cuts[c][k + 1] - cuts[c][k] for c in range(Cn)], np.int32)
            buf = np.zeros((Cn, max(int(cnt.max()), 1)), np.float32)
            for c in range(Cn):
                buf[c, :cnt[c]] = o.syms[c, cuts[c][k]:cuts[c][k + 1]]
            rx.m17_rx_symbols(dev(buf), dev(cnt))
            r = rx.results()
            for c in range(Cn):
                frames[c].append(r["frames"][c, :r["nframes"][c]]); events[c].append(r["events"][c, :r["nevents"][c]])
        for c in range(Cn):
            nf, ne = int(o.counts[c, 2]), int(o.counts[c, 3])
            fa = np.concatenate(frames[c]); ea = np.concatenate(events[c])
            assert len(fa) == nf and len(ea) == ne, ("symbol seam counts", c, len(fa), nf, len(ea), ne)
            for name in REC_DTYPE.names:
                if name != "rsvd":
                    assert bits_eq(fa[name], o.frames[c, :nf][name]), ("symbol seam rec." + name, c, pieces)
            assert np.array_equal(ea, o.events[c, :ne]), ("symbol seam events", c, pieces)
    rx.close()
    return "rx symbols ok (%d ch, %d frames, one call and 3 ragged pieces)" % (Cn, int(o.counts[:, 2].sum()))


def check_rx_bert(ctx, P, seed=47, nchan=8, F=14):
    """BERT receive: carrier, preambles, F BERT frames (m17_fmt_add_bert_frame), EOT at several noise levels.
    With the extension on, records (decoded PRBS bytes) and the m17_prbs9_rx_check state must equal the oracle's; with it off
    (upstream behaviour) BERT records carry no data."""
    import m17_sdr_b200 as m
    rng = np.random.default_rng(seed)
    pre, eot = np.full(192, 0, np.uint8), np.zeros(192, np.uint8)
    P.L.m17o_fmt_preamble(pre.ctypes.data_as(__import__("ctypes").c_void_p)); P.L.m17o_fmt_eot(eot.ctypes.data_as(__import__("ctypes").c_void_p))
    script = np.concatenate([np.full(192, 4, np.uint8), pre, pre] + list(P.fmt_bert(F)) + [eot, np.full(2 * 192, 4, np.uint8)])
    iq = P.mod(script)
    T = len(iq) // 1920 + 2
    X = np.zeros((nchan, T * 1920, 2), np.int16)
    for c in range(nchan):
        d = int(rng.integers(0, 1920))
        X[c, d:d + len(iq)] = iq[: X.shape[1] - d]; X[c, :d] = iq[0]; X[c, d + len(iq):] = iq[-1]
    X = np.stack([signals.add_iq_noise(X[c], [None, 30.0, 26.0, 24.0, 23.0, 22.0, 21.5, 21.0][c % 8], rng) for c in range(nchan)])
    o = P.rx_run(X, seam=0, bert=True)
    rx = m.Rx(ctx, nchan, T)
    rx.set_bert(True)
    for split in (None, [3, 5]):
        rx.reset()
        t0 = 0
        outs = []
        for nb in (split or []) + [T - sum(split or [])]:
            rx.m17_dsp_rx(dev(X[:, t0 * 1920:(t0 + nb) * 1920])); outs.append(rx.results()); t0 += nb
        got = rx.bert().cpu().numpy().view(np.uint32)
        assert np.array_equal(got, o["bert"]), ("bert checker state", got[:3], o["bert"][:3])
        for c in range(nchan):
            fr = np.concatenate([r["frames"][c, :r["nframes"][c]] for r in outs])
            nf = int(o.counts[c, 2])
            assert len(fr) == nf, ("nframes", c)
            for name in ("type", "flags", "nbytes", "data", "crc"):
                assert bits_eq(fr[name], o.frames[c, :nf][name]), ("bert rec." + name, c)
    nb = int((o.frames["type"] == 4).sum())
    best = int(np.argmax(o["bert"][:, 6].astype(np.int64) - 1000 * o["bert"][:, 7].astype(np.int64)))
    assert int(o["bert"][best, 6]) > 150 * (F - 2) and int(o["bert"][best, 7]) == 0, "a high-SNR channel must sync with no bit errors"
    rx.set_bert(False)
    rx.reset(); rx.m17_dsp_rx(dev(X)); off = rx.results()
    gc_off = P.rx_run(X, seam=0)
    compare_chain(off, gc_off, 0, nchan)
    rx.close()
    return "rx bert ok (%d BERT frames, checker states exact; best channel %d bits / %d errors, noisiest %d / %d)" % (
        nb, o["bert"][best, 6], o["bert"][best, 7], o["bert"][nchan - 1, 6], o["bert"][nchan - 1, 7])


def check_rx_afc(ctx, P, seed=31, nchan=12, nframes=30):
    """AFC on (dsp_nco_mixer + radio_afc): records and events exact; discriminator samples / symbols bit-identical unless a
    double sincos result fell on a float rounding boundary (CUDA libm vs glibc), then within 1e-5 relative RMS."""
    X, _ = signals.stream_channels(P, nchan, nframes, seed, ebn0=[None, 30, 26, 24, None, 28] * 2, f0_max=1800.0)
    o = P.rx_run(X, seam=0, afc=True)
    for split in (None, [7, 1, 13]):
        res = run_chain(ctx, X, 0, split, afc=True)
        disc = res["disc_raw"] - res["mean"][:, :, None]
        exact = bits_eq(disc, o.disc)
        rel = float(np.sqrt(np.mean((disc.astype(np.float64) - o.disc) ** 2) / np.mean(o.disc.astype(np.float64) ** 2)))
        assert rel < 1e-5, ("afc disc rms", rel)
        if exact:
            compare_chain(res, o, 0, nchan)
        else:                                   # one-ulp sincos difference somewhere: decisions must still agree
            assert np.array_equal(res["nsym"], o.nsym)
            for c in range(nchan):
                nf = int(o.counts[c, 2]); ns = int(o.counts[c, 1])
                assert int(res["nframes"][c]) == nf
                for name in ("sym_off", "type", "flags", "golay_err", "nbytes", "lich", "data", "crc", "votes", "frame_errors"):
                    assert bits_eq(res["frames"][c, :nf][name], o.frames[c, :nf][name]), (name, c)
                a, b = res["syms"][c, :ns].astype(np.float64), o.syms[c, :ns].astype(np.float64)
                assert np.sqrt(np.mean((a - b) ** 2) / max(np.mean(b ** 2), 1e-30)) < 1e-5
    off = run_chain(ctx, X, 0)
    assert not bits_eq(off["disc_raw"], res["disc_raw"])       # the loop really acted
    return "rx afc ok (%d ch, bit-identical=%s, rel rms %.1e, %d frames)" % (nchan, exact, rel, int(o.counts[:, 2].sum()))


def check_rx_equaliser(ctx, P, seed=61):
    """Equaliser option (m17b_rx_set_equaliser): eq_train_unknown on (half-symbol, symbol) pairs between the timing
    loop and the framer, bit for bit against the oracle's seam flag 64 (itself pinned to the reference's own functions,
    tests/test_oracle_vs_ref.py) -- nsym, the equalised symbol stream, records, events.  Baseband seam with noise down to 0 dB
    (threshold trips and bit slips while unlocked), the IQ path, and both again split over several calls (equaliser state and the
    half-symbol value carried across calls).  The option and AFC exclude each other."""
    import m17_sdr_b200 as m
    D, _ = signals.baseband_channels(P, 8, 14, seed, [None, 12, 10, 6, 4, 2, 0, -3])
    X, _ = signals.stream_channels(P, 6, 12, seed + 1, ebn0=[None, 30, 24, 20, 16, 12], f0_max=1200.0)
    nfr = 0
    for data, seam in ((D, 1), (X, 0)):
        o = P.rx_run(data, seam=seam, eq=True)
        off = P.rx_run(data, seam=seam)
        assert not bits_eq(o.syms, off.syms)
        for split in (None, [2, 1, 5, 1]):
            compare_chain(run_chain(ctx, data, seam, split, eq=True), o, seam, data.shape[0], verbose=True)
        nfr += int(o.counts[:, 2].sum())
    rx = m.Rx(ctx, 2, 2)
    rx.set_afc(True)
    try:
        rx.set_equaliser(True)
        raise AssertionError("equaliser + AFC accepted")
    except m.M17Error:
        pass
    rx.set_afc(False)
    rx.set_equaliser(True)
    try:
        rx.set_afc(True)
        raise AssertionError("AFC + equaliser accepted")
    except m.M17Error:
        pass
    rx.close()
    return "rx equaliser option ok (%d + %d ch, %d frame records)" % (D.shape[0], X.shape[0], nfr)


def run_chain(ctx, X, seam=0, split=None, afc=False, eq=False):
    """Run the CUDA chain over X; split = list of block counts to process in successive calls (state carry)."""
    import m17_sdr_b200 as m
    Cn = X.shape[0]
    per = 1920 if seam == 0 else 384
    T = X.shape[1] // per
    parts = list(split or [])
    assert sum(parts) <= T
    if sum(parts) < T:
        parts.append(T - sum(parts))
    rx = m.Rx(ctx, Cn, max(parts))
    if afc:
        rx.set_afc(True)
    if eq:
        rx.set_equaliser(True)
    out = []
    t0 = 0
    for nb in parts:
        xs = np.ascontiguousarray(X[:, t0 * per:(t0 + nb) * per])
        if seam == 0:
            rx.m17_dsp_rx(torch.from_numpy(xs).cuda())
        else:
            rx.m17_rx_baseband(torch.from_numpy(xs).cuda())
        out.append(rx.results())
        t0 += nb
    rx.close()
    if len(out) == 1:
        return out[0]
    # stitch successive calls into one result in the oracle's layout
    res = {}
    res["nsym"] = np.concatenate([r["nsym"] for r in out], 1)
    res["nframes"] = sum(r["nframes"] for r in out)
    res["nevents"] = sum(r["nevents"] for r in out)
    if seam == 0:
        res["disc_raw"] = np.concatenate([r["disc_raw"] for r in out], 1)
        res["mean"] = np.concatenate([r["mean"] for r in out], 1)
    nsyms = [r["nsym"].sum(1) for r in out]
    mx = max(int(sum(n[c] for n in nsyms)) for c in range(Cn))
    res["syms"] = np.zeros((Cn, mx), np.float32)
    for c in range(Cn):
        v = np.concatenate([out[k]["syms"][c, :nsyms[k][c]] for k in range(len(out))])
        res["syms"][c, :len(v)] = v
    fcap = sum(r["frames"].shape[1] for r in out)
    res["frames"] = np.zeros((Cn, fcap), REC_DTYPE)
    ecap = sum(r["events"].shape[1] for r in out)
    res["events"] = np.zeros((Cn, ecap), out[0]["events"].dtype)
    for c in range(Cn):
        f = np.concatenate([r["frames"][c, :r["nframes"][c]] for r in out]); res["frames"][c, :len(f)] = f
        e = np.concatenate([r["events"][c, :r["nevents"][c]] for r in out]); res["events"][c, :len(e)] = e
    return res


def check_rx_chain(ctx, P, nchan=16, nframes=30, seed=21, verbose=False, split=None):
    eb = [None, None, 30, 28, 26, 24, 23, 22, 21, 20, 26, 26, 24, 24, 22, 22][:nchan] + [26] * max(0, nchan - 16)
    X, pl = signals.stream_channels(P, nchan, nframes, seed, ebn0=eb, f0_max=1000.0)
    o = P.rx_run(X, seam=0)
    res = run_chain(ctx, X, 0, split)
    compare_chain(res, o, 0, nchan, verbose)
    deliv = sum(int(((o.frames[c, :o.counts[c, 2]]["flags"] & F_DELIVERED) != 0).sum()) for c in range(nchan))
    return f"rx chain ok ({nchan} ch x {X.shape[1] // 1920} blocks, {int(o.counts[:, 2].sum())} frames, {deliv} delivered)"


# (|re|, |im|) int16 pairs whose limiter divisor sqrt(re^2 + im^2) has an all-ones fp32 significand: the one input class the
# front end's Newton reciprocal must patch (frontend.cuh); found by enumeration, see the test
FE_ALL_ONES = [(3753, 1810), (4078, 855), (7506, 3620), (7879, 2714), (8156, 1710), (12665, 10834), (14136, 8829), (15012, 7240),
               (16666, 149), (23673, 23467), (24084, 23045), (28738, 16889), (30333, 13821)]


def check_rx_chain_limiter_patch(ctx, P, nchan=40, nframes=8, seed=77):
    """the unpatched fast pass of the front end must detect every sample whose divisor has an all-ones significand and redo
    the warp-unit exactly: such samples are planted (all sign / swap variants) into real signals, in some rows only, at
    positions that include the first and last chunk of a block"""
    X, pl = signals.stream_channels(P, nchan, nframes, seed, ebn0=[26] * nchan, f0_max=500.0)
    X = X.copy()
    rng = np.random.default_rng(seed)
    T = X.shape[1] // 1920
    # check the table really is the class it claims (numpy float32 sqrt is correctly rounded)
    for a, b in FE_ALL_ONES:
        re, im = np.float32(np.float64(a) * 0.00003), np.float32(np.float64(b) * 0.00003)
        mbits = np.sqrt(np.float32(np.float32(re * re) + np.float32(im * im))).view(np.uint32)
        assert (int(mbits) | 0xFF800000) == 0xFFFFFFFF, (a, b)
    planted = 0
    for c in range(nchan):
        if c % 3 == 0:
            continue                                            # rows without any: their units must not change either
        for t in range(T):
            if rng.random() < 0.5:
                continue
            for pos in set([0, 1, 19, 1900, 1919] if (c + t) % 4 == 0 else []) | set(rng.integers(0, 1920, 3).tolist()):
                a, b = FE_ALL_ONES[int(rng.integers(len(FE_ALL_ONES)))]
                if rng.random() < 0.5:
                    a, b = b, a
                i = t * 1920 + pos
                X[c, i, 0] = a * (1 if rng.random() < 0.5 else -1)
                X[c, i, 1] = b * (1 if rng.random() < 0.5 else -1)
                planted += 1
    o = P.rx_run(X, seam=0)
    res = run_chain(ctx, X, 0)
    compare_chain(res, o, 0, nchan)
    return f"limiter patch ok ({planted} planted samples, {int(o.counts[:, 2].sum())} frames)"


def check_rx_baseband(ctx, P, nchan=14, nframes=30, seed=22, verbose=False, split=None):
    eb = ([None, 12, 10, 8, 6, 4, 2, 0, 12, 10, 8, 6, 4, 2] * ((nchan + 13) // 14))[:nchan]
    D, pl = signals.baseband_channels(P, nchan, nframes, seed, eb)
    o = P.rx_run(D, seam=1)
    res = run_chain(ctx, D, 1, split)
    compare_chain(res, o, 1, nchan, verbose)
    return f"rx baseband ok ({nchan} ch, Eb/N0 sweep 0..12 dB, {int(o.counts[:, 2].sum())} frames)"


def check_rx_packet(ctx, P, nchan=12, seed=25, verbose=False):
    """config 3: packet mode with random carrier / fractional-timing offsets; per-frame 26 bytes bit-exact vs the oracle,
    and the host-side reassembly (GPU CRC) recovers every packet the oracle's frames carry"""
    import m17_sdr_b200 as m
    from m17_sdr_b200.api import packets_of
    eb = ([None, None, 30, 28, 26, 24] * 4)[:nchan]
    X, packets = signals.packet_channels(P, nchan, seed, ebn0=eb)
    o = P.rx_run(X, seam=0)
    res = run_chain(ctx, X, 0)
    compare_chain(res, o, 0, nchan, verbose)
    # batched reassembly on the GPU, in one call and with the capture cut into three calls (a packet then spans calls)
    T = X.shape[1] // 1920
    rx = m.Rx(ctx, nchan, T)
    rx.m17_dsp_rx(dev(X))
    whole = rx.reassemble_packets(bytes_cap=8192, max_pkts=64)      # (a noisy channel yields spurious short 'packets': room for all of them)
    rx.reset()
    parts = [[] for _ in range(nchan)]
    cuts = [0, T // 3, T // 3 + 2, T]
    for a, b in zip(cuts[:-1], cuts[1:]):
        rx.m17_dsp_rx(dev(X[:, a * 1920:b * 1920]))
        pb = rx.reassemble_packets(bytes_cap=8192, max_pkts=64)
        for c in range(nchan):
            parts[c] += packets_of(*pb, c)
    rx.close()
    good = 0
    for c in range(nchan):
        got = packets_of(*whole, c)
        assert parts[c] == got, ("reassembly across calls", c)
        # oracle-side reassembly of the oracle's own records with the oracle CRC
        buf, exp = b"", []
        for r in o.frames[c, : o.counts[c, 2]]:
            if r["type"] == 3 and r["flags"] & 2:
                m = int(r["data"][25])
                if m & 0x80:
                    buf += bytes(r["data"][: (m >> 2) & 31]); exp.append((buf[:-2], P.crc(buf) == 0) if len(buf) >= 2 else (b"", False)); buf = b""      # shorter than its CRC: no payload
                else:
                    buf += bytes(r["data"][:25])
        assert got == exp, (c, got, exp)
        good += sum(1 for pl, ok in got if ok and pl == packets[c])
    assert good >= nchan // 2
    return f"rx packet ok ({nchan} ch, {int(o.counts[:, 2].sum())} frames, {good} packets recovered with valid CRC)"


# ---------------------------------------------------------------------------------------------- TX
def check_tx(ctx, P, nchan=6, F=12, seed=31, os_=10):
    import m17_sdr_b200 as m
    rng = np.random.default_rng(seed)
    lsfs = np.stack([P.build_lsf(0xFFFFFFFFFFFF, P.encode_call("G4GUO    "), 5, bytes(rng.integers(0, 256, 14, dtype=np.uint8))) for _ in range(nchan)])
    pl = rng.integers(0, 256, (nchan, F, 16), dtype=np.uint8)
    tx = m.Tx(ctx, nchan, os_)
    assert np.array_equal(tx.fmt_preamble(), P.fmt_preamble()) and np.array_equal(tx.fmt_eot(), P.fmt_eot())
    got = tx.m17_fmt_add_link_setup_frame(dev(lsfs)).cpu().numpy()
    assert np.array_equal(got, np.stack([P.fmt_lsf(l) for l in lsfs])), "fmt lsf"
    # m17_send_packet_frames: CRC + 25-byte chunking, including the exact-fit and single-frame cases
    lens = [1, 22, 23, 24, 48, 73, 100, 200, 798]
    pk = np.zeros((len(lens), 800), np.uint8)
    for i, ln in enumerate(lens):
        pk[i, :ln] = rng.integers(0, 256, ln, dtype=np.uint8)
    dib, nfr = tx.m17_send_packet_frames(dev(pk), dev(np.array(lens, np.int32)))
    dib, nfr = dib.cpu().numpy(), nfr.cpu().numpy()
    for i, ln in enumerate(lens):
        exp = signals.packet_frames(P, bytes(pk[i, :ln]))
        assert nfr[i] == len(exp), ("packet frames", ln, nfr[i], len(exp))
        assert np.array_equal(dib[i, :nfr[i]], np.stack(exp)), ("packet dibits", ln)
        assert np.all(dib[i, nfr[i]:] == 4), "unused packet slots are blank carrier"
    tx.set_lsf(dev(lsfs))
    d1 = tx.m17_fmt_add_stream_frame(dev(pl[:, :F // 2])).cpu().numpy()
    d2 = tx.m17_fmt_add_stream_frame(dev(pl[:, F // 2:])).cpu().numpy()          # state carry: m_fn / m_lich_count
    got = np.concatenate([d1, d2], 1)
    exp = np.stack([P.fmt_stream_frames(lsfs[c], pl[c]) for c in range(nchan)])
    assert np.array_equal(got, exp), ("fmt stream", first_diff(got, exp))
    ch = rng.integers(0, 256, (9, 25), dtype=np.uint8)
    meta = np.array([(k << 2) | (0x80 if k == 8 else 0) for k in range(9)], np.uint8)
    got = tx.m17_fmt_add_packet(dev(ch), dev(meta)).cpu().numpy()
    exp = np.stack([P.fmt_packet(bytes(ch[k]), meta[k] >> 7, (meta[k] >> 2) & 31) for k in range(9)])
    assert np.array_equal(got, exp), "fmt packet"
    got = tx.m17_fmt_add_bert_frame(4).cpu().numpy()
    assert np.array_equal(got[0], P.fmt_bert(4)), "fmt bert"
    # modulator: one over per channel, in two calls (state carry); frequency samples bit-exact, IQ within 1 LSB
    scripts = []
    for c in range(nchan):
        s = [np.full(192, 4, np.uint8), P.fmt_preamble(), P.fmt_lsf(lsfs[c])] + list(exp_stream(P, lsfs[c], pl[c])) + [P.fmt_eot(), np.full(192, 4, np.uint8)]
        scripts.append(np.concatenate(s))
    scripts = np.stack(scripts)
    tx.reset()
    cut = 192 * 5 + 77
    iq1, f1 = tx.m17_mod_dibits(dev(scripts[:, :cut]), want_freq=True)
    iq2, f2 = tx.m17_mod_dibits(dev(scripts[:, cut:]), want_freq=True)
    iq = torch.cat([iq1, iq2], 1).cpu().numpy(); fr = torch.cat([f1, f2], 1).cpu().numpy()
    for c in range(nchan):
        eiq, efr = P.mod(scripts[c], os_, want_freq=True)
        assert bits_eq(fr[c], efr), ("mod freq", c, first_diff(fr[c], efr))
        dmax = np.abs(iq[c].astype(np.int32) - eiq.astype(np.int32)).max()
        assert dmax <= 1, ("mod iq", c, int(dmax))
    tx.close()
    return "tx ok"


def exp_stream(P, lsf, pl):
    return P.fmt_stream_frames(lsf, pl)


def check_equalizer(ctx, P, nchan=5, nsym=400, seed=41):
    import m17_sdr_b200 as m
    rng = np.random.default_rng(seed)
    lv = np.array([1.0, 1 / 3, -1 / 3, -1.0], np.float32)
    tr = lv[rng.integers(0, 4, (nchan, nsym))]
    pairs = np.zeros((nchan, nsym, 2), np.float32)
    for c in range(nchan):
        x = np.repeat(tr[c], 2).astype(np.float64)
        x = np.convolve(x, [0.15, 0.8, 0.25, -0.1])[: 2 * nsym] + rng.normal(0, 0.02, 2 * nsym)
        pairs[c] = x.reshape(nsym, 2).astype(np.float32)
    eq = m.Equalizer(ctx, nchan)
    y1 = eq.eq_train(dev(pairs[:, :150]), dev(np.ascontiguousarray(tr[:, :150])))      # known symbols
    y2 = eq.eq_train(dev(pairs[:, 150:]))                                             # decision directed, state carried
    got = torch.cat([y1, y2], 1).cpu().numpy()
    for c in range(nchan):
        exp = eq_oracle(P, pairs[c], tr[c], 150)
        assert bits_eq(got[c], exp), ("equalizer", c, first_diff(got[c], exp))
    eq.close()
    return "equalizer ok"


def eq_oracle(P, pairs, train, nknown):
    import ctypes as C
    st = np.zeros(64, np.float32)
    L = P.L
    L.m17o_eq_open(st.ctypes.data_as(C.c_void_p))
    y = np.zeros(len(pairs), np.float32)
    for i in range(len(pairs)):
        p = np.ascontiguousarray(pairs[i])
        if i < nknown:
            y[i] = L.m17o_eq_train_known(st.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), float(train[i]))
        else:
            y[i] = L.m17o_eq_train_unknown(st.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
    return y


def check_frontend_math(ctx, P, count=1 << 32):
    """fast limiter arithmetic == reference formulation for EVERY possible int16 IQ sample (proof by enumeration)"""
    bad = ctx.selftest_frontend(0, count)
    assert bad == 0, f"{bad} of {count} IQ samples differ between the fast and the IEEE limiter path; first (raw, m_fast, m_ieee, g_fast, g_ieee): {[[hex(int(x)) for x in r] for r in ctx.last_selftest_dump[:8]]}"
    return f"frontend math ok ({count} samples enumerated)"


CHECKS = [
    ("frontend_math", lambda c, P: check_frontend_math(c, P)),
    ("primitives", lambda c, P: check_primitives(c, P)),
    ("viterbi", lambda c, P: check_viterbi(c, P)),
    ("parse_frames", lambda c, P: check_parse_frames(c, P)),
    ("rx_baseband", lambda c, P: check_rx_baseband(c, P, verbose=True)),
    ("rx_chain", lambda c, P: check_rx_chain(c, P, verbose=True)),
    ("rx_chain_split", lambda c, P: check_rx_chain(c, P, nchan=6, seed=23, verbose=True, split=[1, 7, 2, 1, 13])),
    ("rx_packet", lambda c, P: check_rx_packet(c, P, verbose=True)),
    ("rx_afc", lambda c, P: check_rx_afc(c, P)),
    ("rx_equaliser", lambda c, P: check_rx_equaliser(c, P)),
    ("rx_symbols", lambda c, P: check_rx_symbols(c, P)),
    ("rx_bert", lambda c, P: check_rx_bert(c, P)),
    ("decimator", lambda c, P: check_decimator(c, P)),
    ("udp_frames", lambda c, P: check_udp_frames(c, P)),
    ("tx", lambda c, P: check_tx(c, P)),
    ("tx_os80", lambda c, P: check_tx(c, P, nchan=2, F=3, os_=80)),
    ("equalizer", lambda c, P: check_equalizer(c, P)),
]

if __name__ == "__main__":
    import m17_sdr_b200 as m
    m.build()
    ctx = m.Context(0)
    P = Port()
    only = sys.argv[1:]
    fails = 0
    for name, fn in CHECKS:
        if only and name not in only:
            continue
        try:
            print(f"[{name}] {fn(ctx, P)}", flush=True)
        except Exception as e:  # noqa: BLE001
            fails += 1
            print(f"[{name}] FAILED: {type(e).__name__}: {str(e)[:1500]}", flush=True)
            traceback.print_exc(limit=3)
        torch.cuda.synchronize()
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
