"""CPU tier, only where oracle/_ref exists (the build container and any box the prebuilt _ref travelled to):
the C restatement against the UNMODIFIED reference objects on fresh random inputs -- this is what pins the oracle."""
import numpy as np

from m17_oracles import compare_rx, lsf_for
import signals


def feq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_primitives_random(port, ref):
    rng = np.random.default_rng(101)
    for _ in range(50):
        b = bytes(rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8))
        assert port.crc(b) == ref.crc(b)
    assert np.array_equal(port.golay_errtab(), ref.golay_errtab())
    for w in rng.integers(0, 1 << 24, 3000):
        assert port.golay_decode(w) == ref.golay_decode(w)
    for ln in (296, 420, 488):
        for k in range(25):
            s = rng.normal(0, 1, ln).astype(np.float32)
            if k % 4 == 0:
                s = np.round(s * 2) / 2
            assert np.array_equal(port.viterbi(s), ref.viterbi(s))
    for k in range(100):
        sy = (rng.normal(0, 1, 192) * rng.uniform(0.05, 4)).astype(np.float32)
        assert feq(port.demap_frame(sy), ref.demap_frame(sy))
        a, b = port.sync_check(sy[:8]), ref.sync_check(sy[:8])
        assert a[:2] == b[:2] and feq(a[2], b[2])
    for args in ((0.5, 1240, 80), (0.5, 310, 10), (0.5, 62, 2), (0.5, 2480, 80), (0.35, 101, 4)):
        assert feq(port.rrc(*args), ref.rrc(*args))
    # gps_decode (gps.cpp:8-27) on random LSFs, incl. negative latitudes / longitudes and altitudes below the 1500 m offset
    for _ in range(300):
        lsf = rng.integers(0, 256, 30, dtype=np.uint8)
        assert port.gps_decode(lsf) == ref.gps_decode(lsf)
    # m17_dsp_demap_symbol, m17_dsp_decimating_filter (m17_dsp.cpp:35-42,438-449)
    sy = rng.normal(0, 1, 4000).astype(np.float32); mg = rng.uniform(0.05, 6, 4000).astype(np.float32)
    sy[:4] = [0.0, -0.0, 0.6666, -0.6666]
    assert feq(port.demap_symbol(sy, mg), ref.demap_symbol(sy, mg))
    for stride, flen, ln in ((5, 31, 200), (1, 7, 64), (8, 31, 512), (3, 2, 10)):
        x = rng.normal(0, 1, ln + flen).astype(np.float32); cf = rng.normal(0, 0.3, flen).astype(np.float32)
        assert feq(port.decimating_filter(x, cf, stride, ln), ref.decimating_filter(x, cf, stride, ln))


def test_tx_iq_exact(port, ref):
    rng = np.random.default_rng(102)
    lsf = np.stack([lsf_for(port, src=s) for s in ("G4GUO    ", "M17TEST  ", "AB1CD/P  ")])
    pl = rng.integers(0, 256, (3, 6, 16), dtype=np.uint8)
    iq_r, dib_r = ref.tx_stream_run(lsf, pl)
    for c in range(3):
        iq_p, dib_p, _ = port.tx_stream_over(lsf[c], pl[c])
        assert np.array_equal(dib_p, dib_r[c]) and np.array_equal(iq_p[: iq_r.shape[1]], iq_r[c])


def test_rx_chain_random(port, ref):
    X, _ = signals.stream_channels(port, 10, 14, 103, ebn0=[None, 30, 26, 24, 23, 22, 21, 20, 24, 22], f0_max=1500.0)
    a, b = port.rx_run(X, want_soft=True), ref.rx_run(X, want_soft=True)
    assert feq(a.disc, b.disc) and feq(a.soft, b.soft)
    compare_rx(a, b)


def test_rx_baseband_sweep(port, ref):
    D, _ = signals.baseband_channels(port, 8, 16, 104, [None, 12, 10, 8, 6, 4, 2, 0])
    compare_rx(port.rx_run(D, seam=1), ref.rx_run(D, seam=1))


def test_rx_chain_afc(port, ref):
    """AFC on (dsp_nco_mixer + radio_afc, m17_dsp.cpp:390-408, radio.cpp:196-208): the restatement follows the reference's
    NCO phase, delta updates gated by in_frame, and everything downstream bit for bit."""
    X, _ = signals.stream_channels(port, 8, 24, 105, ebn0=[None, 30, 26, 24, None, 28, 24, 22], f0_max=1800.0)
    a, b = port.rx_run(X, afc=True), ref.rx_run(X, afc=True)
    assert feq(a.disc, b.disc)
    compare_rx(a, b)
    off = port.rx_run(X)
    assert not feq(a.disc, off.disc)          # the NCO really moved the spectrum once a frame was acquired


def test_pluto_decimator(port):
    """the int16 31-tap /8 decimator of the Pluto receive path (radio.cpp:18-51,157-177), restatement
    against the reference's own radio.cpp, including the 31-sample history across 1920-sample chunks and int16 extremes."""
    import pytest
    from m17_oracles import RefRadio
    if not RefRadio.available():
        pytest.skip("oracle/_ref/libm17ref_radio.so not built")
    rng = np.random.default_rng(106)
    X = rng.integers(-32768, 32768, (4, 3 * 8 * 1920, 2)).astype(np.int16)
    X[1] = 32767
    X[2, ::2] = -32768
    r = RefRadio().pluto_run(X)
    assert np.array_equal(port.dec_run(X), r)
    assert np.array_equal(port.dec_run(X, parts=[1, 239, 1920, 3600]), r)


def test_udp_frame_format(port):
    """the 54-byte M17-over-UDP frame, restatement against the reference's m17_net.cpp / m17_tx_routines.cpp
    (datagrams captured at sendto)."""
    import pytest
    from m17_oracles import RefRadio
    if not RefRadio.available():
        pytest.skip("oracle/_ref/libm17ref_radio.so not built")
    R = RefRadio()
    rng = np.random.default_rng(107)
    blank = port.encode_call(" ")                # no reflector connected: the gateway's destination is "<name> <module>" = " "
    for k in range(40):
        lsf = port.build_lsf(int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 48)), int(rng.integers(0, 1 << 16)), rng.integers(0, 256, 14, dtype=np.uint8))
        pld = rng.integers(0, 256, 16, dtype=np.uint8)
        sid, fn = int(rng.integers(0, 1 << 16)), int(rng.integers(0, 1 << 16))
        want = R.net_rx_data(sid, lsf, fn, pld)
        got = port.net_pack(sid, lsf, fn, pld, dst=blank)
        assert np.array_equal(got, want), k
        ok, psid, plsf, pfn, ppld = port.net_parse(want)
        rok, posted = R.net_parse(want)
        assert ok and rok and np.array_equal(posted, want)
        assert (psid, pfn) == (sid, fn) and np.array_equal(ppld, pld)
        assert np.array_equal(plsf, R.lich_from_net(want))
        bad = want.copy(); bad[int(rng.integers(0, 54))] ^= 1 << int(rng.integers(0, 8))
        assert port.net_parse(bad)[0] is False and R.net_parse(bad)[0] is False


def test_prbs9_rx_checker(port):
    """m17_prbs9_rx_check (m17_prbs9.cpp:40-64): the restatement against the reference source itself (its verdict lives in file
    statics, exported by oracle/ref/prbs_shim.cpp): clean sequence, errors while in sync, loss of sync, re-acquisition."""
    import ctypes as C
    import os
    import pytest
    from m17_oracles import ORACLE_DIR, _p
    path = os.path.join(ORACLE_DIR, "_ref", "libm17ref_prbs.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libm17ref_prbs.so not built")
    rng = np.random.default_rng(108)
    seq = np.tile(port.prbs9(), 6)
    cases = [seq[:1500].copy(), seq[3:1200].copy(), rng.integers(0, 2, 900).astype(np.uint8)]
    e = seq[:2500].copy(); e[rng.integers(100, 2500, 40)] ^= 1; cases.append(e)
    b = seq[:3000].copy(); b[700:760] ^= 1; b[1500:1530] = rng.integers(0, 2, 30); cases.append(b)
    L = C.CDLL(path)
    L.refp_init()                                   # once: the reference never clears its counters again
    for k, bits in enumerate(cases):
        L.refp_check(_p(bits), C.c_long(len(bits)))
        st = np.zeros(6, np.uint32)
        L.refp_state(_p(st))
        got = port.prbs_check(np.concatenate(cases[:k + 1]))      # a fresh restatement fed the same history
        assert np.array_equal(got[:6], st), (k, got, st)


def test_channelizer_oracle_pinned_at_the_decimator_point(port):
    """The wideband channeliser's restatement (m17o_chan_run) at M = 1, D = 8 with the reference's own 31 taps IS the Pluto
    decimator of radio.cpp:18-40,157-177 -- compared with the reference object itself, in one call and in ragged calls."""
    from m17_oracles import RefRadio
    import pytest
    if not RefRadio.available():
        pytest.skip("oracle/_ref not built")
    RR = RefRadio()
    rng = np.random.default_rng(105)
    X = rng.integers(-32768, 32768, (2, 3 * 8 * 1920, 2)).astype(np.int16)
    X[1, :9000] = 32767; X[1, 9000:20000, 1] = -32768
    ref = RR.pluto_run(X)
    taps = port.dec_taps()
    for c in range(2):
        assert np.array_equal(port.chan_run(X[c], 1, 8, taps), ref[c][None])
        assert np.array_equal(port.chan_run(X[c], 1, 8, taps, parts=[240, 1, 999, 4520]), ref[c][None])
    # and the 96-channel form is self-consistent across call boundaries (history + window phase)
    h = (np.hamming(96 * 4) * 600).astype(np.int16)
    Y = rng.integers(-20000, 20000, (25 * 90, 2)).astype(np.int16)
    assert np.array_equal(port.chan_run(Y, 96, 25, h), port.chan_run(Y, 96, 25, h, parts=[1, 30, 59]))


def test_equaliser_live_chain_probe_pinned(port):
    """The oracle's equaliser option (seam flag 64: eq_open + eq_train_unknown on T/2 pairs between the timing loop and the framer,
    SURVEY 8f rank 3) against the SAME wiring built from the reference's own functions (oracle/ref/eq_shim.cpp: unmodified
    m17_rx_sync_samples fed one sample per call, rx_sync_filter on its statics for the half-symbol output, eq_train_unknown,
    m17_rx_symbols).  One pristine reference process per channel.  This pins the probe whose numbers DESIGN.md quotes; the option
    is off by default in the product, as upstream (wired literally it breaks the chain: tests/eq_live_chain_probe.py)."""
    import ctypes as C
    import os
    import pytest
    from m17_oracles import ORACLE_DIR, _p, _alloc_rx_out, shared_array, BLOCK
    path = os.path.join(ORACLE_DIR, "_ref", "libm17ref_eq.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libm17ref_eq.so not built")
    D, _ = signals.baseband_channels(port, 4, 12, 109, [None, 10, 6, 2])
    Cn, T = D.shape[0], D.shape[1] // 384
    a = _alloc_rx_out(Cn, T, True, False, np.zeros)
    port.L.m17o_rx_run(_p(D), 1 | 64, Cn, T, 4, _p(a.disc), _p(a.nsym), _p(a.syms), a.symcap, _p(a.frames), a.fcap, _p(a.soft),
                       _p(a.events), a.ecap, _p(a.counts))
    b = _alloc_rx_out(Cn, T, False, False, shared_array)
    for c in range(Cn):
        pid = os.fork()
        if pid == 0:
            try:
                L = C.CDLL(path)
                L.ref_init(10)
                L.refe_open()
                L.ref_trace_begin(None, None, _p(b.syms[c]), C.c_long(b.symcap), _p(b.frames[c]), C.c_long(b.fcap), None, _p(b.events[c]), C.c_long(b.ecap))
                for t in range(T):
                    blk = np.ascontiguousarray(D[c, t * 384:(t + 1) * 384])
                    b.nsym[c, t] = L.refe_block(_p(blk), 384)
                k = np.zeros(4, np.int64)
                L.ref_trace_counts(_p(k))
                b.counts[c] = k
                os._exit(0)
            except BaseException:
                os._exit(1)
        _, st = os.waitpid(pid, 0)
        assert os.WIFEXITED(st) and os.WEXITSTATUS(st) == 0, "reference child failed"
    assert np.array_equal(a.nsym, b.nsym)
    for c in range(Cn):
        n = int(a.nsym[c].sum())
        assert n == int(b.counts[c, 1]) and feq(a.syms[c, :n], b.syms[c, :n]), c
        nf = int(a.counts[c, 2])
        assert nf == int(b.counts[c, 2]), (c, nf, b.counts[c])
        for f in ("type", "flags", "lich", "data", "crc", "votes", "frame_errors", "sym_off"):
            assert np.array_equal(a.frames[c, :nf][f], b.frames[c, :nf][f]), (c, f)
    off = port.rx_run(D, seam=1)
    assert not feq(a.syms, off.syms)               # the equaliser really sat in the chain
