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
