"""CPU tier: the C restatement (oracle/libm17oracle.so) against golden vectors produced by the reference's own
objects (tests/golden/make_golden.py) and against the known-answer table of SURVEY.md section 4."""
import ctypes as C

import numpy as np
import pytest

from m17_oracles import EV_DTYPE, REC_DTYPE, F_DELIVERED
import os

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "m17_golden.npz"))


def feq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_kat_crc(port):
    # identical to the M17 specification's published CRC vectors
    assert [port.crc(m) for m in (b"", b"A", b"123456789", bytes(range(256)))] == [0xFFFF, 0x206E, 0x772B, 0x1C31]
    assert port.crc(bytes(G["crc_msg4"])) == int(G["crc_out"][4])


def test_kat_golay(port):
    assert [port.golay_encode(d) for d in (0, 1, 0x800, 0xABC, 0xFFF)] == [0, 0x0018EB, 0x800C75, 0xABC23C, 0xFFFFFF]
    enc = port.golay_encode(0xABC)
    assert port.golay_decode(enc ^ 0x111000) == (0xABC, 3)
    assert port.golay_decode(enc ^ 0x000007) == (0xABC, 3)
    assert port.golay_decode(enc ^ 0x00F000) == (0x0F3, 4)      # mis-correction reproduced
    assert port.golay_decode(enc ^ 0x101011) == (0x7BD, 4)
    assert np.array_equal(np.array([port.golay_encode(d) for d in range(4096)], np.uint32), G["golay_enc"])
    tab = port.golay_errtab()
    assert np.array_equal(tab, G["golay_errtab"])
    assert np.bincount(tab >> 12, minlength=5).tolist() == [1, 24, 276, 2024, 1771]
    got = [port.golay_decode(int(w)) for w in G["golay_dec_in"]]
    assert [g[0] for g in got] == G["golay_dec_data"].tolist() and [g[1] for g in got] == G["golay_dec_err"].tolist()


def test_kat_conv_viterbi(port):
    out = port.conv_encode_1(np.array([1, 0, 1, 1, 0, 0, 1, 0], np.uint8))
    assert "".join(map(str, out)) == "110110001111101001101100"
    for nb in (18, 26, 30):
        assert np.array_equal(np.stack([port.conv_encode_8(r) for r in G[f"conv8_in_{nb}"]]), G[f"conv8_out_{nb}"])
    assert np.array_equal(np.stack([port.conv_encode_1(r) for r in G["conv1_in"]]), G["conv1_out"])
    bits = port.viterbi(np.zeros(296, np.float32))               # pure tie-break
    assert len(bits) == 148 and bits.sum() == 144 and "".join(map(str, bits[:16])) == "0111111111111111" and "".join(map(str, bits[-8:])) == "11111000"
    for ln in (296, 420, 488):
        assert np.array_equal(np.stack([port.viterbi(r) for r in G[f"vit_in_{ln}"]]), G[f"vit_out_{ln}"])


def test_puncture_interleave_randomise(port):
    for p, ln in ((1, 488), (2, 296), (3, 420)):
        assert np.array_equal(port.punc(p, G[f"punc_in_{p}"]), G[f"punc_out_{p}"])
        assert feq(port.depunc(p, G[f"depunc_in_{p}"], ln), G[f"depunc_out_{p}"])
    assert np.array_equal(port.interleave(G["bits368"]), G["interleave_out"])
    assert feq(port.deinterleave(G["soft368"]), G["deinterleave_out"])
    assert np.array_equal(port.derand_bits(G["bits368"]), G["derand_bits_out"])
    assert feq(port.derand_soft(G["soft368"]), G["derand_soft_out"])
    assert np.array_equal(port.derand_bytes(np.zeros(46, np.uint8)), G["derand_bytes_out"])
    # QPP is a bijection and an involution (same scatter serves TX and RX)
    perm = port.interleave(np.arange(368) % 251).astype(int)
    idx = np.array([(45 * i + 92 * i * i) % 368 for i in range(368)])
    assert sorted(idx) == list(range(368)) and np.array_equal(idx[idx], np.arange(368)) and perm is not None


def test_demap_sync_filters(port):
    assert feq(np.stack([port.demap_frame(r) for r in G["demap_in"]]), G["demap_out"])
    got = [port.sync_check(r) for r in G["sync_in"]]
    assert [g[0] for g in got] == G["sync_type"].tolist() and [g[1] for g in got] == G["sync_votes"].tolist()
    assert feq(np.array([g[2] for g in got], np.float32), G["sync_var"])
    assert feq(port.rrc(0.5, 1240, 80), G["rrc_1240_80"])
    assert feq(port.set_gain(port.rrc(0.5, 310, 10), 10, 1, 310), G["rrc_310_10"])
    assert feq(port.set_gain(port.rrc(0.5, 62, 2), 1.0, 1, 62), G["rrc_62_2"])
    seq = port.prbs9()
    assert np.array_equal(seq, G["prbs9"]) and "".join(map(str, seq[:16])) == "0000100011000010"
    assert port.encode_call("G4GUO    ") == 0x0000025EA29F and port.encode_call("G4GUO/P  ") == 0x00102C8DA29F
    assert [port.encode_call(c) for c in ("G4GUO    ", "G4GUO/P  ", "AB1CD-9 .", "M17      ")] == G["call_enc"].tolist()
    assert port.decode_call(port.encode_call("AB1CD-9 .")) == "AB1CD-9 ." and port.decode_call(0xFFFFFFFFFFFF) == "BROADCAST"


def test_equaliser(port):
    st = np.zeros(64, np.float32)
    port.L.m17o_eq_open(st.ctypes.data_as(C.c_void_p))
    y = np.zeros(300, np.float32)
    for i in range(300):
        p = np.ascontiguousarray(G["eq_pairs"][i])
        if i < 120:
            y[i] = port.L.m17o_eq_train_known(st.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), float(G["eq_train"][i]))
        else:
            y[i] = port.L.m17o_eq_train_unknown(st.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
    assert feq(y, G["eq_out"])


def hexd(d):
    return "".join("%X" % ((d[i] << 2) | d[i + 1]) for i in range(0, 192, 2))


def test_tx_frames_and_modulator(port):
    lsf = port.build_lsf(0xFFFFFFFFFFFF, port.encode_call("G4GUO    "), 0x0005)
    assert bytes(lsf).hex().upper() == "FFFFFFFFFFFF0000025EA29F0005" + "00" * 14 + "8511" and np.array_equal(lsf, G["lsf"])
    d = port.fmt_stream_frames(lsf, G["stream_payload"])
    assert np.array_equal(d, G["dibits_stream"])
    # SURVEY section 4 stream-frame KATs (payload 00..0F, fn 0/1)
    assert hexd(port.fmt_stream_frames(lsf, np.tile(np.arange(16, dtype=np.uint8), (2, 1)))[0]) == \
        "FF5D51A8D399A5CEA4505F2DF4771BFB16B441149EBA2944E37D591D4EA94FE8967167EBBB5E87F57385C996F2FAA744"
    assert hexd(port.fmt_lsf(lsf)) == "55F757BDAAD182D6AD6BFA26D690DAD0F5774C8A5C05D101E4666C373BD804EA4AF2198AD282F3348797F3186CA858C2"
    assert np.array_equal(port.fmt_lsf(lsf), G["dibits_lsf"])
    assert np.array_equal(port.fmt_packet(bytes(G["packet_chunk"]), 1, 22), G["dibits_packet"])
    # SURVEY section 4 packet KAT: 20-byte payload 01,04,..,3A + its CRC 8F3A, meta byte 0xD8 (EOF, 22 bytes used)
    pay = bytes(range(1, 0x3B, 3))
    crc = port.crc(pay)
    assert crc == 0x8F3A
    assert hexd(port.fmt_packet(pay + bytes([crc >> 8, crc & 255]), 1, 22)) == \
        "75FFA5745561B2C9B5A9803E8EDAB80D4A0302CC01E839985EC38E4BB51F7528C71695C014A6329A52EBC33E67B803C8"
    assert np.array_equal(port.fmt_bert(3), G["dibits_bert"])
    assert np.array_equal(port.fmt_preamble(), G["dibits_preamble"]) and np.array_equal(port.fmt_eot(), G["dibits_eot"])
    iq, dib, _ = port.tx_stream_over(lsf, G["stream_payload"][:2], lead=1, npre=1, tail=1)
    assert np.array_equal(dib, G["tx_dibits"])
    assert np.array_equal(iq[: len(G["tx_iq"])], G["tx_iq"])      # same libm: IQ bit-exact between port and reference


def _cmp_rx(o, pre, Cn):
    assert np.array_equal(o.counts, G[pre + "_counts"]) and np.array_equal(o.nsym, G[pre + "_nsym"])
    fr = G[pre + "_frames"].view(REC_DTYPE).reshape(Cn, -1)
    ev = G[pre + "_events"].reshape(Cn, -1, 2)
    for c in range(Cn):
        ns, nf, ne = (int(x) for x in o.counts[c, 1:4])
        assert feq(o.syms[c, :ns], G[pre + "_syms"][c, :ns])
        a, b = o.frames[c, :nf], fr[c, :nf]
        for name in REC_DTYPE.names:
            if name != "rsvd":
                assert feq(a[name], b[name]), (c, name)
        assert np.array_equal(o.events[c, :ne].view(np.int32).reshape(-1, 2), ev[c, :ne])


def test_rx_chain_golden(port):
    o = port.rx_run(G["rx_iq"], seam=0)
    _cmp_rx(o, "rx", 4)
    assert feq(o.disc[0], G["rx_disc_c0"])
    # loopback property: every delivered payload of the clean channel equals what was sent
    f = o.frames[0, : o.counts[0, 2]]
    d = f[(f["type"] == 2) & ((f["flags"] & F_DELIVERED) != 0)]
    assert len(d) >= 1
    for r in d:
        fn = (int(r["data"][0]) << 8) | int(r["data"][1])
        assert np.array_equal(r["data"][2:18], G["rx_payload"][0, fn])


def test_rx_baseband_golden(port):
    _cmp_rx(port.rx_run(G["bb_disc"], seam=1), "bb", 3)


def test_rx_edge_cases(port):
    # empty-ish and ragged inputs: a single block, noise only, and block-by-block == one shot
    rng = np.random.default_rng(5)
    x = G["rx_iq"][:2]
    one = port.rx_run(np.ascontiguousarray(x[:, :1920]), seam=0)
    assert one.counts[:, 0].tolist() == [1, 1]
    full = port.rx_run(x, seam=0)
    T = x.shape[1] // 1920
    assert np.array_equal(full.nsym[:, :1], one.nsym)
    noise = (rng.normal(0, 3000, (1, 1920 * 6, 2))).astype(np.int16)
    noise[(noise[..., 0] == 0) & (noise[..., 1] == 0)] = 1
    o = port.rx_run(noise, seam=0)
    assert o.counts[0, 0] == 6 and 6 * 190 <= o.counts[0, 1] <= 6 * 194 and T > 1


def test_golden_ext(port):
    """Fixtures made by the unmodified reference for the pieces around the hot path (tests/golden/make_golden_ext.py): AFC loop,
    Pluto /8 decimator, M17-over-UDP frames, PRBS9 receive checker."""
    import os
    from m17_oracles import compare_rx  # noqa: F401
    here = os.path.dirname(os.path.abspath(__file__))
    G = np.load(os.path.join(here, "golden", "m17_golden.npz"))
    E = np.load(os.path.join(here, "golden", "m17_golden_ext.npz"))
    o = port.rx_run(np.ascontiguousarray(G["rx_iq"]), seam=0, afc=True)
    assert np.array_equal(o.nsym, E["afc_nsym"]) and np.array_equal(o.counts, E["afc_counts"])
    for c in range(o.nsym.shape[0]):
        ns, nf, ne = (int(x) for x in E["afc_counts"][c, 1:4])
        assert np.array_equal(o.syms[c, :ns].view(np.uint32), E["afc_syms"][c, :ns].view(np.uint32)), c
        assert np.array_equal(o.frames[c, :nf].view(np.uint8).reshape(nf, 64)[:, :56], E["afc_frames"][c, :nf, :56]), c
        assert np.array_equal(o.events[c, :ne].view(np.int32).reshape(ne, 2), E["afc_events"][c, :ne]), c
    assert np.array_equal(o.disc[1].view(np.uint32), E["afc_disc_c1"].view(np.uint32))
    assert np.array_equal(port.dec_run(E["dec_in"]), E["dec_out"])
    assert np.array_equal(port.dec_run(E["dec_in"], parts=[4, 1916, 1920]), E["dec_out"])
    blank = port.encode_call(" ")
    for i in range(len(E["udp_sid"])):
        assert np.array_equal(port.net_pack(int(E["udp_sid"][i]), E["udp_lsf"][i], int(E["udp_fn"][i]), E["udp_pld"][i], dst=blank), E["udp_frames"][i])
        ok, sid, lsf, fn, pld = port.net_parse(E["udp_frames"][i])
        assert ok and sid == E["udp_sid"][i] and fn == E["udp_fn"][i] and np.array_equal(pld, E["udp_pld"][i]) and np.array_equal(lsf, E["udp_lich"][i])
    assert np.array_equal(port.prbs_check(E["prbs_bits"])[:6], E["prbs_state"])
    # m17_dsp_demap_symbol / m17_dsp_decimating_filter outputs of the reference itself
    assert np.array_equal(port.demap_symbol(E["dsym_in"], E["dsym_mag"]).view(np.uint32), E["dsym_out"].view(np.uint32))
    got = np.stack([port.decimating_filter(r, E["dfil_coffs"], 5, 200) for r in E["dfil_in"]])
    assert np.array_equal(got.view(np.uint32), E["dfil_out"].view(np.uint32))
    assert np.array_equal(np.array([port.gps_decode(l) for l in E["gps_lsf"]], np.float64), E["gps_out"])
