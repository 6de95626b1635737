"""CPU tier: the C-ABI library loads and exports every symbol include/m17b200.h declares; entry points that need
a device fail loudly (no CPU fallback); the host-side parts of the ABI (filter design, fixed frames) are correct."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import m17_sdr_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "m17_golden.npz"))


@pytest.fixture(scope="module")
def L():
    m.build()
    return m.load()


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "m17b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(m17b_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(L):
    names = declared_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/m17b200.h but not exported by libm17b200.so"
    assert set(names) == set(m.EXPORTS), set(names) ^ set(m.EXPORTS)


def test_version_and_errors(L):
    assert L.m17b_version() == 100
    assert L.m17b_error_string(0) == b"ok" and L.m17b_error_string(-4) == b"unsupported"


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert L.m17b_ctx_create(0, C.byref(h)) == -2 and not h.value          # M17B_E_CUDA, no context
    assert len(L.m17b_last_cuda_error()) > 0
    with pytest.raises(m.M17Error):
        m.Context(0)
    # null handles are rejected rather than silently ignored
    assert L.m17b_dsp_rx(None, None, 1, None) == -1
    assert L.m17b_viterbi_decode(None, None, 296, 1, None, None) == -1


def test_host_side_filter_design(L):
    for key, (ro, nt, sps, gain) in {"rrc_1240_80": (0.5, 1240, 80, None), "rrc_310_10": (0.5, 310, 10, 10.0), "rrc_62_2": (0.5, 62, 2, 1.0)}.items():
        t = np.zeros(nt, np.float32)
        assert L.m17b_build_rrc_filter(t.ctypes.data_as(C.c_void_p), ro, nt, sps) == 0
        if gain is not None:
            assert L.m17b_set_filter_gain(t.ctypes.data_as(C.c_void_p), gain, 1, nt) == 0
        assert np.array_equal(t.view(np.uint32), G[key].view(np.uint32)), key
    assert L.m17b_build_rrc_filter(None, 0.5, 10, 2) == -1


def test_fixed_frames(L):
    d = np.zeros(192, np.uint8)
    assert L.m17b_fmt_preamble(d.ctypes.data_as(C.c_void_p)) == 0 and np.array_equal(d, G["dibits_preamble"])
    assert L.m17b_fmt_eot(d.ctypes.data_as(C.c_void_p)) == 0 and np.array_equal(d, G["dibits_eot"])


def test_record_layout_matches_oracle():
    from m17_oracles import REC_DTYPE
    from m17_sdr_b200.api import REC_DTYPE as PROD
    assert PROD == REC_DTYPE and PROD.itemsize == 64
    assert [PROD.fields[n][1] for n in ("sym_off", "type", "flags", "golay_err", "nbytes", "lich", "data", "crc", "votes", "frame_errors", "variance", "cor")] == \
        [0, 4, 5, 6, 7, 8, 14, 44, 46, 47, 48, 52]


def test_shim_host_helpers_match_the_reference(L):
    """m17_encode_call / m17_decode_call / m17_pack_type / m17_upack_type of the C++ shim (host code, no GPU) against the
    oracle restatement and -- where oracle/_ref exists -- the reference's own m17_bit_utils.cpp."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from m17_oracles import Port, Ref
    subprocess.run(["bash", os.path.join(ROOT, "tests", "cpp", "build.sh")], check=True)
    rng = np.random.default_rng(5)
    alphabet = " ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-/."
    calls = ["G4GUO    ", "AB1CD/P-.", "         ", "M17-M17 C"] + ["".join(alphabet[i] for i in rng.integers(0, 40, 9)) for _ in range(200)]
    words = [int(w) for w in rng.integers(0, 40 ** 9, 200)] + [0xFFFFFFFFFFFF, 0]
    types = [int(w) for w in rng.integers(0, 1 << 16, 100)] + [0x0005, 0x0002, 0xFFFF]
    req = "".join(f"E {c}\n" for c in calls) + "".join(f"D {w:x}\n" for w in words) + "".join(f"T {t:x}\n" for t in types)
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "bin", "shim_host_helpers")], input=req, capture_output=True, text=True, check=True).stdout.split("\n")
    P = Port()
    refs = [P] + ([Ref()] if Ref.available() else [])
    for R in refs:
        for i, c in enumerate(calls):
            assert int(out[i], 16) == R.encode_call(c), (c, out[i])
        for i, w in enumerate(words):
            assert out[len(calls) + i] == "[" + R.decode_call(w) + "]", (hex(w), out[len(calls) + i])
    for i, t in enumerate(types):
        f = out[len(calls) + len(words) + i].split()
        assert [int(x) for x in f[:6]] == [t & 1, (t >> 1) & 3, (t >> 3) & 3, (t >> 5) & 3, (t >> 7) & 15, (t >> 11) & 31] and int(f[6], 16) == t
