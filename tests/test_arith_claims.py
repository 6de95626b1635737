"""CPU tier: the integer shortcuts the front-end kernel takes, restated in numpy (the GPU tier enumerates the kernel itself:
m17b_selftest_frontend over all 2^32 IQ words, m17b_selftest_limiter over every normal float s)."""
import numpy as np


def test_minus_half_y_by_integer_add():
    """frontend.cuh fe_norm_pair: bits(y) + 0x7F800000 (mod 2^32) is exactly -y/2 when y and y/2 are normal.  y = rsqrt(s) for a
    normal float s lies in [2^-64, 2^63.5]."""
    rng = np.random.default_rng(7)
    e = rng.integers(127 - 64, 127 + 64, 2_000_000).astype(np.uint32)
    m = rng.integers(0, 1 << 23, 2_000_000).astype(np.uint32)
    y = ((e << 23) | m).view(np.float32)
    edge = np.array([2.0 ** -64, 2.0 ** 63, 1.0, 0.70710677, 33333.332, np.nextafter(np.float32(2.0 ** 63), np.float32(0))], np.float32)
    for v in (y, edge):
        got = (v.view(np.uint32) + np.uint32(0x7F800000)).view(np.float32)
        assert np.array_equal(got.view(np.uint32), (np.float32(-0.5) * v).view(np.uint32))


def test_prmt_sign_extension_selector():
    """frontend.cuh fe_limit_pair: prmt.b32 with selector 0x9910 builds {b0, b1, sign(b1), sign(b1)} = the sign-extended low half of
    the raw IQ word; I2FP.F32.S32 of it equals (float)(int16) for every value."""
    raw = (np.arange(1 << 16, dtype=np.uint32) | np.uint32(0xA5A50000))
    b0, b1 = raw & 0xFF, (raw >> 8) & 0xFF
    sgn = np.where(b1 & 0x80, 0xFF, 0x00).astype(np.uint32)
    v = (b0 | (b1 << 8) | (sgn << 16) | (sgn << 24)).view(np.int32)
    assert np.array_equal(v, (raw & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32))
    assert np.array_equal(v.astype(np.float32), (raw & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.float32))


def test_int16_scaling_split():
    """frontend.cuh: (float)((double)x * 0.00003) == fma(x, c_hi, x * c_lo) with c_hi + c_lo the two-float split of 0.00003, for every
    int16 x (fma emulated in float64: both products are exact there -- 16 + 24 significant bits -- and their sum has few enough bits
    that rounding it once to float32 is the fma's own rounding of the exact value unless the double sum itself rounds, which the
    assertion on exactness below rules out)."""
    x = np.arange(-32768, 32768, dtype=np.float64)
    c = 0.00003
    c_hi = np.float32(c)
    c_lo = np.float32(c - float(c_hi))
    ref = (x * c).astype(np.float32)
    lo = (x.astype(np.float32) * c_lo)                       # rounded fp32 product, as the kernel's FMUL2 does
    exact = x * float(c_hi) + lo.astype(np.float64)          # fma's exact argument: representable in float64 here
    from fractions import Fraction
    for k in (0, 1, 12345, 40000, 65535):                    # spot-check that the float64 sum really is exact
        assert Fraction(float(exact[k])) == Fraction(float(x[k])) * Fraction(float(c_hi)) + Fraction(float(lo[k]))
    assert np.array_equal(exact.astype(np.float32).view(np.uint32), ref.view(np.uint32))
