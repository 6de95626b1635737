"""CPU enumeration behind demap_lsb (common.cuh): the five-add fp32 form of (float)((double)a - 0.6666) over every finite a >= 0."""
import numpy as np, time
c = 0.6666
c_hi = np.float32(c); c_lo = np.float32(c - float(c_hi)); print(float(c_hi), float(c_lo), c - float(c_hi) - float(c_lo))
nch = np.float32(-c_hi)
def ref(a):   # a float32 >= 0
    return (a.astype(np.float64) - c).astype(np.float32)
def cand(a):
    s = a - c_hi
    e = (nch - s) + a
    return s + (e - c_lo)
t0 = time.time(); bad = 0; badlist = []
CH = 1 << 24
for hi in range(0, 0x7F800000 >> 24):     # all finite non-negative floats
    bits = (np.arange(CH, dtype=np.uint32) + np.uint32(hi << 24))
    a = bits.view(np.float32)
    with np.errstate(all='ignore'):
        r = ref(a); v = cand(a)
    m = r.view(np.uint32) != v.view(np.uint32)
    if m.any():
        bad += int(m.sum())
        if len(badlist) < 10: badlist += [(float(x), float(y), float(z)) for x, y, z in zip(a[m][:3], r[m][:3], v[m][:3])]
print("mismatches", bad, badlist[:10], time.time() - t0)
