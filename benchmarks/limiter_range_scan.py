import sys; sys.path.insert(0,'/root/repo')
import m17_sdr_b200 as m
ctx = m.Context(0)
tot = 0
for e in range(1, 255):
    bad = ctx.selftest_limiter(e << 23, 1 << 23)
    tot += bad
    if bad: print("exp", e - 127, "bad", bad, [[hex(int(x)) for x in r] for r in ctx.last_selftest_dump[:3]])
print("total bad over normals", tot)
