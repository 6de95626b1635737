"""Channeliser timing: m17b_chan_run on ncap captures x T blocks (CUDA events).   python benchmarks/chan_bench.py [ncap] [T] [P]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m17_sdr_b200 as m
ncap = int(sys.argv[1]) if len(sys.argv) > 1 else 11
T = int(sys.argv[2]) if len(sys.argv) > 2 else 250
P = int(sys.argv[3]) if len(sys.argv) > 3 else 12
m.build()
ctx = m.Context(0)
wide = torch.randint(-8000, 8000, (ncap, T * 1920 * 25, 2), device="cuda", dtype=torch.int16)
ch = m.Channelizer(ctx, ncap, P)
out = torch.empty((ncap * 96, T * 1920, 2), dtype=torch.int16, device="cuda")
for _ in range(3):
    ch.run(wide, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ch.run(wide, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nt = ncap * T * 1920
print(json.dumps({"captures": ncap, "blocks": T, "taps_per_branch": P, "ms": ms, "channel_s_per_s": ncap * 96 * T / 25 / (ms * 1e-3),
                  "gbs": 484 * nt / (ms * 1e-3) / 1e9, "output_times_per_s": nt / (ms * 1e-3)}))
