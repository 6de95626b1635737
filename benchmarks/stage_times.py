"""Per-stage device times of the bench step (front end, sync/framer, decode, post) -- the library marks the stage boundaries with
CUDA events.  Honours M17B_SYNC_IMPL / M17B_FE_IMPL / M17B_LIB for comparing kernel variants."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
rx.set_timing(True)
for _ in range(10): rx.reset(); rx.m17_dsp_rx(iq)
torch.cuda.synchronize()
st = {}
for i in range(10):
    s = rx.stage_ms(i)
    for k in s: st[k] = st.get(k, 0) + s[k] / 10
print(os.environ.get("M17B_LIB", "default"), {k: round(v, 4) for k, v in st.items()})
