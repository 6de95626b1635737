"""End-to-end host path by itself: m17b_dsp_rx_host from pinned host IQ (1024 ch x 250 blocks), records compared with the device path."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
rx.m17_dsp_rx(iq); ref = rx.results()
h = iq.cpu().pin_memory()
for _ in range(2): rx.reset(); fh, nh = rx.m17_dsp_rx_host(h)
torch.cuda.synchronize()
ts = []
for _ in range(8):
    rx.reset(); torch.cuda.synchronize(); t0 = time.perf_counter(); fh, nh = rx.m17_dsp_rx_host(h); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
fr = fh.numpy().view(m.REC_DTYPE).reshape(C, -1)
same = np.array_equal(nh.numpy(), ref["nframes"]) and all(np.array_equal(fr[c, :nh[c]], ref["frames"][c, :nh[c]]) for c in range(C))
print("e2e ms median", round(float(np.median(ts)), 3), "min", round(min(ts), 3), "records equal device path:", same)
