"""Channel-group pipelining of m17b_dsp_rx on the bench workload (1024 channels x 250 blocks): whole-step device time with the
batch run as G independent channel groups on their own streams, and a bit-for-bit comparison of every result with G = 1.
usage: python benchmarks/chan_groups.py [G ...]   (default 1 2 3 4 6 8)"""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
groups = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]
C = int(os.environ.get("CHANNELS", 1024)); T = int(os.environ.get("BLOCKS", 250))
ctx = m.Context(0)
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
ref = None
keys = ("frames", "nframes", "nsym", "syms", "events", "nevents", "stats")
for G in groups:
    rx.set_chan_groups(G)
    for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        rx.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rx.m17_dsp_rx(iq); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = rx.results()
    if ref is None:
        ref = {k: np.ascontiguousarray(res[k]).copy() for k in keys}; same = "reference"
    else:
        bad = [k for k in keys if np.ascontiguousarray(res[k]).tobytes() != ref[k].tobytes()]
        same = "identical" if not bad else "DIFFERS in " + ",".join(bad)
    print(json.dumps({"groups": G, "channels": C, "blocks": T, "sync_impl": os.environ.get("M17B_SYNC_IMPL", "auto"), "ms_per_step_median": round(float(np.median(ts)), 4),
                      "ms_min": round(min(ts), 4), "Mchannel_s_per_s": round(C * T * 0.04 / np.median(ts) / 1e3, 3), "vs_groups_%d" % groups[0]: same}), flush=True)
