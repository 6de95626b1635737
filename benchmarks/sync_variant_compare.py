"""Timing-loop / framer kernel variants on the bench workload (1024 channels x 250 blocks): every variant's results are compared
bit-for-bit with the default one-warp-per-channel kernel (records, symbol streams, symbols per block, events, counters), and the
stage times are taken with CUDA events.  usage: python benchmarks/sync_variant_compare.py [impl ...]   (default: 0 65)"""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
impls = [int(a) for a in sys.argv[1:] if a.lstrip("-").isdigit()] or [0, 65]
C = int(os.environ.get("CHANNELS", 1024)); T = int(os.environ.get("BLOCKS", 250))
ctx = m.Context(0)
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
ref = None
for impl in impls:
    os.environ["M17B_SYNC_IMPL"] = str(impl)
    rx = m.Rx(ctx, C, T)
    for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
    rx.set_timing(True)
    for _ in range(10): rx.reset(); rx.m17_dsp_rx(iq)
    torch.cuda.synchronize()
    st = {}
    for i in range(10):
        s = rx.stage_ms(i)
        for k in s: st[k] = st.get(k, 0) + s[k] / 10
    rx.set_timing(False)
    # a split call sequence from a fresh state must give the same final records as well
    rx.reset(); rx.m17_dsp_rx(iq)
    res = rx.results()
    keys = ("frames", "nframes", "nsym", "syms", "events", "nevents", "stats")
    if ref is None:
        ref = {k: np.ascontiguousarray(res[k]).copy() for k in keys}
        same = "reference"
    else:
        bad = [k for k in keys if np.ascontiguousarray(res[k]).tobytes() != ref[k].tobytes()]
        same = "identical" if not bad else "DIFFERS in " + ",".join(bad)
    print(json.dumps({"impl": impl, "channels": C, "blocks": T, "stage_ms": {k: round(v, 4) for k, v in st.items()}, "vs_impl_%d" % impls[0]: same}), flush=True)
    rx.close()
