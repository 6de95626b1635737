"""Whole-step device time of m17b_dsp_rx on the bench workload under the scheduling knobs that let the stages overlap:
time slices (three streams), front-end CTAs per SM, timing-loop kernel variant, channel groups.  Every configuration's results
are compared byte for byte with the first one.  usage: python benchmarks/overlap_probe.py "SLICE,FE_CTAS,SYNC_IMPL,GROUPS[,OVERLAP,OVERLAP_SLICE]" ..."""
import os, sys, json, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch, bench, m17_sdr_b200 as m
    C = int(os.environ.get("CHANNELS", 1024)); T = int(os.environ.get("BLOCKS", 250))
    ctx = m.Context(0)
    iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
    rx = m.Rx(ctx, C, T)
    for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        rx.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rx.m17_dsp_rx(iq); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = rx.results()
    import hashlib
    h = hashlib.sha1()
    for k in ("frames", "nframes", "nsym", "syms", "events", "nevents", "stats"):
        h.update(np.ascontiguousarray(res[k]).tobytes())
    tl = rx.debug_overlap() if os.environ.get("M17B_OVERLAP", "0") != "0" else None
    print(json.dumps({"ms_median": round(float(np.median(ts)), 4), "ms_min": round(min(ts), 4), "launches": rx.launches(), "sha1": h.hexdigest()[:12], "timeline_ns": tl}))
    sys.exit(0)
first = None
for cfg in sys.argv[1:]:
    sl, fe, impl, grp, ovl, osl = (cfg.split(",") + ["0", "10"])[:6]
    env = dict(os.environ, M17B_SLICE_BLOCKS=sl, M17B_FE_CTAS_PER_SM=fe, M17B_CHAN_GROUPS=grp, M17B_OVERLAP=ovl, M17B_OVERLAP_SLICE=osl)
    if impl != "-1": env["M17B_SYNC_IMPL"] = impl
    out = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
    try:
        r = json.loads(out.stdout.strip().split("\n")[-1])
    except Exception:
        print(cfg, "FAILED", out.stderr[-500:]); continue
    first = first or r["sha1"]
    print(json.dumps({"slice_blocks": int(sl), "fe_ctas_per_sm": int(fe), "sync_impl": int(impl), "groups": int(grp), "overlap": int(ovl), "overlap_slice": int(osl), **r, "same_results": r["sha1"] == first}), flush=True)
