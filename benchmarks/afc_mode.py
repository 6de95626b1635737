"""The bench step with AFC on (radio_set_afc_on): block-serial path, k_frontend_afc + per-block sync launches."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
out = {}
for afc in (False, True):
    rx.set_afc(afc)
    for _ in range(2): rx.reset(); rx.m17_dsp_rx(iq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): rx.reset(); rx.m17_dsp_rx(iq)
    e1.record(); torch.cuda.synchronize()
    res = rx.results()
    ok, tot = bench.payload_check(torch, res["frames"], res["nframes"], payload)
    out["afc_on" if afc else "afc_off"] = {"ms_per_step": e0.elapsed_time(e1) / 5, "launches": rx.launches(), "delivered_exact": f"{ok}/{tot}",
                                           "delivered": int(res["stats"][:, 3].sum())}
print(json.dumps(out))
