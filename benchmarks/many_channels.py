"""BASELINE configs[4]-sized run on ONE GPU: many channels x few blocks (default 65 536 x 25 = 1.64 M channel-frames, 12.6 GB of IQ):
device-resident stage times, throughput, and the loopback sanity check (delivered payloads equal what was sent)."""
import argparse, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=65536)
ap.add_argument("--blocks", type=int, default=25)
a = ap.parse_args()
ctx = m.Context(0)
C, T = a.channels, a.blocks
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=77)
rx = m.Rx(ctx, C, T)
for _ in range(2): rx.reset(); rx.m17_dsp_rx(iq)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): rx.reset(); rx.m17_dsp_rx(iq)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
rx.set_timing(True)
for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
torch.cuda.synchronize()
st = {}
for i in range(3):
    s = rx.stage_ms(i)
    for k in s: st[k] = st.get(k, 0) + s[k] / 3
res = rx.results()
ok, tot = bench.payload_check(torch, res["frames"], res["nframes"], payload)
print(json.dumps({"channels": C, "blocks": T, "ms_per_step": ms, "channel_s_per_s": C * T / 25.0 / (ms * 1e-3), "stages_ms": {k: round(v, 3) for k, v in st.items()},
                  "delivered_payloads_exact": f"{ok}/{tot}", "frames": int(res["stats"][:, 0].sum())}))
