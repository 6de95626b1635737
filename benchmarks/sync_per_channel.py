"""Per-channel cost of the timing-loop kernel on the bench workload: cycles and speculation rounds by noise class."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
for _ in range(2): rx.reset(); rx.m17_dsp_rx(iq)
torch.cuda.synchronize()
d = rx.debug_sync().cpu().numpy()
res = rx.results()
fr = res["stats"]
NAMES = tuple("clean" if e is None else f"{e:g} dB" for e in bench.EBN0_SWEEP)
for k, name in enumerate(NAMES):
    cyc = d[k::5, 0] / 1e6; rnd = d[k::5, 1]
    print(f"{name}: Mcycles min/med/max {cyc.min():.2f}/{np.median(cyc):.2f}/{cyc.max():.2f}  rounds med/max {int(np.median(rnd))}/{int(rnd.max())}  frames med {int(np.median(fr[k::5, 0]))} los max {int(fr[k::5, 5].max())}")
if os.environ.get("M17B_SYNC_IMPL") == "64" and d[:, 2:8].sum() > 0:
    for j, name in enumerate(("A staging", "A timing loop", "A wait for B", "A carry+publish", "B wait for A", "B emission+framer")):
        print(f"phase {name}: median {np.median(d[:, 2 + j]) / T:.0f} cycles/block")
if os.environ.get("M17B_SYNC_IMPL") == "65":
    rr = d[:, 1]
    full, part, miss, unl = rr & 255, (rr >> 8) & 255, (rr >> 16) & 255, (rr >> 24) & 255
    print("blocks of the reporting warp (every second block): predicted + no trip", int(full.sum()), " predicted + trip", int(part.sum()),
          " mispredicted while locked", int(miss.sum()), " unlocked", int(unl.sum()))
    for k, name in enumerate(NAMES):
        print(f"  {name}: median per channel full/trip/miss/unlocked {int(np.median(full[k::5]))}/{int(np.median(part[k::5]))}/{int(np.median(miss[k::5]))}/{int(np.median(unl[k::5]))}")
if os.environ.get("M17B_SYNC_IMPL") == "65" and d[:, 2:8].sum() > 0:
    # the warp that ran the channel's last block reports its own clocks: it handled every second block
    for j, name in enumerate(("staging", "speculative dot products + votes", "wait for the hand-off", "resolve: trip test / commit / fallback rounds", "emission + framer", "hand-off")):
        print(f"phase {name}: median {np.median(d[:, 2 + j]) / (T / 2):.0f} cycles per own block")
elif d[:, 7].sum() > 0:
    # -DM17B_PHASE_UNLOCKED: phases accumulated over the blocks a channel entered unlocked only (count in slot 5)
    nb = d[:, 7].astype(float)
    sel = nb >= 20
    print(f"blocks entered unlocked: total {int(nb.sum())}, channels with >= 20: {int(sel.sum())}")
    for j, name in enumerate(("staging", "timing loop", "emission", "framer", "carry")):
        print(f"unlocked-block phase {name}: median over those channels {np.median(d[sel, 2 + j] / nb[sel]):.0f} cycles/block")
    print(f"unlocked block total: median {np.median(d[sel, 2:7].sum(1) / nb[sel]):.0f} cycles/block")
elif d[:, 2:7].sum() > 0:
    tot = d[:, 0].astype(float)
    for j, name in enumerate(("staging", "timing loop", "emission", "framer", "carry")):
        print(f"phase {name}: median {np.median(d[:, 2 + j]) / T:.0f} cycles/block ({100 * np.median(d[:, 2 + j] / tot):.0f} %)")
w = int(np.argmax(d[:, 0]))
if os.environ.get("M17B_SYNC_IMPL", "0") in ("0", "") and d[:, 2:7].sum() > 0 and d[:, 7].sum() == 0:
    print("slowest channel, cycles per block by phase:", {name: int(d[w, 2 + j] / T) for j, name in enumerate(("staging", "timing loop", "emission", "framer", "carry"))})
print("slowest channel", w, "class", w % 5, "Mcycles", d[w, 0] / 1e6, "rounds", d[w, 1], "frames", fr[w, 0], "aos", fr[w, 4], "los", fr[w, 5])
