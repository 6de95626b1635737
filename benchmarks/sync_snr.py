"""Stage times against the noise level of the channels (all channels at one Eb/N0, then the bench mix): the timing-loop kernel's
time is set by its slowest channel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
for sweep in ((None,), (30.0,), (24.0,), (22.0,), (22.0, 24.0, 26.0, 30.0, None)):
    bench.EBN0_SWEEP = sweep
    iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
    rx = m.Rx(ctx, C, T)
    for _ in range(2): rx.reset(); rx.m17_dsp_rx(iq)
    rx.set_timing(True)
    for _ in range(5): rx.reset(); rx.m17_dsp_rx(iq)
    torch.cuda.synchronize()
    st = {}
    for i in range(5):
        s = rx.stage_ms(i)
        for k in s: st[k] = st.get(k, 0) + s[k] / 5
    stats = rx.results()["stats"].sum(0)
    print(os.environ.get("M17B_SYNC_IMPL", "default"), sweep, {k: round(v, 4) for k, v in st.items()}, "frames", int(stats[0]), "aos", int(stats[4]), "los", int(stats[5]), flush=True)
    rx.close(); del iq
