"""Stage times against the number of channels per GPU (clean channels): shows the timing-loop kernel going from latency-bound
(<= 1184 resident channels) to wave-bound, and the variants selected by M17B_SYNC_IMPL."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
bench.EBN0_SWEEP = (None,)
for C, T in ((1024, 250), (2048, 250), (4096, 250), (8192, 125)):
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
    print(os.environ.get("M17B_SYNC_IMPL", "default"), C, T, {k: round(v, 4) for k, v in st.items()}, flush=True)
    rx.close(); del iq
