"""The bench workload's RX step by itself (for ncu launch lists / captures of single kernels).  usage: python benchmarks/rx_step.py [steps] [groups]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
ctx = m.Context(0)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000)
rx = m.Rx(ctx, C, T)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if len(sys.argv) > 2: rx.set_chan_groups(int(sys.argv[2]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rx.reset(); rx.m17_dsp_rx(iq)
torch.cuda.synchronize(); e0.record()
for _ in range(n): rx.reset(); rx.m17_dsp_rx(iq)
e1.record(); torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / n)
