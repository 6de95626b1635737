"""Per-GPU step time of the bench workload with no inter-GPU traffic at all, every rank on its own GPU at the same time (launch with
torchrun): separates GPU-to-GPU variation of the box from the cost of the per-step counter all-reduce in bench.py."""
import os, sys, json, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m17_sdr_b200 as m
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
C, T = 1024, 250
iq, payload = bench.make_workload(ctx, m, torch, C, T, seed=1000 + rank)
rx = m.Rx(ctx, C, T)
for _ in range(3): rx.reset(); rx.m17_dsp_rx(iq)
torch.cuda.synchronize()
if world > 1: dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): rx.reset(); rx.m17_dsp_rx(iq)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out = [None] * world
if world > 1: dist.all_gather_object(out, round(ms, 4))
else: out = [round(ms, 4)]
if rank == 0: print(json.dumps({"per_gpu_ms_no_collective": out, "max": max(out), "min": min(out)}))
if world > 1: dist.destroy_process_group()
