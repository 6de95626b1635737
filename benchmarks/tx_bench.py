"""TX chain timing: m17b_fmt_stream_frames + m17b_mod_dibits for C channels x F stream frames (CUDA events).
    python benchmarks/tx_bench.py [C] [F] [os]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m17_sdr_b200 as m  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = int(sys.argv[2]) if len(sys.argv) > 2 else 250
OS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
m.build()
ctx = m.Context(0)
tx = m.Tx(ctx, C, OS)
g = torch.Generator(device="cuda"); g.manual_seed(7)
payload = torch.randint(0, 256, (C, F, 16), generator=g, device="cuda", dtype=torch.int32).to(torch.uint8)
lsf = torch.randint(0, 256, (C, 30), generator=g, device="cuda", dtype=torch.int32).to(torch.uint8)
tx.set_lsf(lsf)
iq = torch.empty((C, F * 192 * OS, 2), dtype=torch.int16, device="cuda")
out = {}
for name, fn in (("fmt", lambda: tx.m17_fmt_add_stream_frame(payload)),
                 ("mod", None), ("fmt+mod", None)):
    dib = tx.m17_fmt_add_stream_frame(payload).reshape(C, F * 192)
    if name == "mod":
        fn = lambda: tx.m17_mod_dibits(dib, out=iq)                                            # noqa: E731
    if name == "fmt+mod":
        fn = lambda: tx.m17_mod_dibits(tx.m17_fmt_add_stream_frame(payload).reshape(C, F * 192), out=iq)   # noqa: E731
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[name] = {"ms": round(ms, 4), "Mframes_per_s": round(C * F / ms / 1e3, 2)}
    if name == "mod":
        d = tx.debug_scan()
        nsymc = F * 192 * d["ctas"]
        out["scan"] = {k.replace("_cycles", "") + "_cyc_per_sym": round(v / nsymc, 1) for k, v in d.items() if k.endswith("_cycles")}
        out["scan"].update(chunks=d["chunks"], ctas=d["ctas"])
out["config"] = {"channels": C, "frames": F, "os": OS, "iq_GB": iq.numel() * 2 / 1e9}
out["mod_GBs_out"] = round(iq.numel() * 2 / (out["mod"]["ms"] * 1e-3) / 1e9, 1)
print(json.dumps(out))
