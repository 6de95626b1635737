#!/usr/bin/env python
"""BASELINE.json configs[3]: batched Viterbi-only microbenchmark -- 1M punctured K=5 r=1/2 soft-decision frames
(stream/P2 by default), int8-grid quantised soft values, ACS throughput against the fp32 ALU roofline.
Prints one JSON line.  Inputs are generated on the GPU with the library's own encoder/puncturer."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m17_sdr_b200 as m  # noqa: E402

CFG = {1: (30, 488, 244), 2: (18, 296, 148), 3: (26, 420, 210)}     # pattern -> payload bytes, coded bits used, trellis steps


def make_frames(ctx, pattern, n, ebn0_db, seed=3):
    nb, full, _ = CFG[pattern]
    g = torch.Generator(device=ctx.device); g.manual_seed(seed)
    data = torch.randint(0, 256, (n, nb), generator=g, device=ctx.device, dtype=torch.int32).to(torch.uint8)
    if pattern == 3:
        # packet frames carry eof<<7 | n<<2 in the last byte (m17_tx_routines.cpp:210): its two low bits are 0, which is what lets
        # the reference cut the coded frame at 420 bits (2 of the 4 tail steps) and still trace back from state 0
        data[:, -1] &= 0xFC
    coded = ctx.m17_conv_encode_8(data)[:, :full].contiguous()
    kept = ctx.m17_punc(pattern, coded).float() * 2 - 1
    if ebn0_db is not None:
        sigma = (1.0 / (2 * 0.5 * 10 ** (ebn0_db / 10))) ** 0.5            # rate 1/2, +-1 symbols
        kept = kept + sigma * torch.randn(kept.shape, generator=g, device=ctx.device)
    soft = torch.round(kept.clamp(-2, 2) * 32) / 32                        # int8 grid presented as float
    return data, soft.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1 << 20)
    ap.add_argument("--pattern", type=int, default=2)
    ap.add_argument("--ebn0", type=float, default=3.0)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    m.build()
    ctx = m.Context(0)
    data, soft = make_frames(ctx, a.pattern, a.frames, a.ebn0)
    for _ in range(3):
        out = ctx.viterbi_punctured(a.pattern, soft)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = ctx.viterbi_punctured(a.pattern, soft)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    steps = CFG[a.pattern][2]
    ops = a.frames * steps * 52                        # 4 branch-metric adds + 16 x (2 add + 1 compare) per step
    import numpy as np
    ber = float(np.unpackbits(torch.bitwise_xor(out, data).cpu().numpy()).mean())
    peak = 148 * 128 * 1.965e9                         # fp32 lanes x clock: non-tensor instruction roof
    print(json.dumps({"workload": f"{a.frames} punctured frames, pattern P{a.pattern}, Eb/N0 {a.ebn0} dB", "ms": ms,
                      "frames_per_s": a.frames / (ms * 1e-3), "acs_ops_per_s": ops / (ms * 1e-3), "alu_roof_ops_per_s": peak,
                      "frac_of_alu_roof": ops / (ms * 1e-3) / peak, "bit_error_rate": ber,
                      "bytes_in_per_frame": soft.shape[1] * 4, "gbs_in": soft.numel() * 4 / (ms * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
