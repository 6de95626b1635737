#!/usr/bin/env python
"""Pluto /8 front-end decimator on its own: nchan channels x nblocks 40-ms blocks of 384 kS/s int16 IQ ->
48 kS/s.  HBM-bound: 61 440 B in + 7 680 B out per channel-frame.  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m17_sdr_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    m.build()
    ctx = m.Context(0)
    x = torch.randint(-20000, 20000, (a.channels, a.blocks * 8 * 1920, 2), device=ctx.device, dtype=torch.int16)
    dec = m.Decimator(ctx, a.channels)
    for _ in range(3):
        dec.radio_receive_samples(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        dec.radio_receive_samples(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    nbytes = a.channels * a.blocks * (8 * 7680 + 7680)
    peak = 6553.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    print(json.dumps({"workload": f"{a.channels} channels x {a.blocks} blocks of 384 kS/s IQ ({nbytes / 1e9:.2f} GB per pass)", "ms": ms,
                      "gbs": nbytes / (ms * 1e-3) / 1e9, "hbm_peak_gbs": peak, "frac_of_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
                      "channel_frames_per_s": a.channels * a.blocks / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
