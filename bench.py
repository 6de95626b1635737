#!/usr/bin/env python
"""bench.py -- M17 RX hot-path throughput on B200 (BASELINE.json metric: channel-seconds decoded per second).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (default N=1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU chain on the host cores

Workload (BASELINE.json configs[1]): 1024 concurrent stream-mode channels per GPU, each a 10 s capture
(250 blocks of 1920 int16 IQ samples @48 kHz = one "over": carrier, 2 preambles, LSF, 244 stream frames, EOT),
random start delay, carrier offset U[-1000,1000] Hz and white noise on the IQ at Eb/N0(IQ) swept over
{22,24,26,30,inf} dB -- the range in which the reference's limiter-discriminator front end decodes at all
(the 0..12 dB sweep of configs[1] is applied at the 2-sps baseband seam in the parity tests).
One step = one pass of the whole RX chain (m17_dsp_rx semantics) over the 1024 x 250 channel-frames.
The 1.97 GB of IQ per GPU is far larger than the 126 MB L2, so nothing is cache-resident between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS_PER_GPU = 1024
BLOCKS = 250                      # 10 s per channel
EBN0_SWEEP = (22.0, 24.0, 26.0, 30.0, None)
FRAME_BYTES_FUSED = 7744          # 7680 B IQ in + 64 B record out
STAGE_BYTES = {"frontend": 9216, "sync_frame": 2304, "decode": 768 + 64, "post": 64}   # algorithmic bytes per channel-frame
DECIMATOR_BYTES = 8 * 7680 + 7680   # Pluto /8 front-end decimator: 61 440 B in + 7 680 B out per channel-frame
METRIC = "M17 channel-seconds decoded per second"
UNIT = "channel-s/s"


def clean_env():
    """environment for helper subprocesses (nvidia-smi, the CPU baseline binary): drop profiler injection so that a run
    under ncu does not try to attach to them"""
    bad = ("CUDA_INJECTION", "NV_COMPUTE_PROFILER", "NV_NSIGHT", "NSIGHT", "LD_PRELOAD", "NVTX_INJECTION", "CUPTI")
    return {k: v for k, v in os.environ.items() if not k.startswith(bad)}


def under_profiler():
    """true when a CUDA profiler (ncu) injected itself: helper subprocesses are then skipped altogether -- a number printed
    under a profiler is never a bench value, and ncu on this pool crashes when the traced process forks helpers"""
    bad = ("CUDA_INJECTION", "NV_COMPUTE_PROFILER", "NV_NSIGHT", "NSIGHT", "NVTX_INJECTION")
    return any(k.startswith(bad) for k in os.environ)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        if under_profiler():
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=clean_env())
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# ------------------------------------------------------------------------------------------------ NUMA placement
def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the NUMA node its GPU hangs off, so that
    the end-to-end path's H2D copies do not cross the socket interconnect when several ranks feed their GPUs at once."""
    try:
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            out = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                                 env=clean_env()).stdout.strip().lower()
            if not out:
                return None
            addr = out[-12:] if out.count(":") == 2 else "0000:" + out
        path = f"/sys/bus/pci/devices/{addr}/numa_node"
        if not os.path.exists(path):
            return None
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except (OSError, ValueError, subprocess.SubprocessError, AttributeError):
        pass
    return None


# ------------------------------------------------------------------------------------------------ workload
def make_workload(ctx, m, torch, C, T, seed, ebn0=EBN0_SWEEP, f0_max=1000.0):
    """Synthetic stream-mode channels generated ON THE GPU with the library's own TX path (outside any timed
    region): LSF -> frame formatter -> RRC x10 -> 4FSK -> int16 IQ, then delay / carrier offset / AWGN."""
    dev = ctx.device
    g = torch.Generator(device=dev); g.manual_seed(seed)
    F = T - 6
    payload = torch.randint(0, 256, (C, F, 16), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    lsf = torch.zeros((C, 30), dtype=torch.uint8, device=dev)
    lsf[:, 0:6] = 0xFF                                                      # broadcast destination
    src = torch.randint(1, 40 ** 6, (C,), generator=g, device=dev, dtype=torch.int64)   # some base-40 callsign
    for i in range(6):
        lsf[:, 6 + i] = ((src >> (40 - 8 * i)) & 0xFF).to(torch.uint8)
    lsf[:, 13] = 0x05                                                       # TYPE 0x0005: stream, voice
    crc = ctx.m17_crc_array_encode(lsf[:, :28].contiguous()).view(torch.int16).to(torch.int32) & 0xFFFF
    lsf[:, 28] = ((crc >> 8) & 0xFF).to(torch.uint8)
    lsf[:, 29] = (crc & 0xFF).to(torch.uint8)
    tx = m.Tx(ctx, C, 10)
    tx.set_lsf(lsf)
    pre = torch.from_numpy(tx.fmt_preamble()).to(dev).expand(C, 192)
    eot = torch.from_numpy(tx.fmt_eot()).to(dev).expand(C, 192)
    car = torch.full((C, 192), 4, dtype=torch.uint8, device=dev)
    script = torch.cat([car, pre, pre, tx.m17_fmt_add_link_setup_frame(lsf), tx.m17_fmt_add_stream_frame(payload).reshape(C, F * 192), eot, car], 1).contiguous()
    assert script.shape[1] == T * 192
    iq = tx.m17_mod_dibits(script)
    tx.close()
    del script
    # per-channel start delay 0..1919 samples (channels are not frame aligned)
    delay = torch.randint(0, 1920, (C,), generator=g, device=dev, dtype=torch.int64)
    iq32 = iq.view(torch.int32).reshape(C, T * 1920)
    n = torch.arange(T * 1920, device=dev, dtype=torch.int64)
    for c0 in range(0, C, 64):
        idx = (n[None, :] - delay[c0:c0 + 64, None]).clamp_(min=0)
        iq32[c0:c0 + 64] = torch.gather(iq32[c0:c0 + 64], 1, idx)
    del idx
    # carrier offset + AWGN
    eb = np.array([ebn0[c % len(ebn0)] if ebn0[c % len(ebn0)] is not None else np.inf for c in range(C)])
    sigma = np.where(np.isinf(eb), 0.0, np.sqrt(2.5 * 16383.0 ** 2 / 10 ** (eb / 10))).astype(np.float32)
    f0 = (torch.rand((C,), generator=g, device=dev) * 2.0 - 1.0) * f0_max / 48000.0
    ctx.synth_channel(iq, torch.from_numpy(sigma).to(dev), f0.float().contiguous(), seed=seed)
    torch.cuda.synchronize()
    return iq, payload


def payload_check(torch, res_frames, nframes, payload):
    """loopback sanity (no oracle): delivered stream payloads must equal what was transmitted."""
    fr = res_frames
    ok = tot = 0
    C = fr.shape[0]
    pl = payload.cpu().numpy()
    for c in range(0, C, max(1, C // 64)):
        f = fr[c, :nframes[c]]
        d = f[(f["type"] == 2) & ((f["flags"] & 8) != 0)]
        fn = (d["data"][:, 0].astype(int) << 8) | d["data"][:, 1]
        good = fn < pl.shape[1]
        tot += len(d)
        ok += int(sum(np.array_equal(d["data"][i, 2:18], pl[c, fn[i]]) for i in np.nonzero(good)[0]))
    return ok, tot


# ------------------------------------------------------------------------------------------------ TX record
TX_BYTES_PER_FRAME = 18 + 7680          # 18 B in (FN + payload) + 1920 int16 IQ samples out
TX_FIR_FLOP_PER_FRAME = 192 * 10 * 31 * 2 - 1920   # 119 040 - 1920: 31 products + 30 ordered adds per sample (m17_modulate.cpp:42-48)
FP32_LANE_ROOF = 148 * 128 * 1.965e9    # non-tensor fp32 instruction roof (lane-ops/s): SMs x lanes x clocks.max.sm


def tx_record(ctx, m, torch, C, F, steps, with_cpu):
    """The TX chain on the same batch shape as the RX workload: m17_send_stream_frame for C channels x F frames =
    m17b_fmt_stream_frames (k_fmt<2>) + m17b_mod_dibits (k_mod_fused), CUDA events on the launching stream; output 1.97 GB >> L2."""
    dev = ctx.device
    g = torch.Generator(device=dev); g.manual_seed(77)
    payload = torch.randint(0, 256, (C, F, 16), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    lsf = torch.randint(0, 256, (C, 30), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    tx = m.Tx(ctx, C, 10)
    tx.set_lsf(lsf)
    iq = torch.empty((C, F * 1920, 2), dtype=torch.int16, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]

    def one(k=None):
        if k is not None:
            ev[3 * k].record()
        dib = tx.m17_fmt_add_stream_frame(payload)
        if k is not None:
            ev[3 * k + 1].record()
        tx.m17_mod_dibits(dib.view(C, F * 192), out=iq)
        if k is not None:
            ev[3 * k + 2].record()
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize()
    fmt_ms = sum(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(steps)) / steps
    mod_ms = sum(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(steps)) / steps
    tot_ms = ev[0].elapsed_time(ev[3 * steps - 1]) / steps
    scan = tx.debug_scan()
    tx.close()
    frames = C * F
    fps = frames / (tot_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    rec = {"workload": f"{C} channels x {F} stream frames per step: m17_send_stream_frame = frame formatter + 4FSK modulator (RRC x10, fp32 phase accumulator, cos/sin, int16 IQ)",
           "frames_per_s": fps, "channel_s_per_s": fps / 25.0, "ms_per_step": tot_ms, "launches_per_step": 3,
           "kernels": {"k_fmt<2>": {"ms": round(fmt_ms, 4), "frames_per_s": frames / (fmt_ms * 1e-3)},
                       "k_mod_fused<10>": {"ms": round(mod_ms, 4), "frames_per_s": frames / (mod_ms * 1e-3),
                                           "alg_bytes_per_frame": 192 + 7680, "gbs": round((192 + 7680) * frames / (mod_ms * 1e-3) / 1e9, 1),
                                           "frac_of_hbm": round((192 + 7680) * frames / (mod_ms * 1e-3) / 1e9 / hbm, 4),
                                           "fir_flop_per_frame": TX_FIR_FLOP_PER_FRAME,
                                           "fir_lane_ops_per_s": TX_FIR_FLOP_PER_FRAME * frames / (mod_ms * 1e-3),
                                           "frac_of_fp32_lane_roof_fir_only": round(TX_FIR_FLOP_PER_FRAME * frames / (mod_ms * 1e-3) / FP32_LANE_ROOF, 4),
                                           "bound": "alu / serial phase chain (one fp32 add per sample + a wrap per symbol per channel: ~107 cycles per symbol, "
                                                    "a floor of symbols x 107 cycles whatever the batch; DESIGN.md 4)",
                                           "scan_warp_cycles_per_symbol": round(scan["scan_busy_cycles"] / max(1, scan["ctas"] * F * 192), 1)}},
           "roofline_bytes_per_frame": TX_BYTES_PER_FRAME, "whole_tx_gbs": round(TX_BYTES_PER_FRAME * frames / (tot_ms * 1e-3) / 1e9, 1)}
    if with_cpu:
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "m17ref_bench")
        cores = os.cpu_count() or 1
        if os.path.exists(ref_bin):
            out = json.loads(subprocess.run([ref_bin, "--tx", "20000", str(cores)], capture_output=True, text=True, check=True, env=clean_env()).stdout)
            rec["cpu_baseline"] = {"value": out["frames_per_s"], "unit": "frames/s", "cores": cores, "kind": "reference",
                                   "sample": f"m17_send_stream_frame x 20000 per process, one unmodified-reference process per core ({out['secs_max_worker']:.2f} s)"}
    return rec


# ------------------------------------------------------------------------------------------------ aux: configs[2], [3], [4]
def packet_record(ctx, m, torch, C, steps):
    """BASELINE configs[2]: packet-mode batch.  Per channel: carrier, 2 preambles, LSF (TYPE 0x0002), one packet of 1..200 bytes
    cut into 25-byte frames with its CRC (m17_send_packet_frames), EOT; modulated at 384 kS/s and decimated by 8 from a random
    phase (fractional timing offset), random start delay, carrier offset +-1 kHz, AWGN at 26 / 28 / 30 dB / none.  The RX step
    is timed; packets are reassembled on the GPU and compared with what was sent."""
    dev = ctx.device
    T, MF = 16, 10
    g = torch.Generator(device=dev); g.manual_seed(333)
    L = torch.randint(1, 201, (C,), generator=g, device=dev, dtype=torch.int32)
    pk = torch.randint(0, 256, (C, 256), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    lsf = torch.zeros((C, 30), dtype=torch.uint8, device=dev)
    lsf[:, 0:6] = 0xFF
    lsf[:, 11] = 0x2A
    lsf[:, 13] = 0x02                                                       # TYPE 0x0002: packet, data
    crc = ctx.m17_crc_array_encode(lsf[:, :28].contiguous()).view(torch.int16).to(torch.int32) & 0xFFFF
    lsf[:, 28] = ((crc >> 8) & 0xFF).to(torch.uint8); lsf[:, 29] = (crc & 0xFF).to(torch.uint8)
    tx = m.Tx(ctx, C, 80)
    dib, nf = tx.m17_send_packet_frames(pk, L, max_frames=MF)               # [C][MF][192], frames used per packet (<= 9)
    eot = torch.from_numpy(tx.fmt_eot()).to(dev)
    dib[torch.arange(C, device=dev), nf.long()] = eot                       # EOT directly behind the packet's last frame
    pre = torch.from_numpy(tx.fmt_preamble()).to(dev).expand(C, 192)
    car = torch.full((C, 192), 4, dtype=torch.uint8, device=dev)
    script = torch.cat([car, pre, pre, tx.m17_fmt_add_link_setup_frame(lsf), dib.reshape(C, MF * 192), car, car], 1).contiguous()
    assert script.shape[1] == T * 192
    iq80 = tx.m17_mod_dibits(script).view(torch.int32).reshape(C, T * 192 * 80)
    tx.close()
    ph = torch.randint(0, 8, (C,), generator=g, device=dev, dtype=torch.int64)
    delay = torch.randint(0, 1920, (C,), generator=g, device=dev, dtype=torch.int64)
    n = torch.arange(T * 1920, device=dev, dtype=torch.int64)
    iq = torch.empty((C, T * 1920), dtype=torch.int32, device=dev)
    for c0 in range(0, C, 64):
        idx = ((n[None, :] - delay[c0:c0 + 64, None]).clamp_(min=0)) * 8 + ph[c0:c0 + 64, None]
        iq[c0:c0 + 64] = torch.gather(iq80[c0:c0 + 64], 1, idx)
    del iq80, idx
    iq = iq.view(torch.int16).reshape(C, T * 1920, 2)
    eb = [26.0, 28.0, 30.0, None]
    sigma = np.array([0.0 if eb[c % 4] is None else np.sqrt(2.5 * 16383.0 ** 2 / 10 ** (eb[c % 4] / 10)) for c in range(C)], np.float32)
    f0 = (torch.rand((C,), generator=g, device=dev) * 2000.0 - 1000.0) / 48000.0
    ctx.synth_channel(iq, torch.from_numpy(sigma).to(dev), f0.float().contiguous(), seed=334)
    rx = m.Rx(ctx, C, T)
    for _ in range(3):
        rx.reset(); rx.m17_dsp_rx(iq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rx.reset(); rx.m17_dsp_rx(iq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    by, pkt, npk = rx.reassemble_packets(bytes_cap=256, max_pkts=2)
    stats = rx.view()["stats"].sum(0).cpu().numpy()
    npk_h, pkt_h, by_h, L_h, pk_h = npk.cpu().numpy(), pkt.cpu().numpy(), by.cpu().numpy(), L.cpu().numpy(), pk.cpu().numpy()
    good = sum(1 for c in range(C) if npk_h[c] >= 1 and pkt_h[c, 0, 2] == 1 and pkt_h[c, 0, 1] == L_h[c] and np.array_equal(by_h[c, :L_h[c]], pk_h[c, :L_h[c]]))
    crc_ok = int(sum(int(pkt_h[c, k, 2]) for c in range(C) for k in range(min(int(npk_h[c]), 2))))
    rx.close()
    return {"workload": f"configs[2]: {C} packet-mode channels x {T} blocks (LSF + one packet of 1..200 bytes in 25-byte frames + EOT), fractional timing offset "
                        f"(x8 oversampled TX decimated from a random phase), start delay, f0 +-1 kHz, AWGN on IQ Eb/N0 {{26,28,30,inf}} dB",
            "ms_per_step": ms, "channel_frames_per_s": C * T / (ms * 1e-3), "channel_s_per_s": C * T / (ms * 1e-3) / 25.0,
            "packets_sent": C, "packets_recovered_exact_with_valid_crc": int(good), "packets_with_valid_crc": crc_ok,
            "frames": int(stats[0]), "aos": int(stats[4]), "los": int(stats[5]), "reassembly": "on the GPU (m17b_rx_reassemble_packets)"}


def wideband_record(ctx, m, torch, steps, ncap=11, T=BLOCKS):
    """the same kind of workload arriving as WIDEBAND captures -- ncap captures at 1.2 MS/s, each carrying 96
    stream-mode channels on a 12.5 kHz raster (ncap*96 channels x T blocks) -- channelised on the GPU (m17b_chan_run) straight
    into m17b_dsp_rx.  Reported: the channeliser kernel alone, the device-resident step, and end to end from pinned host memory
    (H2D of the captures, channeliser, RX chain, D2H of the records; three capture groups pipelined on two streams)."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    from wideband import wideband_from_channels
    dev = ctx.device
    C = ncap * 96
    iq, payload = make_workload(ctx, m, torch, C, T, seed=555, ebn0=(None, 30.0, 26.0), f0_max=500.0)
    wide = torch.stack([wideband_from_channels(iq[96 * g:96 * (g + 1)]) for g in range(ncap)]).contiguous()     # [ncap][T*1920*25][2]
    del iq
    torch.cuda.empty_cache()
    ch = m.Channelizer(ctx, ncap, 12)
    rx = m.Rx(ctx, C, T)
    buf = torch.empty((C, T * 1920, 2), dtype=torch.int16, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    chan_ms = step_ms = 0.0
    for k in range(3 + steps):
        ch.reset(); rx.reset()
        ev[0].record()
        ch.run(wide, out=buf)
        ev[1].record()
        rx.m17_dsp_rx(buf)
        ev[2].record()
        torch.cuda.synchronize()
        if k >= 3:
            chan_ms += ev[0].elapsed_time(ev[1]) / steps
            step_ms += ev[0].elapsed_time(ev[2]) / steps
    res = rx.results()
    ok, tot = payload_check(torch, res["frames"], res["nframes"], payload)
    delivered = int(res["stats"][:, 3].sum())
    rx.close(); ch.close()
    del buf
    # ---- end to end from pinned host memory
    groups = [(0, 4), (4, 8), (8, ncap)] if ncap >= 3 else [(0, ncap)]
    wide_host = torch.empty(wide.shape, dtype=torch.int16).pin_memory()
    wide_host.copy_(wide)
    del wide
    torch.cuda.empty_cache()
    chs = [m.Channelizer(ctx, b - a, 12) for a, b in groups]
    rxs = [m.Rx(ctx, (b - a) * 96, T) for a, b in groups]
    wdev = [torch.empty((b - a, T * 1920 * 25, 2), dtype=torch.int16, device=dev) for a, b in groups]
    bufs = [torch.empty(((b - a) * 96, T * 1920, 2), dtype=torch.int16, device=dev) for a, b in groups]
    fr_host = [torch.empty(((b - a) * 96, rxs[i].frame_cap, 64), dtype=torch.uint8).pin_memory() for i, (a, b) in enumerate(groups)]
    nf_host = [torch.empty(((b - a) * 96,), dtype=torch.int32).pin_memory() for a, b in groups]
    copy_s, comp_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    evs = [torch.cuda.Event() for _ in groups]
    ran = [torch.cuda.Event() for _ in groups]                   # the channeliser has consumed wdev[i]: the next step's capture may land

    # A continuous receiver: the steps are queued back to back, the host does not wait between them, so the next step's first
    # captures cross the link while the last group of this step is still being channelised and decoded (the capture buffers
    # are handed back by an event).  Every step's captures are copied and every step's records are read back inside the timed
    # region; the region ends when the last step's records are in host memory.
    def e2e_step():
        for i, (a, b) in enumerate(groups):
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ran[i])
                wdev[i].copy_(wide_host[a:b], non_blocking=True)
                evs[i].record()
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(evs[i])
                chs[i].reset(); rxs[i].reset()
                chs[i].run(wdev[i], out=bufs[i])
                ran[i].record()
                rxs[i].m17_dsp_rx(bufs[i])
                v = rxs[i].view()
                fr_host[i].copy_(v["frames"], non_blocking=True)
                nf_host[i].copy_(v["nframes"], non_blocking=True)
    e2e_steps = max(2, min(steps, 5))
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    nfr_e2e = int(sum(int(x.sum()) for x in nf_host))
    for o in chs + rxs:
        o.close()
    h2d = int(wide_host.numel() * 2)
    d2h = int(sum(f.numel() for f in fr_host) + 4 * C)
    alg = 100 + 384                                              # bytes per output time: 25 input samples in, 96 channel samples out
    nt = ncap * T * 1920
    return {"workload": f"{ncap} wideband captures x {T} blocks at 1.2 MS/s, 96 stream-mode channels each on a 12.5 kHz raster ({C} channels), Eb/N0(IQ) {{inf,30,26}} dB "
                        f"before the raster synthesis, f0 +-500 Hz; m17b_chan_run (96 x 12-tap polyphase FIR + fixed-point 96-point DFT) feeding m17b_dsp_rx on the device",
            "channels": C, "k_chan96": {"ms": round(chan_ms, 4), "alg_bytes_per_output_time": alg, "gbs": round(alg * nt / (chan_ms * 1e-3) / 1e9, 1),
                                        "channel_s_per_s": C * T / (chan_ms * 1e-3) / 25.0, "bound": "integer ALU (fold: 2 x 1152 int MACs, DFT: ~1700 64-bit multiplies per output time)"},
            "device_resident": {"ms_per_step": round(step_ms, 4), "channel_s_per_s": C * T / (step_ms * 1e-3) / 25.0},
            "e2e": {"value": C * T / (e2e_ms * 1e-3) / 25.0, "unit": UNIT, "ms_per_step": round(e2e_ms, 3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_bytes_per_channel_second": h2d / (C * T / 25.0), "note": "per-channel 48 kS/s IQ costs 192 000 B per channel-second on the host link",
                    "schedule": f"{len(groups)} capture groups per step on a copy and a compute stream, {e2e_steps} steps queued back to back (the next step's captures cross the link while the last group is decoded); timed until the last records are in host memory"},
            "check": {"delivered_payloads_exact": f"{ok}/{tot}", "delivered": delivered, "frames_e2e": nfr_e2e}}


def viterbi_record(ctx, m, torch, steps):
    """BASELINE configs[3]: 1M punctured (P2, stream) K=5 r=1/2 soft-decision frames on the int8 grid at Eb/N0 3 dB:
    depuncture + Viterbi + pack (m17b_viterbi_punctured), ACS throughput against the fp32 lane roof."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("viterbi_micro", os.path.join(ROOT, "benchmarks", "viterbi_micro.py"))
    vm = importlib.util.module_from_spec(spec); spec.loader.exec_module(vm)
    n = 1 << 20
    data, soft = vm.make_frames(ctx, 2, n, 3.0)
    for _ in range(3):
        out = ctx.viterbi_punctured(2, soft)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = ctx.viterbi_punctured(2, soft)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ber = float(np.unpackbits(torch.bitwise_xor(out, data).cpu().numpy()).mean())
    ops = n * 148 * 52                                   # 4 branch-metric adds + 16 x (2 add + 1 compare) per trellis step
    return {"workload": "configs[3]: 1 048 576 punctured stream frames (272 soft values each, int8 grid), Eb/N0 3 dB", "ms_per_step": ms,
            "frames_per_s": n / (ms * 1e-3), "acs_lane_ops_per_s": ops / (ms * 1e-3), "fp32_lane_roof": FP32_LANE_ROOF,
            "frac_of_alu_roof": ops / (ms * 1e-3) / FP32_LANE_ROOF, "bytes_in_per_frame": 1088, "gbs_in": n * 1088 / (ms * 1e-3) / 1e9,
            "bit_error_rate": ber}


def config5_record(ctx, m, torch, dist, md, rank, world, steps, total=65536, T=25):
    """BASELINE configs[4]: 65 536 channels sharded over the GPUs (contiguous ranges, dist.shard_range), 25 blocks (1 s) each,
    Eb/N0(IQ) 26 dB, f0 +-1 kHz: 1024 distinct channels tiled x64 with per-copy sample delays.  Every step ends with the NCCL
    gather of all decoded-frame records on rank 0 and the all-reduce of the counters.  Rank 0 re-runs a sample of every other
    rank's channels itself and compares the gathered records byte for byte."""
    dev = ctx.device
    base_iq, _ = make_workload(ctx, m, torch, 1024, T, seed=4242, ebn0=(26.0,))
    base = base_iq.view(torch.int32).reshape(1024, T * 1920)
    n = torch.arange(T * 1920, device=dev, dtype=torch.int64)

    def channels(gidx):                                   # IQ of the global channels in gidx (int64 tensor)
        out = torch.empty((len(gidx), T * 1920), dtype=torch.int32, device=dev)
        for a in range(0, len(gidx), 256):
            gi = gidx[a:a + 256]
            idx = (n[None, :] - ((gi // 1024) * 7)[:, None]).clamp_(min=0)
            out[a:a + 256] = torch.gather(base[gi % 1024], 1, idx)
        return out.view(torch.int16).reshape(len(gidx), T * 1920, 2)
    c0, c1 = md.shard_range(total, rank, world)
    iq = channels(torch.arange(c0, c1, device=dev, dtype=torch.int64))
    rx = m.Rx(ctx, c1 - c0, T)
    fr = nf = tot = None

    def step():
        rx.reset()
        rx.m17_dsp_rx(iq)
        v = rx.view()
        return md.gather_records(v["frames"], v["nframes"], total, dst=0), md.reduce_stats(v["stats"])
    for _ in range(2):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    chain_ms = 0.0
    t_all0 = torch.cuda.Event(enable_timing=True); t_all0.record()
    for _ in range(steps):
        e0.record()
        rx.reset(); rx.m17_dsp_rx(iq)
        e1.record()
        v = rx.view()
        (fr, nf), tot = md.gather_records(v["frames"], v["nframes"], total, dst=0), md.reduce_stats(v["stats"])
        e2.record()
        torch.cuda.synchronize()
        chain_ms += e0.elapsed_time(e1) / steps
    t_all1 = torch.cuda.Event(enable_timing=True); t_all1.record()
    torch.cuda.synchronize()
    tm = torch.tensor([t_all0.elapsed_time(t_all1) / steps, chain_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, chain = float(tm[0]), float(tm[1])
    same, checked = True, 0
    if rank == 0:
        for r in range(1, world):
            a, b = md.shard_range(total, r, world)
            gi = torch.linspace(a, b - 1, 32, device=dev).long()
            rx2 = m.Rx(ctx, len(gi), T)
            rx2.m17_dsp_rx(channels(gi))
            v2 = rx2.view()
            f2, n2 = v2["frames"].cpu().numpy(), v2["nframes"].cpu().numpy()
            fg, ng = fr[gi].cpu().numpy(), nf[gi].cpu().numpy()
            for k in range(len(gi)):
                same = same and ng[k] == n2[k] and np.array_equal(fg[k, :ng[k]], f2[k, :n2[k]])
                checked += 1
            rx2.close()
    cap = rx.frame_cap
    rx.close()
    del iq
    return {"workload": f"configs[4]: {total} channels sharded x{world} ({c1 - c0} per GPU) x {T} blocks, Eb/N0(IQ) 26 dB, f0 +-1 kHz, 1024 distinct channels "
                        f"tiled x{total // 1024} with per-copy sample delays; every step = RX chain + NCCL gather of all records to rank 0 + all-reduce of the counters",
            "ms_per_step": ms, "ms_rx_chain_only": chain, "channel_s_per_s": total * T / (ms * 1e-3) / 25.0,
            "gathered_bytes_per_step": int(total * cap * 64 + total * 4), "frames_total": int(tot[0]) if tot is not None else None,
            "delivered_total": int(tot[3]) if tot is not None else None,
            "gathered_records_equal_local_rerun": bool(same), "channels_rechecked_on_rank0": checked}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(iq_host_sample, T, budget_s=12.0):
    """Time the reference's own RX chain (oracle/_ref/m17ref_bench, unmodified reference objects) on the host cores,
    one process per core, on a bounded sample of the same IQ.  Falls back to the C restatement (oracle port)."""
    cores = os.cpu_count() or 1
    S = iq_host_sample.shape[0]
    frames = S * T
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "m17ref_bench")
    if os.path.exists(ref_bin):
        tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        path = os.path.join(tmpdir, f"m17_bench_iq_{os.getpid()}.bin")
        iq_host_sample.tofile(path)
        try:
            # calibrate with one pass, then size reps to the budget
            out = json.loads(subprocess.run([ref_bin, path, str(S), str(T), str(cores), "1"], capture_output=True, text=True, check=True, env=clean_env()).stdout)
            reps = max(1, int(budget_s / max(out["secs_max_worker"], 1e-3)))
            out = json.loads(subprocess.run([ref_bin, path, str(S), str(T), str(cores), str(reps)], capture_output=True, text=True, check=True, env=clean_env()).stdout)
        finally:
            os.unlink(path)
        fps = out["frames_per_s"]
        return {"value": fps / 25.0, "unit": UNIT, "frames_per_s": fps, "cores": cores, "kind": "reference",
                "sample": f"{S} channels x {T} blocks x {reps} passes of the bench IQ, one reference process per core ({out['frames']} channel-frames, {out['secs_max_worker']:.2f} s)",
                "delivered": out["delivered"]}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from m17_oracles import Port
    P = Port()
    secs = P.rx_time(iq_host_sample, cores)
    reps = max(1, int(budget_s / max(secs, 1e-3)))
    t = sum(P.rx_time(iq_host_sample, cores) for _ in range(reps))
    fps = frames * reps / t
    return {"value": fps / 25.0, "unit": UNIT, "frames_per_s": fps, "cores": cores, "kind": "port",
            "sample": f"{S} channels x {T} blocks x {reps} passes, C restatement, one thread per core"}


def reference_input(S, T, seed):
    """CPU-only synthetic input for --impl reference (no GPU involved): oracle-port TX + numpy channel."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from m17_oracles import Port, add_iq_noise, delay_iq, lsf_for, rotate_iq
    P = Port()
    rng = np.random.default_rng(seed)
    X = np.zeros((S, T * 1920, 2), np.int16)
    for c in range(S):
        pl = rng.integers(0, 256, (T - 6, 16), dtype=np.uint8)
        iq, _, _ = P.tx_stream_over(lsf_for(P), pl, lead=1, npre=2, tail=1)
        x = delay_iq(iq, int(rng.integers(0, 1920)), T * 1920)
        x = rotate_iq(x, float(rng.uniform(-1000, 1000)))
        X[c] = add_iq_noise(x, EBN0_SWEEP[c % len(EBN0_SWEEP)], rng)
    return X


RESULT_OUT = sys.stdout


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    S = min(512, 16 * cores)          # about 0.25 s of all-core work per step
    T = BLOCKS
    X = reference_input(S, T, 1234)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "m17ref_bench")
    vals = []
    kind = "reference" if os.path.exists(ref_bin) else "port"
    if kind == "reference":
        path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), f"m17_ref_iq_{os.getpid()}.bin")
        X.tofile(path)
        try:
            for s in range(args.warmup + args.steps):
                out = json.loads(subprocess.run([ref_bin, path, str(S), str(T), str(cores), "1"], capture_output=True, text=True, check=True, env=clean_env()).stdout)
                if s >= args.warmup:
                    vals.append(out["secs_max_worker"])
        finally:
            os.unlink(path)
    else:
        from m17_oracles import Port
        P = Port()
        for s in range(args.warmup + args.steps):
            t = P.rx_time(X, cores)
            if s >= args.warmup:
                vals.append(t)
    ms = 1e3 * sum(vals) / len(vals)
    fps = S * T / (ms / 1e3)
    val = fps / 25.0
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": fps,
            "config": {"workload": f"configs[1] bounded sample: {S} stream-mode channels x {T} blocks per step, full m17_dsp_rx chain from int16 IQ, "
                                   f"AWGN on IQ Eb/N0 {{22,24,26,30,inf}} dB, f0 +-1 kHz, one unmodified-reference process per host core"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{S} channels x {T} blocks per step"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import m17_sdr_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise m.M17Error("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local) if (world > 1 and not under_profiler()) else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m.build()
    ctx = m.Context(local)
    C, T = args.channels, args.blocks
    t0 = time.time()
    iq, payload = make_workload(ctx, m, torch, C, T, seed=1000 + rank)
    log(f"[rank {rank}] workload {C} ch x {T} blocks generated in {time.time() - t0:.1f}s ({iq.numel() * 2 / 1e9:.2f} GB IQ)")
    rx = m.Rx(ctx, C, T)
    # The only inter-GPU traffic of the step: an all-reduce of the 8 job-wide counters (NCCL over NVLink).  It is off the
    # critical path: the step snapshots its counters (one tiny reduction on the step's stream), and a side stream hands the
    # snapshot to NCCL while the next step's front end already runs; all reductions are complete before the timed region ends.
    side = torch.cuda.Stream(device=ctx.device) if world > 1 else None
    slots = torch.zeros((64, 8), dtype=torch.int64, device=ctx.device)
    pending = []
    stats_view = []

    def step():
        rx.reset()
        rx.m17_dsp_rx(iq)
        if world > 1:
            if not stats_view:
                stats_view.append(rx.view()["stats"])            # (the library's counter buffer: its address does not change)
            k = len(pending) % slots.shape[0]
            torch.sum(stats_view[0], 0, out=slots[k])
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                pending.append(dist.all_reduce(slots[k], async_op=True))

    def drain():
        for w in pending:
            w.wait()
        pending.clear()
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)

    def sync_all():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    if args.slice_blocks is not None:
        rx.set_slice_blocks(args.slice_blocks)
    if args.chan_groups is not None:
        rx.set_chan_groups(args.chan_groups)
    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()                                   # the counter reductions belong to the timed region
    e1.record()
    sync_all()
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    ms_total = e0.elapsed_time(e1)
    tms = torch.tensor([ms_total], device=ctx.device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    launches = (rx.launches() + 1) * args.steps          # chain kernels of every pipeline slice + the state-reset kernel per step
    res = rx.results()
    # ---- per-stage device times: the same step with the stages strictly in sequence (stage timing switches the time-sliced
    #      pipeline off), CUDA events recorded by the library on the launching stream around each stage
    stage = {k: 0.0 for k in ("frontend", "sync_frame", "decode", "post")}
    nst = min(args.steps, 16)
    rx.set_timing(True)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step(); drain(); torch.cuda.synchronize()
    rx.set_timing(True)                      # restart the call counter after the warm-up step
    es0.record()
    for _ in range(nst):
        step()
    drain()
    es1.record()
    torch.cuda.synchronize()
    ms_step_serial = es0.elapsed_time(es1) / nst
    for i in range(nst):
        s = rx.stage_ms(i)
        for k in stage:
            stage[k] += s[k] / nst
    rx.set_timing(False)
    res_serial = rx.results()
    same_serial = bool(np.array_equal(res_serial["frames"].view(np.uint8), res["frames"].view(np.uint8)) and np.array_equal(res_serial["nsym"], res["nsym"]))
    nfr = res["nframes"]
    ok, tot = payload_check(torch, res["frames"], nfr, payload)
    stats = res["stats"].sum(0)

    # ---- the stage ahead of the path: Pluto /8 decimator on its own 384 kS/s input (outside the step above)
    dec_blocks = 32
    dec_in = torch.randint(-20000, 20000, (C, dec_blocks * 8 * 1920, 2), device=ctx.device, dtype=torch.int16)
    dec = m.Decimator(ctx, C)
    for _ in range(3):
        dec.radio_receive_samples(dec_in)
    ed0, ed1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ed0.record()
    for _ in range(10):
        dec.radio_receive_samples(dec_in)
    ed1.record()
    torch.cuda.synchronize()
    dec_ms = ed0.elapsed_time(ed1) / 10
    dec.close()
    del dec_in

    # ---- the TX chain on the same batch shape (outside the RX step; rank 0 reports it)
    txrec = tx_record(ctx, m, torch, C, T, max(3, min(args.steps, 10)), with_cpu=(rank == 0 and not under_profiler()))

    # ---- the other BASELINE configs, so that the driver times them too (rank 0; config 5 needs all ranks)
    aux = {"tx": txrec}
    if rank == 0 and not args.no_aux:
        aux["packet"] = packet_record(ctx, m, torch, 1024, max(3, min(args.steps, 10)))
        aux["viterbi"] = viterbi_record(ctx, m, torch, max(3, min(args.steps, 10)))
        aux["wideband"] = wideband_record(ctx, m, torch, max(3, min(args.steps, 5)))
    if world > 1 and not args.no_aux:
        from m17_sdr_b200 import dist as md
        sync_all()
        aux["config5"] = config5_record(ctx, m, torch, dist, md, rank, world, max(2, min(args.steps, 5)))
        sync_all()

    # ---- end-to-end through the C ABI with HOST buffers ("e2e")
    iq_host = torch.empty(iq.shape, dtype=torch.int16).pin_memory()
    iq_host.copy_(iq)
    fr_host = torch.empty((C, rx.frame_cap, 64), dtype=torch.uint8).pin_memory()
    nf_host = torch.empty((C,), dtype=torch.int32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        rx.reset(); rx.m17_dsp_rx_host(iq_host, fr_host, nf_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rx.reset()
        rx.m17_dsp_rx_host(iq_host, fr_host, nf_host)      # H2D of the IQ, the chain, D2H of the records; returns synchronised
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    # what bounds it: the same bytes copied host -> device by every rank AT THE SAME TIME, nothing else running (pinned memory,
    # one cudaMemcpyAsync per rank, 3 repetitions) -- the per-rank PCIe / host-memory rate when N ranks feed their GPUs at once
    sync_all()
    dev_buf = torch.empty_like(iq)
    cp0, cp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_buf.copy_(iq_host, non_blocking=True); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    cp0.record()
    for _ in range(3):
        dev_buf.copy_(iq_host, non_blocking=True)
    cp1.record()
    torch.cuda.synchronize()
    h2d_ms = cp0.elapsed_time(cp1) / 3
    del dev_buf
    per_rank = torch.tensor([e2e_ms, h2d_ms], device=ctx.device, dtype=torch.float64)
    per_rank_all = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_all, per_rank)
    else:
        per_rank_all = [per_rank]
    e2e_rank_ms = [round(float(x[0]), 3) for x in per_rank_all]
    h2d_rank_gbs = [round(C * T * 7680 / (float(x[1]) * 1e-3) / 1e9, 2) for x in per_rank_all]
    e2e_ms = max(e2e_rank_ms)
    frh = fr_host.numpy().view(m.REC_DTYPE).reshape(C, -1)
    e2e_same = bool(np.array_equal(nf_host.numpy(), nfr) and all(np.array_equal(frh[c, :nfr[c]], res["frames"][c, :nfr[c]]) for c in range(C)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    frames_step = C * T * world
    fps = frames_step / (ms_step / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(stage, key=stage.get)
    stages = {k: {"ms": round(v, 4), "alg_bytes_per_frame": STAGE_BYTES[k], "gbs": round(STAGE_BYTES[k] * C * T / (v * 1e-3) / 1e9, 1) if v > 0 else None}
              for k, v in stage.items()}
    stages["decimator_x8 (ahead of the path, measured separately)"] = {
        "ms": round(dec_ms, 4), "alg_bytes_per_frame": DECIMATOR_BYTES, "channel_frames": C * dec_blocks,
        "gbs": round(DECIMATOR_BYTES * C * dec_blocks / (dec_ms * 1e-3) / 1e9, 1)}
    achieved = STAGE_BYTES[dom] * C * T / (stage[dom] * 1e-3) / 1e9
    kname = {"frontend": "k_frontend", "sync_frame": "k_sync_frame", "decode": "k_stream_acs", "post": "k_post"}
    traffic = None
    try:                                       # dram__bytes_read+write per launch from the ncu --set full capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if (C, T) == (CHANNELS_PER_GPU, BLOCKS):
            traffic = tj.get(kname[dom])
    except (OSError, ValueError):
        pass
    # bounded CPU baseline on the same IQ (rank 0 only)
    S = min(C, max(2 * (os.cpu_count() or 1), 16))
    cb = cpu_baseline(iq_host[:S].numpy(), T) if not under_profiler() else {"value": None, "unit": UNIT, "cores": 0, "kind": "skipped under profiler", "sample": ""}
    gpu_deliv_sample = int(sum(int((((res["frames"][c, :nfr[c]]["flags"] & 8) != 0) & (res["frames"][c, :nfr[c]]["type"] == 2)).sum()) for c in range(S)))
    if "delivered" in cb:
        cb["check"] = f"delivered stream frames on the sample: reference {cb.pop('delivered')} vs CUDA {gpu_deliv_sample}"
    # The dominant kernel's roofline.  k_sync_frame (matched filter + timing loop + framer) is bound by the fp32 lane rate / the
    # latency of each channel's serial chain, not by HBM: its figure is algorithmic lane-ops (23 808 ordered multiplies and adds
    # per channel-frame, no FMA allowed) against the non-tensor fp32 lane roof; its HBM figure is kept beside it.  The HBM-bound
    # stage is the front end: its own roofline object follows.
    ALG_FLOP = {"sync_frame": 23808, "decode": 7696, "frontend": 1920 * 13, "post": 0}
    if dom == "frontend" or ALG_FLOP[dom] == 0:
        roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic}
    else:
        lane_ops = ALG_FLOP[dom] * C * T / (stage[dom] * 1e-3) / 1e12
        roofline = {"bound": "alu", "kernel": kname[dom], "achieved": round(lane_ops, 3), "peak": round(FP32_LANE_ROOF / 1e12, 2), "unit": "TFLOP/s",
                    "frac": round(lane_ops / (FP32_LANE_ROOF / 1e12), 4), "traffic": traffic,
                    "alg_flop_per_channel_frame": ALG_FLOP[dom], "peak_source": "148 SMs x 128 fp32 lanes x 1.965 GHz, one rounded operation per lane and cycle (the reference arithmetic forbids FMA contraction)",
                    "hbm": {"achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "alg_bytes_per_channel_frame": STAGE_BYTES[dom]}}
    fe_gbs = STAGE_BYTES["frontend"] * C * T / (stage["frontend"] * 1e-3) / 1e9
    fe_traffic = None
    try:
        fe_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_frontend") if (C, T) == (CHANNELS_PER_GPU, BLOCKS) else None
    except (OSError, ValueError):
        pass
    roofline.update({
        "note": "dominant kernel by device time (stages in sequence)",
        "hbm_peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
        "frontend": {"bound": "hbm", "kernel": "k_frontend", "achieved": round(fe_gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(fe_gbs / hbm_peak, 4), "traffic": fe_traffic},
        "whole_chain_gbs": round(FRAME_BYTES_FUSED * C * T / (ms_step * 1e-3) / 1e9, 1),
        "stages_measured": "stages strictly in sequence (no channel groups), CUDA events on the launching stream",
        "ms_per_step_in_sequence": round(ms_step_serial, 4), "records_equal_pipelined": same_serial, "stages": stages})
    line = {
        "metric": METRIC, "value": fps / 25.0, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": fps,
        "config": {"workload": f"configs[1]: {C} concurrent stream-mode channels per GPU x {T} blocks (10 s each), full m17_dsp_rx chain from int16 IQ "
                               f"(limiter, discriminator, RRC matched filter + timing loop, sync/framer, demap+gather, Viterbi, Golay, CRC, LICH), "
                               f"AWGN on IQ Eb/N0 {{22,24,26,30,inf}} dB, f0 +-1 kHz, random start delay",
                   "channels_per_gpu": C, "blocks": T, "pipeline_slice_blocks": args.slice_blocks or 0, "channel_groups": "auto (4 independent channel-group chains on their own streams from 512 channels up)" if args.chan_groups is None else args.chan_groups, "l2": "input 1.97 GB per GPU >> 126 MB L2 (no flush needed)", "parallelism": f"channels sharded x{world}, no data-path collective", "host_numa_node_rank0": numa},
        "e2e": {"value": frames_step / (e2e_ms / 1e3) / 25.0, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(C * T * 7680),
                "d2h_bytes_per_step": int(C * rx.frame_cap * 64 + 4 * C), "records_equal_device_path": e2e_same,
                "per_rank_ms": e2e_rank_ms, "h2d_copy_only_gbs_per_rank_all_ranks_at_once": h2d_rank_gbs,
                "bound": "PCIe / host memory: compare per_rank_ms with the copy-only rate of the same bytes; every rank moves its own 1.97 GB per step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cb,
        "aux": aux,
        "e2e_wideband": (dict(aux["wideband"]["e2e"], channels=aux["wideband"]["channels"],
                              note="the same kind of workload delivered as 1.2 MS/s wideband captures (96 channels each on a 12.5 kHz raster) and channelised on the GPU: "
                                   "50 kB instead of 192 kB per channel-second cross the host link; see aux.wideband") if "wideband" in aux else None),
        "check": {"delivered_payloads_exact": f"{ok}/{tot}", "frames": int(stats[0]), "stream_frames": int(stats[1]), "delivered": int(stats[3]),
                  "golay_errors": int(stats[2]), "aos": int(stats[4]), "los": int(stats[5])},
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--channels", type=int, default=CHANNELS_PER_GPU)
    ap.add_argument("--blocks", type=int, default=BLOCKS)
    ap.add_argument("--chan-groups", type=int, default=None, help="independent channel-group chains per call (1 = one chain); default: library default (auto)")
    ap.add_argument("--no-aux", action="store_true", help="skip the packet / Viterbi / config-5 records")
    ap.add_argument("--slice-blocks", type=int, default=None, help="blocks per pipeline slice (0 = stages in sequence); default: library default")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter of child processes) is sent to stderr for the lifetime of the process
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
