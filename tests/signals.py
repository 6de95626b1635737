"""Seeded synthetic M17 channels used by the parity tests (numpy + the C oracle for TX)."""
import numpy as np

from m17_oracles import BLOCK, add_iq_noise, delay_iq, lsf_for, rotate_iq


def stream_channels(port, nchan, nframes, seed, ebn0=None, f0_max=0.0, max_delay=BLOCK - 1, tail_blocks=2):
    """Config-1/2 style: per channel one over (carrier, 2 preambles, LSF, F stream frames, EOT, carrier),
    random start delay, optional white noise on IQ (per-channel Eb/N0(IQ) list or scalar) and carrier offset.
    Returns (iq int16 [C][T*1920][2], payloads [C][F][16])."""
    rng = np.random.default_rng(seed)
    pl = rng.integers(0, 256, (nchan, nframes, 16), dtype=np.uint8)
    chans = []
    for c in range(nchan):
        lsf = lsf_for(port, src="G4GUO    ")
        iq, _, _ = port.tx_stream_over(lsf, pl[c])
        chans.append(iq)
    n = max(len(x) for x in chans)
    T = (n + max_delay) // BLOCK + 1 + tail_blocks
    X = np.zeros((nchan, T * BLOCK, 2), np.int16)
    for c in range(nchan):
        d = int(rng.integers(0, max_delay + 1))
        x = delay_iq(chans[c], d, T * BLOCK)
        e = ebn0[c] if isinstance(ebn0, (list, tuple, np.ndarray)) else ebn0
        x = add_iq_noise(x, e, rng)
        if f0_max > 0:
            x = rotate_iq(x, float(rng.uniform(-f0_max, f0_max)))
        X[c] = x
    return X, pl


def baseband_channels(port, nchan, nframes, seed, ebn0_db):
    """2-sps baseband seam (m17_test.cpp:29-51): dibit levels {+1/3,+1,-1/3,-1} through the 62-tap 2-sps RRC
    (sum-normalised, x2), AWGN with Eb/N0 defined at the matched-filter output.
    Returns (disc float32 [C][T*384], payloads)."""
    rng = np.random.default_rng(seed)
    taps = port.set_gain(port.rrc(0.5, 62, 2), 1.0, 1, 62).astype(np.float64) * 2.0
    lu = np.array([1 / 3, 1.0, -1 / 3, -1.0])
    mf, _ = port.sync_taps()
    br = mf[10].astype(np.float64)
    pl = rng.integers(0, 256, (nchan, nframes, 16), dtype=np.uint8)
    out = []
    for c in range(nchan):
        lsf = lsf_for(port)
        _, dib, _ = port.tx_stream_over(lsf, pl[c])
        lead = np.full(int(rng.integers(8, 200)), 1, np.uint8)
        d = np.concatenate([lead, dib])
        up = np.zeros(2 * len(d))
        up[::2] = lu[d]
        x = np.convolve(up, taps)[: 2 * len(d)]
        e = ebn0_db[c] if isinstance(ebn0_db, (list, tuple, np.ndarray)) else ebn0_db
        if e is not None:
            es = (1 + 1 / 9) / 2
            sig_out = np.sqrt(es / 2 / (2 * 10 ** (e / 10)))
            x = x + rng.normal(0, sig_out / np.sqrt((br ** 2).sum()), len(x))
        out.append(x.astype(np.float32))
    n = max(len(x) for x in out)
    T = n // 384 + 2
    D = np.zeros((nchan, T * 384), np.float32)
    for c in range(nchan):
        D[c, : len(out[c])] = out[c]
    return D, pl


def packet_frames(port, data):
    """m17_send_packet_frames semantics (m17_tx_routines.cpp:323-353): append CRC-16, split in 25-byte chunks,
    non-final frames carry their index, the final frame EOF + the number of bytes used (25 if the split is exact)."""
    crc = port.crc(bytes(data))
    buf = bytes(data) + bytes([crc >> 8, crc & 0xFF])
    frames, left = len(buf) // 25, len(buf) % 25
    out = []
    if left == 0:
        for i in range(frames - 1):
            out.append(port.fmt_packet(buf[i * 25:(i + 1) * 25], 0, i))
        out.append(port.fmt_packet(buf[(frames - 1) * 25:], 1, 25))
    else:
        for i in range(frames):
            out.append(port.fmt_packet(buf[i * 25:(i + 1) * 25], 0, i))
        out.append(port.fmt_packet(buf[frames * 25:], 1, left))
    return out


def packet_channels(port, nchan, seed, ebn0=None, f0_max=1000.0, max_len=200):
    """Config-3 style: preamble x2, LSF (TYPE 0x0002 packet/data), one packet of 1..max_len bytes, EOT; carrier
    offset; fractional timing offset by modulating at os=80 and decimating by 8 from a random phase; integer delay.
    Returns (iq int16 [C][T*1920][2], list of packet payloads)."""
    rng = np.random.default_rng(seed)
    chans, packets = [], []
    for c in range(nchan):
        data = bytes(rng.integers(0, 256, int(rng.integers(1, max_len + 1)), dtype=np.uint8))
        packets.append(data)
        lsf = port.build_lsf(0xFFFFFFFFFFFF, port.encode_call("G4GUO    "), 0x0002)
        script = np.concatenate([np.full(192, 4, np.uint8), port.fmt_preamble(), port.fmt_preamble(), port.fmt_lsf(lsf)] +
                                packet_frames(port, data) + [port.fmt_eot(), np.full(384, 4, np.uint8)])
        iq80 = port.mod(script, 80)
        chans.append(iq80[int(rng.integers(0, 8))::8][: len(script) * 10])
    n = max(len(x) for x in chans)
    T = (n + BLOCK - 1) // BLOCK + 2
    X = np.zeros((nchan, T * BLOCK, 2), np.int16)
    for c in range(nchan):
        x = delay_iq(chans[c], int(rng.integers(0, BLOCK)), T * BLOCK)
        e = ebn0[c] if isinstance(ebn0, (list, tuple, np.ndarray)) else ebn0
        x = add_iq_noise(x, e, rng)
        X[c] = rotate_iq(x, float(rng.uniform(-f0_max, f0_max)))
    return X, packets
