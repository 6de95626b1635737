"""Host-side mirror of the m17gismo function surface, batched, on top of the C ABI (include/m17b200.h).

PyTorch is plumbing only: device memory (torch tensors), streams, and torch.distributed in dist.py.  Every
method hands raw device pointers to libm17b200.so; nothing here computes on the CPU and there is no fallback
path -- without the library or a CUDA device the calls raise.

Naming follows the reference (m17defines.h): m17_dsp_rx, m17_rx_sync_samples+m17_rx_symbols (baseband seam),
m17_viterbi_decode, m_17_golay_decode, m17_crc_array_encode, m17_mod_dibits, ...
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _l

REC_DTYPE = np.dtype([
    ("sym_off", "<i4"), ("type", "u1"), ("flags", "u1"), ("golay_err", "u1"), ("nbytes", "u1"),
    ("lich", "u1", (6,)), ("data", "u1", (30,)), ("crc", "<u2"), ("votes", "u1"), ("frame_errors", "u1"),
    ("variance", "<f4"), ("cor", "<f4"), ("rsvd", "u1", (8,)),
])
EV_DTYPE = np.dtype([("sym_idx", "<i4"), ("kind", "<i4")])
GPS_DTYPE = np.dtype([("lat", "<f8"), ("lon", "<f8"), ("alt", "<i4"), ("course", "<i4"), ("speed", "<i4"), ("object", "<i4")])
BLOCK, DISC_PER_BLOCK, FRAME_SYMS = 1920, 384, 192
STAT_NAMES = ("frames", "stream_frames", "golay_errors", "delivered", "aos", "los", "lsf_events", "symbols")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_dev(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise TypeError(f"{name}: expected a contiguous CUDA tensor of dtype {dtype}")
    return t


class _DevView:
    """zero-copy torch view of a library-owned device buffer (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class Context:
    """Per-GPU tables and filter banks; replaces the reference init chain (main.cpp:108-126)."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _l.M17Error("no CUDA device: m17_sdr_b200 has no CPU path")
        self.L = _l.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        _l.check(self.L.m17b_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.m17b_ctx_destroy(self.h)
            self.h = None

    def _new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ---- filter design (host)
    def m17_dsp_build_rrc_filter(self, rolloff, ntaps, sps):
        out = np.zeros(ntaps, np.float32)
        _l.check(self.L.m17b_build_rrc_filter(out.ctypes.data_as(C.c_void_p), rolloff, ntaps, sps))
        return out

    def m17_dsp_set_filter_gain(self, taps, gain, stride, ntaps):
        a = np.array(taps, np.float32)
        _l.check(self.L.m17b_set_filter_gain(a.ctypes.data_as(C.c_void_p), gain, stride, ntaps))
        return a

    def sync_taps(self):
        mf, md = np.zeros((40, 31), np.float32), np.zeros((40, 31), np.float32)
        _l.check(self.L.m17b_get_sync_taps(self.h, mf.ctypes.data_as(C.c_void_p), md.ctypes.data_as(C.c_void_p)))
        return mf, md

    # ---- M17-over-UDP reflector frames (m17_net.cpp:25-74,203-238; m17_tx_routines.cpp:54-86,298-306)
    def net_pack(self, sid, lsf, fn, payload, dst=None):
        """sid uint16 [n], lsf uint8 [n][>=28], fn uint16 [n], payload uint8 [n][16] -> uint8 [n][54] datagrams."""
        _chk_dev(sid, torch.uint16, "sid"); _chk_dev(lsf, torch.uint8, "lsf"); _chk_dev(fn, torch.uint16, "fn"); _chk_dev(payload, torch.uint8, "payload")
        n = sid.shape[0]
        out = self._new((n, 54), torch.uint8)
        _l.check(self.L.m17b_net_pack(self.h, _ptr(sid), _ptr(lsf), lsf.shape[1], 0 if dst is None else 1, int(dst or 0), _ptr(fn), _ptr(payload), n,
                                      _ptr(out), _stream()))
        return out

    def net_parse(self, frames):
        """uint8 [n][54] -> (ok uint8 [n], sid uint16 [n], lsf uint8 [n][30], fn uint16 [n], payload uint8 [n][16])."""
        _chk_dev(frames, torch.uint8, "frames")
        n = frames.shape[0]
        ok, sid, lsf = self._new((n,), torch.uint8), self._new((n,), torch.uint16), self._new((n, 30), torch.uint8)
        fn, pld = self._new((n,), torch.uint16), self._new((n, 16), torch.uint8)
        _l.check(self.L.m17b_net_parse(self.h, _ptr(frames), n, _ptr(ok), _ptr(sid), _ptr(lsf), _ptr(fn), _ptr(pld), _stream()))
        return ok, sid, lsf, fn, pld

    def m17_dsp_demap_symbol(self, sym, mag):
        _chk_dev(sym, torch.float32, "sym"); _chk_dev(mag, torch.float32, "mag")
        out = self._new((sym.numel(), 2), torch.float32)
        _l.check(self.L.m17b_demap_symbols(self.h, _ptr(sym), _ptr(mag), sym.numel(), _ptr(out), _stream()))
        return out

    def m17_dsp_decimating_filter(self, x, coffs, stride, length):
        """x float32 [n][>= length + len(coffs) - 1] -> [n][ceil(length/stride)]."""
        _chk_dev(x, torch.float32, "x"); _chk_dev(coffs, torch.float32, "coffs")
        n, nout = x.shape[0], (length + stride - 1) // stride
        out = self._new((n, nout), torch.float32)
        ol = C.c_int()
        _l.check(self.L.m17b_dsp_decimating_filter(self.h, _ptr(x), x.shape[1], _ptr(coffs), stride, coffs.numel(), length, n, _ptr(out), C.byref(ol), _stream()))
        assert ol.value == nout
        return out

    def m17_prbs9_rx_check(self, bits, state=None):
        """bits uint8 [n][nbits]; state int32 [n][8] (updated in place, zeros if None) -> state."""
        _chk_dev(bits, torch.uint8, "bits")
        if state is None:
            state = torch.zeros((bits.shape[0], 8), dtype=torch.int32, device=bits.device)
        _l.check(self.L.m17b_prbs9_rx_check(self.h, _ptr(bits), bits.shape[1], bits.shape[0], _ptr(state), _stream()))
        return state

    # ---- bit-domain primitives (CUDA tensors in, CUDA tensors out)
    def m17_crc_array_encode(self, data):
        _chk_dev(data, torch.uint8, "data")
        n, ln = data.shape
        out = self._new((n,), torch.uint16)
        _l.check(self.L.m17b_crc_array_encode(self.h, _ptr(data), ln, ln, n, _ptr(out), _stream()))
        return out

    def m17_golay_encode(self, data12):
        _chk_dev(data12, torch.uint16, "data12")
        out = self._new(data12.shape, torch.int32)
        _l.check(self.L.m17b_golay_encode(self.h, _ptr(data12), data12.numel(), _ptr(out), _stream()))
        return out

    def m_17_golay_decode(self, words24):
        _chk_dev(words24, torch.int32, "words24")
        data = self._new(words24.shape, torch.uint16)
        err = self._new(words24.shape, torch.uint8)
        _l.check(self.L.m17b_golay_decode(self.h, _ptr(words24), words24.numel(), _ptr(data), _ptr(err), _stream()))
        return data, err

    def m17_conv_encode_8(self, data):
        _chk_dev(data, torch.uint8, "data")
        n, nb = data.shape
        out = self._new((n, 2 * (8 * nb + 4)), torch.uint8)
        _l.check(self.L.m17b_conv_encode_8(self.h, _ptr(data), nb, n, _ptr(out), _stream()))
        return out

    def m17_conv_encode_1(self, bits):
        _chk_dev(bits, torch.uint8, "bits")
        n, nb = bits.shape
        out = self._new((n, 2 * (nb + 4)), torch.uint8)
        _l.check(self.L.m17b_conv_encode_1(self.h, _ptr(bits), nb, n, _ptr(out), _stream()))
        return out

    def m17_viterbi_decode(self, soft):
        _chk_dev(soft, torch.float32, "soft")
        n, ln = soft.shape
        out = self._new((n, ln // 2), torch.uint8)
        _l.check(self.L.m17b_viterbi_decode(self.h, _ptr(soft), ln, n, _ptr(out), _stream()))
        return out

    def m17_punc(self, pattern, bits):
        _chk_dev(bits, torch.uint8, "bits")
        n, ln = bits.shape
        kept = C.c_int()
        out = self._new((n, ln), torch.uint8)
        _l.check(self.L.m17b_punc(self.h, pattern, _ptr(bits), ln, n, _ptr(out), C.byref(kept), _stream()))
        # the kernel writes rows of `kept` bytes back to back
        return out.view(-1)[: n * kept.value].view(n, kept.value)

    def m17_de_punc(self, pattern, soft, out_len):
        _chk_dev(soft, torch.float32, "soft")
        n, ln = soft.shape
        out = self._new((n, out_len), torch.float32)
        _l.check(self.L.m17b_de_punc(self.h, pattern, _ptr(soft), ln, out_len, n, _ptr(out), _stream()))
        return out

    def m17_interleave(self, bits):
        _chk_dev(bits, torch.uint8, "bits")
        out = torch.empty_like(bits)
        _l.check(self.L.m17b_interleave(self.h, _ptr(bits), bits.shape[0], _ptr(out), _stream()))
        return out

    def m17_de_interleave(self, soft):
        _chk_dev(soft, torch.float32, "soft")
        out = torch.empty_like(soft)
        _l.check(self.L.m17b_de_interleave(self.h, _ptr(soft), soft.shape[0], _ptr(out), _stream()))
        return out

    def m17_de_correlate_8(self, data):
        _chk_dev(data, torch.uint8, "data")
        out = data.clone()
        _l.check(self.L.m17b_de_correlate_8(self.h, _ptr(out), out.shape[1], out.shape[0], _stream()))
        return out

    def m17_de_correlate_1(self, x):
        out = torch.empty_like(x)
        n, ln = x.shape
        if x.dtype == torch.uint8:
            _l.check(self.L.m17b_de_correlate_1_u8(self.h, _ptr(_chk_dev(x, torch.uint8, "x")), _ptr(out), ln, n, _stream()))
        else:
            _l.check(self.L.m17b_de_correlate_1_f32(self.h, _ptr(_chk_dev(x, torch.float32, "x")), _ptr(out), ln, n, _stream()))
        return out

    def m17_dsp_demap_frame(self, sym):
        _chk_dev(sym, torch.float32, "sym")
        out = self._new((sym.shape[0], 368), torch.float32)
        _l.check(self.L.m17b_demap_frame(self.h, _ptr(sym), sym.shape[0], _ptr(out), _stream()))
        return out

    def m17_sync_check(self, vec8):
        _chk_dev(vec8, torch.float32, "vec8")
        n = vec8.shape[0]
        ty, vo, va = self._new((n,), torch.uint8), self._new((n,), torch.uint8), self._new((n,), torch.float32)
        _l.check(self.L.m17b_sync_check(self.h, _ptr(vec8), n, _ptr(ty), _ptr(vo), _ptr(va), _stream()))
        return ty, vo, va

    def m17_prbs9_tx_load(self, n, length, start=None):
        out = self._new((n, length), torch.uint8)
        _l.check(self.L.m17b_prbs9_tx_load(self.h, _ptr(start), length, n, _ptr(out), _stream()))
        return out

    def m17_rx_parse(self, sym, types, want_soft=False):
        """n independent frames [n][192] + sync types -> records (no cross-frame LICH state)."""
        _chk_dev(sym, torch.float32, "sym"); _chk_dev(types, torch.uint8, "types")
        n = sym.shape[0]
        rec = self._new((n, 64), torch.uint8)
        soft = self._new((n, 368), torch.float32) if want_soft else None
        _l.check(self.L.m17b_rx_parse_frames(self.h, _ptr(sym), _ptr(types), n, _ptr(rec), _ptr(soft), _stream()))
        return rec, soft

    def viterbi_punctured(self, pattern, soft):
        _chk_dev(soft, torch.float32, "soft")
        n = soft.shape[0]
        nb = {1: 30, 2: 18, 3: 26}[pattern]
        out = self._new((n, nb), torch.uint8)
        _l.check(self.L.m17b_viterbi_punctured(self.h, pattern, _ptr(soft), n, _ptr(out), _stream()))
        return out

    def selftest_frontend(self, first=0, count=1 << 32):
        n = C.c_uint64()
        dump = np.zeros(5 * 16, np.uint32)
        _l.check(self.L.m17b_selftest_frontend(self.h, first, count, C.byref(n), dump.ctypes.data_as(C.c_void_p), len(dump), _stream()))
        self.last_selftest_dump = dump.reshape(-1, 5)[: min(16, n.value)]
        return n.value

    def selftest_limiter(self, first=0x00800000, count=0x7F800000 - 0x00800000):
        """fast limiter normaliser vs IEEE sqrt / divide over float bit patterns of s (default: every positive normal float)"""
        n = C.c_uint64()
        dump = np.zeros(5 * 16, np.uint32)
        _l.check(self.L.m17b_selftest_limiter(self.h, first, count, C.byref(n), dump.ctypes.data_as(C.c_void_p), len(dump), _stream()))
        self.last_selftest_dump = dump.reshape(-1, 5)[: min(16, n.value)]
        return n.value

    def gps_decode(self, lsf):
        """gps_decode (gps.cpp:8-27) of the META field: lsf uint8 [n][>=30] -> numpy structured array (lat, lon, alt, course, speed, object)."""
        _chk_dev(lsf, torch.uint8, "lsf")
        n = lsf.shape[0]
        out = torch.empty((n, 32), dtype=torch.uint8, device=lsf.device)
        _l.check(self.L.m17b_gps_decode(self.h, _ptr(lsf), lsf.shape[1], n, _ptr(out), _stream()))
        return out.cpu().numpy().view(GPS_DTYPE).reshape(n)

    def selftest_tx_wrap(self, first=0, count=1 << 32):
        n = C.c_uint64()
        dump = np.zeros(3 * 16, np.uint32)
        _l.check(self.L.m17b_selftest_tx_wrap(self.h, first, count, C.byref(n), dump.ctypes.data_as(C.c_void_p), 16, _stream()))
        self.last_selftest_dump = dump.reshape(-1, 3)[: min(16, n.value)]
        return n.value

    def selftest_demap(self, first=0, count=1 << 32):
        n = C.c_uint64()
        dump = np.zeros(3 * 16, np.uint32)
        _l.check(self.L.m17b_selftest_demap(self.h, first, count, C.byref(n), dump.ctypes.data_as(C.c_void_p), len(dump), _stream()))
        self.last_selftest_dump = dump.reshape(-1, 3)[: min(16, n.value)]
        return n.value

    def synth_channel(self, iq, sigma=None, f0=None, seed=1):
        _chk_dev(iq, torch.int16, "iq")
        nchan, nsamp = iq.shape[0], iq.shape[1]
        _l.check(self.L.m17b_synth_channel(self.h, _ptr(iq), nchan, nsamp, _ptr(sigma), _ptr(f0), seed, _stream()))
        return iq


def records_to_numpy(rec_u8, nframes=None):
    """uint8 tensor [..., 64] -> numpy structured array of m17b_frame_rec."""
    a = rec_u8.cpu().numpy()
    return a.view(REC_DTYPE).reshape(a.shape[:-1])


class Rx:
    """Batched m17_dsp_rx: nchan independent channels, up to max_blocks 40-ms blocks per call."""

    def __init__(self, ctx, nchan, max_blocks):
        self.ctx, self.L = ctx, ctx.L
        self.nchan, self.max_blocks = nchan, max_blocks
        h = C.c_void_p()
        _l.check(self.L.m17b_rx_create(ctx.h, nchan, max_blocks, C.byref(h)))
        self.h = h
        self.frame_cap = self.L.m17b_rx_frame_cap(h)

    def close(self):
        if self.h:
            self.L.m17b_rx_destroy(self.h)
            self.h = None

    def reset(self):
        _l.check(self.L.m17b_rx_reset(self.h, _stream()))

    def m17_rx_lost(self):
        """m17_rx_init / m17_rx_lost (m17_rx_frame.cpp:179-186): clear the sync window and the lock flag only."""
        _l.check(self.L.m17b_rx_framer_reset(self.h, _stream()))

    def set_afc(self, on):
        _l.check(self.L.m17b_rx_set_afc(self.h, int(bool(on)), _stream()))

    def set_equaliser(self, on):
        """Equaliser option: eq_train_unknown (m17_equalize.cpp:185-213) on T/2 pairs between the timing loop and the framer
        (no call site upstream; off = upstream behaviour).  Not together with AFC."""
        _l.check(self.L.m17b_rx_set_equaliser(self.h, int(bool(on)), _stream()))

    def overflow(self):
        """sticky capacity flags since the last reset (bit 0: the symbol seam was given more symbols than max_blocks*200)"""
        v = C.c_int()
        _l.check(self.L.m17b_rx_get_overflow(self.h, C.byref(v)))
        return v.value

    def set_bert(self, on):
        """BERT receive extension: decode BERT frames and run m17_prbs9_rx_check on their bits (off = upstream behaviour)."""
        _l.check(self.L.m17b_rx_set_bert(self.h, int(bool(on))))

    def bert(self):
        """uint32 [nchan][8]: state, idx, bad, good, eq_cnt, dif_cnt (m17_prbs9.cpp:7-12), bits checked in sync, bit errors."""
        out = torch.empty((self.nchan, 8), dtype=torch.int32, device=self.ctx.device)
        _l.check(self.L.m17b_rx_get_bert(self.h, _ptr(out), _stream()))
        return out

    def m17_dsp_rx(self, iq):
        """iq: int16 CUDA tensor [nchan][nblocks*1920][2]."""
        _chk_dev(iq, torch.int16, "iq")
        assert iq.shape[0] == self.nchan and iq.shape[1] % BLOCK == 0 and iq.shape[2] == 2
        _l.check(self.L.m17b_dsp_rx(self.h, _ptr(iq), iq.shape[1] // BLOCK, _stream()))
        return self

    def m17_rx_baseband(self, disc):
        """disc: float32 CUDA tensor [nchan][nblocks*384] (m17_rx_sync_samples + m17_rx_symbols seam)."""
        _chk_dev(disc, torch.float32, "disc")
        assert disc.shape[0] == self.nchan and disc.shape[1] % DISC_PER_BLOCK == 0
        _l.check(self.L.m17b_rx_baseband(self.h, _ptr(disc), disc.shape[1] // DISC_PER_BLOCK, _stream()))
        return self

    def m17_rx_symbols(self, syms, nsym):
        """Symbol seam: syms float32 CUDA [nchan][pitch], nsym int32 CUDA [nchan] (framer + decode + post only)."""
        _chk_dev(syms, torch.float32, "syms"); _chk_dev(nsym, torch.int32, "nsym")
        assert syms.shape[0] == self.nchan
        _l.check(self.L.m17b_rx_symbols(self.h, _ptr(syms), syms.shape[1], _ptr(nsym), _stream()))
        return self

    def m17_dsp_rx_host(self, iq_host, frames_host=None, nframes_host=None):
        """End-to-end form: iq_host is a (preferably pinned) CPU int16 tensor; returns (records uint8 [nchan][cap][64], nframes)."""
        assert iq_host.dtype == torch.int16 and not iq_host.is_cuda and iq_host.is_contiguous()
        nblocks = iq_host.shape[1] // BLOCK
        if frames_host is None:
            frames_host = torch.empty((self.nchan, self.frame_cap, 64), dtype=torch.uint8).pin_memory()
            nframes_host = torch.empty((self.nchan,), dtype=torch.int32).pin_memory()
        _l.check(self.L.m17b_dsp_rx_host(self.h, _ptr(iq_host), nblocks, _ptr(frames_host), _ptr(nframes_host), _stream()))
        return frames_host, nframes_host

    def reassemble_packets(self, bytes_cap=1024, max_pkts=8):
        """Packets carried by the last call's packet frames, per channel (batched on the GPU; a packet may span calls).
        Returns (bytes uint8 [nchan][bytes_cap], pkt int32 [nchan][max_pkts][3] = offset / length / crc_ok, npkt int32 [nchan])."""
        dev = self.ctx.device
        by = torch.empty((self.nchan, bytes_cap), dtype=torch.uint8, device=dev)
        pk = torch.zeros((self.nchan, max_pkts, 3), dtype=torch.int32, device=dev)
        npk = torch.empty((self.nchan,), dtype=torch.int32, device=dev)
        _l.check(self.L.m17b_rx_reassemble_packets(self.h, _ptr(by), bytes_cap, _ptr(pk), max_pkts, _ptr(npk), _stream()))
        return by, pk, npk

    def m17_net_new_rx_data(self, sid, dst=None):
        """Gateway output of the last call: every delivered stream frame as a 54-byte M17-over-UDP datagram.
        sid uint16 [nchan] -> (uint8 [nchan][frame_cap][54], count int32 [nchan])."""
        _chk_dev(sid, torch.uint16, "sid")
        out = torch.empty((self.nchan, self.frame_cap, 54), dtype=torch.uint8, device=sid.device)
        cnt = torch.empty((self.nchan,), dtype=torch.int32, device=sid.device)
        _l.check(self.L.m17b_rx_net_frames(self.h, _ptr(sid), 0 if dst is None else 1, int(dst or 0), _ptr(out), _ptr(cnt), _stream()))
        return out, cnt

    def debug_sync(self):
        """int64 [nchan][8]: SM cycles, speculation rounds, per-phase cycles of the last timing-loop launch, per channel."""
        out = torch.zeros((self.nchan, 8), dtype=torch.int64, device=self.ctx.device)
        _l.check(self.L.m17b_rx_debug_sync(self.h, _ptr(out), _stream()))
        return out

    def set_slice_blocks(self, blocks):
        """Blocks per pipeline slice (0 = run the stages strictly in sequence); results do not depend on it."""
        _l.check(self.L.m17b_rx_set_slice_blocks(self.h, int(blocks)))

    def set_chan_groups(self, groups):
        """Run the batch as `groups` independent channel groups on their own streams (0 / 1 = one chain); results do not depend on it."""
        _l.check(self.L.m17b_rx_set_chan_groups(self.h, int(groups)))

    def set_timing(self, on=True):
        _l.check(self.L.m17b_rx_set_timing(self.h, int(on)))

    def stage_ms(self, call_index):
        out = (C.c_float * 4)()
        _l.check(self.L.m17b_rx_stage_ms(self.h, call_index, out))
        return dict(zip(("frontend", "sync_frame", "decode", "post"), [float(x) for x in out]))

    def launches(self):
        return self.L.m17b_rx_last_launches(self.h)

    def view(self):
        """Zero-copy torch views of the last call's device results."""
        v = _l.RxView()
        _l.check(self.L.m17b_rx_get_view(self.h, C.byref(v)))
        dev = self.ctx.device
        n, T = v.nchan, v.nblocks
        out = {
            "frames": _view(v.d_frames, (n, v.frame_cap, 64), "|u1", dev),
            "nframes": _view(v.d_nframes, (n,), "<i4", dev),
            "syms": _view(v.d_syms, (n, v.sym_pitch), "<f4", dev),
            "nsym": _view(v.d_nsym, (n, T), "<i4", dev),
            "sym_base": _view(v.d_sym_base, (n,), "<i4", dev),
            "events": _view(v.d_events, (n, v.event_cap, 2), "<i4", dev),
            "nevents": _view(v.d_nevents, (n,), "<i4", dev),
            "stats": _view(v.d_stats, (n, 8), "<i8", dev),
            "sym_carry": v.sym_carry,
            "disc": _view(v.d_disc, (n, T, DISC_PER_BLOCK), "<f4", dev) if v.d_disc else None,
            "mean": _view(v.d_mean, (n, T), "<f4", dev) if v.d_mean else None,
        }
        return out

    def results(self):
        """Host copy of the last call's results in the oracle's layout (see tests/m17_oracles.py)."""
        torch.cuda.synchronize()
        v = self.view()
        r = {
            "frames": records_to_numpy(v["frames"]),
            "nframes": v["nframes"].cpu().numpy(),
            "nsym": v["nsym"].cpu().numpy(),
            "syms": v["syms"][:, v["sym_carry"]:].cpu().numpy(),
            "sym_base": v["sym_base"].cpu().numpy(),
            "events": v["events"].cpu().numpy().view(EV_DTYPE).reshape(self.nchan, -1),
            "nevents": v["nevents"].cpu().numpy(),
            "stats": v["stats"].cpu().numpy(),
        }
        if v["disc"] is not None:
            r["disc_raw"] = v["disc"].cpu().numpy()
            r["mean"] = v["mean"].cpu().numpy()
        return r


class Tx:
    """Batched TX formatter + modulator (m17_tx_routines.cpp, m17_modulate.cpp)."""

    def __init__(self, ctx, nchan, oversample=10):
        self.ctx, self.L, self.nchan, self.os = ctx, ctx.L, nchan, oversample
        h = C.c_void_p()
        _l.check(self.L.m17b_tx_create(ctx.h, nchan, oversample, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.m17b_tx_destroy(self.h)
            self.h = None

    def reset(self):
        _l.check(self.L.m17b_tx_reset(self.h, _stream()))

    def set_lsf(self, lsf):
        _chk_dev(lsf, torch.uint8, "lsf")
        assert lsf.shape == (self.nchan, 30)
        _l.check(self.L.m17b_tx_set_lsf(self.h, _ptr(lsf), _stream()))

    def fmt_preamble(self):
        d = np.zeros(192, np.uint8)
        _l.check(self.L.m17b_fmt_preamble(d.ctypes.data_as(C.c_void_p)))
        return d

    def fmt_eot(self):
        d = np.zeros(192, np.uint8)
        _l.check(self.L.m17b_fmt_eot(d.ctypes.data_as(C.c_void_p)))
        return d

    def m17_fmt_add_link_setup_frame(self, lsf):
        _chk_dev(lsf, torch.uint8, "lsf")
        out = torch.empty((lsf.shape[0], 192), dtype=torch.uint8, device=lsf.device)
        _l.check(self.L.m17b_fmt_link_setup_frame(self.ctx.h, _ptr(lsf), lsf.shape[0], _ptr(out), _stream()))
        return out

    def m17_fmt_add_stream_frame(self, payload):
        """payload uint8 [nchan][F][16] -> dibits [nchan][F][192]; advances m_fn / m_lich_count per channel."""
        _chk_dev(payload, torch.uint8, "payload")
        F = payload.shape[1]
        out = torch.empty((self.nchan, F, 192), dtype=torch.uint8, device=payload.device)
        _l.check(self.L.m17b_fmt_stream_frames(self.h, _ptr(payload), F, _ptr(out), _stream()))
        return out

    def m17_fmt_add_packet(self, chunks, meta):
        _chk_dev(chunks, torch.uint8, "chunks"); _chk_dev(meta, torch.uint8, "meta")
        n = chunks.shape[0]
        out = torch.empty((n, 192), dtype=torch.uint8, device=chunks.device)
        _l.check(self.L.m17b_fmt_packet_frames(self.ctx.h, _ptr(chunks), _ptr(meta), n, _ptr(out), _stream()))
        return out

    def m17_send_packet_frames(self, packets, lengths, max_frames=32):
        """packets uint8 [n][stride], lengths int32 [n] -> (dibits uint8 [n][max_frames][192], nframes int32 [n])."""
        _chk_dev(packets, torch.uint8, "packets"); _chk_dev(lengths, torch.int32, "lengths")
        n = packets.shape[0]
        out = torch.empty((n, max_frames, 192), dtype=torch.uint8, device=packets.device)
        nf = torch.empty((n,), dtype=torch.int32, device=packets.device)
        _l.check(self.L.m17b_send_packet_frames(self.ctx.h, _ptr(packets), packets.shape[1], _ptr(lengths), n, max_frames, _ptr(out), _ptr(nf), _stream()))
        return out, nf

    def m17_fmt_add_bert_frame(self, F):
        out = torch.empty((self.nchan, F, 192), dtype=torch.uint8, device=self.ctx.device)
        _l.check(self.L.m17b_fmt_bert_frames(self.h, F, _ptr(out), _stream()))
        return out

    def debug_scan(self):
        """{wait, busy} cycles of the phase-scan warps, chunks and CTAs of the last m17_mod_dibits call (summed over CTAs)."""
        out = (C.c_uint64 * 8)()
        _l.check(self.L.m17b_tx_debug_scan(self.h, out))
        return {"scan_wait_cycles": int(out[0]), "scan_busy_cycles": int(out[1]), "chunks": int(out[2]), "ctas": int(out[3]),
                "worker_fill_cycles": int(out[4]), "worker_fir_cycles": int(out[5]), "worker_wait_cycles": int(out[6]), "worker_emit_cycles": int(out[7])}

    def m17_mod_dibits(self, syms, want_freq=False, out=None):
        """syms uint8 [nchan][nsym] (0..3 dibits, 4 = blank carrier) -> int16 IQ [nchan][nsym*os][2]."""
        _chk_dev(syms, torch.uint8, "syms")
        nsym = syms.shape[1]
        iq = out if out is not None else torch.empty((self.nchan, nsym * self.os, 2), dtype=torch.int16, device=syms.device)
        freq = torch.empty((self.nchan, nsym * self.os), dtype=torch.float32, device=syms.device) if want_freq else None
        _l.check(self.L.m17b_mod_dibits(self.h, _ptr(syms), nsym, _ptr(iq), _ptr(freq), _stream()))
        return (iq, freq) if want_freq else iq


class Decimator:
    """Batched Pluto front-end decimator: radio_receive_samples' /8 int16 FIR (radio.cpp:18-51,157-177)."""

    def __init__(self, ctx, nchan):
        self.ctx, self.L, self.nchan = ctx, ctx.L, nchan
        h = C.c_void_p()
        _l.check(self.L.m17b_dec_create(ctx.h, nchan, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.m17b_dec_destroy(self.h)
            self.h = None

    def reset(self):
        _l.check(self.L.m17b_dec_reset(self.h, _stream()))

    def taps(self):
        t = np.zeros(31, np.int16)
        _l.check(self.L.m17b_dec_get_taps(self.h, t.ctypes.data_as(C.c_void_p)))
        return t

    def radio_receive_samples(self, iq384):
        """iq384: int16 CUDA tensor [nchan][8*nout][2] at 384 kS/s -> int16 [nchan][nout][2] at 48 kS/s."""
        _chk_dev(iq384, torch.int16, "iq384")
        assert iq384.shape[0] == self.nchan and iq384.shape[1] % 32 == 0 and iq384.shape[2] == 2
        nout = iq384.shape[1] // 8
        out = torch.empty((self.nchan, nout, 2), dtype=torch.int16, device=iq384.device)
        _l.check(self.L.m17b_dec_run(self.h, _ptr(iq384), nout, _ptr(out), _stream()))
        return out


class Channelizer:
    """Wideband channeliser: int16 IQ captures at 1.2 MS/s -> 96 channels x 48 kS/s each (the Pluto decimator of radio.cpp:18-40
    generalised to a 12.5 kHz raster; integer arithmetic, bit-exact against the oracle)."""
    M, D = 96, 25

    def __init__(self, ctx, ncap, taps_per_branch=12):
        self.ctx, self.L, self.ncap = ctx, ctx.L, ncap
        h = C.c_void_p()
        _l.check(self.L.m17b_chan_create(ctx.h, ncap, taps_per_branch, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.m17b_chan_destroy(self.h)
            self.h = None

    def reset(self):
        _l.check(self.L.m17b_chan_reset(self.h, _stream()))

    def taps(self):
        t = np.zeros(96 * 16, np.int16)
        n = C.c_int()
        _l.check(self.L.m17b_chan_get_taps(self.h, t.ctypes.data_as(C.c_void_p), C.byref(n)))
        return t[: n.value].copy()

    def run(self, wide, out=None):
        """wide: int16 CUDA tensor [ncap][25*nout][2] -> int16 [ncap*96][nout][2]"""
        _chk_dev(wide, torch.int16, "wide")
        assert wide.shape[0] == self.ncap and wide.shape[1] % 25 == 0 and wide.shape[2] == 2
        nout = wide.shape[1] // 25
        if out is None:
            out = torch.empty((self.ncap * 96, nout, 2), dtype=torch.int16, device=wide.device)
        _l.check(self.L.m17b_chan_run(self.h, _ptr(wide), nout, _ptr(out), out.shape[1], _stream()))
        return out


class Equalizer:
    """Batched eq_train_known / eq_train_unknown (m17_equalize.cpp)."""

    def __init__(self, ctx, nchan):
        self.ctx, self.L, self.nchan = ctx, ctx.L, nchan
        h = C.c_void_p()
        _l.check(self.L.m17b_eq_create(ctx.h, nchan, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.m17b_eq_destroy(self.h)
            self.h = None

    def eq_reset(self):
        _l.check(self.L.m17b_eq_reset(self.h, _stream()))

    def eq_restart(self):
        _l.check(self.L.m17b_eq_restart(self.h, _stream()))

    def eq_train(self, pairs, train=None):
        _chk_dev(pairs, torch.float32, "pairs")
        nsym = pairs.shape[1]
        out = torch.empty((self.nchan, nsym), dtype=torch.float32, device=pairs.device)
        _l.check(self.L.m17b_eq_train(self.h, _ptr(pairs), _ptr(train), nsym, _ptr(out), _stream()))
        return out


def packets_of(bytes_, pkt, npkt, c):
    """host view of one channel's reassembled packets: list of (payload bytes, crc_ok)"""
    b = bytes_[c].cpu().numpy(); p = pkt[c].cpu().numpy(); n = int(npkt[c])
    return [(bytes(b[p[k, 0]: p[k, 0] + p[k, 1]]), bool(p[k, 2])) for k in range(n)]
