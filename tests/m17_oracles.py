"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Port``  -- oracle/libm17oracle.so, the plain-C restatement (oracle/m17_oracle.c); thread-safe.
* ``Ref``   -- oracle/_ref/libm17ref.so, the UNMODIFIED reference objects + our harness
               (oracle/ref/ref_shim.cpp); one channel per forked process.

Both expose the same batch RX contract and produce identical record arrays, so tests compare
``Port`` against ``Ref`` (pinning the restatement) and the CUDA product against ``Port``.
"""
import ctypes as C
import mmap
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

REC_DTYPE = np.dtype([
    ("sym_off", "<i4"), ("type", "u1"), ("flags", "u1"), ("golay_err", "u1"), ("nbytes", "u1"),
    ("lich", "u1", (6,)), ("data", "u1", (30,)), ("crc", "<u2"), ("votes", "u1"), ("frame_errors", "u1"),
    ("variance", "<f4"), ("cor", "<f4"), ("rsvd", "u1", (8,)),
])
assert REC_DTYPE.itemsize == 64
EV_DTYPE = np.dtype([("sym_idx", "<i4"), ("kind", "<i4")])

F_SYNC_OK, F_PARSED, F_LOS, F_DELIVERED, F_LSF_EVENT, F_PKT_EOF = 1, 2, 4, 8, 16, 32
BLOCK = 1920
DISC_PER_BLOCK = 384
SYM_CAP_PER_BLOCK = 200      # >= 192 + slips

_vp = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def build_oracle():
    """Compile oracle/libm17oracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def shared_array(shape, dtype):
    """numpy array backed by MAP_SHARED anonymous memory (visible to forked children)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    mm = mmap.mmap(-1, max(n, 1))
    a = np.frombuffer(mm, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return a


class _RxOut(dict):
    __getattr__ = dict.__getitem__


def _alloc_rx_out(Cn, T, want_disc, want_soft, alloc):
    symcap = T * SYM_CAP_PER_BLOCK + 8
    fcap = T + 4
    ecap = 2 * T + 8
    o = _RxOut()
    o["disc"] = alloc((Cn, T, DISC_PER_BLOCK), np.float32) if want_disc else None
    o["nsym"] = alloc((Cn, T), np.int32)
    o["syms"] = alloc((Cn, symcap), np.float32)
    o["frames"] = alloc((Cn, fcap), REC_DTYPE)
    o["soft"] = alloc((Cn, fcap, 368), np.float32) if want_soft else None
    o["events"] = alloc((Cn, ecap), EV_DTYPE)
    o["counts"] = alloc((Cn, 4), np.int64)
    o["symcap"], o["fcap"], o["ecap"] = symcap, fcap, ecap
    return o


class Port:
    """The C restatement."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libm17oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.m17o_init()
        L.m17o_crc.restype = C.c_uint16
        L.m17o_golay_encode.restype = C.c_uint32
        L.m17o_hard24.restype = C.c_uint32
        L.m17o_encode_call.restype = C.c_uint64
        L.m17o_decode_call.argtypes = [C.c_uint64, C.c_char_p]
        L.m17o_tx_new.restype = _vp
        L.m17o_rx_new.restype = _vp
        L.m17o_mod.restype = C.c_long
        L.m17o_mod.argtypes = [_vp, _vp, C.c_long, _vp, _vp]
        L.m17o_rx_time.restype = C.c_double
        L.m17o_rx_time.argtypes = [_vp, C.c_long, C.c_long, C.c_int]
        L.m17o_rx_run.argtypes = [_vp, C.c_int, C.c_long, C.c_long, C.c_int, _vp, _vp, _vp, C.c_long, _vp, C.c_long, _vp,
                                  _vp, C.c_long, _vp]
        L.m17o_eq_train_known.restype = C.c_float
        L.m17o_eq_train_known.argtypes = [_vp, _vp, C.c_float]
        L.m17o_eq_train_unknown.restype = C.c_float
        L.m17o_build_lsf.argtypes = [C.c_uint64, C.c_uint64, C.c_uint16, _vp, _vp]
        L.m17o_rrc_design.argtypes = [_vp, C.c_float, C.c_int, C.c_int]
        L.m17o_set_gain.argtypes = [_vp, C.c_float, C.c_int, C.c_int]

    # ---- M17-over-UDP frame (m17_net.cpp:25-74,203-238)
    def net_pack(self, sid, lsf, fn, pld, dst=None):
        out = np.zeros(54, np.uint8)
        self.L.m17o_net_pack(C.c_uint16(sid), _p(np.ascontiguousarray(lsf[:28], np.uint8)), 0 if dst is None else 1, C.c_uint64(dst or 0), C.c_uint16(fn),
                             _p(np.ascontiguousarray(pld, np.uint8)), _p(out))
        return out

    def net_parse(self, b54):
        sid, fn = C.c_uint16(), C.c_uint16()
        lsf, pld = np.zeros(30, np.uint8), np.zeros(16, np.uint8)
        ok = self.L.m17o_net_parse(_p(np.ascontiguousarray(b54, np.uint8)), C.byref(sid), _p(lsf), C.byref(fn), _p(pld))
        return bool(ok), sid.value, lsf, fn.value, pld

    # ---- Pluto /8 front-end decimator (radio.cpp:18-51,157-177)
    def dec_taps(self):
        st = np.zeros(256, np.uint8)
        self.L.m17o_dec_open(_p(st))
        return st[:62].view(np.int16).copy()

    def dec_run(self, x, parts=None):
        """x: int16 [C][8*nout][2] at 384 kS/s -> int16 [C][nout][2]; parts = successive call sizes in outputs (state carry)."""
        x = np.ascontiguousarray(x, np.int16)
        Cn, nout = x.shape[0], x.shape[1] // 8
        out = np.zeros((Cn, nout, 2), np.int16)
        parts = list(parts or [nout])
        assert sum(parts) == nout
        for c in range(Cn):
            st = np.zeros(256, np.uint8)
            self.L.m17o_dec_open(_p(st))
            o = 0
            for n in parts:
                xin = np.ascontiguousarray(x[c, 8 * o:8 * (o + n)]); y = np.zeros((n, 2), np.int16)
                self.L.m17o_dec_run(_p(st), _p(xin), C.c_long(n), _p(y))
                out[c, o:o + n] = y
                o += n
        return out

    # ---- primitives
    def crc(self, b):
        a = np.frombuffer(bytes(b), np.uint8) if len(b) else np.zeros(1, np.uint8)
        return self.L.m17o_crc(_p(a), len(b))

    def golay_encode(self, d):
        return self.L.m17o_golay_encode(int(d))

    def golay_decode(self, w):
        o = C.c_uint16()
        e = self.L.m17o_golay_decode(int(w), C.byref(o))
        return o.value, e

    def golay_errtab(self):
        t = np.zeros(4096, np.uint16)
        self.L.m17o_golay_errtab(_p(t))
        return t

    def conv_encode_8(self, data):
        a = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(16 * len(a) + 16, np.uint8)
        n = self.L.m17o_conv_encode_8(_p(a), _p(out), len(a))
        return out[:n].copy()

    def conv_encode_1(self, bits):
        a = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros(2 * len(a) + 16, np.uint8)
        n = self.L.m17o_conv_encode_1(_p(a), _p(out), len(a))
        return out[:n].copy()

    def viterbi(self, soft):
        a = np.ascontiguousarray(soft, np.float32)
        out = np.zeros(len(a) // 2 + 8, np.uint8)
        n = self.L.m17o_viterbi(_p(a), _p(out), len(a))
        return out[:n].copy()

    def punc(self, p, bits):
        a = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros(len(a) + 8, np.uint8)
        n = self.L.m17o_punc(p, _p(a), _p(out), len(a))
        return out[:n].copy()

    def depunc(self, p, soft, outlen):
        a = np.ascontiguousarray(soft, np.float32)
        out = np.zeros(outlen, np.float32)
        self.L.m17o_depunc(p, _p(a), _p(out), outlen)
        return out

    def interleave(self, bits):
        a = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros(368, np.uint8)
        self.L.m17o_interleave(_p(a), _p(out), len(a))
        return out

    def deinterleave(self, soft):
        a = np.ascontiguousarray(soft, np.float32)
        out = np.zeros(368, np.float32)
        self.L.m17o_deinterleave(_p(a), _p(out), len(a))
        return out

    def derand_bytes(self, b):
        a = np.array(b, np.uint8)
        self.L.m17o_derand_bytes(_p(a), len(a))
        return a

    def derand_bits(self, bits):
        a = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros_like(a)
        self.L.m17o_derand_bits(_p(a), _p(out), len(a))
        return out

    def derand_soft(self, soft):
        a = np.ascontiguousarray(soft, np.float32)
        out = np.zeros_like(a)
        self.L.m17o_derand_soft(_p(a), _p(out), len(a))
        return out

    def demap_frame(self, sym192):
        a = np.ascontiguousarray(sym192, np.float32)
        out = np.zeros(368, np.float32)
        self.L.m17o_demap_frame(_p(a), _p(out))
        return out

    def gps_decode(self, lsf30):
        """gps_decode on the META field of one 30-byte LSF -> (lat, lon, alt, course, speed, object)"""
        b = np.ascontiguousarray(lsf30, np.uint8)[14:30].copy()
        ll = np.zeros(2, np.float64); o = np.zeros(4, np.int32)
        self.L.m17o_gps_decode(_p(b), _p(ll), _p(o))
        return (float(ll[0]), float(ll[1]), int(o[0]), int(o[1]), int(o[2]), int(o[3]))

    def chan_run(self, x, M, D, taps, parts=None):
        """wideband channeliser: x int16 [nin][2] (nin a multiple of D) -> int16 [M][nin // D][2]; parts = output counts of
        successive calls (history carried)"""
        x = np.ascontiguousarray(x, np.int16); taps = np.ascontiguousarray(taps, np.int16)
        self.L.m17o_chan_open.restype = C.c_void_p
        h = self.L.m17o_chan_open(M, D, len(taps), _p(taps))
        assert h, "unsupported channeliser geometry"
        nout = x.shape[0] // D
        out = np.zeros((M, nout, 2), np.int16)
        o = 0
        for n in (parts or [nout]):
            y = np.zeros((M, n, 2), np.int16)
            self.L.m17o_chan_run(C.c_void_p(h), _p(np.ascontiguousarray(x[o * D:(o + n) * D])), C.c_long(n), _p(y), C.c_long(n))
            out[:, o:o + n] = y
            o += n
        self.L.m17o_chan_free(C.c_void_p(h))
        return out

    def demap_symbol(self, sym, mag):
        """m17_dsp_demap_symbol for arrays of symbols / normalisers -> [n][2]"""
        sym = np.ascontiguousarray(sym, np.float32); mag = np.ascontiguousarray(mag, np.float32)
        out = np.zeros((len(sym), 2), np.float32)
        self.L.m17o_demap_symbol.argtypes = [C.c_float, C.c_float, C.c_void_p]
        for i in range(len(sym)):
            self.L.m17o_demap_symbol(float(sym[i]), float(mag[i]), out[i].ctypes.data_as(C.c_void_p))
        return out

    def decimating_filter(self, x, coffs, stride, length):
        """m17_dsp_decimating_filter over one row x (>= length + len(coffs) - 1 samples)"""
        x = np.ascontiguousarray(x, np.float32); cf = np.ascontiguousarray(coffs, np.float32)
        out = np.zeros((length + stride - 1) // stride, np.float32)
        n = self.L.m17o_decimating_filter(_p(x), _p(out), _p(cf), stride, len(cf), length)
        assert n == len(out)
        return out

    def hard24(self, soft24):
        a = np.ascontiguousarray(soft24, np.float32)
        return self.L.m17o_hard24(_p(a))

    def sync_check(self, v8):
        a = np.ascontiguousarray(v8, np.float32)
        t, v, var = C.c_int(), C.c_int(), C.c_float()
        self.L.m17o_sync_check(_p(a), C.byref(t), C.byref(v), C.byref(var))
        return t.value, v.value, np.float32(var.value)

    def rrc(self, rolloff, ntaps, sps):
        out = np.zeros(ntaps, np.float32)
        self.L.m17o_rrc_design(_p(out), rolloff, ntaps, sps)
        return out

    def set_gain(self, taps, gain, stride, ntaps):
        a = np.array(taps, np.float32)
        self.L.m17o_set_gain(_p(a), gain, stride, ntaps)
        return a

    def sync_taps(self):
        mf = np.zeros((40, 31), np.float32)
        md = np.zeros((40, 31), np.float32)
        self.L.m17o_get_sync_taps(_p(mf), _p(md))
        return mf, md

    def prbs9(self):
        out = np.zeros(511, np.uint8)
        self.L.m17o_prbs9_seq(_p(out))
        return out

    def encode_call(self, s):
        return self.L.m17o_encode_call(s.encode())

    def decode_call(self, w):
        b = C.create_string_buffer(16)
        self.L.m17o_decode_call(int(w), b)
        return b.value.decode()

    def build_lsf(self, dst, src, typeword, meta=bytes(14)):
        m = np.frombuffer(bytes(meta), np.uint8)
        out = np.zeros(30, np.uint8)
        self.L.m17o_build_lsf(int(dst), int(src), int(typeword), _p(m), _p(out))
        return out

    # ---- equaliser
    def eq_run(self, pairs, train=None):
        """pairs [n][2] float32; train [n] (known) or None (decision-directed). returns y[n]."""
        st = np.zeros(64, np.float32)
        self.L.m17o_eq_open(_p(st))
        pairs = np.ascontiguousarray(pairs, np.float32)
        y = np.zeros(len(pairs), np.float32)
        for i in range(len(pairs)):
            if train is not None:
                y[i] = self.L.m17o_eq_train_known(_p(st), _p(pairs[i]), float(train[i]))
            else:
                y[i] = self.L.m17o_eq_train_unknown(_p(st), _p(pairs[i]))
        return y

    # ---- TX
    def tx_stream_over(self, lsf30, payloads, lead=1, npre=2, tail=2, os_=10, want_freq=False):
        """One over: lead x carrier, npre x preamble, LSF, F stream frames, EOT, tail x carrier.
        Returns (iq int16 [n][2], dibits uint8 [(npre+1+F+1)*192], freq or None)."""
        L = self.L
        payloads = np.ascontiguousarray(payloads, np.uint8).reshape(-1, 16)
        F = len(payloads)
        t = L.m17o_tx_new(os_)
        lsf30 = np.ascontiguousarray(lsf30, np.uint8)
        script, dib = [], []
        d = np.zeros(192, np.uint8)
        script.append(np.full(192 * lead, 4, np.uint8))
        for _ in range(npre):
            L.m17o_fmt_preamble(_p(d)); script.append(d.copy()); dib.append(d.copy())
        L.m17o_fmt_lsf(_p(lsf30), _p(d)); script.append(d.copy()); dib.append(d.copy())
        # tx state: lich = lsf, counters 0  (m17_tx_routines.cpp:98-99)
        self._tx_set_lich(t, lsf30)
        for f in range(F):
            L.m17o_fmt_stream(_vp(t), _p(payloads[f]), _p(d)); script.append(d.copy()); dib.append(d.copy())
        L.m17o_fmt_eot(_p(d)); script.append(d.copy()); dib.append(d.copy())
        script.append(np.full(192 * tail, 4, np.uint8))
        script = np.concatenate(script)
        iq = np.zeros((len(script) * os_, 2), np.int16)
        freq = np.zeros(len(script) * os_, np.float32) if want_freq else None
        n = L.m17o_mod(_vp(t), _p(script), len(script), _p(iq), _p(freq))
        assert n == len(iq)
        L.m17o_tx_free(_vp(t))
        return iq, np.concatenate(dib), freq

    class _TxStruct(C.Structure):
        _fields_ = [("os", C.c_int), ("taps", _vp), ("s", C.c_float * 31), ("acc", C.c_float), ("lich", C.c_uint8 * 30),
                    ("lich_count", C.c_int), ("fn", C.c_int), ("prbs_idx", C.c_int)]

    def _tx_set_lich(self, t, lsf30, lich_count=0, fn=0):
        s = C.cast(t, C.POINTER(Port._TxStruct)).contents
        for i in range(30):
            s.lich[i] = int(lsf30[i])
        s.lich_count = lich_count
        s.fn = fn

    def fmt_stream_frames(self, lsf30, payloads, lich_count=0, fn=0):
        L = self.L
        t = L.m17o_tx_new(10)
        self._tx_set_lich(t, np.asarray(lsf30, np.uint8), lich_count, fn)
        payloads = np.ascontiguousarray(payloads, np.uint8).reshape(-1, 16)
        out = np.zeros((len(payloads), 192), np.uint8)
        for f in range(len(payloads)):
            L.m17o_fmt_stream(_vp(t), _p(payloads[f]), _p(out[f]))
        L.m17o_tx_free(_vp(t))
        return out

    def fmt_lsf(self, lsf30):
        d = np.zeros(192, np.uint8)
        self.L.m17o_fmt_lsf(_p(np.ascontiguousarray(lsf30, np.uint8)), _p(d))
        return d

    def fmt_packet(self, chunk, eof, nf):
        d = np.zeros(192, np.uint8)
        c = np.zeros(25, np.uint8); c[:len(chunk)] = np.frombuffer(bytes(chunk), np.uint8)
        self.L.m17o_fmt_packet(_p(c), len(chunk), int(eof), int(nf), _p(d))
        return d

    def fmt_bert(self, nframes):
        t = self.L.m17o_tx_new(10)
        out = np.zeros((nframes, 192), np.uint8)
        for f in range(nframes):
            self.L.m17o_fmt_bert(_vp(t), _p(out[f]))
        self.L.m17o_tx_free(_vp(t))
        return out

    def fmt_preamble(self):
        d = np.zeros(192, np.uint8); self.L.m17o_fmt_preamble(_p(d)); return d

    def fmt_eot(self):
        d = np.zeros(192, np.uint8); self.L.m17o_fmt_eot(_p(d)); return d

    def mod(self, script, os_=10, want_freq=False):
        t = self.L.m17o_tx_new(os_)
        script = np.ascontiguousarray(script, np.uint8)
        iq = np.zeros((len(script) * os_, 2), np.int16)
        freq = np.zeros(len(script) * os_, np.float32) if want_freq else None
        self.L.m17o_mod(_vp(t), _p(script), len(script), _p(iq), _p(freq))
        self.L.m17o_tx_free(_vp(t))
        return (iq, freq) if want_freq else iq

    # ---- RX
    def prbs_check(self, bits):
        """m17_prbs9_rx_check over a bit sequence from a fresh checker -> [state, idx, bad, good, eq, dif, bits, errs]."""
        self.L.m17o_rx_new.restype = _vp
        r = self.L.m17o_rx_new()
        b = np.ascontiguousarray(bits, np.uint8)
        for v in b:
            self.L.m17o_prbs9_rx_check(_vp(r), int(v))
        out = np.zeros(8, np.uint32)
        self.L.m17o_rx_get_bert(_vp(r), _p(out))
        self.L.m17o_rx_free(_vp(r))
        return out

    def rx_run(self, x, seam=0, nthreads=None, want_disc=True, want_soft=False, afc=False, bert=False, eq=False):
        x = np.ascontiguousarray(x)
        if seam == 0:
            assert x.dtype == np.int16 and x.ndim == 3 and x.shape[2] == 2 and x.shape[1] % BLOCK == 0
            Cn, T = x.shape[0], x.shape[1] // BLOCK
        else:
            assert x.dtype == np.float32 and x.ndim == 2 and x.shape[1] % DISC_PER_BLOCK == 0
            Cn, T = x.shape[0], x.shape[1] // DISC_PER_BLOCK
        o = _alloc_rx_out(Cn, T, want_disc, want_soft, np.zeros)
        nthreads = nthreads or min(os.cpu_count() or 1, Cn)
        if bert:
            o["bert"] = np.zeros((Cn, 8), np.uint32)
            self.L.m17o_set_bert_out(_p(o["bert"]))
        seam_flags = seam | (16 if afc else 0) | (32 if bert else 0) | (64 if eq else 0)      # 64: equaliser option (checker of m17b_rx_set_equaliser)
        self.L.m17o_rx_run(_p(x), seam_flags, Cn, T, nthreads, _p(o.disc), _p(o.nsym), _p(o.syms), o.symcap, _p(o.frames), o.fcap,
                           _p(o.soft), _p(o.events), o.ecap, _p(o.counts))
        self.L.m17o_set_bert_out(None)
        return o

    def rx_time(self, iq, nthreads):
        iq = np.ascontiguousarray(iq)
        Cn, T = iq.shape[0], iq.shape[1] // BLOCK
        return self.L.m17o_rx_time(_p(iq), Cn, T, nthreads)


class Ref:
    """The unmodified reference objects (only where oracle/_ref/libm17ref.so exists)."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libm17ref.so")

    @staticmethod
    def available():
        return os.path.exists(Ref.PATH)

    def __init__(self, oversample=10):
        L = self.L = C.CDLL(Ref.PATH)
        L.ref_init(oversample)
        self.os = oversample
        L.ref_crc.restype = C.c_uint16
        L.ref_golay_encode.restype = C.c_uint32
        L.ref_hard24.restype = C.c_uint32
        L.ref_encode_call.restype = C.c_uint64
        L.ref_decode_call.argtypes = [C.c_uint64, C.c_char_p]
        L.ref_pack_type.restype = C.c_uint16
        L.ref_eq_train_known.restype = C.c_float
        L.ref_eq_train_known.argtypes = [_vp, C.c_float]
        L.ref_eq_train_unknown.restype = C.c_float
        L.ref_build_lsf.argtypes = [C.c_uint64, C.c_uint64, C.c_uint16, _vp, _vp]
        L.ref_rrc.argtypes = [_vp, C.c_float, C.c_int, C.c_int]
        L.ref_set_gain.argtypes = [_vp, C.c_float, C.c_int, C.c_int]
        L.ref_rx_run.argtypes = [_vp, C.c_int, C.c_long, C.c_long, C.c_int, _vp, _vp, _vp, C.c_long, _vp, C.c_long, _vp,
                                 _vp, C.c_long, _vp]
        L.ref_tx_stream_run.argtypes = [C.c_long, C.c_int, _vp, _vp, C.c_long, C.c_int, C.c_int, C.c_int, _vp, C.c_long, _vp, _vp]
        L.ref_tx_dibits_run.argtypes = [C.c_long, C.c_int, _vp, C.c_long, _vp, C.c_long, _vp]
        L.ref_rx_time.argtypes = [_vp, C.c_long, C.c_long, C.c_int, _vp, _vp]

    def crc(self, b):
        a = np.frombuffer(bytes(b), np.uint8).copy() if len(b) else np.zeros(1, np.uint8)
        return self.L.ref_crc(_p(a), len(b))

    def golay_encode(self, d):
        return self.L.ref_golay_encode(int(d))

    def golay_decode(self, w):
        o = C.c_uint16()
        e = self.L.ref_golay_decode(int(w), C.byref(o))
        return o.value, e

    def golay_errtab(self):
        t = np.zeros(4096, np.uint16)
        self.L.ref_golay_errtab(_p(t))
        return t

    def conv_encode_8(self, data):
        a = np.array(data, np.uint8)
        out = np.zeros(16 * len(a) + 16, np.uint8)
        n = self.L.ref_conv_encode_8(_p(a), _p(out), len(a))
        return out[:n].copy()

    def conv_encode_1(self, bits):
        a = np.array(bits, np.uint8)
        out = np.zeros(2 * len(a) + 16, np.uint8)
        n = self.L.ref_conv_encode_1(_p(a), _p(out), len(a))
        return out[:n].copy()

    def viterbi(self, soft):
        a = np.array(soft, np.float32)
        out = np.zeros(len(a) // 2 + 8, np.uint8)
        n = self.L.ref_viterbi(_p(a), _p(out), len(a))
        return out[:n].copy()

    def punc(self, p, bits):
        a = np.array(bits, np.uint8)
        out = np.zeros(len(a) + 8, np.uint8)
        n = self.L.ref_punc(p, _p(a), _p(out), len(a))
        return out[:n].copy()

    def depunc(self, p, soft, outlen):
        a = np.zeros(outlen + 8, np.float32); a[:len(soft)] = soft
        out = np.zeros(outlen, np.float32)
        self.L.ref_depunc(p, _p(a), _p(out), outlen)
        return out

    def interleave(self, bits):
        a = np.array(bits, np.uint8)
        out = np.zeros(368, np.uint8)
        self.L.ref_interleave(_p(a), _p(out), len(a))
        return out

    def deinterleave(self, soft):
        a = np.array(soft, np.float32)
        out = np.zeros(368, np.float32)
        self.L.ref_deinterleave(_p(a), _p(out), len(a))
        return out

    def derand_bytes(self, b):
        a = np.array(b, np.uint8)
        self.L.ref_derand_bytes(_p(a), len(a))
        return a

    def derand_bits(self, bits):
        a = np.array(bits, np.uint8)
        out = np.zeros_like(a)
        self.L.ref_derand_bits(_p(a), _p(out), len(a))
        return out

    def derand_soft(self, soft):
        a = np.array(soft, np.float32)
        out = np.zeros_like(a)
        self.L.ref_derand_soft(_p(a), _p(out), len(a))
        return out

    def demap_frame(self, sym192):
        a = np.array(sym192, np.float32)
        out = np.zeros(368, np.float32)
        self.L.ref_demap_frame(_p(a), _p(out))
        return out

    def gps_decode(self, lsf30):
        b = np.ascontiguousarray(lsf30, np.uint8)[14:30].copy()
        ll = np.zeros(2, np.float64); o = np.zeros(4, np.int32)
        self.L.ref_gps_decode(_p(b), _p(ll), _p(o))
        return (float(ll[0]), float(ll[1]), int(o[0]), int(o[1]), int(o[2]), int(o[3]))

    def demap_symbol(self, sym, mag):
        sym = np.ascontiguousarray(sym, np.float32); mag = np.ascontiguousarray(mag, np.float32)
        out = np.zeros((len(sym), 2), np.float32)
        self.L.ref_demap_symbol.argtypes = [C.c_float, C.c_float, C.c_void_p]
        for i in range(len(sym)):
            self.L.ref_demap_symbol(float(sym[i]), float(mag[i]), out[i].ctypes.data_as(C.c_void_p))
        return out

    def decimating_filter(self, x, coffs, stride, length):
        x = np.array(x, np.float32); cf = np.array(coffs, np.float32)
        out = np.zeros((length + stride - 1) // stride, np.float32)
        n = self.L.ref_decimating_filter(_p(x), _p(out), _p(cf), stride, len(cf), length)
        assert n == len(out)
        return out

    def hard24(self, soft24):
        a = np.array(soft24, np.float32)
        return self.L.ref_hard24(_p(a))

    def sync_check(self, v8):
        a = np.array(v8, np.float32)
        t, v, var = C.c_int(), C.c_int(), C.c_float()
        self.L.ref_sync_check(_p(a), C.byref(t), C.byref(v), C.byref(var))
        return t.value, v.value, np.float32(var.value)

    def rrc(self, rolloff, ntaps, sps):
        out = np.zeros(ntaps, np.float32)
        self.L.ref_rrc(_p(out), rolloff, ntaps, sps)
        return out

    def set_gain(self, taps, gain, stride, ntaps):
        a = np.array(taps, np.float32)
        self.L.ref_set_gain(_p(a), gain, stride, ntaps)
        return a

    def prbs9(self, n=511):
        out = np.zeros(n, np.uint8)
        self.L.ref_prbs9_reset()
        self.L.ref_prbs9_load(_p(out), n)
        return out

    def encode_call(self, s):
        return self.L.ref_encode_call(s.encode())

    def decode_call(self, w):
        b = C.create_string_buffer(16)
        self.L.ref_decode_call(int(w), b)
        return b.value.decode()

    def build_lsf(self, dst, src, typeword, meta=bytes(14)):
        m = np.frombuffer(bytes(meta), np.uint8).copy()
        out = np.zeros(30, np.uint8)
        self.L.ref_build_lsf(int(dst), int(src), int(typeword), _p(m), _p(out))
        return out

    def fmt_lsf(self, lsf30):
        d = np.zeros(192, np.uint8)
        self.L.ref_fmt_lsf_safe(_p(np.array(lsf30, np.uint8)), _p(d))
        return d

    def fmt_stream_frames(self, lsf30, payloads, lich_count=0, fn=0):
        self.L.ref_set_tx_state(_p(np.array(lsf30, np.uint8)), lich_count, fn)
        payloads = np.array(payloads, np.uint8).reshape(-1, 16)
        out = np.zeros((len(payloads), 192), np.uint8)
        for f in range(len(payloads)):
            self.L.ref_fmt_stream(_p(payloads[f]), _p(out[f]))
        return out

    def fmt_packet(self, chunk, eof, nf):
        d = np.zeros(192, np.uint8)
        c = np.zeros(25, np.uint8); c[:len(chunk)] = np.frombuffer(bytes(chunk), np.uint8)
        self.L.ref_fmt_packet_safe(_p(c), len(chunk), int(eof), int(nf), _p(d))
        return d

    def fmt_bert(self, nframes):
        self.L.ref_prbs9_reset()
        out = np.zeros((nframes, 192), np.uint8)
        for f in range(nframes):
            self.L.ref_fmt_bert_safe(_p(out[f]))
        return out

    def fmt_preamble(self):
        d = np.zeros(192, np.uint8); self.L.ref_fmt_preamble(_p(d)); return d

    def fmt_eot(self):
        d = np.zeros(192, np.uint8); self.L.ref_fmt_eot(_p(d)); return d

    def eq_run(self, pairs, train=None):
        """NOTE: equaliser state is process-global in the reference; call once per process or accept carry-over."""
        self.L.ref_eq_open()
        pairs = np.ascontiguousarray(pairs, np.float32)
        y = np.zeros(len(pairs), np.float32)
        for i in range(len(pairs)):
            if train is not None:
                y[i] = self.L.ref_eq_train_known(_p(pairs[i]), float(train[i]))
            else:
                y[i] = self.L.ref_eq_train_unknown(_p(pairs[i]))
        return y

    def tx_stream_run(self, lsf, payloads, lead=1, npre=2, tail=2, nproc=None):
        """lsf [C][30], payloads [C][F][16] -> iq int16 [C][n][2], dibits [C][(npre+1+F+1)*192]"""
        lsf = np.ascontiguousarray(lsf, np.uint8)
        payloads = np.ascontiguousarray(payloads, np.uint8)
        Cn, F = payloads.shape[0], payloads.shape[1]
        nsym = (lead + npre + 1 + F + 1 + tail) * 192
        cap = nsym * self.os
        iq = shared_array((Cn, cap, 2), np.int16)
        nout = shared_array((Cn,), np.int64)
        dib = shared_array((Cn, (npre + 1 + F + 1) * 192), np.uint8)
        fails = self.L.ref_tx_stream_run(Cn, nproc or os.cpu_count(), _p(lsf), _p(payloads), F, lead, npre, tail, _p(iq), cap,
                                         _p(nout), _p(dib))
        assert fails == 0
        # the modulator flushes only whole 1920-sample blocks (m17_modulate.cpp:30-33)
        n = int(nout.min())
        assert (nout == n).all()
        return np.array(iq[:, :n]), np.array(dib)

    def tx_dibits_run(self, script, nproc=None):
        script = np.ascontiguousarray(script, np.uint8)
        Cn, nsym = script.shape
        cap = nsym * self.os
        iq = shared_array((Cn, cap, 2), np.int16)
        nout = shared_array((Cn,), np.int64)
        fails = self.L.ref_tx_dibits_run(Cn, nproc or os.cpu_count(), _p(script), nsym, _p(iq), cap, _p(nout))
        assert fails == 0
        n = int(nout.min())
        return np.array(iq[:, :n])

    def rx_run(self, x, seam=0, nproc=None, want_disc=True, want_soft=False, afc=False):
        x = np.ascontiguousarray(x)
        self.L.ref_set_afc(1 if afc else 0)           # inherited by the per-channel child processes
        if seam == 0:
            Cn, T = x.shape[0], x.shape[1] // BLOCK
        else:
            Cn, T = x.shape[0], x.shape[1] // DISC_PER_BLOCK
        o = _alloc_rx_out(Cn, T, want_disc, want_soft, shared_array)
        fails = self.L.ref_rx_run(_p(x), seam, Cn, T, nproc or os.cpu_count(), _p(o.disc), _p(o.nsym), _p(o.syms), o.symcap,
                                  _p(o.frames), o.fcap, _p(o.soft), _p(o.events), o.ecap, _p(o.counts))
        self.L.ref_set_afc(0)
        assert fails == 0, "a reference child process crashed"
        return o

    def rx_time(self, iq, nproc):
        iq = np.ascontiguousarray(iq)
        Cn, T = iq.shape[0], iq.shape[1] // BLOCK
        secs = shared_array((nproc,), np.float64)
        nfr = shared_array((nproc,), np.int64)
        self.L.ref_rx_time(_p(iq), Cn, T, nproc, _p(secs), _p(nfr))
        return float(secs.max())


class RefRadio:
    """The reference's radio.cpp (Pluto /8 decimator) linked with the hot-path objects and SDR driver stubs."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libm17ref_radio.so")

    @staticmethod
    def available():
        return os.path.exists(RefRadio.PATH)

    def __init__(self):
        self.L = C.CDLL(RefRadio.PATH)
        self.L.ref_init(10)                       # the reference's init chain (CRC table etc., main.cpp:108-126)

    def net_rx_data(self, frame_id, lsf30, fn, pld):
        out = np.zeros(54, np.uint8)
        n = self.L.ref_net_rx_data(int(frame_id), _p(np.ascontiguousarray(lsf30, np.uint8)), int(fn), _p(np.ascontiguousarray(pld, np.uint8)), _p(out))
        assert n == 54, n
        return out

    def net_parse(self, b54):
        posted = np.zeros(54, np.uint8)
        ok = self.L.ref_net_parse(_p(np.ascontiguousarray(b54, np.uint8)), _p(posted))
        return bool(ok), posted

    def lich_from_net(self, b54):
        out = np.zeros(30, np.uint8)
        self.L.ref_lich_from_net(_p(np.ascontiguousarray(b54, np.uint8)), _p(out))
        return out

    def pluto_run(self, x, nproc=None):
        """x: int16 [C][nblk*8*1920][2] -> radio_receive_samples() output int16 [C][nblk*1920][2]."""
        x = np.ascontiguousarray(x, np.int16)
        Cn, nblk = x.shape[0], x.shape[1] // (8 * BLOCK)
        out = shared_array((Cn, nblk * BLOCK, 2), np.int16)
        fails = self.L.ref_pluto_run(_p(x), C.c_long(Cn), C.c_long(nblk), nproc or os.cpu_count(), _p(out))
        assert fails == 0
        return np.array(out)


# ---------------------------------------------------------------------------- synthetic channels
def lsf_for(port, src="G4GUO    ", dst=0xFFFFFFFFFFFF, typeword=0x0005):
    return port.build_lsf(dst, port.encode_call(src), typeword)


def add_iq_noise(iq, ebn0_db, rng):
    """White noise on int16 IQ; Eb/N0(IQ) per SURVEY 8d: sigma^2 = 2.5 A^2 / 10^(EbN0/10) per component."""
    if ebn0_db is None:
        return iq
    A = float(0x3FFF)
    sigma = np.sqrt(2.5 * A * A / 10 ** (ebn0_db / 10))
    x = iq.astype(np.float64) + rng.normal(0, sigma, iq.shape)
    x = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    z = (x[..., 0] == 0) & (x[..., 1] == 0)        # dsp_limit divides by |z|
    x[z, 0] = 1
    return x


def rotate_iq(iq, f0_hz, fs=48000.0):
    n = np.arange(iq.shape[-2])
    z = (iq[..., 0].astype(np.float64) + 1j * iq[..., 1]) * np.exp(2j * np.pi * f0_hz * n / fs)
    out = np.stack([np.clip(np.rint(z.real), -32768, 32767), np.clip(np.rint(z.imag), -32768, 32767)], -1).astype(np.int16)
    z0 = (out[..., 0] == 0) & (out[..., 1] == 0)
    out[z0, 0] = 1
    return out


def delay_iq(iq, d, total):
    """prepend d samples of unmodulated carrier, then pad/truncate to `total` samples (multiple of 1920)."""
    out = np.zeros((total, 2), np.int16)
    out[:, 0] = 0x3FFF
    n = min(len(iq), total - d)
    out[d:d + n] = iq[:n]
    return out


def compare_rx(a, b, sym_tol=0.0):
    """Assert two rx_run outputs agree: counts, nsym, records, events exact; symbols within sym_tol rel RMS."""
    assert (a.counts == b.counts).all(), (a.counts[:4], b.counts[:4])
    assert (a.nsym == b.nsym).all()
    for c in range(a.counts.shape[0]):
        ns, nf, ne = int(a.counts[c, 1]), int(min(a.counts[c, 2], a.fcap)), int(min(a.counts[c, 3], a.ecap))
        sa, sb = a.syms[c, :ns], b.syms[c, :ns]
        if sym_tol == 0.0:
            assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32)), f"channel {c}: symbol stream differs"
        else:
            den = np.sqrt(np.mean(sb.astype(np.float64) ** 2)) + 1e-30
            assert np.sqrt(np.mean((sa.astype(np.float64) - sb) ** 2)) / den <= sym_tol
        fa, fb = a.frames[c, :nf], b.frames[c, :nf]
        for name in REC_DTYPE.names:
            if name in ("variance", "cor"):
                assert np.array_equal(fa[name].view(np.uint32), fb[name].view(np.uint32)), (c, name)
            elif name != "rsvd":
                assert np.array_equal(fa[name], fb[name]), (c, name, fa[name][:8], fb[name][:8])
        assert np.array_equal(a.events[c, :ne], b.events[c, :ne]), (c, "events")
