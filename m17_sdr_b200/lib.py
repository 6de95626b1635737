"""ctypes binding of the C ABI in include/m17b200.h.  No CPU fallback: a missing library or a failing
CUDA call raises."""
import ctypes as C
import os

from .build import LIB

_vp, _i64, _i32, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64


class M17Error(RuntimeError):
    pass


class RxView(C.Structure):
    _fields_ = [("nchan", _i64), ("nblocks", _i64), ("d_frames", _vp), ("frame_cap", _i64), ("d_nframes", _vp),
                ("d_syms", _vp), ("sym_pitch", _i64), ("sym_carry", _i64), ("d_nsym", _vp), ("d_sym_base", _vp),
                ("d_disc", _vp), ("d_mean", _vp), ("d_events", _vp), ("event_cap", _i64), ("d_nevents", _vp), ("d_stats", _vp)]


_SIGS = {
    "m17b_version": ([], _i32),
    "m17b_ctx_create": ([_i32, C.POINTER(_vp)], _i32),
    "m17b_ctx_destroy": ([_vp], _i32),
    "m17b_build_rrc_filter": ([_vp, C.c_float, _i32, _i32], _i32),
    "m17b_set_filter_gain": ([_vp, C.c_float, _i32, _i32], _i32),
    "m17b_get_sync_taps": ([_vp, _vp, _vp], _i32),
    "m17b_crc_array_encode": ([_vp, _vp, _i64, _i32, _i64, _vp, _vp], _i32),
    "m17b_golay_encode": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_golay_decode": ([_vp, _vp, _i64, _vp, _vp, _vp], _i32),
    "m17b_conv_encode_8": ([_vp, _vp, _i32, _i64, _vp, _vp], _i32),
    "m17b_conv_encode_1": ([_vp, _vp, _i32, _i64, _vp, _vp], _i32),
    "m17b_viterbi_decode": ([_vp, _vp, _i32, _i64, _vp, _vp], _i32),
    "m17b_punc": ([_vp, _i32, _vp, _i32, _i64, _vp, C.POINTER(_i32), _vp], _i32),
    "m17b_de_punc": ([_vp, _i32, _vp, _i32, _i32, _i64, _vp, _vp], _i32),
    "m17b_interleave": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_de_interleave": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_de_correlate_8": ([_vp, _vp, _i32, _i64, _vp], _i32),
    "m17b_de_correlate_1_u8": ([_vp, _vp, _vp, _i32, _i64, _vp], _i32),
    "m17b_de_correlate_1_f32": ([_vp, _vp, _vp, _i32, _i64, _vp], _i32),
    "m17b_demap_frame": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_sync_check": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp], _i32),
    "m17b_prbs9_tx_load": ([_vp, _vp, _i32, _i64, _vp, _vp], _i32),
    "m17b_rx_parse_frames": ([_vp, _vp, _vp, _i64, _vp, _vp, _vp], _i32),
    "m17b_viterbi_punctured": ([_vp, _i32, _vp, _i64, _vp, _vp], _i32),
    "m17b_rx_create": ([_vp, _i64, _i64, C.POINTER(_vp)], _i32),
    "m17b_rx_destroy": ([_vp], _i32),
    "m17b_rx_reset": ([_vp, _vp], _i32),
    "m17b_rx_framer_reset": ([_vp, _vp], _i32),
    "m17b_rx_set_afc": ([_vp, _i32, _vp], _i32),
    "m17b_rx_set_equaliser": ([_vp, _i32, _vp], _i32),
    "m17b_rx_get_overflow": ([_vp, C.POINTER(_i32)], _i32),
    "m17b_rx_set_bert": ([_vp, _i32], _i32),
    "m17b_rx_get_bert": ([_vp, _vp, _vp], _i32),
    "m17b_dsp_rx": ([_vp, _vp, _i64, _vp], _i32),
    "m17b_rx_baseband": ([_vp, _vp, _i64, _vp], _i32),
    "m17b_rx_symbols": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_rx_get_view": ([_vp, C.POINTER(RxView)], _i32),
    "m17b_rx_frame_cap": ([_vp], _i64),
    "m17b_dsp_rx_host": ([_vp, _vp, _i64, _vp, _vp, _vp], _i32),
    "m17b_rx_last_launches": ([_vp], _i32),
    "m17b_rx_debug_sync": ([_vp, _vp, _vp], _i32),
    "m17b_rx_set_slice_blocks": ([_vp, _i32], _i32),
    "m17b_rx_set_chan_groups": ([_vp, _i32], _i32),
    "m17b_rx_set_timing": ([_vp, _i32], _i32),
    "m17b_rx_stage_ms": ([_vp, _i64, _vp], _i32),
    "m17b_selftest_frontend": ([_vp, _u64, _u64, C.POINTER(_u64), _vp, _i32, _vp], _i32),
    "m17b_selftest_limiter": ([_vp, _u64, _u64, C.POINTER(_u64), _vp, _i32, _vp], _i32),
    "m17b_selftest_tx_wrap": ([_vp, _u64, _u64, C.POINTER(_u64), _vp, _i32, _vp], _i32),
    "m17b_selftest_demap": ([_vp, _u64, _u64, C.POINTER(_u64), _vp, _i32, _vp], _i32),
    "m17b_tx_debug_scan": ([_vp, _vp], _i32),
    "m17b_tx_create": ([_vp, _i64, _i32, C.POINTER(_vp)], _i32),
    "m17b_tx_destroy": ([_vp], _i32),
    "m17b_tx_reset": ([_vp, _vp], _i32),
    "m17b_tx_set_lsf": ([_vp, _vp, _vp], _i32),
    "m17b_fmt_preamble": ([_vp], _i32),
    "m17b_fmt_eot": ([_vp], _i32),
    "m17b_fmt_link_setup_frame": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_fmt_stream_frames": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_fmt_packet_frames": ([_vp, _vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_send_packet_frames": ([_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp], _i32),
    "m17b_fmt_bert_frames": ([_vp, _i64, _vp, _vp], _i32),
    "m17b_mod_dibits": ([_vp, _vp, _i64, _vp, _vp, _vp], _i32),
    "m17b_build_lpf_filter": ([_vp, C.c_float, _i32], _i32),
    "m17b_float_to_short": ([_vp, _vp, _i32], _i32),
    "m17b_dec_create": ([_vp, _i64, C.POINTER(_vp)], _i32),
    "m17b_dec_destroy": ([_vp], _i32),
    "m17b_dec_reset": ([_vp, _vp], _i32),
    "m17b_dec_get_taps": ([_vp, _vp], _i32),
    "m17b_dec_run": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_chan_create": ([_vp, _i64, _i32, C.POINTER(_vp)], _i32),
    "m17b_chan_destroy": ([_vp], _i32),
    "m17b_chan_reset": ([_vp, _vp], _i32),
    "m17b_chan_get_taps": ([_vp, _vp, C.POINTER(_i32)], _i32),
    "m17b_chan_run": ([_vp, _vp, _i64, _vp, _i64, _vp], _i32),
    "m17b_net_pack": ([_vp, _vp, _vp, _i64, _i32, _u64, _vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_net_parse": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "m17b_rx_net_frames": ([_vp, _vp, _i32, _u64, _vp, _vp, _vp], _i32),
    "m17b_demap_symbols": ([_vp, _vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_dsp_decimating_filter": ([_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i64, _vp, C.POINTER(_i32), _vp], _i32),
    "m17b_prbs9_rx_check": ([_vp, _vp, _i32, _i64, _vp, _vp], _i32),
    "m17b_rx_reassemble_packets": ([_vp, _vp, _i64, _vp, _i32, _vp, _vp], _i32),
    "m17b_gps_decode": ([_vp, _vp, _i64, _i64, _vp, _vp], _i32),
    "m17b_eq_restart": ([_vp, _vp], _i32),
    "m17b_eq_create": ([_vp, _i64, C.POINTER(_vp)], _i32),
    "m17b_eq_destroy": ([_vp], _i32),
    "m17b_eq_reset": ([_vp, _vp], _i32),
    "m17b_eq_train": ([_vp, _vp, _vp, _i64, _vp, _vp], _i32),
    "m17b_synth_channel": ([_vp, _vp, _i64, _i64, _vp, _vp, _u64, _vp], _i32),
}
EXPORTS = sorted(_SIGS) + ["m17b_error_string", "m17b_last_cuda_error"]

_lib = None


def load():
    """dlopen libm17b200.so (built in-tree by m17_sdr_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise M17Error(f"{LIB} is missing: run `python -m m17_sdr_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB)
    for name, (args, res) in _SIGS.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = args, res
    L.m17b_error_string.argtypes, L.m17b_error_string.restype = [_i32], C.c_char_p
    L.m17b_last_cuda_error.argtypes, L.m17b_last_cuda_error.restype = [], C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc != 0:
        L = load()
        msg = L.m17b_error_string(rc).decode()
        if rc == -2:
            msg += ": " + L.m17b_last_cuda_error().decode()
        raise M17Error(f"libm17b200: {msg} (code {rc})")
