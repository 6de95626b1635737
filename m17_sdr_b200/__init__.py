"""m17_sdr_b200 -- batched, B200-native (sm_100a) M17 baseband hot path behind the m17gismo function names.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/m17b200.h), the ctypes loader
(lib.py), the host-side mirror of the reference interface (api.py) and channel sharding across GPUs (dist.py).
There is no CPU implementation in this package.
"""
from .lib import M17Error, EXPORTS, load  # noqa: F401
from .build import build, LIB  # noqa: F401


def __getattr__(name):
    # api needs torch; keep `import m17_sdr_b200` light for the build check
    if name in ("Context", "Rx", "Tx", "Equalizer", "Decimator", "Channelizer", "REC_DTYPE", "EV_DTYPE", "GPS_DTYPE", "records_to_numpy", "packets_of", "STAT_NAMES"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
