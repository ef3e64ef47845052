"""Loader for librt_b200.so.  Fails loudly: no fallback of any kind."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_DEBUG=1: the bounds-checked build of the same sources (csrc/rt_debug.h, `make librt_b200_debug.so`).
# RT_B200_LIBRARY: another build of the same library (csrc/Makefile `variant`), for A/B measurements.
_DEFAULT = "librt_b200_debug.so" if os.environ.get("RT_B200_DEBUG", "0") not in ("", "0") else "librt_b200.so"
LIB_PATH = os.environ.get("RT_B200_LIBRARY") or os.path.join(_HERE, "csrc", _DEFAULT)

_lib = None


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librt_b200: {msg} (rt_status {code})")
        self.code = code


def load():
    """dlopen the in-tree library and bind every prototype of include/rt_b200.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU path.")
        _lib = abi.bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
        if _lib.rt_abi_version() != abi.RT_B200_ABI_VERSION:
            raise RuntimeError("librt_b200.so ABI version mismatch; rebuild")
    return _lib


def check(code):
    if code != abi.RT_OK:
        msg = load().rt_last_error()
        raise RtError(code, msg.decode() if msg else "unknown error")
