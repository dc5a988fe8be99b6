"""ctypes loader for oracle/_build/liboracle.so (the C restatement in oracle.c).  TEST INFRASTRUCTURE."""
import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_here, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def quantize_weights(w):
    w = np.ascontiguousarray(w, dtype=np.float32)
    N, K = w.shape
    packed = np.empty((N, K // 2), dtype=np.uint8)
    scales = np.empty(N, dtype=np.float32)
    zps = np.empty(N, dtype=np.float32)
    load().oracle_quantize_weights(_p(w), ctypes.c_int64(N), ctypes.c_int64(K), _p(packed), _p(scales), _p(zps))
    return packed, scales, zps


def dequantize_weights(packed, scales, zps):
    N, Kh = packed.shape
    out = np.empty((N, Kh * 2), dtype=np.float32)
    load().oracle_dequantize_weights(_p(np.ascontiguousarray(packed)), _p(np.ascontiguousarray(scales, dtype=np.float32)),
                                     _p(np.ascontiguousarray(zps, dtype=np.float32)), ctypes.c_int64(N),
                                     ctypes.c_int64(Kh * 2), _p(out))
    return out


class Linear:
    """Holds the scratch fp32 weight matrix so repeated calls time the algorithm, not malloc."""

    def __init__(self, packed, scales, zps):
        self.packed = np.ascontiguousarray(packed)
        self.scales = np.ascontiguousarray(scales, dtype=np.float32)
        self.zps = np.ascontiguousarray(zps, dtype=np.float32)
        self.N, self.K = packed.shape[0], packed.shape[1] * 2
        self.scratch = np.empty((self.N, self.K), dtype=np.float32)

    def __call__(self, x):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
        y = np.empty((x.shape[0], self.N), dtype=np.float32)
        load().oracle_reference_quantized_linear(_p(x), ctypes.c_int64(x.shape[0]), _p(self.packed), _p(self.scales),
                                                 _p(self.zps), ctypes.c_int64(self.N), ctypes.c_int64(self.K),
                                                 _p(self.scratch), _p(y))
        return y


def max_threads():
    return int(load().oracle_max_threads())
