"""ctypes loader of oracle/liboracle.so (ORACLE — test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = C.CDLL(path)
        lib.oracle_knn_ip_heap.restype = None
        lib.oracle_knn_ip_heap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p]
        lib.oracle_synth_rows.restype = None
        lib.oracle_synth_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_float]
        _LIB = lib
    return _LIB


def knn_ip_heap(x: np.ndarray, xb: np.ndarray, k: int):
    x = np.ascontiguousarray(x, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    D = np.empty((x.shape[0], k), dtype=np.float32)
    I = np.empty((x.shape[0], k), dtype=np.int64)
    load().oracle_knn_ip_heap(x.ctypes.data, xb.ctypes.data, x.shape[1], x.shape[0], xb.shape[0], k,
                              D.ctypes.data, I.ctypes.data)
    return D, I


def synth_block(first_row: int, n: int, seed: int = 0, stream: int = 0, norm: float = 1.0) -> np.ndarray:
    out = np.empty((n, 768), dtype=np.float32)
    load().oracle_synth_rows(out.ctypes.data, first_row, n, seed, stream, norm)
    return out
