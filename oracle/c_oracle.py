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
        lib.oracle_synth_rows2.restype = None
        lib.oracle_synth_rows2.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_float, C.c_int]
        lib.oracle_topk_synth_f64.restype = None
        lib.oracle_topk_synth_f64.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_uint64,
                                              C.c_uint64, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
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


def synth_block(first_row: int, n: int, seed: int = 0, stream: int = 0, norm: float = 1.0,
                mean_shift: int = 0) -> np.ndarray:
    out = np.empty((n, 768), dtype=np.float32)
    load().oracle_synth_rows2(out.ctypes.data, first_row, n, seed, stream, norm, mean_shift)
    return out


def set_threads(n: int) -> None:
    """OpenMP threads of the C oracle (torchrun sets OMP_NUM_THREADS=1 for every rank)."""
    lib = load()
    lib.oracle_set_threads.restype = None
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_set_threads(int(n))


def topk_synth_f64(Q: np.ndarray, k: int, first_row: int, n_rows: int, seed: int = 0, stream: int = 0,
                   norm: float = 1.0, mean_shift: int = 0):
    """float64 ground truth (D [nq,k], I [nq,k] global rows, order (score desc, row asc)) over rows
    [first_row, first_row + n_rows) of a synthetic stream, regenerated block by block on all host threads."""
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    D = np.empty((Q.shape[0], k), dtype=np.float64)
    I = np.empty((Q.shape[0], k), dtype=np.int64)
    load().oracle_topk_synth_f64(Q.ctypes.data, Q.shape[0], k, first_row, n_rows, seed, stream, norm, mean_shift,
                                 D.ctypes.data, I.ctypes.data)
    return D, I
