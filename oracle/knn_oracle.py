"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/knn_oracle.c (see its header: k-NN
parity is UNPINNED, torch_cluster is absent).  Also a tiny pure-numpy version of the same
spec used to cross-check the C build on small inputs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libknn_oracle.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libknn_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.egspr_oracle_knn_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _LIB.egspr_oracle_knn_batch.restype = None
    return _LIB


def knn(x, k, threads=None):
    """x [N,3] or [C,N,3] float32 -> nbr int32 [N,k] / [C,N,k], nearest first, ties -> lower id."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    single = x.ndim == 2
    if single:
        x = x[None]
    c, n, _ = x.shape
    assert k <= 64
    nbr = np.empty((c, n, k), dtype=np.int32)
    _lib().egspr_oracle_knn_batch(x.ctypes.data, c, n, k, nbr.ctypes.data,
                                  int(threads or os.cpu_count() or 1))
    return nbr[0] if single else nbr


def knn_numpy(x, k):
    """Same spec in numpy (float64 emulation of the fp32 FMA chain; exact because every
    partial result is rounded to fp32 from an exactly-representable float64 value except in
    astronomically rare double-rounding cases).  Small N only."""
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    d = (x[None, :, :] - x[:, None, :]).astype(np.float32)      # [i, j, :] = x_j - x_i
    dx, dy, dz = (d[..., 0].astype(np.float64), d[..., 1].astype(np.float64), d[..., 2].astype(np.float64))
    t0 = (dx * dx).astype(np.float32)
    t1 = (dy * dy + t0.astype(np.float64)).astype(np.float32)
    d2 = (dz * dz + t1.astype(np.float64)).astype(np.float32)
    order = np.argsort(d2, axis=1, kind="stable")               # ties -> lower j
    out = order[:, :k].astype(np.int32)
    if n < k:
        out = np.concatenate([out, -np.ones((n, k - n), np.int32)], 1)
    return out
