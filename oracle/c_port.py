"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/liboracle.so (the C restatement)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'liboracle.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(['make', '-s', '-C', _HERE])
        L = ctypes.CDLL(_PATH)
        L.oracle_rgcn_block_layer.restype = ctypes.c_int
        L.oracle_rgcn_block_layer.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_int] * 4
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def rgcn_block_layer(H, W, Wloop, src, dst, etype, norm, relu, num_bases):
    H = np.ascontiguousarray(H, np.float32); W = np.ascontiguousarray(W, np.float32)
    Wl = None if Wloop is None else np.ascontiguousarray(Wloop, np.float32)
    src, dst, etype = (np.ascontiguousarray(a, np.int64) for a in (src, dst, etype))
    norm = np.ascontiguousarray(norm, np.float32)
    N, d_in = H.shape
    d_out = num_bases * (W.shape[1] // (num_bases * (d_in // num_bases)))
    out = np.empty((N, d_out), np.float32)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().oracle_rgcn_block_layer(p(H), p(W), p(Wl), p(src), p(dst), p(etype), p(norm), p(out), N, len(src),
                                       d_in, d_out, num_bases, int(relu))
    if rc != 0:
        raise MemoryError('oracle_rgcn_block_layer')
    return out
