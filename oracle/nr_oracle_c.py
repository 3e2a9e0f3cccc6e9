"""ctypes front end of the C oracle (oracle/nr_oracle_c.c) -- test infrastructure, see that file's header."""
import ctypes
import os
import subprocess

import numpy as np

import nr_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libnr_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.nr_oracle_crc.restype = ctypes.c_uint32
        _LIB.nr_oracle_crc.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_uint32, ctypes.c_int]
    return _LIB


def flat_graph(bg, zc, ils, num_rows=None):
    P, n, _ = O.bg_dims(bg)
    if num_rows is not None:
        P = num_rows
    h = O.base_graph(bg, zc, ils)
    deg, col, sh = [], [], []
    for i in range(P):
        cols = np.nonzero(h[i] >= 0)[0]
        deg.append(len(cols))
        col += list(cols)
        sh += list(h[i, cols])
    return P, n, np.asarray(deg, np.int32), np.asarray(col, np.int32), np.asarray(sh, np.int32)


def decode_beliefs(rx, bg, zc, ils, num_iter, dtype=np.float32, num_rows=None):
    """[C, ncols_in*Z] -> beliefs [C, n*Z] in `dtype`, same arithmetic as nr_oracle.decode."""
    dtype = np.dtype(dtype)
    rx = np.ascontiguousarray(rx, dtype=dtype)
    C = rx.shape[0]
    ncols_in = rx.shape[1] // zc
    P, n, deg, col, sh = flat_graph(bg, zc, ils, num_rows)
    out = np.empty((C, n * zc), dtype)
    fn = lib().nr_oracle_decode_f32 if dtype == np.float32 else lib().nr_oracle_decode_f64
    fn.restype = ctypes.c_int
    rc = fn(rx.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(C), ctypes.c_int(ncols_in), ctypes.c_int(P),
            ctypes.c_int(n), ctypes.c_int(zc), deg.ctypes.data_as(ctypes.c_void_p),
            col.ctypes.data_as(ctypes.c_void_p), sh.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(num_iter),
            out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def crc(bits, poly):
    bits = np.ascontiguousarray(bits, dtype=np.int8)
    c = O.crc_len(poly)
    return lib().nr_oracle_crc(bits.ctypes.data_as(ctypes.c_void_p), bits.size, O.CRC_POLYS[poly] & ((1 << c) - 1), c)
