"""ctypes binding of libnrldpc (include/nrldpc.h) and its in-tree build.

The library is built with nvcc for sm_100a only and kept next to this file (neoradium_b200/libnrldpc.so) so that it
travels with the repository snapshot.  There is no CPU fallback anywhere: if the library is missing it is built, if
it cannot be built or loaded an ImportError/RuntimeError is raised, and every compute entry point fails with
NRLDPC_ERR_CUDA when no B200 is present.
"""
import ctypes
import os
import subprocess
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_INCLUDE = os.path.join(_HERE, "..", "include")
LIB_PATH = os.path.join(_HERE, "libnrldpc.so")
_LIB_OVERRIDE = os.environ.get("NRLDPC_LIB")   # A/B measurements: load another build of the same library
_BUILD_DIR = os.path.join(_HERE, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM = 0, 1, 2, 3
CRC_IDS = {"6": 0, "11": 1, "16": 2, "24A": 3, "24B": 4, "24C": 5}
F32, F64, F16 = 0, 1, 2
DEC_EARLY_STOP, DEC_ALL_ROWS, DEC_ES_AUTO = 1, 2, 4


def dec_flags(early_stop, es_from=1):
    """decoder flags word: NRLDPC_DEC_EARLY_STOP | NRLDPC_DEC_ES_FROM(es_from) (include/nrldpc.h); es_from="auto" sets
    NRLDPC_DEC_ES_AUTO: the first tested iteration follows the previous launch on the same handle"""
    if not early_stop:
        return 0
    if isinstance(es_from, str):
        if es_from != "auto":
            raise ValueError("earlyStopFrom must be an iteration number or 'auto'")
        return DEC_EARLY_STOP | DEC_ES_AUTO
    return DEC_EARLY_STOP | ((max(0, min(255, int(es_from))) & 0xff) << 8)


class NrldpcError(RuntimeError):
    pass


def _sources():
    return sorted(f for f in os.listdir(_CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_INCLUDE, "nrldpc.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into neoradium_b200/libnrldpc.so (object files compiled in parallel)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(_BUILD_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(os.path.join(_CSRC, f)) for f in os.listdir(_CSRC)
                   if f.endswith((".cuh", ".h")))
    hdr_time = max(hdr_time, os.path.getmtime(os.path.join(_INCLUDE, "nrldpc.h")))

    def compile_one(src):
        obj = os.path.join(_BUILD_DIR, src[:-3] + ".o")
        sp = os.path.join(_CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(sp)
                and os.path.getmtime(obj) > hdr_time):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NRLDPC_NVCC_EXTRA", "").split() + ["-c", sp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise NrldpcError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    r = subprocess.run([nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise NrldpcError("link failed:\n" + r.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64


class TbConfig(ctypes.Structure):
    """struct nrldpc_tb_config"""
    _fields_ = [("bg", ctypes.c_int32), ("zc", ctypes.c_int32), ("K", ctypes.c_int32), ("F", ctypes.c_int32),
                ("C", ctypes.c_int32), ("qm", ctypes.c_int32), ("nl", ctypes.c_int32), ("ncb", ctypes.c_int32),
                ("rv", ctypes.c_int32), ("reserved", ctypes.c_int32), ("G", ctypes.c_int64)]


_cfgp = ctypes.POINTER(TbConfig)


class TbGroup(ctypes.Structure):
    """struct nrldpc_tb_group"""
    _fields_ = [("cfg", TbConfig), ("in_dtype", ctypes.c_int32), ("reserved", ctypes.c_int32), ("llr", ctypes.c_void_p),
                ("num_tb", ctypes.c_int64), ("llr_len", ctypes.c_int64), ("llr_stride", ctypes.c_int64),
                ("soft_buffer", ctypes.c_void_p), ("tb_bits", ctypes.c_void_p), ("tb_bits_stride", ctypes.c_int64),
                ("cb_crc_ok", ctypes.c_void_p), ("tb_crc_ok", ctypes.c_void_p), ("iters", ctypes.c_void_p)]

# every symbol include/nrldpc.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "nrldpc_version": (_i32, []),
    "nrldpc_last_error": (ctypes.c_char_p, []),
    "nrldpc_base_graph": (_i32, [_i32, _i32, _i32, _vp]),
    "nrldpc_lifting_set_index": (_i32, [_i32]),
    "nrldpc_graph_info": (_i32, [_i32, _vp, _vp, _vp, _vp]),
    "nrldpc_create": (_i32, [_i32, ctypes.POINTER(_vp)]),
    "nrldpc_destroy": (_i32, [_vp]),
    "nrldpc_managed_supported": (_i32, [_vp]),
    "nrldpc_managed_alloc": (_i32, [_vp, ctypes.c_uint64, _i32, ctypes.POINTER(_vp), _vp]),
    "nrldpc_managed_free": (_i32, [_vp, _vp]),
    "nrldpc_managed_clear": (_i32, [_vp, _vp, ctypes.c_uint64, _vp]),
    "nrldpc_managed_prefetch": (_i32, [_vp, _vp, ctypes.c_uint64, _i32, _vp]),
    "nrldpc_crc": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    "nrldpc_crc_attach": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "nrldpc_crc_check": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "nrldpc_segment": (_i32, [_vp, _cfgp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "nrldpc_encode": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _i32, _vp]),
    "nrldpc_rate_match": (_i32, [_vp, _cfgp, _vp, _i64, _vp, _i64, _vp]),
    "nrldpc_parity_check": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "nrldpc_parity_check_rows": (_i32, [_vp, _i32, _i32, _vp, _i64, _i32, _vp, _vp]),
    "nrldpc_rate_recover": (_i32, [_vp, _cfgp, _i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "nrldpc_decode": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp,
                             _vp]),
    "nrldpc_decode2": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _i32, _i32, ctypes.c_double, _i32, _i32, _vp,
                              _vp, _vp, _vp]),
    "nrldpc_decode2_offset": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _i32, _i32, ctypes.c_double, ctypes.c_double, _i32,
                                     _i32, _vp, _vp, _vp, _vp]),
    "nrldpc_decode_tb": (_i32, [_vp, _cfgp, _i32, _i32, _vp, _i64, _i64, _i64, _vp, _i32, _i32, _vp, _i64, _vp,
                                _vp, _vp, _vp]),
    "nrldpc_decode_tb_symbols": (_i32, [_vp, _cfgp, _vp, _i64, _i64, _i64, ctypes.c_double, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "nrldpc_decode_tb_groups": (_i32, [_vp, ctypes.POINTER(TbGroup), _i32, _i32, _i32, _i32, _vp]),
    "nrldpc_check_crc_and_merge": (_i32, [_vp, _cfgp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "nrldpc_accumulate_counters": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "nrldpc_accumulate_counters_ref": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "nrldpc_random_bits": (_i32, [_vp, ctypes.c_uint64, ctypes.c_uint64, _vp, _i64, _vp]),
    "nrldpc_modulate": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "nrldpc_demap_maxlog": (_i32, [_vp, _i32, _i32, _vp, _i64, ctypes.c_double, _i32, _vp, _vp]),
    "nrldpc_awgn_llr": (_i32, [_vp, _i32, _vp, _i64, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, _vp, _vp]),
    "nrldpc_gold_sequence": (_i32, [_vp, ctypes.c_uint32, _i64, _vp, _vp]),
    "nrldpc_scramble_bits": (_i32, [_vp, ctypes.c_uint32, _vp, _i64, _vp, _vp]),
    "nrldpc_scramble_llrs": (_i32, [_vp, ctypes.c_uint32, _i32, _vp, _i64, _vp, _vp]),
}


def lib():
    """Load (building first if needed) and return the ctypes library with typed signatures."""
    global _lib
    with _lock:
        if _lib is None:
            if _LIB_OVERRIDE is None and _stale():
                build()
            try:
                L = ctypes.CDLL(_LIB_OVERRIDE or LIB_PATH)
            except OSError as e:   # no silent fallback
                raise ImportError("libnrldpc.so could not be loaded (%s); neoradium_b200 has no CPU fallback" % e)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().nrldpc_last_error().decode("utf-8", "replace")
        if rc == ERR_ARG:
            raise ValueError(msg)
        raise NrldpcError("libnrldpc error %d: %s" % (rc, msg))


_handles = {}


def new_handle(device_index):
    """A private library handle (own scratch / temporaries) for work issued concurrently on another stream."""
    p = _vp()
    check(lib().nrldpc_create(int(device_index), ctypes.byref(p)))
    return p


def handle(device_index):
    """One library handle per device (the Python layer issues work on torch's current stream)."""
    h = _handles.get(device_index)
    if h is None:
        p = _vp()
        check(lib().nrldpc_create(int(device_index), ctypes.byref(p)))
        h = p
        _handles[device_index] = h
    return h
