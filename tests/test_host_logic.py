"""CPU: host-side logic of the product package and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import nr_oracle as O
import nr_tables
from neoradium_b200 import ChanCodeBase, LdpcDecoder, LdpcEncoder, _native, params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nrldpc.h")).read()
    declared = set(re.findall(r"\b(nrldpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nrldpc_version() == 100


def test_params_match_oracle():
    for bg in (1, 2):
        for B in list(range(25, 4000, 3)) + list(range(4000, 200000, 1231)) + [3840, 3841, 8448, 8449]:
            p = O.derive_params(bg, B)
            assert params.segmentation_params(bg, B) == (p["C"], p["Zc"], p["iLS"], p["K"])
    for g, c, nl, qm in [(22807, 2, 1, 2), (100, 1, 1, 1), (30000, 7, 2, 6), (123457, 13, 4, 8), (5000, 3, 1, 10)]:
        assert params.rate_matched_cb_lens(g, c, nl, qm) == list(O.rm_cb_lens(g, c, nl, qm))
    for bg in (1, 2):
        n = (66 if bg == 1 else 50) * 96
        for rv in range(4):
            for ncb in (n, n - 500, 3000):
                assert params.k0_start(bg, rv, ncb, n, 96) == O.k0_start(bg, rv, ncb, n, 96)


def test_base_graph_tables_through_c_abi():
    """The product's table (csrc/nr_bg_tables.h, read through nrldpc_base_graph) against the oracle's copy, for every
    base graph and lifting size, including the reference's verbatim 880 entry."""
    lib = _native.lib()
    for bg in (1, 2):
        P, n, k = O.bg_dims(bg)
        rows, cols, sys_cols, edges = (ctypes.c_int() for _ in range(4))
        assert lib.nrldpc_graph_info(bg, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(sys_cols),
                                     ctypes.byref(edges)) == 0
        assert (rows.value, cols.value, sys_cols.value, edges.value) == (P, n, k, len(nr_tables.EDGES[bg]))
        for ils, zs in enumerate(nr_tables.LIFTING_SETS):
            for z in zs:
                assert lib.nrldpc_lifting_set_index(z) == ils
                out = np.empty((P, n), np.int16)
                assert lib.nrldpc_base_graph(bg, -1, z, out.ctypes.data) == 0
                assert np.array_equal(out, O.base_graph(bg, z, ils))
    assert lib.nrldpc_lifting_set_index(100) == -1
    # the 880 quirk: BG1 row 6 col 0, iLS 4
    out = np.empty((46, 68), np.int16)
    lib.nrldpc_base_graph(1, -1, 288, out.ctypes.data)
    assert out[6, 0] == 880 % 288


def test_constructor_errors_and_attributes():
    with pytest.raises(ValueError):
        LdpcEncoder(3)
    with pytest.raises(ValueError):
        LdpcDecoder(1, "8PSK")
    with pytest.raises(ValueError):
        LdpcDecoder(1, precision="fp16")
    e = LdpcEncoder(baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0, targetRate=449 / 1024)
    assert (e.qm, e.maxCodeBlockSize, e.txBlockSize, e.numCodeBlocks, e.liftingSize, e.setIndex) == (2, 8448, 0, 0, 0, -1)
    assert ChanCodeBase.LARGE_LLR == 1e20 and e.LARGE_LLR == 1e20
    assert [ChanCodeBase.getCrcLen(p) for p in ("6", "11", "16", "24A", "24B", "24C")] == [6, 11, 16, 24, 24, 24]
    e.initialize(10024)
    assert (e.numCodeBlocks, e.liftingSize, e.setIndex, e.codeBlockSize) == (2, 240, 7, 5280)
    assert e.baseGraph.shape == (46, 68) and e.baseGraph.dtype == np.int16
    # the base graph follows a later re-initialisation (the reference's cache goes stale, ldpc.py:777)
    e.initialize(3000)
    assert e.liftingSize == 144 and e.baseGraph[0, 0] == O.base_graph(1, 144, 4)[0, 0]
    d = e.getDecoder()
    assert (d.baseGraphNo, d.modulation, d.txLayers, d.nRef) == (1, 'QPSK', 1, 0)
    assert "LDPC Encoder Properties" in repr(e) and "Target Rate" in repr(e) and "LDPC Decoder Properties" in repr(d)
    assert list(e.getRateMatchedCbLens(22807, 2)) == [11404, 11404]
    with pytest.raises(AssertionError):
        LdpcEncoder(2).baseGraph


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NrldpcError):
        ChanCodeBase.getCrc(np.array([1, 0, 1], np.int8), "24A")
    with pytest.raises(_native.NrldpcError):
        LdpcEncoder(1).getRateMatchedCodeBlocks(np.zeros(100, np.int8))
    h = ctypes.c_void_p()
    rc = _native.lib().nrldpc_create(0, ctypes.byref(h))
    assert rc == _native.ERR_CUDA and b"no CPU fallback" in _native.lib().nrldpc_last_error()


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the oracle is a checker only)."""
    pkg = os.path.join(ROOT, "neoradium_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "nr_oracle" not in src and "import oracle" not in src and "ref_loader" not in src, f


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The descriptor structs cross the ABI by value / as arrays: sizes and field offsets of the ctypes mirrors
    (neoradium_b200/_native.py) must equal what a C compiler makes of include/nrldpc.h."""
    import subprocess
    src = tmp_path / "layout.c"
    fields_cfg = ["bg", "zc", "K", "F", "C", "qm", "nl", "ncb", "rv", "reserved", "G"]
    fields_grp = ["cfg", "in_dtype", "reserved", "llr", "num_tb", "llr_len", "llr_stride", "soft_buffer", "tb_bits",
                  "tb_bits_stride", "cb_crc_ok", "tb_crc_ok", "iters"]
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrldpc.h"', 'int main(void) {',
            'printf("%zu %zu\\n", sizeof(nrldpc_tb_config), sizeof(nrldpc_tb_group));']
    body += ['printf("%%zu\\n", offsetof(nrldpc_tb_config, %s));' % f for f in fields_cfg]
    body += ['printf("%%zu\\n", offsetof(nrldpc_tb_group, %s));' % f for f in fields_grp]
    body += ['return 0; }']
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    sizes, offs = [int(v) for v in out[:2]], [int(v) for v in out[2:]]
    assert sizes == [ctypes.sizeof(_native.TbConfig), ctypes.sizeof(_native.TbGroup)]
    want = [getattr(_native.TbConfig, f).offset for f in fields_cfg] + [getattr(_native.TbGroup, f).offset for f in fields_grp]
    assert offs == want


def test_decoder_flags_word_matches_the_header():
    """_native.dec_flags builds the `flags` word of include/nrldpc.h: NRLDPC_DEC_EARLY_STOP | NRLDPC_DEC_ES_FROM(k), or
    NRLDPC_DEC_ES_AUTO for earlyStopFrom='auto'; the constants are the header's."""
    hdr = open(os.path.join(ROOT, "include", "nrldpc.h")).read()
    enum = {m.group(1): int(m.group(2)) for m in re.finditer(r"(NRLDPC_DEC_[A-Z_]+) = (\d+)", hdr)}
    assert enum == {"NRLDPC_DEC_EARLY_STOP": _native.DEC_EARLY_STOP, "NRLDPC_DEC_ALL_ROWS": _native.DEC_ALL_ROWS,
                    "NRLDPC_DEC_ES_AUTO": _native.DEC_ES_AUTO}
    assert _native.dec_flags(False, 5) == 0 and _native.dec_flags(False, "auto") == 0
    assert _native.dec_flags(True) == 1 | (1 << 8)
    assert _native.dec_flags(True, 6) == 1 | (6 << 8) and _native.dec_flags(True, 1000) == 1 | (255 << 8) and _native.dec_flags(True, -3) == 1
    assert _native.dec_flags(True, "auto") == 1 | 4
    with pytest.raises(ValueError):
        _native.dec_flags(True, "sometimes")
