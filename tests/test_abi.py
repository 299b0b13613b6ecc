"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/mrn_b200.h declares
(no compute calls without a GPU), the struct layout matches, and host-only helpers agree with the oracle's layout."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mrn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mrn_b200.build import build
        build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from mrn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mrn_b200.h")).read()
    declared = set(re.findall(r"\b(mrnb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mrnb_version() >= 100


def test_header_enum_values_match_binding():
    from mrn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mrn_b200.h")).read()
    assert "MRNB_P_BLOCK0 = 13" in header and _lib.P_BLOCK0 == 13
    assert _lib.P_SUB0 == 13 + 12 * 12 and _lib.P_SEQ_W == 13 + 12 * 12 + 3 * 4 and _lib.P_COUNT == _lib.P_SEQ_B + 1


def test_router_param_offsets_follow_module_parameter_order(lib):
    from mrn_b200 import ops
    from oracle import synth
    for I in (2, 3, 6):
        n, off = ops.router_param_offsets(I)
        shapes = synth.router_shapes(I)
        assert list(shapes) == list(ops.ROUTER_PARAM_NAMES)
        sizes = [int(__import__("numpy").prod(s)) for s in shapes.values()]
        assert [off[k + 1] - off[k] for k in range(20)] == [(s + 7) // 8 * 8 for s in sizes]      # 32-byte aligned slots
        assert all(o % 8 == 0 for o in off) and n == off[-1]


def test_argument_errors_are_reported_without_a_gpu(lib):
    rc = lib.mrnb_clip_adam(None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, None, None, None)
    assert rc < 0 and b"clip_adam" in lib.mrnb_last_error()
    assert lib.mrnb_svtr_workspace_bytes(6, 256, 32, 1) > 0


def test_struct_layouts_match_the_c_header(tmp_path):
    """include/mrn_b200.h is plain C: compile a probe with gcc and compare sizeof / offsetof of the pack structs with the
    ctypes mirrors in mrn_b200/_lib.py (a silent mismatch would hand the kernels garbage pointers)."""
    import ctypes as C
    import shutil
    import subprocess
    from mrn_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mrn_b200.h"\n'
                   'int main(void) {\n'
                   '  printf("%zu %zu %zu %zu\\n", sizeof(MrnbSvtrPack), offsetof(MrnbSvtrPack, h), offsetof(MrnbSvtrPack, fc_w), offsetof(MrnbSvtrPack, n_class));\n'
                   '  printf("%zu %zu %zu %zu\\n", sizeof(MrnbCrnnPack), offsetof(MrnbCrnnPack, h), offsetof(MrnbCrnnPack, fc_w), offsetof(MrnbCrnnPack, n_class));\n'
                   '  printf("%zu %zu %zu %zu\\n", sizeof(MrnbCrnnTrainPack), offsetof(MrnbCrnnTrainPack, h), offsetof(MrnbCrnnTrainPack, bn_mean), offsetof(MrnbCrnnTrainPack, n_class));\n'
                   '  printf("%d %d %d %d\\n", MRNB_P_COUNT, MRNB_C_COUNT, MRNB_T_COUNT, MRNB_T_FC_W);\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = [[int(v) for v in line.split()] for line in out if line.strip()]
    for row, st, third in ((rows[0], _lib.MrnbSvtrPack, "fc_w"), (rows[1], _lib.MrnbCrnnPack, "fc_w"),
                           (rows[2], _lib.MrnbCrnnTrainPack, "bn_mean")):
        assert row == [C.sizeof(st), getattr(st, "h").offset, getattr(st, third).offset, getattr(st, "n_class").offset], st
    assert rows[3] == [_lib.P_COUNT, _lib.C_COUNT, _lib.T_COUNT, _lib.T_FC_W]
