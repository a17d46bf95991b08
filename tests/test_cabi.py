"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/brb.h declares; with no CUDA
device the compute entry points fail loudly (BRB_ECUDA) instead of falling back."""
import ctypes as C
import pathlib
import re

import numpy as np
import pytest
import torch

from balance_robot_b200 import _cabi, mjcf, model

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    _cabi.build()
    return _cabi.lib()


def test_exports_match_header(lib):
    header = (ROOT / "include" / "brb.h").read_text()
    declared = set(re.findall(r"\b(brb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_consts_struct_layout_matches_header():
    header = (ROOT / "include" / "brb.h").read_text()
    body = header[header.index("typedef struct BrbModelConsts {"):header.index("} BrbModelConsts;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"(?:float|int)\s+([^;]+);", body):
        for item in decl.split(","):
            m = re.match(r"\s*(\w+)((?:\[\d+\])*)", item)
            dims = [int(x) for x in re.findall(r"\[(\d+)\]", m.group(2))]
            names.append((m.group(1), int(np.prod(dims)) if dims else 1))
    fields = [(n, (C.sizeof(t) // 4)) for n, t in model.BrbModelConsts._fields_]
    assert names == fields
    assert C.sizeof(model.BrbModelConsts) == 4 * sum(k for _, k in names)


def test_version_and_strerror(lib):
    assert lib.brb_version() >= 100
    assert b"invalid" in lib.brb_strerror(-22)
    assert b"CUDA" in lib.brb_strerror(-5)


def test_argument_validation(lib):
    out = C.c_void_p()
    assert lib.brb_model_create(None, None, 0, 0, C.byref(out)) == -22
    assert lib.brb_env_create(None, 10, 0, 0, C.byref(out)) == -22
    assert lib.brb_env_step(None, None, None, None, None, None, None, None, None, None, None) == -22


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    rm = model.compile_model(mjcf.parse("scene_env01.xml"), 1, 6000)
    tt = np.ascontiguousarray(rm.time_table)
    out = C.c_void_p()
    rc = lib.brb_model_create(C.byref(rm.consts), tt.ctypes.data, len(tt), 0, C.byref(out))
    assert rc == -5                                  # BRB_ECUDA: no device -> error, never a CPU path
    from balance_robot_b200 import make_vec
    with pytest.raises(_cabi.BrbError):
        make_vec("Env01-v2", 8)
    with pytest.raises(_cabi.BrbError):
        make_vec("Env01-v2", 8, device="cpu")


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "balance_robot_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "libbrb_ref" not in src, py
    for cu in (ROOT / "balance_robot_b200" / "csrc").glob("*"):
        if cu.suffix in (".cu", ".h"):
            assert "brb_ref" not in cu.read_text(), cu
