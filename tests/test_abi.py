"""CPU: the C-ABI library builds, loads and exports every symbol include/cgx.h declares;
without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers
from new_cg_variants_b200 import _lib, build


def header_functions():
    src = open(os.path.join(helpers.ROOT, "include", "cgx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgx_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_for_sm100a():
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 20
    lib = C.CDLL(build.build())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cgx.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes table out of sync with include/cgx.h"
    assert _lib.load().cgx_version() == 100


def test_info_struct_layout():
    assert C.sizeof(_lib.CgxInfo) == 4 * 8 + 8 + 4 * 4


def _no_gpu():
    return _lib.load().cgx_device_count() == 0


def test_fails_loudly_without_gpu():
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.cgx_ctx_create(0, C.byref(ctx))
    assert rc == _lib.ERR_CUDA and not ctx
    assert lib.cgx_last_error()
    from new_cg_variants_b200 import Session
    from new_cg_variants_b200.cg_variants import pr_pcg
    A = helpers.load_matrix("nos4")
    with pytest.raises(_lib.CgxError):
        Session(A)
    with pytest.raises(_lib.CgxError):
        pr_pcg(A, np.ones(100), np.zeros(100), 5)


def test_bad_arguments_are_reported():
    lib = _lib.load()
    assert lib.cgx_ctx_create(0, None) == _lib.ERR_ARG
    assert lib.cgx_run(None, 0, 10, 0, 0, None) == _lib.ERR_ARG
    assert b"ctx" in lib.cgx_last_error()
    assert lib.cgx_ctx_destroy(None) == _lib.OK


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU solver)."""
    pkg = os.path.join(helpers.ROOT, "new_cg_variants_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "cg_oracle" not in src, f
