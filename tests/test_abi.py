"""The C-ABI shared library: builds, loads, exports every symbol include/cleverrec_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from cleverrec_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cleverrec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(crb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    from cleverrec_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_abi_version_and_struct_layout(lib):
    from cleverrec_b200 import _lib
    assert lib.crb_abi_version() == 2
    assert C.sizeof(_lib.CrbTable) == 48 and C.sizeof(_lib.CrbOpt) == 48


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.crb_create(0, C.byref(h))
    assert rc == -2 and b"no CPU path" in lib.crb_last_error()
    from cleverrec_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_does_not_import_oracle():
    # the product package must never reach into oracle/ (parity claims depend on it)
    pkg = os.path.join(ROOT, "cleverrec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/_ref" not in src and "liboracle" not in src, f
