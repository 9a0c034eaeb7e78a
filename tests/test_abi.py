"""The C-ABI library loads and exports exactly what include/nidx_b200.h declares; without a CUDA device
every entry point fails loudly (no CPU fallback).  CPU only: no compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nucliadb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "nidx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nidx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    L = _lib.load()
    names = header_functions()
    assert names, "no functions parsed from the header"
    for n in names:
        assert hasattr(L, n), f"libnidx_b200.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names


def test_library_does_not_link_the_oracle():
    out = os.popen(f"nm -D --defined-only {_lib.LIB_PATH}").read()
    assert "oracle_" not in out
    for f in ("vector.py", "text.py", "segment.py", "dist.py", "_lib.py", "__init__.py"):
        src = open(os.path.join(ROOT, "nucliadb_b200", f)).read()
        assert "import oracle" not in src and "from oracle" not in src


def test_no_device_fails_loudly():
    L = _lib.load()
    if L.nidx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = _lib.VecConfig(8, _lib.NIDX_SIM_COSINE, 0, 0, 0, 0, 0, 0)
    h = C.c_void_p()
    v = np.zeros((4, 8), dtype=np.float32)
    rc = L.nidx_vec_create(C.byref(cfg), _lib.ptr(v), C.c_uint64(4), C.c_int32(8), _lib.NIDX_MEM_HOST, None, C.byref(h))
    assert rc == -2 and b"no CUDA device" in L.nidx_last_error()
    with pytest.raises(_lib.NidxError):
        _lib.require_device()
    from nucliadb_b200.segment import VectorSegment
    with pytest.raises(_lib.NidxError):
        VectorSegment.create(v, 8)
