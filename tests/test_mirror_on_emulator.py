"""The mirror of the reference interface (nucliadb_b200/vector.py, text.py) on a machine without a GPU: the reference's own test
flows, which tests/test_gpu_mirror*.py run against the CUDA library, run here against tests/abi_emulator.py (the same ctypes
calls answered by the oracle).  What this covers is the HOST logic above the C ABI -- formula evaluation, prefilters, Fssc,
MaxSim re-scoring, deletions by sequence, merges with and without graph reuse, the segment directory round trip, BM25 request
handling and search-after; the kernels are covered by the same functions in the GPU suite."""
import inspect

import pytest

import abi_emulator
import test_gpu_mirror
import test_gpu_mirror_segment
import test_gpu_zz_open_dir
from nucliadb_b200 import _lib

SKIP = {"test_segment_files_round_trip": "compares device-built files with the oracle's writer: a test of the library, not of the mirror"}


def _cases():
    for mod in (test_gpu_mirror, test_gpu_mirror_segment, test_gpu_zz_open_dir):
        for name, fn in inspect.getmembers(mod, inspect.isfunction):
            if name.startswith("test_") and fn.__module__ == mod.__name__ and name not in SKIP:
                marks = getattr(fn, "pytestmark", [])
                params = [m for m in marks if m.name == "parametrize"]
                if params:
                    for value in params[0].args[1]:
                        yield pytest.param(fn, {params[0].args[0]: value}, id=f"{mod.__name__}.{name}[{value}]")
                else:
                    yield pytest.param(fn, {}, id=f"{mod.__name__}.{name}")


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", abi_emulator.EmulatedLib())


@pytest.mark.parametrize("fn,kwargs", list(_cases()))
def test_mirror_flow_on_the_emulated_abi(fn, kwargs, emulated, tmp_path, monkeypatch):
    sig = inspect.signature(fn).parameters
    if "tmp_path" in sig:
        kwargs = dict(kwargs, tmp_path=tmp_path)
    if "monkeypatch" in sig:
        kwargs = dict(kwargs, monkeypatch=monkeypatch)
    fn(**kwargs)
