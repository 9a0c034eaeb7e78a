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
import test_gpu_binding
import test_gpu_zz_open_dir
from nucliadb_b200 import _lib

SKIP = {"test_segment_files_round_trip": "compares device-built files with the oracle's writer: a test of the library, not of the mirror"}


def _cases():
    for mod in (test_gpu_mirror, test_gpu_mirror_segment, test_gpu_zz_open_dir, test_gpu_binding):
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


# ---- more of searcher.rs:150-199, 241-343 (cross-segment behaviour), on the emulated ABI only ---------------------------------------
import numpy as np

from nucliadb_b200 import vector as V

RID = "9cb39c75f8d9498d8f82d92b173011f5"


def _unit(rng, d):
    v = rng.standard_normal(d).astype(np.float32)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_cross_segment_top_k_dedup_and_id_collisions(emulated):
    rng = np.random.default_rng(77)
    cfg = V.VectorConfig(dimension=16, similarity=V.Similarity.Dot)
    vecs = [_unit(rng, 16) for _ in range(60)]
    segs = [V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/s{s}/{i}-{i + 1}", [vecs[20 * s + i]]) for i in range(20)], cfg) for s in range(3)]
    searcher = V.VectorSearcher.open(cfg, [(seg, i + 1) for i, seg in enumerate(segs)])
    q = _unit(rng, 16)
    docs = searcher.search(V.VectorSearchRequest(vector=q, result_per_page=5, min_score=-1.0)).documents
    exact = sorted(((float(np.float32(np.dot(v.astype(np.float64), q.astype(np.float64)))), i) for i, v in enumerate(vecs)), reverse=True)[:5]
    assert [d.doc_id for d in docs] == [f"{RID}/f/s{i // 20}/{i % 20}-{i % 20 + 1}" for _, i in exact]      # the union's exact top-5
    assert all(abs(d.score - s) < 1e-6 for d, (s, _) in zip(docs, exact)) and [d.score for d in docs] == sorted((d.score for d in docs), reverse=True)
    # the same vector under different ids in two segments: one result unless duplicates are asked for (Fssc.seen, searcher.rs:175-183)
    a = V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/a/0-1", [vecs[0]])], cfg)
    b = V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/b/0-1", [vecs[0]]), V.Elem(f"{RID}/f/b/1-2", [vecs[1]])], cfg)
    two = V.VectorSearcher.open(cfg, [(a, 1), (b, 2)])
    n = lambda dup: len(two.search(V.VectorSearchRequest(vector=vecs[0], result_per_page=10, min_score=-1.0, with_duplicates=dup)).documents)
    assert n(True) == 3 and n(False) == 2
    # the same paragraph id in two segments (an update whose deletion has not been applied): the buffer is keyed by id
    c = V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/a/0-1", [vecs[2]])], cfg)
    same_id = V.VectorSearcher.open(cfg, [(a, 1), (c, 2)])
    docs = same_id.search(V.VectorSearchRequest(vector=vecs[0], result_per_page=10, min_score=-1.0, with_duplicates=True)).documents
    assert [d.doc_id for d in docs] == [f"{RID}/f/a/0-1"] and docs[0].score > 0.9999      # first segment's entry stays (HashSet::insert keeps the old one)
    # ... and with its deletion applied (seq 2 > segment a's seq 1) only the new version is left
    updated = V.VectorSearcher.open(cfg, [(V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/a/0-1", [vecs[0]])], cfg), 1), (c, 2)], deletions=[(f"{RID}/f/a", 2)])
    docs = updated.search(V.VectorSearchRequest(vector=vecs[0], result_per_page=10, min_score=-1.0)).documents
    assert len(docs) == 1 and docs[0].score < 0.9


def test_normalize_vectors_at_index_and_query_time(emulated):
    """config.normalize_vectors (searcher.rs:246-252, indexer.rs:94-146, utils.rs:20-23): Dot over normalised vectors."""
    rng = np.random.default_rng(78)
    cfg = V.VectorConfig(dimension=8, similarity=V.Similarity.Dot, normalize_vectors=True)
    base = [_unit(rng, 8) for _ in range(10)]
    seg = V.VectorIndexer.index_elems([V.Elem(f"{RID}/f/x/{i}-{i + 1}", [base[i] * np.float32(3.0 + i)]) for i in range(10)], cfg)
    searcher = V.VectorSearcher.open(cfg, [(seg, 1)])
    docs = searcher.search(V.VectorSearchRequest(vector=base[4] * np.float32(0.01), result_per_page=3, min_score=-1.0)).documents
    assert docs[0].doc_id == f"{RID}/f/x/4-5" and abs(docs[0].score - 1.0) < 1e-6 and docs[1].score < 0.999
