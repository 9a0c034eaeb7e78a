"""The CUDA path, through the C ABI, against the committed fixtures of tests/golden/ (frozen oracle outputs, see
tests/golden/make_golden.py): ids, edges and vector scores bit-exact, BM25 scores within the stated 1e-5.  Independent of what
is compiled on the GPU box.  (Sorted last on purpose: the live oracle comparisons of test_gpu_*.py run first.)"""
import os

import numpy as np
import pytest

import oracle as O
from test_golden_fixtures import BM25_CASES, load, postings_of

pytestmark = pytest.mark.gpu


def test_cuda_path_reproduces_the_vector_fixture():
    from nucliadb_b200 import _lib
    from nucliadb_b200.segment import VectorSegment

    fx = load("vector_small.npz")
    v, q = fx["vectors"], fx["queries"]
    m, m0, efc, seed, max_batch = (int(x) for x in fx["build"])
    for name, sim in (("cos", _lib.NIDX_SIM_COSINE), ("dot", _lib.NIDX_SIM_DOT)):
        seg = VectorSegment.create(v, v.shape[1], similarity=sim, m=m, m0=m0, ef_construction=efc)
        ids, sc, cnt = seg.search(q, 10, method=_lib.NIDX_METHOD_BRUTE)
        assert (ids == fx[f"bf_{name}_ids"]).all() and np.array_equal(sc, fx[f"bf_{name}_scores"]) and (cnt == fx[f"bf_{name}_counts"]).all()
    seg = VectorSegment.create(v, v.shape[1], similarity=_lib.NIDX_SIM_COSINE, m=m, m0=m0, ef_construction=efc)
    seg.set_graph(fx["level"], fx["adj0"], fx["adjU"], fx["w0"], fx["wU"])                 # search on the frozen graph
    ids, sc, cnt = seg.search(q, 10, ef=40, method=_lib.NIDX_METHOD_HNSW)
    assert (ids == fx["hnsw_ids"]).all() and np.array_equal(sc, fx["hnsw_scores"]) and (cnt == fx["hnsw_counts"]).all()
    ids, sc, cnt = seg.search(q, 10, ef=40, min_score=0.0, with_duplicates=False, method=_lib.NIDX_METHOD_HNSW, filter_bits=fx["filter_bits"])
    assert (ids == fx["filt_ids"]).all() and np.array_equal(sc, fx["filt_scores"]) and (cnt == fx["filt_counts"]).all()
    seg = VectorSegment.create(v, v.shape[1], similarity=_lib.NIDX_SIM_COSINE, m=m, m0=m0, ef_construction=efc)
    seg.build_hnsw(seed=seed, max_batch=max_batch)                                            # and the build gives the frozen graph
    g = seg.get_graph()
    rows = int(fx["level"].astype(np.int64).sum())
    assert (g["level"] == fx["level"]).all() and (g["adj0"] == fx["adj0"]).all() and np.array_equal(g["w0"], fx["w0"])
    assert (g["adjU"][:rows] == fx["adjU"][:rows]).all() and [g["entry_node"], g["entry_layer"]] == fx["entry"].tolist()


def test_cuda_path_reproduces_the_bm25_and_rabitq_fixtures():
    from nucliadb_b200 import _lib
    from nucliadb_b200.segment import TextSegment, VectorSegment

    fx = load("bm25_small.npz")
    P = postings_of(fx)
    ts = TextSegment.create(P.n_docs, P.n_terms, P.term_off, P.post_doc, P.post_tf, P.fieldnorm_id)
    ts.set_stats(P.n_docs, P.total_tokens, P.doc_freq)
    for name, mode, use_tf in BM25_CASES:
        queries = fx[f"{name}_queries"]
        qoff = (np.arange(len(queries) + 1) * queries.shape[1]).astype(np.uint32)
        docs, sc, cnt, total = ts.search(queries.reshape(-1).astype(np.uint32), qoff, 20, mode={O.BM25_OR: _lib.NIDX_BM25_OR, O.BM25_AND: _lib.NIDX_BM25_AND}[mode],
                                         use_tf=use_tf)
        od, osc, oc = fx[f"{name}_docs"], fx[f"{name}_scores"], fx[f"{name}_counts"]
        assert (total == fx[f"{name}_total"]).all() and (cnt == oc).all()
        assert np.allclose(sc, osc, rtol=1e-5, atol=1e-5)          # stated tolerance: fixed-point accumulation vs the oracle's f32 sum
        for i in range(len(queries)):                               # ids wherever the frozen scores are separated by more than that
            c = int(oc[i])
            if c == 0:
                continue
            gaps = np.abs(np.diff(osc[i, :c])) > 2e-5 * np.maximum(1.0, np.abs(osc[i, 1:c]))
            strict = np.concatenate([[True], gaps]) & np.concatenate([gaps, [True]])
            assert (docs[i, :c][strict] == od[i, :c][strict]).all()
    fx = load("rabitq_small.npz")
    v = fx["vectors"]
    seg = VectorSegment.create(v, v.shape[1], similarity=_lib.NIDX_SIM_DOT)
    seg.rabitq_encode()
    assert (seg.rabitq_codes() == fx["codes"]).all()
    est, err = seg.rabitq_estimate(fx["queries"])
    assert np.array_equal(est, fx["estimate"]) and np.array_equal(err, fx["error"])
    ids, sc, cnt = seg.search(fx["queries"], 10, min_score=0.0, method=_lib.NIDX_METHOD_BRUTE_RABITQ)
    assert (ids == fx["scan_ids"]).all() and np.array_equal(sc, fx["scan_scores"]) and (cnt == fx["scan_counts"]).all()
