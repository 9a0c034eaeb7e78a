"""The committed fixtures of tests/golden/ (made by tests/golden/make_golden.py from the oracle on small seeded inputs).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path, through the C ABI, reproduces them -- ids, edges and
vector scores bit-exact, BM25 scores within the stated 1e-5."""
import os

import numpy as np
import pytest

import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def graph_of(fx):
    n = len(fx["vectors"])
    m, m0 = int(fx["build"][0]), int(fx["build"][1])
    g = O.Graph(n, m, m0, fx["level"])
    rows = fx["adjU"].shape[0] if int(fx["level"].astype(np.int64).sum()) else 0
    g.adj0[:], g.w0[:] = fx["adj0"], fx["w0"]
    g.adjU[:rows], g.wU[:rows] = fx["adjU"][:rows], fx["wU"][:rows]
    g.entry_node, g.entry_layer = int(fx["entry"][0]), int(fx["entry"][1])
    return g


def postings_of(fx):
    return O.Postings(fx["doc_off"], fx["tokens"], int(fx["n_terms"][0]))


BM25_CASES = (("or_tf1", O.BM25_OR, False), ("or_tf", O.BM25_OR, True), ("and_tf", O.BM25_AND, True))


# ---- CPU: the oracle against its own frozen outputs -------------------------------------------------------------------
def test_oracle_reproduces_the_vector_fixture():
    fx = load("vector_small.npz")
    v, q = fx["vectors"], fx["queries"]
    for name, sim in (("cos", O.SIM_COSINE), ("dot", O.SIM_DOT)):
        ids, sc, cnt = O.brute_force(v, q, 10, sim=sim, min_score=-1.0)
        assert (ids == fx[f"bf_{name}_ids"]).all() and np.array_equal(sc, fx[f"bf_{name}_scores"]) and (cnt == fx[f"bf_{name}_counts"]).all()
    m, m0, efc, seed, max_batch = (int(x) for x in fx["build"])
    g = O.hnsw_build(v, M=m, M0=m0, efC=efc, seed=seed, max_batch=max_batch, nthreads=4)      # thread count must not matter
    rows = int(g.level.astype(np.int64).sum())
    assert (g.level == fx["level"]).all() and (g.adj0 == fx["adj0"]).all() and np.array_equal(g.w0, fx["w0"])
    assert (g.adjU[:rows] == fx["adjU"][:rows]).all() and [g.entry_node, g.entry_layer] == fx["entry"].tolist()
    ids, sc, cnt, counters = O.hnsw_search(v, graph_of(fx), q, 10, 40, nthreads=2)
    assert (ids == fx["hnsw_ids"]).all() and np.array_equal(sc, fx["hnsw_scores"]) and (cnt == fx["hnsw_counts"]).all()
    assert (counters == fx["hnsw_counters"]).all()
    ids, sc, cnt, _ = O.hnsw_search(v, graph_of(fx), q, 10, 40, min_score=0.0, with_duplicates=False, filter_bits=fx["filter_bits"])
    assert (ids == fx["filt_ids"]).all() and np.array_equal(sc, fx["filt_scores"]) and (cnt == fx["filt_counts"]).all()
    assert (fx["filt_counts"][-4:] <= fx["hnsw_counts"][-4:]).all()


def test_oracle_reproduces_the_bm25_and_rabitq_fixtures():
    fx = load("bm25_small.npz")
    P = postings_of(fx)
    for name, mode, use_tf in BM25_CASES:
        d, s, c, tot = O.bm25_search(P, [list(x) for x in fx[f"{name}_queries"]], 20, mode=mode, use_tf=use_tf, nthreads=2)
        assert (d == fx[f"{name}_docs"]).all() and np.array_equal(s, fx[f"{name}_scores"]) and (c == fx[f"{name}_counts"]).all()
        assert (tot == fx[f"{name}_total"]).all()
    fx = load("rabitq_small.npz")
    enc = O.rabitq_encode(fx["vectors"])
    assert (enc == fx["codes"]).all()
    est, err = O.rabitq_estimate(enc, fx["vectors"].shape[1], fx["queries"])
    assert np.array_equal(est, fx["estimate"]) and np.array_equal(err, fx["error"])
    ids, sc, cnt, evals = O.rabitq_brute_force(fx["vectors"], enc, fx["queries"], 10, min_score=0.0)
    assert (ids == fx["scan_ids"]).all() and np.array_equal(sc, fx["scan_scores"]) and (evals == fx["scan_exact_evals"]).all()


# ---- GPU: the CUDA path against the same frozen outputs ---------------------------------------------------------------
@pytest.mark.gpu
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


@pytest.mark.gpu
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
