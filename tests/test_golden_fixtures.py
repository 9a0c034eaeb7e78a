"""The committed fixtures of tests/golden/ (made by tests/golden/make_golden.py from the oracle on small seeded inputs).
Here (CPU): the oracle still reproduces them bit for bit.  tests/test_gpu_zz_golden.py: the CUDA path reproduces them."""
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
