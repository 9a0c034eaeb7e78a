"""GPU parity tests for the RaBitQ slice (vector_types/rabitq.rs): codes, estimator and the quantised exact scan,
against oracle/rabitq.hpp (which is pinned to the reference's own test_rabitq_estimate)."""
import numpy as np
import pytest

import oracle as O
from conftest import make_queries, make_vectors
from nucliadb_b200 import _lib
from nucliadb_b200.segment import VectorSegment

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [64, 256, 768])
def test_codes_and_estimates_are_bit_identical(d):
    v = make_vectors(3000, d, seed=51)
    v[5, :7] = 0.0                      # zeros quantise to the negative side (v > 0.0 is false)
    q = make_queries(v, 9)
    seg = VectorSegment.create(v, d, similarity=_lib.NIDX_SIM_DOT)
    seg.rabitq_encode()
    codes = seg.rabitq_codes()
    want = O.rabitq_encode(v, nthreads=4)
    assert codes.shape == want.shape == (3000, d // 8 + 8)        # the reference's vectors.quant record
    assert (codes == want).all()
    est, err = seg.rabitq_estimate(q)
    oest, oerr = O.rabitq_estimate(want, d, q, nthreads=4)
    assert np.array_equal(est, oest) and np.array_equal(err, oerr)
    exact = q @ v.T
    assert np.mean(np.abs(exact - est) < err) > 0.9               # the bound is probabilistic (EPSILON = 1.9)


def test_rabitq_scan_matches_oracle_and_exact_top_k():
    v = make_vectors(30000, 256, seed=52)
    q = make_queries(v, 40)
    seg = VectorSegment.create(v, 256, similarity=_lib.NIDX_SIM_DOT)
    seg.rabitq_encode()
    alive = np.ones(len(v), dtype=bool)
    alive[::7] = False
    words = np.zeros((len(v) + 63) // 64 * 8, dtype=np.uint8)
    pb = np.packbits(alive, bitorder="little")
    words[: len(pb)] = pb
    for k, ms in ((10, 0.0), (50, 0.3), (1, 0.0)):
        ids, sc, cnt = seg.search(q, k, min_score=ms, method=_lib.NIDX_METHOD_BRUTE_RABITQ)
        oi, os_, oc, evals = O.rabitq_brute_force(v, O.rabitq_encode(v, nthreads=4), q, k, min_score=ms, nthreads=4)
        assert (cnt == oc).all() and (ids == oi).all() and np.array_equal(sc, os_)          # the sequential rerank semantics, exactly
        bi, bs, bc = seg.search(q, k, min_score=ms, method=_lib.NIDX_METHOD_BRUTE)
        assert np.mean([len(set(a[:c]) & set(b[:c])) / max(c, 1) for a, b, c in zip(ids, bi, bc)]) >= 0.98
    seg.set_alive(words.view(np.uint64))
    ids, sc, cnt = seg.search(q, 10, min_score=0.0, method=_lib.NIDX_METHOD_BRUTE_RABITQ)
    assert all(alive[i] for i in ids[ids != 0xFFFFFFFF])


def test_rabitq_needs_dot_and_dim_multiple_of_64():
    v = make_vectors(100, 96, seed=1)
    seg = VectorSegment.create(v, 96, similarity=_lib.NIDX_SIM_DOT)
    with pytest.raises(_lib.NidxError):
        seg.rabitq_encode()
    seg = VectorSegment.create(make_vectors(100, 128, seed=1), 128, similarity=_lib.NIDX_SIM_COSINE)
    with pytest.raises(_lib.NidxError):
        seg.rabitq_encode()
