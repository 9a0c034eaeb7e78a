"""GPU parity tests for the BM25 kernel (through the C ABI) against the oracle's restatement of tantivy.
Parity with tantivy itself is unpinned: these tests pin the CUDA path to the oracle."""
import numpy as np
import pytest

import oracle as O
from nucliadb_b200 import _lib
from nucliadb_b200.segment import TextSegment

pytestmark = pytest.mark.gpu


def corpus(n_docs, n_terms, seed, mean_len=40):
    rng = np.random.default_rng(seed)
    lens = np.maximum(1, rng.lognormal(np.log(mean_len), 0.6, n_docs).astype(np.int64))
    doc_off = np.concatenate([[0], np.cumsum(lens)])
    tokens = (rng.zipf(1.2, int(doc_off[-1])) - 1) % n_terms
    return O.Postings(doc_off, tokens.astype(np.uint32), n_terms)


def run(P, queries, k, mode, use_tf, min_score=0.0, alive=None):
    ts = TextSegment.create(P.n_docs, P.n_terms, P.term_off, P.post_doc, P.post_tf, P.fieldnorm_id)
    ts.set_stats(P.n_docs, P.total_tokens, P.doc_freq)
    if alive is not None:
        ts.set_alive(alive)
    qoff = np.concatenate([[0], np.cumsum([len(x) for x in queries])]).astype(np.uint32)
    qt = np.concatenate([np.asarray(x, dtype=np.uint32) for x in queries]) if qoff[-1] else np.zeros(0, np.uint32)
    return ts.search(qt, qoff, k, mode=mode, use_tf=use_tf, min_score=min_score)


@pytest.mark.parametrize("mode,use_tf", [(_lib.NIDX_BM25_OR, False), (_lib.NIDX_BM25_OR, True), (_lib.NIDX_BM25_AND, True)])
def test_bm25_matches_oracle(mode, use_tf):
    P = corpus(60000, 5000, seed=7)
    rng = np.random.default_rng(1)
    nterms = 3 if mode == _lib.NIDX_BM25_AND else 12
    queries = [list(rng.choice(400, nterms, replace=False) + (0 if mode == _lib.NIDX_BM25_AND else 20)) for _ in range(40)]
    docs, sc, cnt, total = run(P, queries, 100, mode, use_tf)
    od, osc, oc, otot = O.bm25_search(P, queries, 100, mode=mode, use_tf=use_tf, nthreads=4)
    assert (total == otot).all()           # Count collector: exact
    assert (cnt == oc).all()
    assert np.allclose(sc, osc, rtol=1e-5, atol=1e-5)   # stated tolerance (fixed-point accumulation vs f32 sum)
    # ids: identical wherever the oracle's scores are separated by more than the tolerance
    for q in range(len(queries)):
        c = cnt[q]
        if c == 0:
            continue
        gaps = np.abs(np.diff(osc[q, :c])) > 2e-5 * np.maximum(1.0, np.abs(osc[q, 1:c]))
        strict = np.concatenate([[True], gaps]) & np.concatenate([gaps, [True]])
        assert (docs[q, :c][strict] == od[q, :c][strict]).all()
        assert set(docs[q, :c].tolist()) == set(od[q, :c].tolist()) or not strict.all()


def test_bm25_ties_keep_doc_order():
    # tf == 1 and equal lengths => exactly equal scores: TopDocs orders by doc id ascending
    n_docs, n_terms = 5000, 50
    doc_off = np.arange(0, (n_docs + 1) * 8, 8)
    rng = np.random.default_rng(3)
    tokens = np.concatenate([rng.choice(n_terms, 8, replace=False) for _ in range(n_docs)]).astype(np.uint32)
    P = O.Postings(doc_off, tokens, n_terms)
    queries = [[1, 2, 3], [10], [4, 40]]
    docs, sc, cnt, total = run(P, queries, 50, _lib.NIDX_BM25_OR, False)
    od, osc, oc, otot = O.bm25_search(P, queries, 50, mode=O.BM25_OR, use_tf=False)
    assert (docs == od).all() and (cnt == oc).all() and (total == otot).all()
    assert np.allclose(sc, osc, rtol=1e-5, atol=1e-6)


def test_bm25_min_score_alive_and_missing_terms():
    P = corpus(20000, 2000, seed=9)
    alive = np.ones(P.n_docs, dtype=bool)
    alive[::2] = False
    words = np.zeros((P.n_docs + 63) // 64 * 8, dtype=np.uint8)
    pb = np.packbits(alive, bitorder="little")
    words[: len(pb)] = pb
    bits = words.view(np.uint64)
    queries = [[5, 6, 7], [1999999], [], [3, 1999999]]
    docs, sc, cnt, total = run(P, queries, 20, _lib.NIDX_BM25_OR, True, alive=bits)
    od, osc, oc, otot = O.bm25_search(P, queries, 20, mode=O.BM25_OR, use_tf=True, alive_bits=bits)
    assert (cnt == oc).all() and (total == otot).all()
    assert np.allclose(sc, osc, rtol=1e-5, atol=1e-5)
    assert all(alive[d] for d in docs[docs != 0xFFFFFFFF])
    # AND with an unknown term matches nothing (tantivy: empty term => empty intersection)
    d2, s2, c2, t2 = run(P, [[3, 1999999]], 20, _lib.NIDX_BM25_AND, True)
    assert c2[0] == 0 and t2[0] == 0
    # min_score cut after top-k (nidx_text/src/reader.rs:302-305)
    d0, s0, c0, _ = run(P, [[5, 6, 7]], 20, _lib.NIDX_BM25_OR, True)
    thr = float(s0[0, 7])
    d3, s3, c3, _ = run(P, [[5, 6, 7]], 20, _lib.NIDX_BM25_OR, True, min_score=thr)
    assert c3[0] == int((s0[0, : c0[0]] >= thr).sum()) and (s3[0, : c3[0]] >= thr).all()
    assert (d3[0, : c3[0]] == d0[0, : c3[0]]).all()


def check_against_oracle(P, queries, k, mode, use_tf, **kw):
    docs, sc, cnt, total = run(P, queries, k, mode, use_tf, **kw)
    od, osc, oc, otot = O.bm25_search(P, queries, k, mode=mode, use_tf=use_tf, nthreads=4)
    assert (total == otot).all() and (cnt == oc).all()
    assert np.allclose(sc, osc, rtol=1e-5, atol=1e-5)
    for q in range(len(queries)):   # the same documents wherever the oracle's k-th score is separated from the next
        c = cnt[q]
        if c and (c < k or True):
            assert set(docs[q, :c].tolist()) == set(od[q, :c].tolist()) or abs(osc[q, c - 1] - osc[q, max(c - 2, 0)]) < 2e-5
    return docs, sc, cnt, total


def test_bm25_dense_tiles_fall_back_to_fine_tiles():
    """Few documents, long documents, frequent terms: a fine tile (4096 docs) holds far more postings than a round has slots,
    so the tile span drops to one fine tile processed in several rounds (postings re-read in phase C, runs located by binary
    search instead of the octet map)."""
    P = corpus(30000, 60, seed=21, mean_len=120)
    rng = np.random.default_rng(4)
    queries = [list(rng.choice(60, 40, replace=False)) for _ in range(6)]
    check_against_oracle(P, queries, 100, _lib.NIDX_BM25_OR, True)
    check_against_oracle(P, queries, 100, _lib.NIDX_BM25_OR, False)
    check_against_oracle(P, [q[:3] for q in queries], 100, _lib.NIDX_BM25_AND, True)


def test_bm25_sparse_query_spans_many_fine_tiles_per_tile():
    """Rare terms over many documents: one tile covers the maximum span of fine tiles (terms without a skip row are walked
    linearly), and the threshold-crossing candidate list carries the top-k from tile to tile."""
    P = corpus(300000, 40000, seed=23, mean_len=30)
    df = np.diff(P.term_off.astype(np.int64))
    rng = np.random.default_rng(5)
    rare = np.nonzero((df >= 3) & (df < 200))[0]
    mid = np.nonzero(df >= 300)[0]
    queries = [list(rng.choice(rare, 30, replace=False)) for _ in range(8)] + [list(rng.choice(mid, 20, replace=False)) for _ in range(8)]
    queries += [list(rng.choice(rare, 10, replace=False)) + list(rng.choice(mid, 10, replace=False)) for _ in range(8)]
    check_against_oracle(P, queries, 100, _lib.NIDX_BM25_OR, False)
    check_against_oracle(P, queries, 10, _lib.NIDX_BM25_OR, True)
    check_against_oracle(P, [[int(q[-1]), int(q[-2])] for q in queries[8:]], 50, _lib.NIDX_BM25_AND, True)
