"""Second opinions on the oracle: small, independent pure-Python / float64 restatements of the same published formulas, written
without looking at oracle/*.hpp's structure.  They cannot pin the oracle to the reference (test_oracle_golden.py does what the
reference's own tests allow), but a slip in the C++ restatement would have to be repeated here to go unnoticed."""
import math

import numpy as np

import oracle as O
from conftest import make_queries, make_vectors


def naive_bm25(docs, n_terms, query, mode, use_tf, k1=1.2, b=0.75):
    """tantivy's Bm25Weight from its published formula, float64, document at a time.  docs: list of token-id lists."""
    n = len(docs)
    avg = sum(len(d) for d in docs) / n
    df = [sum(1 for d in docs if t in d) for t in range(n_terms)]
    out = []
    for i, d in enumerate(docs):
        dl = O.fieldnorm_id_to_value(O.fieldnorm_to_id(len(d)))          # the 1-byte quantised length is what tantivy scores with
        hits = [t for t in query if t < n_terms and t in d]
        if not hits or (mode == O.BM25_AND and len(hits) < len(query)):
            continue
        s = 0.0
        for t in hits:
            tf = d.count(t) if use_tf else 1
            idf = math.log(1.0 + (n - df[t] + 0.5) / (df[t] + 0.5))
            s += idf * (1.0 + k1) * tf / (tf + k1 * (1.0 - b + b * dl / avg))
        out.append((s, i))
    out.sort(key=lambda x: (-x[0], x[1]))
    return out


def test_bm25_oracle_equals_a_naive_restatement():
    rng = np.random.default_rng(31)
    n_terms = 40
    docs = [list(rng.integers(0, n_terms, int(rng.integers(1, 60)))) for _ in range(300)]
    doc_off = np.concatenate([[0], np.cumsum([len(d) for d in docs])])
    P = O.Postings(doc_off, np.concatenate(docs).astype(np.uint32), n_terms)
    for mode, use_tf, nt in ((O.BM25_OR, True, 4), (O.BM25_OR, False, 4), (O.BM25_AND, True, 2)):
        for _ in range(6):
            query = [int(t) for t in rng.choice(n_terms, nt, replace=False)]
            want = naive_bm25(docs, n_terms, query, mode, use_tf)
            got_docs, got_sc, got_cnt, got_total = O.bm25_search(P, [query], 25, mode=mode, use_tf=use_tf)
            assert got_total[0] == len(want) and got_cnt[0] == min(25, len(want))
            c = int(got_cnt[0])
            assert np.allclose(got_sc[0, :c], [s for s, _ in want[:c]], rtol=2e-6, atol=1e-6)
            for j in range(c):                                    # same document wherever the naive scores are clearly apart
                apart = (j == 0 or want[j - 1][0] - want[j][0] > 1e-5) and (j + 1 >= len(want) or want[j][0] - want[j + 1][0] > 1e-5)
                if apart:
                    assert got_docs[0, j] == want[j][1]


def test_brute_force_oracle_equals_float64_argsort():
    v = make_vectors(800, 96, seed=33)
    v[7] = 0.0                                                    # simsimd: both norms 0 -> distance 0 only if the query is 0 too
    q = make_queries(v, 12, seed=34)
    for sim in (O.SIM_DOT, O.SIM_COSINE):
        exact = q.astype(np.float64) @ v.astype(np.float64).T
        if sim == O.SIM_COSINE:
            exact /= np.maximum(np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(v.astype(np.float64), axis=1)[None, :], 1e-300)
        ids, sc, cnt = O.brute_force(v, q, 10, sim=sim, min_score=-1.0)
        order = np.argsort(-exact, axis=1, kind="stable")[:, :10]
        assert (cnt == 10).all()
        assert np.abs(sc - np.take_along_axis(exact, order, axis=1)).max() < 2e-6
        gaps = np.abs(np.diff(np.take_along_axis(exact, np.argsort(-exact, axis=1)[:, :11], axis=1), axis=1)) > 1e-5
        clear = np.concatenate([np.ones((len(q), 1), bool), gaps[:, :9]], axis=1) & gaps[:, :10]
        assert (ids[clear] == order[clear]).all()


def test_built_graph_invariants():
    """hnsw/params.rs:24-31 and build.rs:104-119: degree caps per layer, links stay inside their layer, rows are left packed,
    weights are the similarities, no duplicates and no self links -- except on the entry point: `insert` starts every search at
    the entry point, so the entry point finds ITSELF (similarity 1), links to itself and gets the reverse link too
    (build.rs:104-119 has no self check).  The reference does the same; the visited set makes it harmless."""
    v = make_vectors(1500, 48, seed=35)
    g = O.hnsw_build(v, M=6, M0=12, efC=30, max_batch=16, nthreads=4)
    assert (g.level == O.assign_levels(len(v), 6, 2)).all() and g.level[g.entry_node] == g.entry_layer == g.level.max()
    for node in range(len(v)):
        for layer in range(int(g.level[node]) + 1):
            row = g.adj0[node] if layer == 0 else g.adjU[int(g.upper_off[node]) + layer - 1]
            w = g.w0[node] if layer == 0 else g.wU[int(g.upper_off[node]) + layer - 1]
            k = int((row != O.NIL).sum())
            assert (row[:k] != O.NIL).all() and (row[k:] == O.NIL).all()
            assert k <= (12 if layer == 0 else 6)
            if node == g.entry_node:
                assert row[:k].tolist().count(node) == 2 and len(set(row[:k].tolist())) == k - 1
            else:
                assert node not in row[:k] and len(set(row[:k].tolist())) == k
            assert (g.level[row[:k]] >= layer).all()
            for t, wt in zip(row[:k], w[:k]):
                assert abs(float(wt) - O.cosine(v[node], v[int(t)])) < 1e-6
    deg0 = (g.adj0 != O.NIL).sum(1)
    assert deg0.min() >= 1                                                          # nobody is left without a link
