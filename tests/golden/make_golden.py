"""Generates the committed fixtures of tests/golden/ from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is Rust and cannot be built or imported here, so these are not outputs of the reference itself: they are the
oracle's outputs (oracle/, pinned to the reference's own known-answer tests by tests/test_oracle_golden.py) on small seeded
inputs, frozen so that
  * `-m "not gpu"`: a change of the oracle that moves any id, bit of a score or edge is caught (test_golden_fixtures.py), and
  * `-m gpu`: the CUDA path is compared with values that do not depend on what is compiled on the GPU box.
The INPUTS are stored too: the generators use BLAS, whose low bits differ between hosts."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
from conftest import make_queries, make_vectors  # noqa: E402


def pack_bits(mask):
    words = np.zeros((len(mask) + 63) // 64 * 8, dtype=np.uint8)
    pb = np.packbits(mask, bitorder="little")
    words[: len(pb)] = pb
    return words.view(np.uint64)


def vector_fixture():
    v = make_vectors(1200, 64, seed=11)
    v[700:705] = v[10:15]                                  # exact duplicates for the with_duplicates = false case
    q = np.concatenate([make_queries(v, 20, seed=3), v[10:14]])
    out = dict(vectors=v, queries=q)
    for name, sim in (("cos", O.SIM_COSINE), ("dot", O.SIM_DOT)):
        ids, sc, cnt = O.brute_force(v, q, 10, sim=sim, min_score=-1.0)
        out[f"bf_{name}_ids"], out[f"bf_{name}_scores"], out[f"bf_{name}_counts"] = ids, sc, cnt
    g = O.hnsw_build(v, M=8, M0=16, efC=40, seed=2, max_batch=32)
    rows = int(g.level.astype(np.int64).sum())
    out.update(level=g.level, adj0=g.adj0, w0=g.w0, adjU=g.adjU[: max(rows, 1)], wU=g.wU[: max(rows, 1)],
               entry=np.asarray([g.entry_node, g.entry_layer], dtype=np.uint32), build=np.asarray([8, 16, 40, 2, 32], dtype=np.int32))
    ids, sc, cnt, counters = O.hnsw_search(v, g, q, 10, 40)
    out.update(hnsw_ids=ids, hnsw_scores=sc, hnsw_counts=cnt, hnsw_counters=counters)
    keep = np.random.default_rng(3).random(len(v)) < 0.4
    bits = pack_bits(keep)
    ids, sc, cnt, _ = O.hnsw_search(v, g, q, 10, 40, min_score=0.0, with_duplicates=False, filter_bits=bits)
    out.update(filter_bits=bits, filt_ids=ids, filt_scores=sc, filt_counts=cnt)
    np.savez_compressed(os.path.join(HERE, "vector_small.npz"), **out)


def bm25_fixture():
    rng = np.random.default_rng(17)
    n_docs, n_terms = 4000, 300
    lens = np.maximum(1, rng.lognormal(np.log(30), 0.6, n_docs).astype(np.int64))
    doc_off = np.concatenate([[0], np.cumsum(lens)])
    tokens = ((rng.zipf(1.2, int(doc_off[-1])) - 1) % n_terms).astype(np.uint32)
    P = O.Postings(doc_off, tokens, n_terms)
    out = dict(doc_off=doc_off.astype(np.int64), tokens=tokens, n_terms=np.asarray([n_terms], dtype=np.int32))
    cases = (("or_tf1", O.BM25_OR, False, 8), ("or_tf", O.BM25_OR, True, 8), ("and_tf", O.BM25_AND, True, 2))
    for name, mode, use_tf, nt in cases:
        queries = np.stack([rng.choice(60, nt, replace=False) + 5 for _ in range(12)]).astype(np.uint32)
        d, s, c, tot = O.bm25_search(P, [list(x) for x in queries], 20, mode=mode, use_tf=use_tf)
        out.update({f"{name}_queries": queries, f"{name}_docs": d, f"{name}_scores": s, f"{name}_counts": c, f"{name}_total": tot})
    np.savez_compressed(os.path.join(HERE, "bm25_small.npz"), **out)


def rabitq_fixture():
    v = make_vectors(400, 128, seed=23)
    v[5, :7] = 0.0
    q = make_queries(v, 6, seed=4)
    enc = O.rabitq_encode(v)
    est, err = O.rabitq_estimate(enc, 128, q)
    ids, sc, cnt, evals = O.rabitq_brute_force(v, enc, q, 10, min_score=0.0)
    np.savez_compressed(os.path.join(HERE, "rabitq_small.npz"), vectors=v, queries=q, codes=enc, estimate=est, error=err, scan_ids=ids,
                        scan_scores=sc, scan_counts=cnt, scan_exact_evals=evals)


if __name__ == "__main__":
    vector_fixture()
    bm25_fixture()
    rabitq_fixture()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
