"""GPU parity: rank fusion on the device (nidx_rank_fusion_rrf) against the reference's own outputs (tests/golden/rank_fusion.json, made by
running the reference's ReciprocalRankFusion) and against the oracle on random batches; the fused shard search (nidx_shard_search)
against the three searches run one by one + the oracle's fusion."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from nucliadb_b200 import _lib
from nucliadb_b200.rank_fusion import ReciprocalRankFusion, shard_search
from nucliadb_b200.segment import TextSegment, VectorSegment
from oracle.rank_fusion import rrf_fuse
from test_rank_fusion import load_cases

pytestmark = pytest.mark.gpu


def test_device_rrf_reproduces_the_reference_outputs():
    for c in load_cases():
        algo = ReciprocalRankFusion(k=c["k"], window=100, weights=c["weights"])
        got = algo.fuse({name: [(int(i), float(s)) for i, s in c[name]] for name in c["order"]})
        single = sum(1 for name in c["order"] if c[name]) == 1
        want = c["fused"]
        assert [g[0] for g in got] == [w[0] for w in want]                       # keys in the reference's order (ties: first insertion)
        assert [g[2] for g in got] == [w[2] for w in want]                       # BM25 / VECTOR / BOTH
        if single:   # fusion skipped: the source's own scores, which cross the C ABI as f32
            assert [g[1] for g in got] == [float(np.float32(w[1])) for w in want]
        else:        # IEEE doubles, bit for bit
            assert [g[1] for g in got] == [w[1] for w in want]


def test_device_rrf_batch_matches_oracle():
    """A batch of queries with three sources, ragged counts, duplicate keys inside a source, device buffers."""
    import torch

    L = _lib.require_device()
    rng = np.random.default_rng(3)
    nq, ks, weights, kk = 257, [20, 7, 50], [1.0, 2.5, 0.3], 13.0
    keys = [rng.integers(0, 60, (nq, k)).astype(np.uint64) for k in ks]
    scores = [-np.sort(-rng.random((nq, k)).astype(np.float32), axis=1) for k in ks]
    counts = [rng.integers(0, k + 1, nq).astype(np.int32) for k in ks]
    counts[1][:5] = 0
    counts[0][3], counts[2][3] = 0, 0                        # query 3: a single source with results
    dk = [torch.from_numpy(a.view(np.int64)).cuda() for a in keys]
    ds = [torch.from_numpy(a).cuda() for a in scores]
    dc = [torch.from_numpy(a).cuda() for a in counts]
    src = (_lib.RrfSource * 3)(*[_lib.RrfSource(dk[i].data_ptr(), ds[i].data_ptr(), dc[i].data_ptr(), ks[i], weights[i]) for i in range(3)])
    cap = sum(ks)
    ok = torch.empty((nq, cap), dtype=torch.int64, device="cuda")
    osc = torch.empty((nq, cap), dtype=torch.float64, device="cuda")
    orf = torch.empty((nq, cap), dtype=torch.int32, device="cuda")
    ocn = torch.empty((nq,), dtype=torch.int32, device="cuda")
    _lib.check(L.nidx_rank_fusion_rrf(0, src, 3, nq, C.c_double(kk), _lib.NIDX_MEM_DEVICE, _lib.ptr(ok), _lib.ptr(osc), _lib.ptr(orf), _lib.ptr(ocn),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ok, osc, orf, ocn = ok.cpu().numpy().view(np.uint64), osc.cpu().numpy(), orf.cpu().numpy().view(np.uint32), ocn.cpu().numpy()
    for q in range(nq):
        srcs = [[(int(keys[i][q, j]), float(scores[i][q, j])) for j in range(counts[i][q])] for i in range(3)]
        want = rrf_fuse(srcs, weights, k=kk)
        n = int(ocn[q])
        assert n == len(want)
        assert ok[q, :n].tolist() == [w[0] for w in want]
        assert osc[q, :n].tolist() == [w[1] for w in want]
        assert [(int(r) >> 28, int(r) & 0xFFFFFF, (int(r) >> 24) & 0xF) for r in orf[q, :n]] == [(w[2], w[3], w[4]) for w in want]
        assert (ok[q, n:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()


def _text_corpus(rng, n_docs, n_terms):
    lens = rng.integers(5, 40, n_docs)
    doc_off = np.concatenate([[0], np.cumsum(lens)])
    tokens = (rng.zipf(1.3, doc_off[-1]) % n_terms).astype(np.uint32)
    return O.Postings(doc_off, tokens, n_terms)


def test_shard_search_equals_the_three_searches_and_the_oracle_fusion():
    """shard_search.rs:176-241: vector + paragraph + document searches of one batch in one call, fused on the device; every part equals
    the stand-alone entry point, the fused list equals the oracle's fusion of those parts (keys: paragraph p <-> document p)."""
    rng = np.random.default_rng(11)
    n, d, nq, kv, kp, kd = 3000, 64, 33, 10, 20, 5
    v = rng.standard_normal((n, d)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    q = v[rng.integers(0, n, nq)] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
    seg = VectorSegment.create(v, d, similarity=_lib.NIDX_SIM_COSINE, m=16, m0=32, ef_construction=64)
    seg.build_hnsw(seed=2, max_batch=256)
    keys = rng.permutation(n).astype(np.uint64) + 1000          # paragraph p of the vector index and document p of the paragraph index share keys[p]
    seg.set_paragraph_keys(keys)
    P = _text_corpus(rng, n, 200)
    par = TextSegment.create(P.n_docs, P.n_terms, P.term_off, P.post_doc, P.post_tf, P.fieldnorm_id)
    par.set_stats(P.n_docs, P.total_tokens, P.doc_freq)
    par.set_doc_keys(keys)
    D = _text_corpus(rng, 500, 120)
    doc = TextSegment.create(D.n_docs, D.n_terms, D.term_off, D.post_doc, D.post_tf, D.fieldnorm_id)
    doc.set_stats(D.n_docs, D.total_tokens, D.doc_freq)
    pq = [list(rng.integers(0, 200, 4)) for _ in range(nq)]
    dq = [list(rng.integers(0, 120, 2)) for _ in range(nq)]
    poff = np.concatenate([[0], np.cumsum([len(x) for x in pq])]).astype(np.uint32)
    doff = np.concatenate([[0], np.cumsum([len(x) for x in dq])]).astype(np.uint32)
    pterms, dterms = np.concatenate(pq).astype(np.uint32), np.concatenate(dq).astype(np.uint32)
    vp = _lib.VecSearchParams(kv, 64, -1.0, 1, _lib.NIDX_METHOD_HNSW, None, 0)
    pp = _lib.TxtSearchParams(kp, _lib.NIDX_BM25_OR, 0, 0.0, 0, 0.0, 0, 0)
    dp = _lib.TxtSearchParams(kd, _lib.NIDX_BM25_AND, 1, 0.0, 0, 0.0, 0, 0)
    for semantic_first in (False, True):
        r = shard_search(nq, vec=seg, queries=q, vec_params=vp, par=par, par_terms=pterms, par_off=poff, par_params=pp, doc=doc, doc_terms=dterms,
                         doc_off=doff, doc_params=dp, rrf_k=60.0, weight_keyword=1.0, weight_semantic=2.0, semantic_first=semantic_first)
        ids, sc, cnt = seg.search(q, kv, ef=64, method=_lib.NIDX_METHOD_HNSW, min_score=-1.0, with_duplicates=True)
        assert (r["vec_ids"] == ids).all() and np.array_equal(r["vec_scores"], sc) and (r["vec_counts"] == cnt).all()
        pd_, ps_, pc_, pt_ = par.search(pterms, poff, kp, mode=_lib.NIDX_BM25_OR, use_tf=False)
        assert (r["par_docs"] == pd_).all() and np.array_equal(r["par_scores"], ps_) and (r["par_counts"] == pc_).all() and (r["par_total"] == pt_).all()
        dd_, ds_, dc_, dt_ = doc.search(dterms, doff, kd, mode=_lib.NIDX_BM25_AND, use_tf=True)
        assert (r["doc_docs"] == dd_).all() and np.array_equal(r["doc_scores"], ds_) and (r["doc_counts"] == dc_).all() and (r["doc_total"] == dt_).all()
        overlaps = 0
        for i in range(nq):
            kw = [(int(keys[pd_[i, j]]), float(ps_[i, j])) for j in range(pc_[i])]
            sem = [(int(keys[ids[i, j]]), float(sc[i, j])) for j in range(cnt[i])]
            want = rrf_fuse([sem, kw] if semantic_first else [kw, sem], [2.0, 1.0] if semantic_first else [1.0, 2.0], k=60.0)
            m = int(r["fused_counts"][i])
            assert m == len(want)
            assert r["fused_keys"][i, :m].tolist() == [w[0] for w in want]
            assert r["fused_scores"][i, :m].tolist() == [w[1] for w in want]
            overlaps += sum(1 for w in want if w[4] == 3)
    # a request with only one kind of index, and one without fusion
    r = shard_search(nq, par=par, par_terms=pterms, par_off=poff, par_params=pp)
    assert (r["par_docs"] == pd_).all() and "vec_ids" not in r and "fused_keys" not in r
    r = shard_search(nq, vec=seg, queries=q, vec_params=vp, doc=doc, doc_terms=dterms, doc_off=doff, doc_params=dp)
    assert (r["vec_ids"] == ids).all() and (r["doc_docs"] == dd_).all() and "fused_keys" not in r
