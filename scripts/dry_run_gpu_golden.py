"""Runs the GPU half of the golden-fixture tests (tests/test_gpu_zz_golden.py) on CPU against a stand-in of the device API that
is backed by the oracle: checks the TEST code (keys, shapes, call conventions), not the kernels.  python scripts/dry_run_gpu_golden.py"""
import os
import sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['NIDX_B200_UNVERIFIED_GPU_TESTS'] = '1'
import numpy as np
import oracle as O
from nucliadb_b200 import _lib
import nucliadb_b200.segment as S

class FakeVec:
    def __init__(s, v, d, sim, m, m0, efc):
        s.v, s.d, s.sim, s.m, s.m0, s.efc, s.g, s.enc = v, d, sim, m, m0, efc, None, None
    @classmethod
    def create(cls, v, d, similarity=_lib.NIDX_SIM_COSINE, m=30, m0=60, ef_construction=100, **kw):
        return cls(np.asarray(v, np.float32), d, similarity, m, m0, ef_construction)
    def _osim(s): return O.SIM_COSINE if s.sim == _lib.NIDX_SIM_COSINE else O.SIM_DOT
    def search(s, q, k, ef=0, min_score=-1.0, with_duplicates=True, method=0, filter_bits=None, **kw):
        if method == _lib.NIDX_METHOD_BRUTE:
            return O.brute_force(s.v, q, k, sim=s._osim(), min_score=min_score)
        if method == _lib.NIDX_METHOD_BRUTE_RABITQ:
            return O.rabitq_brute_force(s.v, s.enc, q, k, min_score=min_score)[:3]
        return O.hnsw_search(s.v, s.g, q, k, ef, sim=s._osim(), min_score=min_score, with_duplicates=with_duplicates, filter_bits=filter_bits)[:3]
    def set_graph(s, level, adj0, adjU, w0=None, wU=None):
        g = O.Graph(len(s.v), s.m, s.m0, np.asarray(level))
        g.adj0[:] = adj0; g.w0[:] = w0
        r = min(len(adjU), len(g.adjU)); g.adjU[:r] = adjU[:r]; g.wU[:r] = wU[:r]
        s.g = g
    def build_hnsw(s, seed=2, max_batch=4096):
        s.g = O.hnsw_build(s.v, sim=s._osim(), M=s.m, M0=s.m0, efC=s.efc, seed=seed, max_batch=max_batch)
    def get_graph(s):
        g = s.g
        return dict(level=g.level, adj0=g.adj0, w0=g.w0, adjU=g.adjU, wU=g.wU, entry_node=g.entry_node, entry_layer=g.entry_layer)
    def rabitq_encode(s): s.enc = O.rabitq_encode(s.v)
    def rabitq_codes(s): return s.enc
    def rabitq_estimate(s, q): return O.rabitq_estimate(s.enc, s.d, q)

class FakeTxt:
    @classmethod
    def create(cls, n_docs, n_terms, term_off, post_doc, post_tf, fieldnorm_id, device=0):
        o = cls(); o.args = (n_docs, n_terms, term_off, post_doc, post_tf, fieldnorm_id); return o
    def set_stats(s, *a): pass
    def search(s, qt, qoff, k, mode=0, use_tf=True, min_score=0.0, **kw):
        queries = [list(qt[qoff[i]:qoff[i+1]]) for i in range(len(qoff)-1)]
        return O.bm25_search(s.P, queries, k, mode=mode, use_tf=use_tf)

S.VectorSegment = FakeVec
S.TextSegment = FakeTxt
import test_gpu_zz_golden as T
import test_golden_fixtures as TF
FakeTxt.P = TF.postings_of(TF.load("bm25_small.npz"))
T.test_cuda_path_reproduces_the_vector_fixture()
T.test_cuda_path_reproduces_the_bm25_and_rabitq_fixtures()
print("dry run of the gpu tests against an oracle-backed fake: ok")
