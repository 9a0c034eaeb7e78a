"""Rank fusion: the oracle restatement against golden vectors produced by the reference's own Python class
(tests/golden/rank_fusion.json, generator tests/golden/make_rank_fusion_golden.py)."""
import json
import os

import numpy as np

from oracle.rank_fusion import fused_score_type, rrf_fuse

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rank_fusion.json")


def load_cases():
    return json.load(open(GOLD))["cases"]


TYPE_OF = {"keyword": "BM25", "semantic": "VECTOR", "graph": "RELATION_RELEVANCE"}


def case_sources(c):
    srcs = [[(int(i), float(s)) for i, s in c[name]] for name in c["order"]]
    weights = [float(c["weights"][name]) for name in c["order"]]
    types = [TYPE_OF[name] for name in c["order"]]
    return srcs, weights, types


def score_type(first, mask, types):
    return fused_score_type(types[first], [t for i, t in enumerate(types) if mask >> i & 1])


def test_oracle_reproduces_the_reference_rank_fusion():
    cases = load_cases()
    assert len(cases) >= 50 and sum(1 for c in cases if len(c["order"]) == 3) >= 20
    seen_single = seen_both = 0
    for c in cases:
        srcs, weights, types = case_sources(c)
        got = rrf_fuse(srcs, weights, k=c["k"])
        assert [[key, sc, score_type(first, mask, types)] for key, sc, first, _, mask in got] == c["fused"]    # ids, f64 scores bit for bit, types
        seen_single += sum(1 for s in srcs if s) == 1
        seen_both += any(t == "BOTH" for _, _, t in c["fused"])
    assert seen_single >= 2 and seen_both >= 10


def test_mirror_rank_fusion_on_the_emulated_abi(monkeypatch):
    """The host mirror (nucliadb_b200/rank_fusion.py: sorting of the sources, packing for nidx_rank_fusion_rrf, score-type rule) against
    the reference's outputs, on the oracle-backed emulation of the C ABI (no GPU: host-logic coverage only)."""
    import abi_emulator
    from nucliadb_b200 import _lib
    from nucliadb_b200.rank_fusion import ReciprocalRankFusion

    monkeypatch.setattr(_lib, "_lib", abi_emulator.EmulatedLib())
    for c in load_cases():
        got = ReciprocalRankFusion(k=c["k"], window=100, weights=c["weights"]).fuse({name: [(int(i), float(s)) for i, s in c[name]] for name in c["order"]})
        single = sum(1 for name in c["order"] if c[name]) == 1
        assert [[g[0], g[2]] for g in got] == [[w[0], w[2]] for w in c["fused"]]
        want_scores = [float(np.float32(w[1])) if single else w[1] for w in c["fused"]]      # a skipped fusion reports the source's f32 scores
        assert [g[1] for g in got] == want_scores
