"""Rank fusion: the oracle restatement against golden vectors produced by the reference's own Python class
(tests/golden/rank_fusion.json, generator tests/golden/make_rank_fusion_golden.py)."""
import json
import os

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
