"""Rank fusion: the oracle restatement against golden vectors produced by the reference's own Python class
(tests/golden/rank_fusion.json, generator tests/golden/make_rank_fusion_golden.py)."""
import json
import os

from oracle.rank_fusion import rrf_fuse

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rank_fusion.json")


def load_cases():
    return json.load(open(GOLD))["cases"]


def case_sources(c):
    srcs = [[(int(i), float(s)) for i, s in c[name]] for name in c["order"]]
    weights = [float(c["weights"][name]) for name in c["order"]]
    type_bits = [1 if name == "keyword" else 2 for name in c["order"]]
    return srcs, weights, type_bits


def score_type(mask, type_bits):
    t = 0
    for i, b in enumerate(type_bits):
        if mask >> i & 1:
            t |= b
    return {1: "BM25", 2: "VECTOR", 3: "BOTH"}[t]


def test_oracle_reproduces_the_reference_rank_fusion():
    cases = load_cases()
    assert len(cases) >= 40
    seen_single = seen_both = 0
    for c in cases:
        srcs, weights, type_bits = case_sources(c)
        got = rrf_fuse(srcs, weights, k=c["k"])
        assert [[key, sc, score_type(mask, type_bits)] for key, sc, _, _, mask in got] == c["fused"]    # ids, f64 scores bit for bit, types
        seen_single += sum(1 for s in srcs if s) == 1
        seen_both += any(t == "BOTH" for _, _, t in c["fused"])
    assert seen_single >= 2 and seen_both >= 10
