"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's rank fusion
(nucliadb/src/nucliadb/search/search/rank_fusion.py):

    RankFusionAlgorithm.fuse        lines 78-96    one non-empty source -> that source as it is, else _fuse; stable sort by score desc
    ReciprocalRankFusion._fuse      lines 143-186  score(d) = sum over sources r, in the order given, of 1 / (k + rank_r(d)) * w(r)

Pinned: tests/golden/rank_fusion.json holds inputs and outputs of the reference's own class, run unmodified in the build container
(tests/golden/make_rank_fusion_golden.py); tests/test_rank_fusion.py checks this restatement against every case, scores bit for bit
(Python floats are IEEE doubles on both sides).

Items are (key, score) pairs; a fused item is (key, score, first_source, first_rank, source_mask): the reference keeps the item
object of a key's FIRST occurrence (`scores[id] = (RrfScore, item)`) and widens its score_type to BOTH when a BM25 and a VECTOR
item meet -- `source_mask` (bit i = source i contributed) carries the same information for any number of sources.
"""
from __future__ import annotations


def rrf_fuse(sources, weights, k=60.0):
    """sources: list of lists of (key, score), in fusion order; weights: one float per source."""
    non_empty = [i for i, s in enumerate(sources) if len(s) > 0]
    if len(non_empty) == 1:                                   # rank_fusion.py:86-89: fusion skipped, no de-duplication either
        i = non_empty[0]
        merged = [(key, float(score), i, r, 1 << i) for r, (key, score) in enumerate(sources[i])]
    else:
        ranked = [sorted(enumerate(s), key=lambda t: t[1][1], reverse=True) for s in sources]   # :151-154 (stable)
        acc, order = {}, []
        for i, ranking in enumerate(ranked):                                                      # :159-176
            for rank, (pos, (key, _score)) in enumerate(ranking):
                term = 1 / (k + rank) * weights[i]
                if key not in acc:
                    acc[key] = [term, i, pos, 1 << i]
                    order.append(key)
                else:
                    acc[key][0] += term
                    acc[key][3] |= 1 << i
        merged = [(key, acc[key][0], acc[key][1], acc[key][2], acc[key][3]) for key in order]     # :178-184 (dict order = first insertion)
    merged.sort(key=lambda t: t[1], reverse=True)                                                 # :92-93 (stable)
    return merged


def fused_score_type(first_type: str, types_of_contributors) -> str:
    """rank_fusion.py:166-174: the surviving item (a key's first occurrence) becomes BOTH when a BM25 and a VECTOR item meet; an
    item of any other type (RELATION_RELEVANCE) keeps its type whatever joins it."""
    if first_type in ("BM25", "VECTOR") and {"BM25", "VECTOR"} <= set(types_of_contributors):
        return "BOTH"
    return first_type
