"""More of the reference's own tests restated against the Python mirror (nucliadb_b200.vector) on the GPU:
nidx_vector/src/segment/tests.rs, nidx_vector/src/searcher.rs (tests module), nidx_vector/tests/test_hidden.rs and
tests/test_paragraph_merge.rs.  Each test names the Rust test it follows."""
import uuid

import numpy as np
import pytest

from nucliadb_b200 import vector as V

pytestmark = pytest.mark.gpu

CONFIG = dict(dimension=128, similarity=V.Similarity.Cosine)      # segment/tests.rs:28-36
RNG = np.random.default_rng(20240923)


def create_query(d=128):  # segment/tests.rs:38-50: uniform(-1, 1), L2-normalised
    v = RNG.uniform(-1.0, 1.0, d).astype(np.float32)
    return (v / np.sqrt((v * v).sum(dtype=np.float32))).astype(np.float32)


def seg_search(seg, query, clauses=(), top_k=10, min_score=-1.0, with_duplicates=True):
    return seg.search(query, list(clauses), True, with_duplicates, top_k, min_score)


def test_simple_flow():  # segment/tests.rs:53-79
    labels = [f"LABEL_{i}" for i in range(50)]
    elems = [V.Elem(f"9cb39c75f8d9498d8f82d92b173011f5/f/field/0-{i}", [np.full(128, RNG.random(), np.float32)], labels) for i in range(50)]
    seg = V.VectorIndexer.index_elems(elems, V.VectorConfig(**CONFIG))
    addrs, _ = seg_search(seg, np.full(128, RNG.random(), np.float32), [V.Literal(l) for l in labels[:20]], top_k=10)
    assert len(addrs) == 10


def test_single_graph():  # segment/tests.rs:123-143
    key = "9cb39c75f8d9498d8f82d92b173011f5/f/field/0-100"
    vector = create_query()
    cfg = V.VectorConfig(**CONFIG)
    seg = V.VectorIndexer.index_elems([V.Elem(key, [vector])], cfg)
    seg.apply_deletions([key])
    assert len(seg_search(seg, vector, top_k=5)[0]) == 0
    seg = V.VectorIndexer.index_elems([V.Elem(key, [vector])], cfg)
    addrs, scores = seg_search(seg, vector, top_k=5)
    assert len(addrs) == 1 and scores[0] >= 0.9 and seg.keys[seg.paragraph_of(int(addrs[0]))] == key


def test_data_merge_v2():  # segment/tests.rs:197-243
    cfg = V.VectorConfig(**CONFIG)
    key0, key1 = "9cb39c75f8d9498d8f82d92b173011f5/f/field/0-100", "29ee1f6e4585423585f31ded0202ee3a/f/field/0-100"
    vector0, vector1 = create_query(), create_query()
    dp0 = V.VectorIndexer.index_elems([V.Elem(key0, [vector0])], cfg)
    dp1 = V.VectorIndexer.index_elems([V.Elem(key1, [vector1])], cfg)
    dp = V.VectorIndexer.merge(cfg, [(dp1, 1), (dp0, 1)])
    for vec, key in ((vector1, key1), (vector0, key0)):
        addrs, scores = seg_search(dp, vec, top_k=1)
        assert len(addrs) == 1 and scores[0] >= 0.9 and dp.keys[dp.paragraph_of(int(addrs[0]))] == key
    for s in (dp0, dp1):
        s.apply_deletions([key0])
        s.apply_deletions([key1])
    assert V.VectorIndexer.merge(cfg, [(dp1, 1), (dp0, 1)]).records == 0


def test_label_filtering():  # segment/tests.rs:296-337
    cfg = V.VectorConfig(**CONFIG)
    elems = [V.Elem(f"6e5a546a9a5c480f8579472016b1ee14/f/field/{i}-{i + 1}", [create_query()], [f"LABEL_{i}"]) for i in range(100)]
    seg = V.VectorIndexer.index_elems(elems, cfg)
    query = create_query()
    for i in range(5):
        assert len(seg_search(seg, query, [V.Literal(f"LABEL_{i}")])[0]) == 1      # LABEL_1 is not a prefix match of LABEL_10
    seg.apply_deletions(["6e5a546a9a5c480f8579472016b1ee14/f/field"])
    for i in range(5):
        assert len(seg_search(seg, query, [V.Literal(f"LABEL_{i}")])[0]) == 0


def test_label_prefix_search():  # segment/tests.rs:340-377
    cfg = V.VectorConfig(**CONFIG)
    elems = [V.Elem(f"6e5a546a9a5c480f8579472016b1ee14/f/field/{i}-{i + 1}", [create_query()], [f"/l/labelset/LABEL_{i}"]) for i in range(5)]
    seg = V.VectorIndexer.index_elems(elems, cfg)
    query = create_query()
    assert len(seg_search(seg, query, [V.Literal("/l/labelset")])[0]) == 5          # the labelset: everything
    assert len(seg_search(seg, query, [V.Literal("/l/labelset/LABEL_0")])[0]) == 1  # one label
    assert len(seg_search(seg, query, [V.Literal("/l/labelset/LABEL")])[0]) == 0    # a prefix of a label name: nothing


def test_fast_data_merge():  # segment/tests.rs:380-470 (without the wall-clock comparison)
    cfg = V.VectorConfig(**CONFIG)
    sv = [create_query() for _ in range(4)]
    big = [V.Elem(f"75a6eed3f94e456daa3f2d578a2254b7/t/trash/0-{k}", [create_query()]) for k in range(100)]
    big += [V.Elem("00000000000000000000000000000000/f/file/0-100", [sv[0]]), V.Elem("00000000000000000000000000000001/f/file/0-100", [sv[1]])]
    small = [V.Elem("00000000000000000000000000000002/f/file/0-100", [sv[2]]), V.Elem("00000000000000000000000000000003/f/file/0-100", [sv[3]])]
    big_segment, small_segment = V.VectorIndexer.index_elems(big, cfg), V.VectorIndexer.index_elems(small, cfg)
    dp = V.VectorIndexer.merge(cfg, [(big_segment, 1), (small_segment, 1)])          # no deletions: the big segment's graph is reused
    assert dp.records == 104
    for i, v in enumerate(sv):
        addrs, scores = seg_search(dp, v, top_k=1, min_score=0.999)
        assert len(addrs) == 1 and scores[0] >= 0.999
        assert dp.keys[dp.paragraph_of(int(addrs[0]))] == f"0000000000000000000000000000000{i}/f/file/0-100"
    big_segment.apply_deletions(["00000000000000000000000000000000/f/file/0-100"])
    small_segment.apply_deletions(["00000000000000000000000000000002/f/file/0-100"])
    dp = V.VectorIndexer.merge(cfg, [(big_segment, 1), (small_segment, 1)])          # deletions: full rebuild
    assert dp.records == 102
    for i, v in enumerate(sv):
        addrs, scores = seg_search(dp, v, top_k=1, min_score=0.999)
        if i in (0, 2):
            assert len(addrs) == 0
        else:
            assert len(addrs) == 1 and dp.keys[dp.paragraph_of(int(addrs[0]))] == f"0000000000000000000000000000000{i}/f/file/0-100"


def test_key_prefix_search():  # searcher.rs:411-498
    cfg = V.VectorConfig(dimension=3, similarity=V.Similarity.Dot)
    rid = "6c5fc1f7a69042d4b24b7f18ea354b4a"
    raw = [([1.0, 3.0, 4.0], 1), ([2.0, 4.0, 5.0], 2), ([3.0, 5.0, 6.0], 3), ([3.0, 5.0, 6.0], 4)]
    seg = V.VectorIndexer.index_elems([V.Elem(f"{rid}/f/field1/{i}", [v], ["/1"]) for v, i in raw], cfg)
    searcher = V.VectorSearcher.open(cfg, [(seg, 1)])
    request = V.VectorSearchRequest(vector=[4.0, 6.0, 7.0], result_per_page=20, with_duplicates=True)
    hit = V.PrefilterResult.some([V.FieldId(uuid.UUID(rid), "/f/field1")])
    assert len(searcher.search(request, hit).documents) == 4
    miss = V.PrefilterResult.some([V.FieldId(uuid.UUID(rid), "/f/field2")])
    assert len(searcher.search(request, miss).documents) == 0


def test_new_vector_reader():  # searcher.rs:501-606
    cfg = V.VectorConfig(dimension=3, similarity=V.Similarity.Cosine)
    rid = "9cb39c75f8d9498d8f82d92b173011f5"
    raw = [([1.0, 3.0, 4.0], "0-1"), ([2.0, 4.0, 5.0], "1-2"), ([3.0, 5.0, 6.0], "2-3"), ([3.0, 5.0, 6.0], "3-4")]
    seg = V.VectorIndexer.index_elems([V.Elem(f"{rid}/f/field/{r}", [v]) for v, r in raw], cfg)
    searcher = V.VectorSearcher.open(cfg, [(seg, 1)])
    n = lambda **kw: len(searcher.search(V.VectorSearchRequest(vector=[4.0, 6.0, 7.0], result_per_page=20, **kw)).documents)
    assert n(with_duplicates=True) == 4
    assert n(with_duplicates=False) == 3                   # the two identical vectors collapse
    assert n(with_duplicates=False, min_score=900.0) == 0
    with pytest.raises(V.NidxError):
        searcher.search(V.VectorSearchRequest(vector=[4.0, 6.0], result_per_page=20))


def test_hidden_search():  # tests/test_hidden.rs:24-79: segment tags + segment_filtering_formula
    cfg = V.VectorConfig(dimension=4, similarity=V.Similarity.Cosine)
    hidden_id, visible_id = uuid.uuid4().hex, uuid.uuid4().hex
    vec = lambda: np.asarray([0.5, 0.5, 0.5, RNG.random()], np.float32)
    hidden = V.VectorIndexer.index_elems([V.Elem(f"{hidden_id}/a/title/0-5", [vec()])], cfg, tags={"/q/h"})
    visible = V.VectorIndexer.index_elems([V.Elem(f"{visible_id}/a/title/0-5", [vec()])], cfg)
    searcher = V.VectorSearcher.open(cfg, [(hidden, 1), (visible, 2)])
    request = V.VectorSearchRequest(vector=[0.5, 0.5, 0.5, 0.5], min_score=-1.0, result_per_page=10)
    assert {d.doc_id for d in searcher.search(request).documents} == {f"{hidden_id}/a/title/0-5", f"{visible_id}/a/title/0-5"}
    request.segment_filtering_formula = V.Not(V.Literal("/q/h"))
    docs = searcher.search(request).documents
    assert [d.doc_id for d in docs] == [f"{visible_id}/a/title/0-5"]


def test_paragraph_merge_with_deletions():  # tests/test_paragraph_merge.rs:70-174
    cfg = V.VectorConfig(dimension=4, similarity=V.Similarity.Cosine)
    uuid1, uuid2 = "00112233445566778899aabbccddeeff", "ffeeddccbbaa99887766554433221100"
    axis = lambda i: np.eye(4, dtype=np.float32)[i]
    resource = lambda u, a, b: [V.Elem(f"{u}/a/title/0-10", [axis(a)]), V.Elem(f"{u}/t/body/0-10", [axis(b)])]
    seg1, seg2 = V.VectorIndexer.index_elems(resource(uuid1, 0, 1), cfg), V.VectorIndexer.index_elems(resource(uuid2, 2, 3), cfg)
    assert seg1.records == 2 and seg2.records == 2
    # every deletion carries its own segment's seq: it only removes what was indexed BEFORE that seq -- nothing here
    deletions = [(f"{uuid1}/a/title", 2), (f"{uuid1}/t/body", 2), (f"{uuid2}/a/title", 4), (f"{uuid2}/t/body", 4)]
    merged = V.VectorIndexer.merge(cfg, [(seg1, 2), (seg2, 4)], deletions=deletions)
    assert merged.records == 4
    searcher = V.VectorSearcher.open(cfg, [(merged, 5)])
    for i in (0, 2, 3):
        docs = searcher.search(V.VectorSearchRequest(vector=axis(i), result_per_page=10, with_duplicates=True)).documents
        assert len(docs) == 4 and docs[0].score > 0.9999
