"""The oracle pinned to the reference's own known-answer tests.  Every test names the
reference test it restates.  CPU only."""
import struct

import numpy as np

import oracle as O
from oracle import disk_v2


# ---- nidx_vector/src/vector_types/dense_f32.rs:52-84 ---------------------------------------------------
def test_cosine_test():
    v0 = np.arange(758, dtype=np.float32) * 2.0
    v1 = np.arange(758, dtype=np.float32) + 1.0
    naive = lambda a, b: float(np.dot(a, b) / (np.sqrt(np.dot(a, a)) * np.sqrt(np.dot(b, b))))
    assert abs(naive(v0, v0) - O.cosine(v0, v0)) < 0.01
    assert abs(naive(v0, v1) - O.cosine(v0, v1)) < 0.01


def test_dot_test():
    v0 = np.arange(758, dtype=np.float32) * 0.002
    v1 = np.arange(758, dtype=np.float32) * 0.002 + 0.05
    assert abs(float(np.dot(v0, v0)) - O.dot(v0, v0)) < 0.01
    assert abs(float(np.dot(v0, v1)) - O.dot(v0, v1)) < 0.01


def test_ordered_dot_is_within_1e6_of_exact():
    rng = np.random.default_rng(0)
    for d in (3, 64, 100, 384, 768, 1536, 2048):
        a = rng.standard_normal(d).astype(np.float32)
        b = rng.standard_normal(d).astype(np.float32)
        a /= np.linalg.norm(a)
        b /= np.linalg.norm(b)
        assert abs(O.dot(a, b) - O.dot_f64(a, b)) < 1e-6


def test_simsimd_edge_cases():
    z = np.zeros(8, dtype=np.float32)
    e0 = np.eye(8, dtype=np.float32)[0]
    e1 = np.eye(8, dtype=np.float32)[1]
    assert O.cosine(z, z) == 1.0      # both norms 0 -> distance 0
    assert O.cosine(e0, e1) == 0.0    # ab == 0 -> distance 1
    assert O.cosine(e0, e0) == 1.0
    assert O.cosine(e0 * 3, e0 * 0.5) <= 1.0  # distance clamped at 0


# ---- nidx_vector/src/utils.rs:140-155 -------------------------------------------------------------------
def test_vector_normalization():
    assert O.normalize(np.zeros(0, np.float32)).size == 0
    assert O.normalize([3.0, 0.0, 4.0, 0.0]).tolist() == [np.float32(3.0) / np.float32(5.0), 0.0, np.float32(4.0) / np.float32(5.0), 0.0]
    assert O.normalize([-1.0, -1.0, 0.0, 1.0, 1.0]).tolist() == [-0.5, -0.5, 0.0, 0.5, 0.5]
    big = O.normalize(np.full(10000, 100.0, np.float32))
    assert big[0] == np.float32(0.01) and (big == np.float32(0.01)).all()


# ---- nidx_vector/tests/test_basic_search.rs:39-143 -------------------------------------------------------
def _one_hot(dim=64):
    return np.eye(dim, dtype=np.float32)


def _check_basic(search):
    ids, sc = search(_one_hot()[5])
    assert len(ids) == 10 and ids[0] == 5 and sc[0] > 0.9999 and sc[1] < 0.0001
    q = np.zeros(64, np.float32)
    q[42], q[43], q[44], q[45] = 0.7, 0.59, 0.35, 0.2
    ids, sc = search(q)
    assert len(ids) == 10
    assert list(ids[:4]) == [42, 43, 44, 45]
    assert sc[0] > 0.6 and sc[1] > 0.5 and sc[2] > 0.3 and sc[3] > 0.15
    assert sc[5] == 0.0


def test_basic_search_brute_force():
    v = _one_hot()
    for sim in (O.SIM_DOT, O.SIM_COSINE):
        def search(q, sim=sim):
            i, s, c = O.brute_force(v, q, 10, sim=sim, min_score=-1.0)
            return i[0, : c[0]], s[0, : c[0]]
        _check_basic(search)


def test_basic_search_hnsw_reference_constants():
    v = _one_hot()
    for sim in (O.SIM_DOT, O.SIM_COSINE):
        g = O.hnsw_build(v, sim=sim, M=30, M0=60, efC=100)   # params.rs:34-46
        def search(q, sim=sim, g=g):
            i, s, c, _ = O.hnsw_search(v, g, q, 10, 30, sim=sim, min_score=-1.0, with_duplicates=False)
            return i[0, : c[0]], s[0, : c[0]]
        _check_basic(search)


# ---- nidx_vector/tests/test_min_score.rs + request_types.rs Default ----------------------------------------
def test_min_score_default_drops_negative_similarities():
    v = np.stack([np.eye(4, dtype=np.float32)[0], -np.eye(4, dtype=np.float32)[0], np.eye(4, dtype=np.float32)[1]])
    q = np.eye(4, dtype=np.float32)[0]
    i, s, c = O.brute_force(v, q, 10, sim=O.SIM_DOT, min_score=0.0)
    assert c[0] == 2 and set(i[0, :2]) == {0, 2}          # >= 0.0 keeps the orthogonal one (segment.rs:594 uses >=)
    i, s, c = O.brute_force(v, q, 10, sim=O.SIM_DOT, min_score=-1.0)
    assert c[0] == 3


# ---- nidx_vector/src/segment.rs:626-660 (100k unfiltered => cost 508 vs 100 000) ------------------
def test_use_hnsw_cost_model():
    assert O.use_hnsw(100_000, 100_000, 10)
    assert not O.use_hnsw(100_000, 50, 10)          # very selective filter => brute force
    assert not O.use_hnsw(100, 100, 10)             # tiny segment: 10*30*100/100 = 300 > 100


# ---- nidx_vector/src/hnsw/params.rs + build.rs:97-101 ---------------------------------------------
def test_level_distribution_round_not_floor():
    lv = O.assign_levels(400_000, 30, 2)
    assert abs((lv >= 1).mean() - 30 ** -0.5) < 0.003       # P(level >= 1) = M^-0.5 = 18.3 %
    assert abs((lv >= 2).mean() - 30 ** -1.5) < 0.001
    assert O.lib().oracle_prune_m(60) == 57 and O.lib().oracle_prune_m(30) == 28   # params.rs:29-31


# ---- nidx_vector/src/segment.rs:841-912 test_recall_clustered_data ---------------------------------------------
def test_recall_clustered_data():
    rng = np.random.default_rng(1234567890)
    D = 256

    def random_vector():
        v = rng.uniform(-1, 1, D).astype(np.float32)
        return v / np.sqrt((v * v).sum())

    def nearby(base, dist):
        v = base + random_vector() * np.float32(dist)
        return (v / np.sqrt((v * v).sum())).astype(np.float32)

    elems, center = [], random_vector()
    for _ in range(4):
        elems += [nearby(center, 0.01) for _ in range(80)]
        elems += [nearby(center, 0.03) for _ in range(80)]
        center = nearby(center, 0.1)
    v = np.stack(elems)
    g = O.hnsw_build(v, sim=O.SIM_DOT, M=30, M0=60, efC=100)
    queries = np.stack([nearby(v[rng.integers(0, len(v))], 0.05) for _ in range(100)])
    bi, _, _ = O.brute_force(v, queries, 5, sim=O.SIM_DOT, min_score=0.0)
    hi, _, _, _ = O.hnsw_search(v, g, queries, 5, 30, sim=O.SIM_DOT, min_score=0.0, with_duplicates=False)
    recall = np.mean([len(set(a) & set(b)) / 5 for a, b in zip(hi, bi)])
    assert recall >= 0.95      # "Expected ~0.98", asserted >= 0.95 in the reference


def test_recall_clustered_data_takes_the_rabitq_path():
    """The same reference test, the way the reference actually runs it: for_paragraphs() is Dot and DIMENSION = 256 is a
    multiple of 64, so the segment has vectors.quant and `_search` builds a RaBitQ query (segment.rs:506-513); with 640
    vectors use_hnsw(.., has_rabitq = true) picks the quantised exact scan + rerank (segment.rs:581-608)."""
    rng = np.random.default_rng(1234567890)
    d = 256

    def random_vector():
        v = rng.uniform(-1.0, 1.0, d).astype(np.float32)
        return (v / np.sqrt((v * v).sum())).astype(np.float32)

    def nearby(base, dist):
        v = base + random_vector() * np.float32(dist)
        return (v / np.sqrt((v * v).sum())).astype(np.float32)

    elems, center = [], random_vector()
    for _ in range(4):
        elems += [nearby(center, 0.01) for _ in range(80)]
        elems += [nearby(center, 0.03) for _ in range(80)]
        center = nearby(center, 0.1)
    v = np.stack(elems)
    assert not O.use_hnsw(len(v), len(v), 5, has_rabitq=True) and O.use_hnsw(len(v), len(v), 5, has_rabitq=False)
    enc = O.rabitq_encode(v)
    queries = np.stack([nearby(v[rng.integers(0, len(v))], 0.05) for _ in range(100)])
    bi, _, _ = O.brute_force(v, queries, 5, sim=O.SIM_DOT, min_score=0.0)
    ri, _, _, evals = O.rabitq_brute_force(v, enc, queries, 5, min_score=0.0)
    assert np.mean([len(set(a) & set(b)) / 5 for a, b in zip(ri, bi)]) >= 0.95
    assert (evals <= len(v)).all()     # candidates arrive in address order (no sort before rerank_top, segment.rs:600-608): on
                                       # clusters this tight every upper bound beats the running k-th score, nothing is spared


def test_hnsw_search_with_a_rabitq_query():
    """hnsw/search.rs:306-383, RaBitQ branch (taken from ~100 k vectors up, segment.rs:626-660): walk on estimates with
    min(k * 100, 2000) results at layer 0, exact rerank, exact closest_up_nodes.  The scores that come out are exact dots."""
    from conftest import make_queries, make_vectors

    v = make_vectors(6000, 128, seed=71)
    g = O.hnsw_build(v, sim=O.SIM_DOT, M=16, M0=32, efC=100, max_batch=128, nthreads=4)
    enc = O.rabitq_encode(v, nthreads=4)
    q = make_queries(v, 64, seed=5)
    bi, bs, _ = O.brute_force(v, q, 10, sim=O.SIM_DOT, min_score=0.0, nthreads=4)
    ri, rs, rc, counters = O.hnsw_search_rabitq(v, enc, g, q, 10, min_score=0.0, nthreads=4)
    assert (rc == 10).all()
    assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ri, bi)]) >= 0.99
    same = ri == bi
    assert np.array_equal(rs[same], bs[same])              # reranked with the raw vectors: no estimate leaks out
    assert (np.diff(rs, axis=1) <= 0).all()
    assert counters[3] > 5 * counters[0]                   # far more 24-byte estimates than 512-byte exact similarities
    # a filter is honoured by closest_up_nodes (exact phase)
    bits = np.zeros((len(v) + 63) // 64, dtype=np.uint64)
    keep = np.arange(0, len(v), 3)
    np.bitwise_or.at(bits, keep // 64, np.uint64(1) << (keep % 64).astype(np.uint64))
    fi, _, fc, _ = O.hnsw_search_rabitq(v, enc, g, q[:8], 5, min_score=0.0, filter_bits=bits, nthreads=2)
    assert all((fi[i, : fc[i]] % 3 == 0).all() for i in range(8))


# ---- nidx_vector/src/searcher.rs:600-688 test_vectors_deduplication + Fssc -------------------------------------
def test_fssc_dedup_and_replacement():
    vb = np.asarray([1.0, 2.0, 3.0], np.float32).tobytes()
    f = O.Fssc(20, with_duplicates=True)
    f.add("r/f/field/0-100", 32.0, 0, 0, vb)
    f.add("r/f/field/100-200", 32.0, 0, 1, vb)
    assert len(f.result()[2]) == 2
    f = O.Fssc(20, with_duplicates=False)
    f.add("r/f/field/0-100", 32.0, 0, 0, vb)
    f.add("r/f/field/100-200", 32.0, 0, 1, vb)
    assert len(f.result()[2]) == 1
    # full: replaces the smallest element that is smaller than the candidate (searcher.rs:184-195)
    f = O.Fssc(2, with_duplicates=True)
    f.add("a", 0.5, 0, 0, b"a")
    f.add("b", 0.7, 0, 1, b"b")
    f.add("c", 0.6, 0, 2, b"c")
    f.add("d", 0.1, 0, 3, b"d")
    seg, addr, sc = f.result()
    assert list(addr) == [1, 2] and list(sc) == [np.float32(0.7), np.float32(0.6)]


# ---- nidx_vector/src/hnsw/ram_hnsw.rs:173-198 test_fix_broken_links -------------------------------------------------
def test_fix_broken_links():
    g = O.Graph(2, 30, 60, np.array([1, 0], dtype=np.uint8))       # node 0 lives in layers 0-1, node 1 in layer 0 only
    g.adj0[0, 0], g.adj0[1, 0] = 1, 0
    g.w0[0, 0] = g.w0[1, 0] = 0.5
    g.adjU[0, 0], g.wU[0, 0] = 1, 0.5                              # broken: layer-1 link to a node that is not in layer 1
    assert O.fix_broken_links(g) == 1
    assert (g.adjU[0] == O.NIL).all() and g.adj0[0, 0] == 1 and g.adj0[1, 0] == 0


# ---- merge with graph reuse (segment.rs:143-167) when the new vectors raise the top layer ------------------------------
def test_extend_keeps_the_reused_graph_reachable():
    """The reference moves the entry point to the new, still unlinked top-layer node before inserting anything (build.rs:49-55);
    restated literally that cuts the reused graph off (recall 0.31 on this input).  The oracle and the CUDA path insert that
    node first from the old entry point instead -- the one deliberate deviation on this path."""
    from conftest import make_queries, make_vectors

    v = make_vectors(3000, 32, seed=63)
    n0 = 2000
    g0 = O.hnsw_build(v[:n0], M=8, M0=16, efC=40, max_batch=64, nthreads=4)
    seed = next(s for s in range(3, 500) if O.assign_levels(len(v) - n0, 8, s).max() > g0.entry_layer)
    g = O.hnsw_extend(v, g0, efC=40, seed=seed, max_batch=64, nthreads=4)
    assert g.entry_layer > g0.entry_layer and g.entry_node >= n0
    q = make_queries(v, 100)
    bi, _, _ = O.brute_force(v, q, 10, nthreads=4)
    hi, _, _, _ = O.hnsw_search(v, g, q, 10, 64, nthreads=4)
    assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(hi, bi)]) >= 0.97
    assert (hi < n0).mean() > 0.5                      # two thirds of the data are the reused segment


# ---- nidx_vector/tests/test_maxsim.rs:22-150 ------------------------------------------------------------------------
def test_maxsim_exact_scores():
    e = np.eye(5, dtype=np.float32)
    query = [e[0], e[3]]
    docs = [np.stack([e[1], e[2], e[4]]), np.stack([e[0], e[1], e[2]]), np.stack([e[0], e[2], e[3]])]      # d0 (0), d1 (1), d2 (2)
    assert O.multi_vector_search(docs, query, 1, -10.0) == [(2, 2.0)]
    assert O.multi_vector_search(docs, query, 10, 1.5) == [(2, 2.0)]          # min_score applies to the MaxSim score only
    assert O.multi_vector_search(docs, query, 2, -10.0) == [(2, 2.0), (1, 1.0)]
    assert O.maxsim_similarity(query, docs[0]) == 0.0                          # negative / zero similarities never add


# ---- nidx_vector/src/hnsw/disk/v2.rs:16-49 worked example + 349-398 hnsw_test ----------------------------------
def test_disk_v2_worked_example_bytes():
    layers = [{0: [(1, 0.1), (17, 0.2), (5433, 0.3), (45, 0.4), (667, 0.5)]}, {0: [(45, 1.0), (666, 2.0), (22, 3.0)]}, {}]
    graph, edges = disk_v2.serialize_graph(layers, 1, entry_node=0, entry_layer=2)
    words = struct.unpack(f"<{len(graph) // 4}I", graph)
    # the node of the format doc: 5 1 17 5433 45 667 | 3 45 666 22 | 0 | 16 32 <layer-0 offset> | node end | layer node.
    # The doc comment prints 52 for the layer-0 offset, but serialize_node (v2.rs:146-154) measures offsets from the
    # end of the node INCLUDING the offset table, which gives 56 -- and get_out_edges (v2.rs:159-174) only
    # round-trips with 56.  The code is the contract; the comment is off by one word.
    assert words == (5, 1, 17, 5433, 45, 667, 3, 45, 666, 22, 0, 16, 32, 56, 56, 2, 0)
    assert len(edges) == 8 * 4
    assert disk_v2.get_out_edges(graph, 0, 0) == [1, 17, 5433, 45, 667]
    assert disk_v2.get_out_edges(graph, 0, 1) == [45, 666, 22]
    assert disk_v2.get_out_edges(graph, 0, 2) == []
    assert disk_v2.entrypoint(graph) == (0, 2)


def test_disk_v2_hnsw_test():
    cnx0 = {0: [(1, 1.0)], 1: [(2, 2.0)], 2: [(3, 3.0)]}
    cnx1 = {0: [(1, 4.0)], 1: [(2, 5.0)]}
    cnx2 = {0: [(1, 6.0)]}
    graph, edges = disk_v2.serialize_graph([cnx0, cnx1, cnx2], 3, 0, 2)
    assert disk_v2.entrypoint(graph) == (0, 2)
    for layer, cnx in enumerate([cnx0, cnx1, cnx2]):
        for node in range(3):
            assert disk_v2.get_out_edges(graph, node, layer) == [t for t, _ in cnx.get(node, [])]
    assert disk_v2.serialize_graph([], 0, 0, 0) == (b"", b"")   # empty_hnsw


# ---- BM25 (tantivy; parity unpinned: only the published algorithm's fixed points) ---------------------------------
def test_fieldnorm_table_and_bm25_fixed_points():
    table = [O.fieldnorm_id_to_value(i) for i in range(256)]
    assert table[:41] == list(range(40)) + [40]
    assert table[40:49] == [40, 42, 44, 46, 48, 50, 52, 54, 56]
    assert table[56:58] == [88, 96] and table[255] == 2_013_265_944
    assert all(O.fieldnorm_to_id(v) == i for i, v in enumerate(table))
    assert O.fieldnorm_to_id(41) == 40 and O.fieldnorm_to_id(43) == 41
    assert abs(O.bm25_idf(1, 1) - np.log(1 + 0.5 / 1.5)) < 1e-7
    # one doc, one term, tf = fieldnorm = avg = 1: score = idf * 2.2 * 1 / (1 + 1.2)
    assert abs(O.bm25_term_score(1, 1, 1, 1, 1) - O.bm25_idf(1, 1) * 2.2 / 2.2) < 1e-6


def test_bm25_min_score_counts_like_reference_tests():
    # nidx_text/tests/test_search.rs:311-332: min_score 0 -> hits, min_score 100 -> none
    doc_off = [0, 4, 9, 12]
    tokens = [0, 1, 2, 3, 0, 0, 4, 5, 6, 7, 8, 9]
    P = O.Postings(doc_off, tokens, 10)
    docs, sc, cnt, total = O.bm25_search(P, [[0]], 10, mode=O.BM25_AND, use_tf=True)
    assert total[0] == 2 and cnt[0] == 2 and sc[0, 0] > 0 and sc[0, 0] < 100
    assert docs[0, 0] == 1          # tf = 2 in a 5-token doc beats tf = 1 in a 4-token doc
    docs, sc, cnt, total = O.bm25_search(P, [[0, 4]], 10, mode=O.BM25_AND, use_tf=True)
    assert list(docs[0, : cnt[0]]) == [1]
    docs, sc, cnt, total = O.bm25_search(P, [[0, 7]], 10, mode=O.BM25_OR, use_tf=False)
    assert total[0] == 3


# ---- nidx_vector/src/vector_types/rabitq.rs:284-306 test_rabitq_estimate + layout 70-106 ---------------------------
def test_rabitq_estimate_error_bound_and_layout():
    rng = np.random.default_rng(123)
    D = 2048

    def random_vector():
        v = rng.uniform(-1, 1, D).astype(np.float32)
        return (v / np.sqrt((v * v).sum())).astype(np.float32)

    v1 = random_vector()
    fuzz = random_vector()
    v2 = v1 + fuzz * np.float32(0.1)
    v2 = (v2 / np.sqrt((v2 * v2).sum())).astype(np.float32)
    v3 = random_vector()
    enc = O.rabitq_encode(v1[None, :])
    assert enc.shape == (1, D // 8 + 8)                                   # encoded_len
    dqo, sum_bits = struct.unpack_from("<fI", enc[0].tobytes(), 0)
    assert sum_bits == int((v1 > 0).sum())
    bits = np.unpackbits(enc[0, 8:], bitorder="little")
    assert (bits.astype(bool) == (v1 > 0)).all()                          # bit i of word i/64 = sign of element i
    assert abs(dqo - float(np.abs(v1).sum() / np.sqrt(D))) < 1e-4         # <v, sign(v)/sqrt(D)>
    for other in (v2, v3):                                                # high and low similarity
        est, err = O.rabitq_estimate(enc, D, other[None, :])
        actual = float(np.dot(v1.astype(np.float64), other.astype(np.float64)))
        assert abs(actual - float(est[0, 0])) < float(err[0, 0]) and float(err[0, 0]) < 0.05
    planes, low, delta, sq = O.rabitq_query(v2)
    wq = np.floor((v2 - np.float32(low)) / np.float32(delta)).astype(np.int64)
    assert wq.min() == 0 and wq.max() == 15 and sq == int(wq.sum())
    for p in range(4):
        assert (np.unpackbits(planes[p].view(np.uint8), bitorder="little").astype(np.int64) == ((wq >> p) & 1)).all()


def test_rabitq_brute_force_finds_the_exact_top_k():
    from conftest import make_queries, make_vectors
    v = make_vectors(4000, 256, seed=31)
    q = make_queries(v, 16)
    enc = O.rabitq_encode(v, nthreads=4)
    ids, sc, cnt, evals = O.rabitq_brute_force(v, enc, q, 10, min_score=0.0, nthreads=4)
    bi, bs, bc = O.brute_force(v, q, 10, sim=O.SIM_DOT, min_score=0.0, nthreads=4)
    assert (cnt == bc).all()
    assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ids, bi)]) >= 0.99     # the bound holds with ~97 % confidence per pair
    assert (evals < 4000).all() and evals.mean() < 1500                              # most raw vectors are never touched
