"""OpenSegment.save / OpenSegment.open: a segment written in the reference's data-store-v2 layout (vectors.bin, hnsw.graph,
hnsw.edges through the library; paragraphs.bin / paragraphs.pos through paragraph_store.py) comes back with the same
paragraphs and answers searches identically (segment.rs:39-90 open, data_store/v2/*.rs)."""
import os

import numpy as np
import pytest

from nucliadb_b200 import vector as V

pytestmark = pytest.mark.gpu

RID = "9cb39c75f8d9498d8f82d92b173011f5"


def test_segment_directory_round_trip(tmp_path):
    rng = np.random.default_rng(8)
    cfg = V.VectorConfig(dimension=32, similarity=V.Similarity.Dot)
    elems = []
    for i in range(200):
        v = rng.standard_normal(32).astype(np.float32)
        v /= np.linalg.norm(v)
        elems.append(V.Elem(f"{RID}/f/field{i % 4}/{i}-{i + 1}", [v], labels=[f"/l/set/{i % 5}"] if i % 3 else [], metadata=bytes([i % 251]) if i % 2 else None))
    seg = V.VectorIndexer.index_elems(elems, cfg, tags={"/q/h"})
    seg.save(str(tmp_path))
    assert {"vectors.bin", "hnsw.graph", "hnsw.edges", "paragraphs.bin", "paragraphs.pos"} <= set(os.listdir(tmp_path))
    assert os.path.getsize(tmp_path / "vectors.bin") == 200 * (32 * 4 + 4) and os.path.getsize(tmp_path / "paragraphs.pos") == 200 * 4
    back = V.OpenSegment.open(cfg, str(tmp_path), tags={"/q/h"})
    assert back.keys == seg.keys and back.labels == seg.labels and back.metadata == seg.metadata
    assert (back.first_vec == seg.first_vec).all() and np.array_equal(back.host_vectors, seg.host_vectors)
    a, b = V.VectorSearcher.open(cfg, [(seg, 1)]), V.VectorSearcher.open(cfg, [(back, 1)])
    for i in (0, 17, 123):
        for formula in (None, V.Literal("/l/set/2"), V.Not(V.Literal("/l/set"))):
            req = V.VectorSearchRequest(vector=elems[i].vectors[0], result_per_page=7, min_score=-1.0, filtering_formula=formula)
            ra, rb = a.search(req), b.search(req)
            assert [(d.doc_id, d.score, d.labels, d.metadata) for d in ra.documents] == [(d.doc_id, d.score, d.labels, d.metadata) for d in rb.documents]
    back.apply_deletions([f"{RID}/f/field1"])
    r = b.search(V.VectorSearchRequest(vector=elems[1].vectors[0], result_per_page=200, min_score=-1.0))
    assert len(r.documents) == 150 and all("/f/field1/" not in d.doc_id for d in r.documents)
