"""The outer boundary: ``nidx_binding.NidxBinding`` (nidx_binding.pyi:15-71) and ``NidxSearcher.Search`` over gRPC
(nidx.proto:20-21, nodereader.proto:388-437 / 476-488).  A serialised ``nodewriter.IndexMessage`` goes in through ``index``,
a serialised ``nodereader.SearchRequest`` through a real gRPC channel to ``searcher_port``; the ``SearchResponse`` must carry
the vector, document and paragraph results the searchers of this package return (checked against numpy)."""
import uuid

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIM = 16


def _resource(P, shard, rid, texts, vec_rng, labels=()):
    res = P.Resource()
    res.resource.uuid, res.resource.shard_id, res.shard_id = rid, shard, shard
    res.labels.extend(labels)
    vectors = {}
    for fid, (text, sentences) in texts.items():
        res.texts[fid].text = text
        start = 0
        for i, sent in enumerate(sentences):
            end = start + len(sent)
            pid = f"{rid}/{fid}/{start}-{end}"
            par = res.paragraphs[fid].paragraphs[pid]
            par.start, par.end, par.field, par.index = start, end, fid, i
            par.labels.append(f"/k/par{i % 2}")
            v = vec_rng.standard_normal(DIM).astype(np.float32)
            v /= np.linalg.norm(v)
            s = par.vectorsets_sentences["en"].sentences[pid]
            s.vector.extend(v.tolist())
            s.metadata.position.index, s.metadata.position.start, s.metadata.position.end = i, start, end
            vectors[pid] = v
            start = end + 1
    return res, vectors


def test_index_sync_and_search_over_grpc(tmp_path):
    import grpc

    from nidx_binding import NidxBinding
    from nucliadb_b200 import nidx_protos as P

    binding = NidxBinding({"INDEXER__OBJECT_STORE": "file", "INDEXER__FILE_PATH": str(tmp_path), "METADATA__DATABASE_URL": "unused"})
    assert isinstance(binding.searcher_port, int) and binding.searcher_port > 0 and binding.api_port > 0
    # NidxApi.NewShard over gRPC (nidx.proto:9)
    api = grpc.insecure_channel(f"127.0.0.1:{binding.api_port}")
    req = P.NewShardRequest(kbid="kb")
    req.vectorsets_configs["en"].vector_dimension = DIM
    req.vectorsets_configs["en"].similarity = 1          # DOT
    shard = api.unary_unary(P.NEW_SHARD_METHOD, request_serializer=lambda m: m.SerializeToString(), response_deserializer=P.ShardCreated.FromString)(req).id
    rng = np.random.default_rng(5)
    rids = [uuid.UUID(int=i + 1).hex for i in range(3)]
    all_vectors = {}
    sentences = [["the quick brown fox", "jumps over the lazy dog"], ["graph search on hbm", "the fox likes postings"], ["nothing in common here"]]
    for i, rid in enumerate(rids):
        res, vecs = _resource(P, shard, rid, {"a/title": (" ".join(sentences[i]), sentences[i])}, rng, labels=[f"/l/r{i}"])
        all_vectors.update(vecs)
        key = f"index/{rid}"
        (tmp_path / "index").mkdir(exist_ok=True)
        (tmp_path / key).write_bytes(res.SerializeToString())
        seq = binding.index(P.IndexMessage(shard=shard, resource=rid, typemessage=0, storage_key=key, kbid="kb").SerializeToString())
        assert seq == i + 1
    binding.wait_for_sync()

    chan = grpc.insecure_channel(f"127.0.0.1:{binding.searcher_port}")
    search = chan.unary_unary(P.SEARCH_METHOD, request_serializer=lambda m: m.SerializeToString(), response_deserializer=P.SearchResponse.FromString)
    target = list(all_vectors)[2]
    sreq = P.SearchRequest(shard_ids=[shard], body="fox", vector=all_vectors[target].tolist(), vectorset="en", result_per_page=3, paragraph=True, document=True,
                           min_score_semantic=-1.0, with_duplicates=True)
    resp = search(sreq)
    assert list(resp.shard_ids) == [shard]
    # vectors: exact top-3 by dot product (3 tiny segments, cross-segment Fssc), metadata and labels travel
    want = sorted(((float(np.dot(v, all_vectors[target])), k) for k, v in all_vectors.items()), reverse=True)[:3]
    got = [(d.score, d.doc_id.id) for d in resp.vector.documents]
    assert [k for _, k in got] == [k for _, k in want] and all(abs(a - b) < 1e-5 for (a, _), (b, _) in zip(got, want))
    assert resp.vector.documents[0].doc_id.id == target and resp.vector.documents[0].metadata.position.end > 0
    assert any(l.startswith("/k/par") for l in resp.vector.documents[0].labels)
    # BM25: "fox" is in resources 0 and 1 (documents = fields, paragraphs = sentences)
    assert {r.uuid for r in resp.document.results} == {rids[0], rids[1]} and resp.document.total == 2
    assert {r.uuid for r in resp.paragraph.results} == {rids[0], rids[1]} and all(r.score.bm25 > 0 for r in resp.paragraph.results)
    # paragraph filter (labels) + field filter (prefilter on the text index)
    freq = P.SearchRequest(shard_ids=[shard], vector=all_vectors[target].tolist(), vectorset="en", result_per_page=10, min_score_semantic=-1.0, with_duplicates=True)
    freq.paragraph_filter.facet.facet = "/k/par0"
    r2 = search(freq)
    assert len(r2.vector.documents) == 3 and all("/k/par0" in d.labels for d in r2.vector.documents)
    freq.field_filter.facet.facet = "/l/r1"
    r3 = search(freq)
    assert [d.doc_id.id.split("/")[0] for d in r3.vector.documents] == [rids[1]]
    # deletion of a resource (IndexMessage DELETION) and re-sync
    binding.index(P.IndexMessage(shard=shard, resource=rids[1], typemessage=1, kbid="kb").SerializeToString())
    binding.wait_for_sync()
    r4 = search(sreq)
    assert all(not d.doc_id.id.startswith(rids[1]) for d in r4.vector.documents) and {r.uuid for r in r4.document.results} == {rids[0]}
    # errors: unknown shard -> NOT_FOUND, unknown vectorset -> INVALID_ARGUMENT (shard_search.rs:95-99)
    with pytest.raises(grpc.RpcError) as e:
        search(P.SearchRequest(shard_ids=["nope"], vector=[0.0] * DIM, vectorset="en", result_per_page=1))
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    with pytest.raises(grpc.RpcError) as e:
        search(P.SearchRequest(shard_ids=[shard], vector=[0.0] * DIM, vectorset="other", result_per_page=1))
    assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
    binding.close()
