"""The outer boundary of the hot path: the ``nidx_binding.NidxBinding`` Python surface and the ``NidxSearcher.Search`` gRPC
service, over the GPU searchers of this package.

Reference:
  nidx/nidx_binding/nidx_binding.pyi:15-71, src/lib.rs:53-127    NidxBinding(settings), index(bytes) -> seq, wait_for_sync(),
                                                                   searcher_port, api_port
  nidx/nidx_protos/nidx.proto:9,20-21                             NidxApi.NewShard, NidxSearcher.Search
  nidx/src/searcher/shard_search.rs:60-241                        one SearchRequest -> prefilter -> vector / paragraph / document searches
  nidx/src/searcher/shard_merge.rs:177-348                        merge of the per-shard responses
  nidx/nidx_vector/src/indexer.rs:96-146                          Resource -> vector Elems (key = sentence id, labels = paragraph labels)
  nidx/nidx_text/src/resource_indexer.rs:22-91                    Resource.texts -> one document per field
  nidx/nidx_paragraph/src/resource_indexer.rs:33-131              Resource.paragraphs -> one document per paragraph (text[start:end])

What is kept of the reference's machinery is the INTERFACE: metadata lives in memory (no PostgreSQL), every index message
becomes one immutable segment per index (as in the reference), deletions are (key, seq) pairs applied to older segments, and
"sync" re-opens the searchers (index_cache.rs:180-200).  Scheduler, worker, merges-in-the-background, NATS, object stores other
than the local file store, relations / JSON / graph / suggest are outside the hot path and answer UNIMPLEMENTED.
"""
from __future__ import annotations

import os
import threading
import uuid as _uuid
from concurrent import futures
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import nidx_protos as P
from . import text as T
from . import vector as V


@dataclass
class _VectorIndex:
    config: V.VectorConfig
    segments: list = field(default_factory=list)     # [(OpenSegment, seq)]
    deletions: list = field(default_factory=list)    # [(key prefix, seq)]
    searcher: Optional[V.VectorSearcher] = None


@dataclass
class _Shard:
    kbid: str
    vectorsets: dict = field(default_factory=dict)   # name -> _VectorIndex
    text_segments: list = field(default_factory=list)        # [[TextDoc]] one list per index message
    paragraph_segments: list = field(default_factory=list)   # [[TextDoc]] (+ paragraph positions in .field / extra)
    paragraph_meta: dict = field(default_factory=dict)       # paragraph id -> (field, start, end, index, split, labels, metadata bytes)
    deleted_resources: set = field(default_factory=set)
    text_searcher: Optional[T.TextSearcher] = None
    paragraph_searcher: Optional[T.ParagraphSearcher] = None


def _expr_to_boolean(e) -> Optional[V.BooleanExpression]:
    """nodereader.FilterExpression (paragraph_filter) -> BooleanExpression over labels (query_io::map_expression's input)."""
    kind = e.WhichOneof("expr")
    if kind == "facet":
        return V.Literal(e.facet.facet)
    if kind == "bool_not":
        inner = _expr_to_boolean(e.bool_not)
        return V.Not(inner) if inner is not None else None
    if kind in ("bool_and", "bool_or"):
        ops = [x for x in (_expr_to_boolean(o) for o in getattr(e, kind).operands) if x is not None]
        return V.Operation("and" if kind == "bool_and" else "or", tuple(ops)) if ops else None
    return None


def _doc_matches(e, doc: T.TextDoc) -> bool:
    """nodereader.FilterExpression (field_filter) evaluated on one text document: the prefilter of nidx_text (reader.rs:148-180)."""
    kind = e.WhichOneof("expr")
    if kind == "facet":
        return any(l == e.facet.facet or l.startswith(e.facet.facet + "/") for l in doc.labels)
    if kind == "resource":
        return doc.uuid == e.resource.resource_id
    if kind == "field":
        ft, fid = e.field.field_type, e.field.field_id if e.field.HasField("field_id") else None
        return doc.field.startswith(f"/{ft}/") and (fid is None or doc.field == f"/{ft}/{fid}")
    if kind == "bool_not":
        return not _doc_matches(e.bool_not, doc)
    if kind == "bool_and":
        return all(_doc_matches(o, doc) for o in e.bool_and.operands)
    if kind == "bool_or":
        return any(_doc_matches(o, doc) for o in e.bool_or.operands)
    return True


class NidxBinding:
    """nidx_binding.pyi:15-71.  ``settings`` mirrors the reference's environment schema; the keys read here are
    ``INDEXER__OBJECT_STORE`` (must be ``file``), ``INDEXER__FILE_PATH`` (where IndexMessage.storage_key points into) and
    ``NIDX_B200__DEVICE`` (CUDA ordinal, default 0)."""

    def __init__(self, settings: dict):
        import grpc

        settings = dict(settings)
        settings["INDEXER__NATS_SERVER"] = ""                      # lib.rs:73: always the in-process indexer
        self.settings = settings
        self.device = int(settings.get("NIDX_B200__DEVICE", "0"))
        self._lock = threading.RLock()
        self._shards: dict = {}
        self._seq = 1                                               # lib.rs:128-140: the sequence every index message consumes
        self._dirty = set()
        self._searcher = grpc.server(futures.ThreadPoolExecutor(max_workers=8))
        self._searcher.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("nidx.NidxSearcher", {
            "Search": grpc.unary_unary_rpc_method_handler(self._grpc_search, request_deserializer=P.SearchRequest.FromString,
                                                          response_serializer=lambda m: m.SerializeToString())}),))
        self.searcher_port = self._searcher.add_insecure_port("127.0.0.1:0")
        self._api = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
        self._api.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("nidx.NidxApi", {
            "NewShard": grpc.unary_unary_rpc_method_handler(self._grpc_new_shard, request_deserializer=P.NewShardRequest.FromString,
                                                            response_serializer=lambda m: m.SerializeToString())}),))
        self.api_port = self._api.add_insecure_port("127.0.0.1:0")
        self._searcher.start()
        self._api.start()

    # ---- NidxApi.NewShard (nidx.proto:9, grpc.rs) --------------------------------------------------------------------------
    def new_shard(self, kbid: str, vectorsets: dict) -> str:
        """vectorsets: name -> VectorConfig."""
        sid = str(_uuid.uuid4())
        with self._lock:
            self._shards[sid] = _Shard(kbid, {n: _VectorIndex(c) for n, c in vectorsets.items()})
        return sid

    def _grpc_new_shard(self, request, context):
        cfgs = {}
        for name, c in request.vectorsets_configs.items():
            if not c.HasField("vector_dimension"):
                import grpc

                context.abort(grpc.StatusCode.INVALID_ARGUMENT, f"vectorset {name}: vector_dimension is required")
            cfgs[name] = V.VectorConfig(dimension=int(c.vector_dimension), similarity=V.Similarity.Cosine if c.similarity == 0 else V.Similarity.Dot,
                                        normalize_vectors=bool(c.normalize_vectors), device=self.device)
        return P.ShardCreated(id=self.new_shard(request.kbid, cfgs))

    # ---- index (lib.rs:83-111 -> process_index_message) ---------------------------------------------------------------------
    def index(self, bytes: bytes) -> int:  # noqa: A002  (the reference names the parameter `bytes`: nidx_binding.pyi:45, callers may pass it by keyword)
        msg = P.IndexMessage.FromString(memoryview(bytes).tobytes())
        with self._lock:
            seq = self._seq
            self._seq += 1                                          # lib.rs:104-105: always incremented, even on failure
            try:
                self._process(msg, seq)
            except Exception as e:  # noqa: BLE001
                raise Exception(f"Error indexing {e}") from e
            self._dirty.add(msg.shard)
        return seq

    def _load_resource(self, storage_key: str):
        if self.settings.get("INDEXER__OBJECT_STORE", "file") != "file":
            raise ValueError("only the local file object store is supported (INDEXER__OBJECT_STORE=file)")
        path = os.path.join(self.settings.get("INDEXER__FILE_PATH", ""), storage_key)
        with open(path, "rb") as f:
            return P.Resource.FromString(f.read())

    def _process(self, msg, seq: int):
        shard = self._shards.get(msg.shard)
        if shard is None:
            raise KeyError(f"shard {msg.shard} not found")
        if msg.typemessage == 1:                                    # DELETION: every index drops the resource (indexer.rs delete_resource)
            for vi in shard.vectorsets.values():
                vi.deletions.append((msg.resource, seq))
            shard.deleted_resources.add((msg.resource, seq))
            return
        res = self._load_resource(msg.storage_key)
        rid = res.resource.uuid
        # a re-indexed resource replaces its older copies: prefixes to delete, applied to OLDER segments only (seq rule, lib.rs:188-199)
        for vi in shard.vectorsets.values():
            for key in list(res.vectors_to_delete_in_all_vectorsets) or [rid]:
                vi.deletions.append((key, seq))
        shard.deleted_resources.add((rid, seq))
        # vectors: one segment per vectorset (indexer.rs:96-146)
        for name, vi in shard.vectorsets.items():
            elems = []
            for _, paragraphs in res.paragraphs.items():
                for _, par in paragraphs.paragraphs.items():
                    sentences = par.vectorsets_sentences[name].sentences if name in par.vectorsets_sentences else par.sentences
                    for key, sentence in sentences.items():
                        if len(sentence.vector) == 0:
                            continue
                        meta = sentence.metadata.SerializeToString() if sentence.HasField("metadata") else None
                        elems.append(V.Elem(key, [np.asarray(sentence.vector, dtype=np.float32)], labels=list(par.labels), metadata=meta))
            if elems:
                vi.segments.append((V.VectorIndexer.index_elems(elems, vi.config), seq))
        # documents (nidx_text/src/resource_indexer.rs:22-91): one per field; paragraphs (nidx_paragraph): one per paragraph
        if not res.skip_texts:
            docs = [T.TextDoc(rid, "/" + fid if not fid.startswith("/") else fid, ti.text, tuple(list(res.labels) + list(ti.labels))) for fid, ti in res.texts.items()]
            if docs:
                shard.text_segments.append((docs, seq))
        if not res.skip_paragraphs:
            pdocs = []
            for fid, paragraphs in res.paragraphs.items():
                text = res.texts[fid].text if fid in res.texts else ""
                for pid, par in paragraphs.paragraphs.items():
                    labels = tuple(list(res.labels) + list(res.texts[fid].labels if fid in res.texts else ()) + list(par.labels))
                    pdocs.append(T.TextDoc(rid, "/" + fid if not fid.startswith("/") else fid, text[par.start:par.end], labels))
                    shard.paragraph_meta[(rid, "/" + fid if not fid.startswith("/") else fid, len(pdocs) - 1, seq)] = (pid, par)
            if pdocs:
                shard.paragraph_segments.append((pdocs, seq))

    # ---- sync (lib.rs:113-126; searcher/sync.rs + index_cache.rs:180-200: searchers are re-opened, never mutated) --------------
    def wait_for_sync(self) -> None:
        with self._lock:
            for sid in list(self._dirty):
                self._reopen(self._shards[sid])
            self._dirty.clear()

    def _alive(self, docs_segments, deleted):
        out = []
        for docs, seq in docs_segments:
            keep = [d for d in docs if not any(d.uuid == rid and dseq > seq for rid, dseq in deleted)]
            if keep:
                out.append(keep)
        return out

    def _reopen(self, shard: _Shard):
        for vi in shard.vectorsets.values():
            vi.searcher = V.VectorSearcher.open(vi.config, vi.segments, vi.deletions) if vi.segments else None
        ts = self._alive(shard.text_segments, shard.deleted_resources)
        ps = self._alive(shard.paragraph_segments, shard.deleted_resources)
        shard.text_searcher = T.TextSearcher.open(ts, device=self.device) if ts else None
        shard.paragraph_searcher = T.ParagraphSearcher.open(ps, device=self.device) if ps else None

    # ---- NidxSearcher.Search (shard_search.rs:60-241 + shard_merge.rs) ---------------------------------------------------------
    def search(self, request):
        """nodereader.SearchRequest -> nodereader.SearchResponse (in-process twin of the gRPC method)."""
        parts = []
        with self._lock:
            for sid in request.shard_ids:
                shard = self._shards.get(sid)
                if shard is None:
                    raise KeyError(f"shard {sid} not found")
                parts.append((sid, self._search_shard(shard, request)))
        return self._merge(request, parts)

    def _grpc_search(self, request, context):
        import grpc

        try:
            return self.search(request)
        except KeyError as e:
            context.abort(grpc.StatusCode.NOT_FOUND, str(e))
        except V.NidxError as e:
            context.abort(grpc.StatusCode.INTERNAL, str(e))
        except ValueError as e:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))

    def _search_shard(self, shard: _Shard, req):
        k = int(req.result_per_page)
        out = {}
        # prefilter (shard_search.rs:108-137): field_filter on the documents -> the fields that may answer
        prefilter = V.PrefilterResult.all()
        if req.HasField("field_filter") and shard.text_searcher is not None:
            fields = [V.FieldId(_uuid.UUID(d.uuid), d.field) for seg in shard.text_searcher.segments for d in seg.docs if _doc_matches(req.field_filter, d)]
            prefilter = V.PrefilterResult.some(fields) if fields else V.PrefilterResult.none()
        if len(req.vector):
            name = req.vectorset
            if name not in shard.vectorsets:
                raise ValueError(f"vectorset {name!r} not found")          # shard_search.rs:95-99 InvalidArgument
            vi = shard.vectorsets[name]
            formula = _expr_to_boolean(req.paragraph_filter) if req.HasField("paragraph_filter") else None
            vreq = V.VectorSearchRequest(vector=list(req.vector), result_per_page=k, with_duplicates=bool(req.with_duplicates), vector_set=name,
                                         min_score=float(req.min_score_semantic), filtering_formula=formula,
                                         filter_operator=V.FilterOperator.Or if req.filter_operator == P.FILTER_OR else V.FilterOperator.And)
            out["vector"] = vi.searcher.search(vreq, prefilter).documents if vi.searcher is not None else []
        if req.document and shard.text_searcher is not None:
            out["document"] = shard.text_searcher.search(T.DocumentSearchRequest(body=req.body, result_per_page=k, min_score=float(req.min_score_bm25)))
        if req.paragraph and shard.paragraph_searcher is not None:
            after = None
            if req.HasField("search_after"):
                after = T.SearchAfter(score=req.search_after.score, tie_break="keep_after", docaddr=int(req.search_after.docaddr))
            out["paragraph"] = shard.paragraph_searcher.search(T.DocumentSearchRequest(body=req.body, result_per_page=k, min_score=float(req.min_score_bm25), search_after=after))
        return out

    def _merge(self, req, parts):
        k = int(req.result_per_page)
        resp = P.SearchResponse()
        resp.shard_ids.extend(sid for sid, _ in parts)
        # vectors: kmerge_by(score >=), take(k) (shard_merge.rs:332-348)
        vec = sorted(((-d.score, i, j, d) for i, (_, p) in enumerate(parts) for j, d in enumerate(p.get("vector", []))), key=lambda t: t[:3])[:k]
        for _, _, _, d in vec:
            ds = resp.vector.documents.add()
            ds.doc_id.id, ds.score = d.doc_id, d.score
            ds.labels.extend(d.labels)
            if d.metadata:
                ds.metadata.CopyFrom(P.SentenceMetadata.FromString(d.metadata))
        # documents / paragraphs: bm25 desc, then shard, then lower docaddr (shard_merge.rs:227-231)
        for kind, target in (("document", resp.document), ("paragraph", resp.paragraph)):
            found = [(sid, p[kind]) for sid, p in parts if kind in p]
            if not found:
                continue
            rows = sorted(((-r.score.bm25, i, r.score.docaddr, sid, r) for i, (sid, rs) in enumerate(found) for r in rs.results), key=lambda t: t[:3])
            target.total = sum(rs.total for _, rs in found)
            target.next_page = any(rs.next_page for _, rs in found) or len(rows) > k
            target.query = req.body
            for _, _, _, sid, r in rows[:k]:
                o = target.results.add()
                o.uuid, o.field = r.uuid, r.field
                o.score.bm25, o.score.docaddr = r.score.bm25, r.score.docaddr
                o.labels.extend(r.labels)
                o.shard_id = sid.encode()
        return resp

    def close(self):
        """Stop the servers and release every device-resident segment (the reference's Drop cancels its runtime)."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        self._searcher.stop(0)
        self._api.stop(0)
        with self._lock:
            for shard in self._shards.values():
                for vi in shard.vectorsets.values():
                    vi.searcher = None
                    for seg, _ in vi.segments:
                        seg.close()
                    vi.segments.clear()
                for searcher in (shard.text_searcher, shard.paragraph_searcher):
                    if searcher is not None:
                        for seg in searcher.segments:
                            if seg._gpu is not None:
                                seg._gpu.close()
                shard.text_searcher = shard.paragraph_searcher = None
            self._shards.clear()

    def __del__(self):   # lib.rs Drop: the cancellation token
        try:
            self.close()
        except Exception:
            pass
