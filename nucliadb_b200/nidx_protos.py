"""The wire schema of the nidx searcher / indexer that the hot path needs, built at import time from descriptors written by
hand (there is no protoc in this image): same packages, message names, field names, numbers and types as the reference's
``nidx/nidx_protos/{nidx,nodereader,noderesources,nodewriter}.proto`` for the SUBSET of fields the search path reads or writes
(cited per message below).  Fields that are not declared here are skipped by the protobuf runtime as unknown fields, so requests
serialised by the reference's clients (``nidx_protos`` / ``nucliadb_protos``) decode, and the responses decode on their side.

    SearchRequest / SearchResponse                         nodereader.proto:388-437, 476-488
    DocumentSearchResponse / DocumentResult / ResultScore  nodereader.proto:48-81
    ParagraphSearchResponse / ParagraphResult              nodereader.proto:83-124
    VectorSearchResponse / DocumentScored                  nodereader.proto:126-142
    FilterExpression, FilterOperator, SearchAfter          nodereader.proto:287-336, 382-386
    IndexMessage, TypeMessage                              nodewriter.proto:26-43
    Resource, IndexParagraph(s), VectorSentence, ...       noderesources.proto:8-180
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_T = {"string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES, "int32": _F.TYPE_INT32, "int64": _F.TYPE_INT64, "uint32": _F.TYPE_UINT32, "uint64": _F.TYPE_UINT64,
      "float": _F.TYPE_FLOAT, "bool": _F.TYPE_BOOL}


def _field(msg, name, number, typ, repeated=False, oneof=None, optional=False):
    f = msg.field.add()
    f.name, f.number = name, number
    f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
    if typ in _T:
        f.type = _T[typ]
    elif typ.startswith("enum:"):
        f.type, f.type_name = _F.TYPE_ENUM, typ[5:]
    else:
        f.type, f.type_name = _F.TYPE_MESSAGE, typ
    if oneof is not None:
        f.oneof_index = oneof
    if optional:  # proto3 `optional`: a synthetic one-field oneof
        msg.oneof_decl.add().name = "_" + name
        f.oneof_index = len(msg.oneof_decl) - 1
        f.proto3_optional = True
    return f


def _map(msg, pkg_path, name, number, key_type, value_type):
    """map<key, value> name = number  ==  repeated NameEntry (map_entry) with key = 1, value = 2."""
    entry = msg.nested_type.add()
    entry.name = "".join(p.capitalize() for p in name.split("_")) + "Entry"
    entry.options.map_entry = True
    _field(entry, "key", 1, key_type)
    _field(entry, "value", 2, value_type)
    _field(msg, name, number, f"{pkg_path}.{entry.name}", repeated=True)


def _build():
    pool = descriptor_pool.DescriptorPool()

    # ---- noderesources.proto ---------------------------------------------------------------------------------------------
    fd = descriptor_pb2.FileDescriptorProto(name="nidx_protos/noderesources.proto", package="noderesources", syntax="proto3")
    m = fd.message_type.add(name="TextInformation")           # :8-11
    _field(m, "text", 1, "string"); _field(m, "labels", 2, "string", repeated=True)
    m = fd.message_type.add(name="ResourceID")                # :36-39
    _field(m, "shard_id", 1, "string"); _field(m, "uuid", 2, "string")
    m = fd.message_type.add(name="Position")                  # :53-67
    _field(m, "index", 1, "uint64"); _field(m, "start", 2, "uint64"); _field(m, "end", 3, "uint64"); _field(m, "page_number", 4, "uint64")
    _field(m, "start_seconds", 5, "uint32", repeated=True); _field(m, "end_seconds", 6, "uint32", repeated=True); _field(m, "in_page", 7, "bool")
    m = fd.message_type.add(name="Representation")            # :69-72
    _field(m, "is_a_table", 1, "bool"); _field(m, "file", 2, "string")
    for name in ("SentenceMetadata", "ParagraphMetadata"):    # :74-78, 89-93
        m = fd.message_type.add(name=name)
        _field(m, "position", 1, ".noderesources.Position"); _field(m, "page_with_visual", 2, "bool"); _field(m, "representation", 3, ".noderesources.Representation")
    m = fd.message_type.add(name="VectorSentence")            # :80-83
    _field(m, "vector", 1, "float", repeated=True); _field(m, "metadata", 9, ".noderesources.SentenceMetadata")
    m = fd.message_type.add(name="VectorsetSentences")        # :85-87
    _map(m, ".noderesources.VectorsetSentences", "sentences", 1, "string", ".noderesources.VectorSentence")
    m = fd.message_type.add(name="IndexParagraph")            # :95-106
    _field(m, "start", 1, "int32"); _field(m, "end", 2, "int32"); _field(m, "labels", 3, "string", repeated=True)
    _map(m, ".noderesources.IndexParagraph", "sentences", 4, "string", ".noderesources.VectorSentence")
    _field(m, "field", 5, "string"); _field(m, "split", 6, "string"); _field(m, "index", 7, "uint64"); _field(m, "repeated_in_field", 8, "bool")
    _field(m, "metadata", 9, ".noderesources.ParagraphMetadata")
    _map(m, ".noderesources.IndexParagraph", "vectorsets_sentences", 10, "string", ".noderesources.VectorsetSentences")
    m = fd.message_type.add(name="IndexParagraphs")           # :118-121
    _map(m, ".noderesources.IndexParagraphs", "paragraphs", 1, "string", ".noderesources.IndexParagraph")
    m = fd.message_type.add(name="Resource")                  # :123-180
    _field(m, "resource", 1, ".noderesources.ResourceID")
    _map(m, ".noderesources.Resource", "texts", 3, "string", ".noderesources.TextInformation")
    _field(m, "labels", 4, "string", repeated=True)
    _map(m, ".noderesources.Resource", "paragraphs", 6, "string", ".noderesources.IndexParagraphs")
    _field(m, "paragraphs_to_delete", 7, "string", repeated=True)
    _field(m, "vectors_to_delete_in_all_vectorsets", 8, "string", repeated=True)
    _field(m, "shard_id", 11, "string")
    _field(m, "texts_to_delete", 17, "string", repeated=True)
    _field(m, "skip_texts", 18, "bool"); _field(m, "skip_paragraphs", 19, "bool")
    pool.Add(fd)

    # ---- nodereader.proto ------------------------------------------------------------------------------------------------
    fd = descriptor_pb2.FileDescriptorProto(name="nidx_protos/nodereader.proto", package="nodereader", syntax="proto3",
                                            dependency=["nidx_protos/noderesources.proto"])
    e = fd.enum_type.add(name="FilterOperator")               # :333-336
    e.value.add(name="AND", number=0); e.value.add(name="OR", number=1)
    m = fd.message_type.add(name="ResultScore")               # :48-53
    _field(m, "bm25", 1, "float"); _field(m, "docaddr", 3, "uint64")
    m = fd.message_type.add(name="DocumentResult")            # :55-64
    m.oneof_decl.add().name = "sort_value"
    _field(m, "uuid", 1, "string"); _field(m, "score", 3, ".nodereader.ResultScore", oneof=0); _field(m, "field", 4, "string")
    _field(m, "labels", 5, "string", repeated=True); _field(m, "shard_id", 7, "bytes")
    m = fd.message_type.add(name="DocumentSearchResponse")    # :66-81
    _field(m, "total", 1, "int32"); _field(m, "results", 2, ".nodereader.DocumentResult", repeated=True); _field(m, "query", 6, "string")
    _field(m, "next_page", 7, "bool")
    m = fd.message_type.add(name="ParagraphResult")           # :83-104
    m.oneof_decl.add().name = "sort_value"
    _field(m, "uuid", 1, "string"); _field(m, "field", 3, "string"); _field(m, "start", 4, "uint64"); _field(m, "end", 5, "uint64")
    _field(m, "paragraph", 6, "string"); _field(m, "split", 7, "string"); _field(m, "index", 8, "uint64")
    _field(m, "score", 9, ".nodereader.ResultScore", oneof=0); _field(m, "matches", 10, "string", repeated=True)
    _field(m, "metadata", 11, ".noderesources.ParagraphMetadata"); _field(m, "labels", 12, "string", repeated=True); _field(m, "shard_id", 14, "bytes")
    m = fd.message_type.add(name="ParagraphSearchResponse")   # :106-124
    _field(m, "total", 1, "int32"); _field(m, "results", 2, ".nodereader.ParagraphResult", repeated=True); _field(m, "query", 6, "string")
    _field(m, "next_page", 7, "bool"); _field(m, "ematches", 9, "string", repeated=True)
    m = fd.message_type.add(name="DocumentVectorIdentifier")  # :126-128
    _field(m, "id", 1, "string")
    m = fd.message_type.add(name="DocumentScored")            # :130-135
    _field(m, "doc_id", 1, ".nodereader.DocumentVectorIdentifier"); _field(m, "score", 2, "float")
    _field(m, "metadata", 3, ".noderesources.SentenceMetadata"); _field(m, "labels", 4, "string", repeated=True)
    m = fd.message_type.add(name="VectorSearchResponse")      # :137-142
    _field(m, "documents", 1, ".nodereader.DocumentScored", repeated=True)
    m = fd.message_type.add(name="FilterExpression")          # :287-331
    lst = m.nested_type.add(name="FilterExpressionList"); _field(lst, "operands", 1, ".nodereader.FilterExpression", repeated=True)
    r = m.nested_type.add(name="ResourceFilter"); _field(r, "resource_id", 1, "string")
    ff = m.nested_type.add(name="FieldFilter"); _field(ff, "field_type", 1, "string"); _field(ff, "field_id", 2, "string", optional=True)
    kw = m.nested_type.add(name="KeywordFilter"); _field(kw, "keyword", 1, "string")
    fc = m.nested_type.add(name="FacetFilter"); _field(fc, "facet", 1, "string")
    m.oneof_decl.add().name = "expr"
    _field(m, "bool_and", 1, ".nodereader.FilterExpression.FilterExpressionList", oneof=0)
    _field(m, "bool_or", 2, ".nodereader.FilterExpression.FilterExpressionList", oneof=0)
    _field(m, "bool_not", 3, ".nodereader.FilterExpression", oneof=0)
    _field(m, "resource", 4, ".nodereader.FilterExpression.ResourceFilter", oneof=0)
    _field(m, "field", 5, ".nodereader.FilterExpression.FieldFilter", oneof=0)
    _field(m, "keyword", 6, ".nodereader.FilterExpression.KeywordFilter", oneof=0)
    _field(m, "facet", 8, ".nodereader.FilterExpression.FacetFilter", oneof=0)
    m = fd.message_type.add(name="SearchAfter")               # :382-386
    _field(m, "score", 1, "float"); _field(m, "shard_id", 2, "bytes"); _field(m, "docaddr", 3, "uint64")
    m = fd.message_type.add(name="SearchRequest")             # :388-437
    _field(m, "shard_ids", 1, "string", repeated=True); _field(m, "body", 3, "string"); _field(m, "result_per_page", 8, "int32")
    _field(m, "vector", 10, "float", repeated=True); _field(m, "paragraph", 12, "bool"); _field(m, "document", 13, "bool")
    _field(m, "with_duplicates", 14, "bool"); _field(m, "vectorset", 15, "string"); _field(m, "only_faceted", 16, "bool")
    _field(m, "min_score_semantic", 23, "float"); _field(m, "min_score_bm25", 25, "float")
    _field(m, "field_filter", 26, ".nodereader.FilterExpression", optional=True); _field(m, "paragraph_filter", 27, ".nodereader.FilterExpression", optional=True)
    _field(m, "filter_operator", 28, "enum:.nodereader.FilterOperator"); _field(m, "search_after", 35, ".nodereader.SearchAfter", optional=True)
    m = fd.message_type.add(name="SearchResponse")            # :476-488
    _field(m, "document", 1, ".nodereader.DocumentSearchResponse"); _field(m, "paragraph", 2, ".nodereader.ParagraphSearchResponse")
    _field(m, "vector", 3, ".nodereader.VectorSearchResponse"); _field(m, "shard_ids", 6, "string", repeated=True)
    pool.Add(fd)

    # ---- nodewriter.proto ------------------------------------------------------------------------------------------------
    fd = descriptor_pb2.FileDescriptorProto(name="nidx_protos/nodewriter.proto", package="nodewriter", syntax="proto3")
    e = fd.enum_type.add(name="TypeMessage")                  # :26-29
    e.value.add(name="CREATION", number=0); e.value.add(name="DELETION", number=1)
    m = fd.message_type.add(name="IndexMessage")              # :32-43
    _field(m, "node", 1, "string"); _field(m, "shard", 2, "string"); _field(m, "txid", 3, "uint64"); _field(m, "resource", 4, "string")
    _field(m, "typemessage", 5, "enum:.nodewriter.TypeMessage"); _field(m, "reindex_id", 6, "string"); _field(m, "storage_key", 8, "string")
    _field(m, "kbid", 9, "string")
    m = fd.message_type.add(name="OpStatus")
    _field(m, "detail", 2, "string")
    m = fd.message_type.add(name="VectorIndexConfig")         # :49-54 (similarity: utils.VectorSimilarity COSINE = 0, DOT = 1, utils.proto:96-99)
    _field(m, "similarity", 1, "int32"); _field(m, "normalize_vectors", 2, "bool"); _field(m, "vector_type", 3, "int32")
    _field(m, "vector_dimension", 4, "uint32", optional=True)
    m = fd.message_type.add(name="NewShardRequest")           # :56-71
    _field(m, "kbid", 2, "string")
    _map(m, ".nodewriter.NewShardRequest", "vectorsets_configs", 6, "string", ".nodewriter.VectorIndexConfig")
    pool.Add(fd)
    # noderesources.ShardCreated / ShardId live in noderesources.proto; declared in a side file of the same package
    fd = descriptor_pb2.FileDescriptorProto(name="nidx_protos/noderesources_shards.proto", package="noderesources", syntax="proto3")
    m = fd.message_type.add(name="ShardCreated")              # noderesources.proto:30-34
    _field(m, "id", 1, "string")
    m = fd.message_type.add(name="ShardId")                   # :22-24
    _field(m, "id", 1, "string")
    pool.Add(fd)
    return pool


POOL = _build()


def _cls(name):
    return message_factory.GetMessageClass(POOL.FindMessageTypeByName(name))


SearchRequest = _cls("nodereader.SearchRequest")
SearchResponse = _cls("nodereader.SearchResponse")
FilterExpression = _cls("nodereader.FilterExpression")
DocumentSearchResponse = _cls("nodereader.DocumentSearchResponse")
ParagraphSearchResponse = _cls("nodereader.ParagraphSearchResponse")
VectorSearchResponse = _cls("nodereader.VectorSearchResponse")
SentenceMetadata = _cls("noderesources.SentenceMetadata")
Resource = _cls("noderesources.Resource")
IndexMessage = _cls("nodewriter.IndexMessage")
NewShardRequest = _cls("nodewriter.NewShardRequest")
ShardCreated = _cls("noderesources.ShardCreated")
NEW_SHARD_METHOD = "/nidx.NidxApi/NewShard"           # nidx.proto:9
FILTER_AND, FILTER_OR = 0, 1
SEARCH_METHOD = "/nidx.NidxSearcher/Search"   # nidx.proto:20-21: package nidx, service NidxSearcher, rpc Search
