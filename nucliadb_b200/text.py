"""Host-side mirror of the BM25 part of ``nidx_text`` / ``nidx_paragraph`` over the CUDA library.

Reference call shape (the scoring itself is tantivy's, restated in oracle/bm25.hpp and done on the GPU by
``bm25_kernel``):

* ``TextSearcher.search(DocumentSearchRequest)``      nidx_text/src/lib.rs:178-227 -> reader.rs:367-451
  body parsed with ``QueryParser::set_conjunction_by_default`` (AND of terms, real tf), ``TopDocs(k+1)``,
  ``next_page = len > k``, hits below ``min_score`` dropped (reader.rs:289-355).
* ``ParagraphSearcher.search(ParagraphSearchRequest)``  nidx_paragraph/src/lib.rs:117-147 -> reader.rs:244-392
  keyword query = OR of ``TermQuery(IndexRecordOption::Basic)`` (keyword_parser.rs:27-67): tf == 1.
* statistics over the union of all segments (nidx_tantivy/src/index_reader.rs:39-77); results ordered by
  (score desc, segment_ord asc, doc asc) with ``docaddr = (segment_ord << 32) + doc`` (reader.rs:310, Q13).

Tokenisation mirrors tantivy's "default" analyzer (SimpleTokenizer + RemoveLongFilter(40) + LowerCaser)
[recalled]; fuzzy fallback, facets, filters and stop words are query-preparation, outside the hot path.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .segment import TextSegment

_TOKEN = re.compile(r"[^\W_]+", re.UNICODE)


def tokenize(text: str) -> list:
    """tantivy's "default" analyzer [recalled]: SimpleTokenizer (alphanumeric runs) -> RemoveLongFilter::limit(40), which keeps a
    token iff `token.text.len() < 40` -- a length in UTF-8 BYTES, strictly below the limit -- -> LowerCaser."""
    return [t.lower() for t in _TOKEN.findall(text) if len(t.encode("utf-8")) < 40]


def fieldnorm_to_id(n: int) -> int:
    """tantivy's 1-byte fieldnorm code (Lucene SmallFloat.intToByte4) [recalled]."""
    if n < 24:
        return n
    x = n - 24
    nbits = x.bit_length()
    if nbits <= 3:
        return 24 + x
    shift = nbits - 4
    enc = ((x >> shift) & 7) | ((shift + 1) << 3)
    return min(255, 24 + enc)


@dataclass
class ResultScore:  # nodereader.proto:48-53
    bm25: float
    docaddr: int


@dataclass
class DocumentResult:
    uuid: str
    field: str
    score: ResultScore
    labels: list


@dataclass
class DocumentSearchRequest:  # nidx_text/src/request_types.rs:17-28
    body: str = ""
    result_per_page: int = 20
    min_score: float = 0.0
    only_faceted: bool = False
    search_after: Optional[SearchAfter] = None  # ParagraphSearchRequest.search_after (nidx_paragraph only)


@dataclass
class SearchAfter:  # nidx_paragraph/src/request_types.rs:20-31
    score: float
    tie_break: str = "drop"   # "drop" | "keep" | "keep_after"
    docaddr: int = 0          # payload of KeepAfter


@dataclass
class DocumentSearchResponse:
    results: list = field(default_factory=list)
    total: int = 0
    next_page: bool = False
    query: str = ""


@dataclass
class TextDoc:
    uuid: str
    field: str
    text: str
    labels: Sequence[str] = ()


class TextIndexSegment:
    """One immutable segment: term dictionary + postings (host build, HBM resident)."""

    def __init__(self, docs: Sequence[TextDoc], vocab: dict, device=0):
        self.docs = list(docs)
        toks = [[vocab.setdefault(t, len(vocab)) for t in tokenize(d.text)] for d in self.docs]
        self.n_docs = len(self.docs)
        self.lens = np.asarray([len(t) for t in toks], dtype=np.int64)
        self.total_tokens = int(self.lens.sum())
        pairs = sorted({(t, i) for i, ts in enumerate(toks) for t in ts})
        tf = {}
        for i, ts in enumerate(toks):
            for t in ts:
                tf[(t, i)] = tf.get((t, i), 0) + 1
        self.n_terms = len(vocab)
        self.post_term = np.asarray([p[0] for p in pairs], dtype=np.int64)
        self.post_doc = np.asarray([p[1] for p in pairs], dtype=np.uint32)
        self.post_tf = np.asarray([tf[p] for p in pairs], dtype=np.uint32)
        self.fieldnorm_id = np.asarray([fieldnorm_to_id(int(x)) for x in self.lens], dtype=np.uint8)
        self._gpu: Optional[TextSegment] = None
        self.device = device

    def doc_freq(self, n_terms: int) -> np.ndarray:
        return np.bincount(self.post_term, minlength=n_terms).astype(np.uint64)

    def upload(self, n_terms: int):
        term_off = np.zeros(n_terms + 1, dtype=np.uint64)
        term_off[1:] = np.cumsum(np.bincount(self.post_term, minlength=n_terms))
        self._gpu = TextSegment.create(self.n_docs, n_terms, term_off, self.post_doc, self.post_tf, self.fieldnorm_id, device=self.device)
        return self._gpu


class TextSearcher:
    """nidx_text TextSearcher over one or more segments sharing one term dictionary."""

    conjunction = True   # QueryParser::set_conjunction_by_default (reader.rs:372-377)
    use_tf = True

    def __init__(self, segments: Sequence[TextIndexSegment], vocab: dict):
        _lib.require_device()
        self.segments, self.vocab = list(segments), vocab
        n_terms = len(vocab)
        total_docs = sum(s.n_docs for s in self.segments)
        total_tokens = sum(s.total_tokens for s in self.segments)
        df = np.zeros(n_terms, dtype=np.uint64)
        for s in self.segments:
            df += s.doc_freq(n_terms)
        for s in self.segments:  # union statistics on every segment (index_reader.rs:39-77)
            s.upload(n_terms).set_stats(max(total_docs, 1), max(total_tokens, 1), df)

    @classmethod
    def open(cls, docs_per_segment: Sequence[Sequence[TextDoc]], device=0):
        vocab: dict = {}
        segs = [TextIndexSegment(d, vocab, device) for d in docs_per_segment]
        return cls(segs, vocab)

    def _terms(self, body: str):
        terms = []
        for t in tokenize(body):
            terms.append(self.vocab.get(t, 0xFFFFFFF0))  # unknown term: matches nothing
        return terms

    def search(self, request: DocumentSearchRequest) -> DocumentSearchResponse:
        terms = self._terms(request.body)
        k = request.result_per_page
        resp = DocumentSearchResponse(query=request.body)
        if not terms or k <= 0:
            return resp
        qt = np.asarray(terms, dtype=np.uint32)
        qo = np.asarray([0, len(terms)], dtype=np.uint32)
        merged = []
        for ord_, seg in enumerate(self.segments):
            after = None
            if request.search_after is not None:
                sa = request.search_after
                after = (sa.score, {"drop": 1, "keep_after": 2, "keep": 3}[sa.tie_break], sa.docaddr)
            docs, scores, counts, total = seg._gpu.search(qt, qo, k + 1, mode=_lib.NIDX_BM25_AND if self.conjunction else _lib.NIDX_BM25_OR,
                                                          use_tf=self.use_tf, min_score=0.0, after=after, docaddr_base=ord_ << 32)
            resp.total += int(total[0])
            merged += [(-float(scores[0, i]), ord_, int(docs[0, i])) for i in range(int(counts[0]))]
        merged.sort()  # score desc, then segment_ord, then doc: lower docaddr first
        resp.next_page = len(merged) > k  # reader.rs:300-301
        for neg, ord_, doc in merged[:k]:
            score = -neg
            if score < request.min_score:  # reader.rs:302-305
                continue
            d = self.segments[ord_].docs[doc]
            resp.results.append(DocumentResult(d.uuid, d.field, ResultScore(score, (ord_ << 32) + doc), list(d.labels)))
        return resp


class ParagraphSearcher(TextSearcher):
    """nidx_paragraph keyword search: OR of TermQuery(Basic) => tf == 1 (keyword_parser.rs:27-67)."""

    conjunction = False
    use_tf = False
