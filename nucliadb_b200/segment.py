"""Thin object wrappers over the C ABI handles (``include/nidx_b200.h``).

``VectorSegment``  = nidx_vec_segment: vectors + HNSW graph resident in HBM, exact scan / HNSW search /
                     GPU graph build.  Accepts numpy arrays (host path: copies inside the call) or torch
                     CUDA tensors (device path: zero copy, asynchronous on the current torch stream).
``TextSegment``    = nidx_txt_segment: postings resident in HBM, BM25 top-k.
torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import NIL, NidxError, TxtSearchParams, VecConfig, VecSearchParams, check, ptr


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _torch_stream(device: int):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class VectorSegment:
    def __init__(self, handle, cfg: VecConfig):
        self._h, self.cfg = handle, cfg

    # ---- lifecycle ----------------------------------------------------------------------------------
    @classmethod
    def create(cls, vectors, dimension: int, similarity=_lib.NIDX_SIM_COSINE, m=30, m0=60, ef_construction=100, ef_search=30, device=0,
               multi_vector=False, paragraph_of: Optional[np.ndarray] = None) -> "VectorSegment":
        L = _lib.require_device()
        cfg = VecConfig(dimension, similarity, int(multi_vector), m, m0, ef_construction, ef_search, device)
        h = C.c_void_p()
        if _is_torch(vectors):
            assert vectors.is_cuda and vectors.is_contiguous() and vectors.dtype.is_floating_point
            n, ld = vectors.shape
            mem = _lib.NIDX_MEM_DEVICE
        else:
            vectors = np.ascontiguousarray(vectors, dtype=np.float32)
            n, ld = vectors.shape if vectors.ndim == 2 else (0, dimension)
            mem = _lib.NIDX_MEM_HOST
        par = None if paragraph_of is None else np.ascontiguousarray(paragraph_of, dtype=np.uint32)
        check(L.nidx_vec_create(C.byref(cfg), ptr(vectors) if n else None, C.c_uint64(n), C.c_int32(ld), mem, ptr(par), C.byref(h)))
        return cls(h, cfg)

    @classmethod
    def open(cls, directory: str, dimension: int, similarity=_lib.NIDX_SIM_COSINE, m=30, m0=60, ef_construction=100, ef_search=30, device=0,
             multi_vector=False) -> "VectorSegment":
        L = _lib.require_device()
        cfg = VecConfig(dimension, similarity, int(multi_vector), m, m0, ef_construction, ef_search, device)
        h = C.c_void_p()
        check(L.nidx_vec_open(C.byref(cfg), directory.encode(), C.byref(h)))
        return cls(h, cfg)

    def save(self, directory: str):
        check(_lib.load().nidx_vec_save(self._h, directory.encode()))

    def close(self):
        if self._h is not None:
            _lib.load().nidx_vec_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(_lib.load().nidx_vec_len(self._h))

    # ---- graph ----------------------------------------------------------------------------------------
    def build_hnsw(self, seed=2, max_batch=4096):
        check(_lib.load().nidx_vec_build_hnsw(self._h, C.c_uint64(seed), C.c_int32(max_batch), None))

    def extend_hnsw(self, n_existing, level, adj0, adjU, w0, wU, entry_node, entry_layer, seed=2, max_batch=4096):
        """Reuse the graph of the first n_existing vectors and insert the rest (segment.rs:143-167)."""
        level = np.ascontiguousarray(level, dtype=np.uint8)
        adj0 = np.ascontiguousarray(adj0, dtype=np.uint32)
        adjU = np.ascontiguousarray(adjU, dtype=np.uint32)
        w0 = np.ascontiguousarray(w0, dtype=np.float32)
        wU = np.ascontiguousarray(wU, dtype=np.float32)
        check(_lib.load().nidx_vec_extend_hnsw(self._h, C.c_uint64(n_existing), ptr(level), ptr(adj0), ptr(w0), ptr(adjU), ptr(wU), C.c_uint32(entry_node),
                                               C.c_uint32(entry_layer), C.c_uint64(seed), C.c_int32(max_batch), None))

    def graph_dims(self):
        s0, su, rows, en, el = C.c_int32(), C.c_int32(), C.c_uint64(), C.c_uint32(), C.c_uint32()
        check(_lib.load().nidx_vec_graph_dims(self._h, C.byref(s0), C.byref(su), C.byref(rows), C.byref(en), C.byref(el)))
        return s0.value, su.value, rows.value, en.value, el.value

    def set_graph(self, level, adj0, adjU, w0=None, wU=None):
        level = np.ascontiguousarray(level, dtype=np.uint8)
        adj0 = np.ascontiguousarray(adj0, dtype=np.uint32)
        adjU = np.ascontiguousarray(adjU, dtype=np.uint32)
        w0 = None if w0 is None else np.ascontiguousarray(w0, dtype=np.float32)
        wU = None if wU is None else np.ascontiguousarray(wU, dtype=np.float32)
        check(_lib.load().nidx_vec_set_graph(self._h, ptr(level), ptr(adj0), ptr(w0), ptr(adjU), ptr(wU)))

    def get_graph(self):
        """-> dict(level, adj0, w0, adjU, wU, entry_node, entry_layer, s0, su)."""
        s0, su, rows, en, el = self.graph_dims()
        n = len(self)
        level = np.empty(n, dtype=np.uint8)
        adj0 = np.empty((n, s0), dtype=np.uint32)
        w0 = np.empty((n, s0), dtype=np.float32)
        adjU = np.full((max(rows, 1), su), NIL, dtype=np.uint32)
        wU = np.zeros((max(rows, 1), su), dtype=np.float32)
        check(_lib.load().nidx_vec_get_graph(self._h, ptr(level), ptr(adj0), ptr(w0), ptr(adjU), ptr(wU)))
        return dict(level=level, adj0=adj0, w0=w0, adjU=adjU, wU=wU, entry_node=en, entry_layer=el, s0=s0, su=su, upper_rows=rows)

    def set_alive(self, alive_bits: Optional[np.ndarray]):
        check(_lib.load().nidx_vec_set_alive(self._h, ptr(alive_bits), _lib.NIDX_MEM_HOST))

    def set_paragraph_keys(self, keys: Optional[np.ndarray]):
        """64-bit keys of the paragraph ids, for the cross-segment de-duplication of a sharded search (Fssc, searcher.rs:150-199)."""
        keys = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
        check(_lib.load().nidx_vec_set_paragraph_keys(self._h, ptr(keys)))

    # ---- search ----------------------------------------------------------------------------------------
    def search(self, queries, k: int, ef: int = 0, min_score: float = -1.0, with_duplicates=True, method=_lib.NIDX_METHOD_AUTO,
               filter_bits=None, filter_matching: int = 0, out=None, stream: Optional[int] = None):
        """Batch search.  numpy in -> numpy out (host path, synchronous); torch CUDA tensors in -> torch
        CUDA tensors out (device path, asynchronous on the current stream).  `stream` (a cudaStream_t as int) lets concurrent
        host-path callers overlap their copies with each other's kernels.  Returns (ids, scores, counts)."""
        L = _lib.load()
        p = VecSearchParams(k, ef, min_score, int(with_duplicates), method, None, filter_matching)
        if _is_torch(queries):
            import torch

            assert queries.is_cuda and queries.is_contiguous() and queries.dtype == torch.float32
            nq, ldq = queries.shape
            dev = queries.device
            if out is None:
                out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
                       torch.empty((nq,), dtype=torch.int32, device=dev))
            if filter_bits is not None:
                p.filter_bits = filter_bits.data_ptr()
            check(L.nidx_vec_search(self._h, ptr(queries), C.c_int32(nq), C.c_int32(ldq), _lib.NIDX_MEM_DEVICE, C.byref(p), ptr(out[0]), ptr(out[1]),
                                    ptr(out[2]), _torch_stream(self.cfg.device)))
            return out
        queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        nq, ldq = queries.shape
        ids = np.empty((nq, k), dtype=np.uint32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        keep = None
        if filter_bits is not None:
            keep = np.ascontiguousarray(filter_bits, dtype=np.uint64)
            p.filter_bits = keep.ctypes.data
        check(L.nidx_vec_search(self._h, ptr(queries), C.c_int32(nq), C.c_int32(ldq), _lib.NIDX_MEM_HOST, C.byref(p), ptr(ids), ptr(scores), ptr(counts),
                                C.c_void_p(stream) if stream else None))
        return ids, scores, counts

    # ---- RaBitQ (vector_types/rabitq.rs) ----------------------------------------------------------------
    def rabitq_encode(self):
        check(_lib.load().nidx_vec_rabitq_encode(self._h, None))

    def rabitq_codes(self) -> np.ndarray:
        out = np.empty((len(self), self.cfg.dimension // 8 + 8), dtype=np.uint8)
        check(_lib.load().nidx_vec_rabitq_codes(self._h, ptr(out)))
        return out

    def rabitq_estimate(self, queries):
        queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        nq = queries.shape[0]
        est = np.empty((nq, len(self)), dtype=np.float32)
        err = np.empty((nq, len(self)), dtype=np.float32)
        check(_lib.load().nidx_vec_rabitq_estimate(self._h, ptr(queries), C.c_int32(nq), C.c_int32(queries.shape[1]), _lib.NIDX_MEM_HOST, ptr(est), ptr(err), None))
        return est, err

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(_lib.load().nidx_vec_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def counters_ex(self):
        out = (C.c_uint64 * 6)()
        check(_lib.load().nidx_vec_counters_ex(self._h, out))
        return dict(similarities=out[0], expansions=out[1], overflows=out[2] + out[3], estimates=out[4], rerank_needed=out[5])

    def counters(self):
        out = (C.c_uint64 * 3)()
        check(_lib.load().nidx_vec_counters(self._h, out))
        return dict(similarities=out[0], expansions=out[1], overflows=out[2])


def merge_topk(ids, scores, device=0, part_stride=0, out=None):
    """[n_parts, nq, k] torch CUDA tensors (each part sorted desc, NIL padded) -> merged (ids, scores, part).
    ids / scores may be strided views of one all-gather buffer: part_stride = elements between parts."""
    import torch

    n_parts, nq, k = ids.shape
    if out is None:
        out = (torch.empty((nq, k), dtype=ids.dtype, device=ids.device), torch.empty((nq, k), dtype=torch.float32, device=ids.device),
               torch.empty((nq, k), dtype=torch.int32, device=ids.device))
    check(_lib.load().nidx_merge_topk(C.c_int32(device), ptr(ids), ptr(scores), C.c_int32(n_parts), C.c_int64(part_stride), C.c_int32(nq), C.c_int32(k),
                                      ptr(out[0]), ptr(out[1]), ptr(out[2]), _torch_stream(device)))
    return out


class TextSegment:
    def __init__(self, handle, n_docs, n_terms, device):
        self._h, self.n_docs, self.n_terms, self.device = handle, n_docs, n_terms, device

    @classmethod
    def create(cls, n_docs, n_terms, term_off, post_doc, post_tf, fieldnorm_id, device=0) -> "TextSegment":
        L = _lib.require_device()
        term_off = np.ascontiguousarray(term_off, dtype=np.uint64)
        post_doc = np.ascontiguousarray(post_doc, dtype=np.uint32)
        post_tf = np.ascontiguousarray(post_tf, dtype=np.uint32)
        fieldnorm_id = np.ascontiguousarray(fieldnorm_id, dtype=np.uint8)
        h = C.c_void_p()
        check(L.nidx_txt_create(C.c_int32(device), C.c_uint32(n_docs), C.c_uint32(n_terms), ptr(term_off), ptr(post_doc), ptr(post_tf), ptr(fieldnorm_id),
                                C.byref(h)))
        return cls(h, n_docs, n_terms, device)

    def set_stats(self, total_docs: int, total_tokens: int, doc_freq: Optional[np.ndarray] = None):
        df = None if doc_freq is None else np.ascontiguousarray(doc_freq, dtype=np.uint64)
        check(_lib.load().nidx_txt_set_stats(self._h, C.c_uint64(total_docs), C.c_uint64(total_tokens), ptr(df)))

    def set_alive(self, alive_bits: Optional[np.ndarray]):
        check(_lib.load().nidx_txt_set_alive(self._h, ptr(alive_bits)))

    def search(self, query_terms, query_off, k, mode=_lib.NIDX_BM25_OR, use_tf=True, min_score=0.0, out=None, after=None, docaddr_base=0):
        """query i = query_terms[query_off[i]:query_off[i+1]].  numpy -> host path, torch CUDA int32 -> device path.
        Returns (docs, scores, counts, total)."""
        L = _lib.load()
        # after = (score, mode, docaddr) with mode 1 Drop / 2 KeepAfter / 3 Keep (nidx_paragraph SearchAfter)
        p = TxtSearchParams(k, mode, int(use_tf), min_score, 0, 0.0, 0, docaddr_base)
        if after is not None:
            p.after_score, p.after_mode, p.after_docaddr = float(after[0]), int(after[1]), int(after[2])
        if _is_torch(query_terms):
            import torch

            nq = query_off.numel() - 1
            dev = query_terms.device
            if out is None:
                out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
                       torch.empty((nq,), dtype=torch.int32, device=dev), torch.empty((nq,), dtype=torch.int64, device=dev))
            check(L.nidx_txt_search(self._h, ptr(query_terms), ptr(query_off), C.c_int32(nq), _lib.NIDX_MEM_DEVICE, C.byref(p), ptr(out[0]), ptr(out[1]),
                                    ptr(out[2]), ptr(out[3]), _torch_stream(self.device)))
            return out
        query_terms = np.ascontiguousarray(query_terms, dtype=np.uint32)
        query_off = np.ascontiguousarray(query_off, dtype=np.uint32)
        nq = len(query_off) - 1
        docs = np.empty((nq, k), dtype=np.uint32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        total = np.empty(nq, dtype=np.uint64)
        check(L.nidx_txt_search(self._h, ptr(query_terms), ptr(query_off), C.c_int32(nq), _lib.NIDX_MEM_HOST, C.byref(p), ptr(docs), ptr(scores), ptr(counts),
                                ptr(total), None))
        return docs, scores, counts, total

    def set_doc_keys(self, keys: Optional[np.ndarray]):
        """Caller keys of the documents (paragraph ids) for rank fusion; None = the document number."""
        k = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
        check(_lib.load().nidx_txt_set_doc_keys(self._h, ptr(k)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(_lib.load().nidx_txt_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h is not None:
            _lib.load().nidx_txt_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
