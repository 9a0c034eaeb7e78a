"""Multi-GPU search: one process per GPU, segments sharded across ranks, per-rank partial top-k
exchanged with ``all_gather`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) and merged by
``parts_merge_kernel`` -- the B200-native form of the reference's gRPC scatter-gather + k-way merge
(nidx/src/searcher/grpc.rs:253-431, shard_merge.rs:332-348 / 177-207).

The exchange is [nq, k] (u32 id, f32 score) per rank = 8*nq*k bytes (80 KB at nq=1024, k=10): latency
bound, so it is one collective per batch, not per query.
"""
from __future__ import annotations


def gather_partials(ids, scores, group=None):
    """All ranks contribute their [nq, k] partial results; every rank receives [world, nq, k] in rank order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    ids_all = [torch.empty_like(ids) for _ in range(world)]
    sc_all = [torch.empty_like(scores) for _ in range(world)]
    dist.all_gather(ids_all, ids.contiguous(), group=group)
    dist.all_gather(sc_all, scores.contiguous(), group=group)
    return torch.stack(ids_all), torch.stack(sc_all)


def global_ids(local_ids, part, vectors_per_rank: int):
    """(segment rank, local vector address) -> global address; NIL (-1 as int32) stays NIL."""
    import torch

    g = local_ids.to(torch.int64) + part.to(torch.int64) * int(vectors_per_rank)
    return torch.where(local_ids.to(torch.int64) < 0, torch.full_like(g, -1), g)


def sharded_search(segment, queries, k, ef, device, group=None, **kw):
    """Search this rank's segment, exchange partial top-k, merge on the GPU.  Returns (local ids, scores, part)."""
    from . import _lib
    from .segment import merge_topk

    ids, scores, _ = segment.search(queries, k, ef=ef, method=_lib.NIDX_METHOD_HNSW, **kw)
    ids_all, sc_all = gather_partials(ids, scores, group)
    return merge_topk(ids_all, sc_all, device=device)
