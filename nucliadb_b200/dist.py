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


class ShardedSearcher:
    """One rank's view of a segment-sharded index.  Buffers are allocated once: the local result is written
    straight into this rank's slot of the exchange buffer ([2, nq, k]: ids, score bits), ONE all_gather moves
    every rank's slot, and parts_merge_kernel merges the gathered buffer in place (part_stride = 2*nq*k)."""

    def __init__(self, segment, nq, k, device, group=None):
        import torch
        import torch.distributed as dist

        self.segment, self.nq, self.k, self.device, self.group = segment, nq, k, device, group
        self.world = dist.get_world_size(group)
        dev = torch.device("cuda", device)
        self.local = torch.empty((2, nq, k), dtype=torch.int32, device=dev)
        self.counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.gathered = torch.empty((self.world, 2, nq, k), dtype=torch.int32, device=dev)
        self.out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
                    torch.empty((nq, k), dtype=torch.int32, device=dev))

    def search(self, queries, ef, **kw):
        """-> (local ids, scores, part) of the merged top-k, identical on every rank."""
        import torch
        import torch.distributed as dist

        from . import _lib
        from .segment import merge_topk

        self.segment.search(queries, self.k, ef=ef, method=_lib.NIDX_METHOD_HNSW, out=(self.local[0], self.local[1].view(torch.float32), self.counts), **kw)
        dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        g = self.gathered
        return merge_topk(g[:, 0], g[:, 1].view(torch.float32), device=self.device, part_stride=2 * self.nq * self.k, out=self.out)
