"""The N>1 exchange path on CPU: world_size 2 over gloo.  Checks the all_gather plumbing and the global
id mapping of nucliadb_b200.dist against a numpy restatement of shard_merge.rs:332-348 (the CUDA merge
kernel itself is covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nucliadb_b200.dist import gather_partials, global_ids

    rng = np.random.default_rng(100 + rank)
    nq, k = 7, 5
    scores = np.sort(rng.random((nq, k)).astype(np.float32), axis=1)[:, ::-1].copy()
    ids = rng.integers(0, 1000, (nq, k)).astype(np.int32)
    ids_all, sc_all = gather_partials(torch.from_numpy(ids), torch.from_numpy(scores))
    assert ids_all.shape == (world, nq, k)
    # numpy restatement of merge_vector_responses: kmerge_by(score >=) of per-rank lists, take k
    merged_ids, merged_part = [], []
    for q in range(nq):
        items = sorted(((-float(sc_all[r, q, j]), r, j) for r in range(world) for j in range(k)))[:k]
        merged_ids.append([int(ids_all[r, q, j]) for _, r, j in items])
        merged_part.append([r for _, r, j in items])
    g = global_ids(torch.tensor(merged_ids, dtype=torch.int32), torch.tensor(merged_part, dtype=torch.int32), 1000)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gather_and_global_ids(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert (a == b).all()                      # every rank ends with the same merged answer
    assert ((a >= 0) & (a < 2000)).all() and (a >= 1000).any() and (a < 1000).any()
