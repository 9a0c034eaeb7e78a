"""recall@10 / QPS over ef on one index (BASELINE.md 3's generators):   python scripts/ef_sweep.py clustered|gauss|latent [n]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from nucliadb_b200 import _lib  # noqa: E402
from nucliadb_b200.segment import VectorSegment  # noqa: E402


class A:
    data, latent, noise = sys.argv[1], 16, 0.15


n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
d, k, nq = 768, 10, 1024
dev = torch.device("cuda", 0)
vecs = B.make_vectors(A, n, d, dev, seed=1234567890)
qs = [B.gen_queries(vecs, nq, seed=123 + i) for i in range(8)]
seg = VectorSegment.create(vecs, d, similarity=_lib.NIDX_SIM_COSINE, m=16, m0=32, ef_construction=200, device=0)
del vecs
seg.build_hnsw(seed=2, max_batch=8192)
torch.cuda.synchronize()
gt = seg.search(qs[0], k, method=_lib.NIDX_METHOD_BRUTE)[0].cpu().numpy()
out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq,), dtype=torch.int32, device=dev))
for ef in (30, 64, 128, 256, 512, 1024):
    try:
        for i in range(2):
            seg.search(qs[i], k, ef=ef, method=_lib.NIDX_METHOD_HNSW, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(8):
            seg.search(qs[i], k, ef=ef, method=_lib.NIDX_METHOD_HNSW, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        ids = seg.search(qs[0], k, ef=ef, method=_lib.NIDX_METHOD_HNSW)[0].cpu().numpy()
        c = seg.counters()
        print(json.dumps({"data": A.data, "n": n, "ef": ef, "recall_at_10": B.recall_at_k(ids, gt), "qps": nq / ms * 1e3, "ms_per_batch": ms,
                          "similarities_per_query": c["similarities"] / nq, "overflows": c["overflows"]}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"data": A.data, "ef": ef, "error": str(e)}), flush=True)
