"""Experiment: per-launch time of the quantised HNSW walk, visited-table size sweep.   python scripts/exp_rq.py [n]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gen_queries, gen_vectors, recall_at_k  # noqa: E402
from nucliadb_b200 import _lib  # noqa: E402
from nucliadb_b200.segment import VectorSegment  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, nq, k = 768, 1024, 10
dev = torch.device("cuda", 0)
vecs = gen_vectors(n, d, dev, seed=1234567890, latent=16, noise=0.15)
queries = [gen_queries(vecs, nq, seed=123 + i) for i in range(6)]
seg = VectorSegment.create(vecs, d, similarity=_lib.NIDX_SIM_DOT, m=16, m0=32, ef_construction=200)
del vecs
seg.build_hnsw(seed=2, max_batch=8192)
seg.rabitq_encode()
gt = seg.search(queries[0], k, method=_lib.NIDX_METHOD_BRUTE)[0].cpu().numpy()
out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq,), dtype=torch.int32, device=dev))
for bits, rqw, pf in (("", "8", "1"), ("", "8", "0"), ("", "4", "1"), ("", "4", "0"), ("17", "4", "0")):
    os.environ["NIDX_B200_RQ_W"], os.environ["NIDX_B200_RQ_PREFETCH"] = rqw, pf
    if bits:
        os.environ["NIDX_B200_RQ_VISITED_BITS"] = bits
    times = []
    for i in range(6):
        seg.search(queries[i], k, method=_lib.NIDX_METHOD_HNSW_RABITQ, out=out)
        torch.cuda.synchronize()
        times.append(round(seg.last_kernel_ms(), 3))
    seg.search(queries[0], k, method=_lib.NIDX_METHOD_HNSW_RABITQ, out=out)
    c = seg.counters_ex()
    print(json.dumps({"prefetch": pf, "warps": rqw, "visited_bits": bits or "default(16)", "kernel_ms": times, "recall": recall_at_k(out[0].cpu().numpy(), gt), "overflows": c["overflows"],
                      "estimates_per_q": c["estimates"] / nq, "expansions_per_q": c["expansions"] / nq}), flush=True)
