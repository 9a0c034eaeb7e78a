"""How the CPU port's throughput scales with threads on the GPU box's host (is os.cpu_count() real?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from bench import gen_vectors, gen_queries, export_graph_for_oracle
from nucliadb_b200 import _lib
from nucliadb_b200.segment import VectorSegment
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
dev = torch.device("cuda", 0)
n, d = 2_000_000, 768
v = gen_vectors(n, d, dev, 1, 16, 0.15); q = gen_queries(v, 2048, 5)
seg = VectorSegment.create(v, d, similarity=_lib.NIDX_SIM_COSINE, m=16, m0=32, ef_construction=200)
seg.build_hnsw(2, 8192)
hv, hq = v.cpu().numpy(), q.cpu().numpy()
og = export_graph_for_oracle(seg, O, n, 16, 32)
O.build(native=True)
norms = O.norms(hv, nthreads=64)
for t in (1, 8, 32, 64, 128):
    nq = min(2048, 64 * t)
    O.hnsw_search(hv, og, hq[:nq], 10, 128, nthreads=t, native=True, norms_=norms)
    t0 = time.perf_counter(); O.hnsw_search(hv, og, hq[:nq], 10, 128, nthreads=t, native=True, norms_=norms); dt = time.perf_counter() - t0
    print("threads", t, "queries", nq, "qps", round(nq / dt, 1), "per-thread qps", round(nq / dt / t, 2), flush=True)
