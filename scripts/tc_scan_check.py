import os, sys, numpy as np, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from conftest import make_vectors, make_queries
from nucliadb_b200 import _lib
from nucliadb_b200.segment import VectorSegment
for (n, d, nq) in ((5000, 128, 200), (100000, 384, 1000), (30000, 768, 300)):
    v = make_vectors(n, d, seed=3); q = make_queries(v, nq)
    for sim in (_lib.NIDX_SIM_COSINE, _lib.NIDX_SIM_DOT):
        seg = VectorSegment.create(v, d, similarity=sim)
        os.environ['NIDX_B200_SCAN'] = 'exact'
        t=time.time(); ei, es, ec = seg.search(q, 10, method=_lib.NIDX_METHOD_BRUTE); te=time.time()-t
        os.environ['NIDX_B200_SCAN'] = 'tensor'
        t=time.time(); ti, ts, tc = seg.search(q, 10, method=_lib.NIDX_METHOD_BRUTE); tt=time.time()-t
        kms = seg.last_kernel_ms()
        print(n, d, nq, 'sim', sim, 'ids equal frac', float((ei == ti).mean()), 'max |dscore|', float(np.abs(es - ts).max()), 'exact s', round(te,4), 'tensor s', round(tt,4), 'tc kernel ms', round(kms,3), flush=True)
