import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
m = engine.B200Model(p, w, max_batch=B)
pk = m.pack_rows(rows)
dev = m.alloc_device(pk.nbytes); m.memcpy_h2d(dev, pk)
ob, oq = m.alloc_device(B * 120), m.alloc_device(B * 120)
def run(n, prof):
    m.set_profile(prof)
    tot = 0.0
    pend = None
    for i in range(n):
        t = m.submit_packed_raw(dev, B, 3, ob, oq)
        if pend is not None:
            m.wait_raw(pend); tot += m.last_forward_ms()
        pend = t
    m.wait_raw(pend); tot += m.last_forward_ms()
    pr = m.get_profile() if prof else None
    m.set_profile(False)
    return tot / n, pr
run(20, False)
import time
for prof in (False, True, False, True):
    t0 = time.perf_counter(); d, pr = run(100, prof); m.synchronize(); wall = (time.perf_counter() - t0) / 100 * 1e3
    ks = sum(v["ms"] for v in pr["kernels"].values()) / 100 if pr else None
    print("profile", prof, "wall/step %.4f  ev0->ev1 %.4f  sum of kernels %s" % (wall, d, "%.4f" % ks if ks else "-"))
