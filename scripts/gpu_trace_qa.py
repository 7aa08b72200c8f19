"""Cycle trace of qkv_attn_pair_kernel (needs -DDCB_TRACE): runs ONLY that kernel last by using a 1-layer model."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
m = engine.B200Model(p, w, max_batch=B)
for _ in range(2): m.forward(rows)
lib = engine.load_library()
buf = (ctypes.c_uint64 * (256 * 16))()
lib.dcb_debug_trace(buf, 256 * 16)
a = np.array(buf[:], dtype=np.float64).reshape(256, 16)[148:256]
lead = a[a[:, 0] > 0]
rounds = lead[:, 7].max()
print("rounds", rounds, "tiles/CTA; heads", rounds * 2)
for i, n in [(0, "mma_total"), (1, "mma_wait_afull"), (2, "mma_wait_accfree"), (3, "mma_wait_full"), (4, "mma_issue")]:
  print("%-18s mean %9.0f per-head %7.0f" % (n, lead[:, i].mean(), lead[:, i].mean() / (rounds * 2)))
wk = a[a[:, 10] > 0]
for i, n in [(8, "wrk_wait_accfull"), (9, "wrk_epilogue"), (10, "wrk_attention")]:
  print("%-18s mean %9.0f per-head %7.0f" % (n, wk[:, i].mean(), wk[:, i].mean() / (rounds * 2)))
