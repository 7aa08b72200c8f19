"""FFN cycle trace (needs a -DDCB_TRACE build)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
m = engine.B200Model(p, w, max_batch=B)
for _ in range(3): m.forward(rows)
print("device ms", m.last_ms)
lib = engine.load_library()
buf = (ctypes.c_uint64 * (256 * 16))()
lib.dcb_debug_trace(buf, 256 * 16)
a = np.array(buf[:], dtype=np.float64).reshape(256, 16)[:148]
a = a[a[:, 0] > 0]
names = ["mma_total", "mma_wait_hfree", "mma_wait_full", "mma_wait_hsfull", "mma_issue", "mma_wait_afull", "mma_wait_yempty", "mma_wait_a2full",
         "hepi_wait_hfull", "hepi_wait_hsfree", "hepi_body", "row_wait_yfull", "row_body", "row_ldtm_in_body", "row_phaseA", "mma_oproj_loop"]
rounds = float(os.environ.get('TRACE_ROUNDS', '7'))
print("rounds", rounds, "chunks", rounds * 16)
for i, nme in enumerate(names):
    col = a[:, i]
    print("%-18s mean %10.0f  per-chunk %8.0f   min %10.0f max %10.0f" % (nme, col.mean(), col.mean() / (rounds * 16), col.min(), col.max()))
