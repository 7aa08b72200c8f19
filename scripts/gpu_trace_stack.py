"""Cycle trace of stack_pair_kernel's UMMA issuer (needs a -DDCB_TRACE build)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
LN = "--ln" in sys.argv      # pre-LayerNorm checkpoint-like config instead of C2
p = P.synthetic_params(20, 100, use_ccs_bq=True, num_hidden_layers=5, rezero=False) if LN else P.synthetic_params(20, 120)
w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
# DCB_OUT=libdcb200_trace.so DCB_EXTRA_FLAGS=-DDCB_TRACE bash deepconsensus_b200/csrc/build.sh
libname = [a for a in sys.argv[1:] if a.endswith(".so")] or ["libdcb200_trace.so"]
lib = engine._load(os.path.join(os.path.dirname(engine.library_path()), libname[0]))
m = engine.B200Model(p, w, max_batch=B, library=lib)
for _ in range(3): m.forward(rows)
print("device ms", m.last_ms)
buf = (ctypes.c_uint64 * (256 * 16))()
lib.dcb_debug_trace(buf, 256 * 16)
a = np.array(buf[:], dtype=np.float64).reshape(256, 16)[:148:2]
names = ["total", "wait_a_ready_P1", "wait_acc_free", "qkv_issue", "wait_att_ready", "oproj_issue", "wait_a_ready_P5", "ffn_g1_loop",
         "W row passes (2/layer)", "W q/k/v staging incl. waits (2 heads)", "W attention: q frags + QK^T", "W softmax", "W PV",
         "W att_h store + arrive", "W wait s_free", "W hidden epilogue (16 chunks, excl. waits)"]
units = 7 * p.num_hidden_layers
for i, nme in enumerate(names):
    col = a[:, i]
    print("%-18s mean %10.0f  per tile-layer %8.0f   min %10.0f max %10.0f" % (nme, col.mean(), col.mean() / units, col.min(), col.max()))
