"""Cycle trace of embed_condense_kernel (needs a -DDCB_TRACE build; run with DCB_STACK=0 DCB_FUSE_QA=0 to keep later kernels
from overwriting the trace rows -- here the engine is stopped after the embedding by using 0 layers is not possible, so the
stack kernel's issuer rows (col 0-7 of even blocks) are overwritten; we read cols that only the embed kernel writes)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
libname = [a for a in sys.argv[1:] if a.endswith(".so")] or ["libdcb200_trace.so"]
lib = engine._load(os.path.join(os.path.dirname(engine.library_path()), libname[0]))
m = engine.B200Model(p, w, max_batch=B, library=lib)
packed = "--packed" in sys.argv
pk = m.pack_rows(rows)
for _ in range(3):
  m.forward_packed(pk) if packed else m.forward(rows)
print("packed rows" if packed else "float32 rows")
buf = (ctypes.c_uint64 * (256 * 16))()
lib.dcb_debug_trace(buf, 256 * 16)
a = np.array(buf[:], dtype=np.float64).reshape(256, 16)[1:148:2]   # odd blocks: the stack kernel's issuer writes even blocks only
names = ["builder_total", "ids_phase", "wait_a_empty", "build", "epi_wait_acc_full", "epi_body", "prologue (entry -> builders start)", "kernel total (entry -> exit)"]
for i, nme in enumerate(names):
    col = a[:, i]
    print("%-18s mean %10.0f  per tile %8.0f" % (nme, col.mean(), col.mean() / 7))
