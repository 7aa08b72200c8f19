"""Attention kernel cycle trace (needs a -DDCB_TRACE build)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
p = P.synthetic_params(20, 120); w = W.init_weights(p, seed=1)
B = 1024
rows = synthetic.make_rows(p, B, seed=7)
m = engine.B200Model(p, w, max_batch=B)
for _ in range(3): m.forward(rows)
lib = engine.load_library()
buf = (ctypes.c_uint64 * (256 * 16))()
lib.dcb_debug_trace(buf, 256 * 16)
a = np.array(buf[:], dtype=np.float64).reshape(256, 16)
print("attention CTAs 0..255 (first wave): staging wait cycles mean %.0f min %.0f max %.0f" % (a[:, 13].mean(), a[:, 13].min(), a[:, 13].max()))
print("compute cycles mean %.0f min %.0f max %.0f" % (a[:, 14].mean(), a[:, 14].min(), a[:, 14].max()))
print("distinct SMs among first 256 CTAs:", len(set(a[:, 15].astype(int).tolist())))
m.set_profile(True)
for _ in range(5): m.forward(rows)
print(m.get_profile()["kernels"])
