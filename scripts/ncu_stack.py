"""Short workload for ncu captures: a few forwards of the bench workload (C2, or C5 with --c5) with resident inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
c5 = "--c5" in sys.argv
packed = "--packed" in sys.argv
p = P.synthetic_params(32, 200) if c5 else P.synthetic_params(20, 120)
B = 2048 if c5 else 1024
w = W.init_weights(p, seed=1)
rows = synthetic.make_rows(p, B, seed=7)
m = engine.B200Model(p, w, max_batch=B)
L = p.max_length
if packed:
  pk = m.pack_rows(rows)
  dev = m.alloc_device(pk.nbytes); m.memcpy_h2d(dev, pk)
else:
  dev = m.alloc_device(rows.nbytes); m.memcpy_h2d(dev, rows[..., 0])
ob, oq = m.alloc_device(B * L), m.alloc_device(B * L)
for _ in range(4):
  if packed:
    t = m.submit_packed_raw(dev, B, 3, ob, oq); m.wait_raw(t)
  else:
    m.forward_raw(dev, B, 3, ob, oq)
print("device ms", m.last_forward_ms())
