"""A/B of two builds of the library on the same box: device time of single forwards, alternating between the builds.

usage: python scripts/gpu_ab.py libdcb200_main.so libdcb200.so [--packed]   (file names under deepconsensus_b200/csrc)
--packed: packed rows as the resident input (include/dcb200.h), with per-kernel CUDA-event times
"""
import json, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine

CONFIGS = [
    ("C2 20x120 rezero 6L", P.synthetic_params(20, 120), 1024),
    ("C3 L=100 bq LN 5L", P.synthetic_params(20, 100, use_ccs_bq=True, num_hidden_layers=5, rezero=False), 4096),
    ("C5 32x200 rezero 6L", P.synthetic_params(32, 200), 2048),
]
PACKED = "--packed" in sys.argv
libs = [engine._load(os.path.join(os.path.dirname(engine.library_path()), n)) for n in sys.argv[1:3]]
for tag, p, B in CONFIGS:
  w = W.init_weights(p, seed=1)
  rows = synthetic.make_rows(p, B, seed=3)
  models, bufs, logits = [], [], []
  for lib in libs:
    m = engine.B200Model(p, w, max_batch=B, library=lib)
    logits.append(m.forward(rows, want_logits=True, strict_input=False)["logits"])
    if PACKED:
      pk = m.pack_rows(rows)
      dev = m.alloc_device(pk.nbytes); m.memcpy_h2d(dev, pk)
    else:
      dev = m.alloc_device(rows.nbytes); m.memcpy_h2d(dev, rows[..., 0])
    bufs.append((dev, m.alloc_device(B * p.max_length), m.alloc_device(B * p.max_length)))
    models.append(m)
  ts = [[], []]
  for rep in range(12):
    for i, m in enumerate(models):
      if PACKED:
        m.wait_raw(m.submit_packed_raw(bufs[i][0], B, 3, bufs[i][1], bufs[i][2]))
      else:
        m.forward_raw(bufs[i][0], B, 3, bufs[i][1], bufs[i][2])
      if rep >= 2: ts[i].append(m.last_forward_ms())
  kern = []
  if PACKED:
    for i, m in enumerate(models):
      m.set_profile(True)
      for _ in range(5): m.wait_raw(m.submit_packed_raw(bufs[i][0], B, 3, bufs[i][1], bufs[i][2]))
      pr = m.get_profile(); m.set_profile(False)
      kern.append({k: round(v["ms"] / 5, 4) for k, v in pr["kernels"].items() if v["ms"] > 0})
  clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks_throttle_reasons.active", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip()
  print(json.dumps(dict(config=tag, batch=B, ms=[float(np.median(t)) for t in ts], libs=sys.argv[1:3],
                        max_logit_diff=float(np.abs(logits[0] - logits[1]).max()), kernels_ms=kern, clocks=clk)), flush=True)
  for m in models: m.close()
