"""Runs the other BASELINE.json configurations once for the record (windows/s, parity spot check)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
from oracle import model as omodel

def run(tag, p, B, reps=5, check=8):
  w = W.init_weights(p, seed=1)
  rows = synthetic.make_rows(p, B, seed=3)
  m = engine.B200Model(p, w, max_batch=B)
  out = m.forward(rows, want_logits=True, strict_input=False)
  ref = omodel.forward(rows[:check], p, w)["logits"]
  err = float(np.abs(out["logits"][:check] - ref).max())
  dev = m.alloc_device(rows.nbytes); m.memcpy_h2d(dev, rows[..., 0])
  ob, oq = m.alloc_device(B * p.max_length), m.alloc_device(B * p.max_length)
  ts = []
  for _ in range(reps):
    m.forward_raw(dev, B, 3, ob, oq); ts.append(m.last_forward_ms())
  ms = float(np.median(ts))
  print(json.dumps(dict(config=tag, batch=B, L=p.max_length, P=p.max_passes, layers=p.num_hidden_layers, rezero=bool(p.rezero), bq=bool(p.use_ccs_bq),
                        device_ms=ms, windows_per_s=B / ms * 1e3, max_logit_err_vs_fp32_oracle=err)), flush=True)
  m.close()

run("C2 20x120 B=1024", P.synthetic_params(20, 120), 1024)
run("C3 ckpt-like L=100 bq LN 5L B=4096", P.synthetic_params(20, 100, use_ccs_bq=True, num_hidden_layers=5, rezero=False), 4096)
run("C3' L=100 rezero 6L B=4096", P.synthetic_params(20, 100), 4096)
run("C5 32x200 B=8192", P.synthetic_params(32, 200), 8192, reps=3)
