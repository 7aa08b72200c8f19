"""Default-path error against the reference-code goldens for one or two builds of the library (debug aid).

usage: python scripts/gpu_golden_stats.py [libA.so [libB.so]]   (file names under deepconsensus_b200/csrc)
"""
import ast, glob, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import engine, params as params_lib, weights as weights_lib

names = sys.argv[1:] or ["libdcb200.so"]
libs = [engine._load(os.path.join(os.path.dirname(engine.library_path()), n)) for n in names]
for f in sorted(glob.glob("tests/golden/ref_model_*.npz")):
  z = np.load(f)
  p = params_lib.get_config(str(z["config"]))
  for k, v in ast.literal_eval(str(z["overrides"])).items():
    p[k] = v
  params_lib.modify_params(p, max_length=int(z["max_length"]))
  w = weights_lib.init_weights(p, seed=int(z["seed"]))
  rows = z["rows"]
  rec = dict(golden=os.path.basename(f), positions=int(rows.shape[0] * p.max_length), rezero=bool(p.rezero))
  for n, lib in zip(names, libs):
    m = engine.B200Model(p, w, max_batch=rows.shape[0], library=lib)
    out = m.forward(rows, want_logits=True)
    m.close()
    d = out["logits"] - z["logits"]
    rec[n] = dict(max=float(np.abs(d).max()), rms=float(np.sqrt((d * d).mean())),
                  base_mismatches=int((out["logits"].argmax(-1) != z["logits"].argmax(-1)).sum()))
  print(json.dumps(rec), flush=True)
