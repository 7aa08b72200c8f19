"""Race evidence for the mbarrier / tcgen05.commit hand-overs.

libdcb200_jitter.so is the same source built with -DDCB_JITTER: a pseudo-random __nanosleep (25 % of the time, up to
~2 us) in front of every mbarrier wait / arrive, bulk-copy issue and tcgen05.commit (csrc/sm100.cuh).  That shifts
producer, UMMA-issuer, relay and worker warps against each other differently on every run.  This script scores the same
batches N times with the jittered library and requires every output byte and every logit to be identical to the plain
build's -- for the one-kernel stack path and (developer switches) the per-layer path.

  DCB_OUT=libdcb200_jitter.so DCB_EXTRA_FLAGS="-DDCB_JITTER -DDCB_DEV_SWITCHES" bash deepconsensus_b200/csrc/build.sh   (here)
  gpurun -- python scripts/gpu_jitter.py [runs]                                                                     (GPU)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import engine, params as P, synthetic, weights as W  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
jit = engine._load(os.path.join(os.path.dirname(engine.library_path()), "libdcb200_jitter.so"))
report = []
for tag, env, kw in (("stack kernel (rezero)", {}, {}),
                     ("stack kernel (pre-LN, CCS-BQ, 5 layers, L=100)", {}, dict(use_ccs_bq=True, rezero=False, num_hidden_layers=5, L=100)),
                     ("per-layer kernels", {"DCB_STACK": "0"}, {})):
  kw = dict(kw)
  L = kw.pop("L", 120)
  p = P.synthetic_params(20, L, **kw)
  w = W.init_weights(p, seed=7)
  for B in (1, 37, 1024):
    rows = synthetic.make_rows(p, B, seed=100 + B)
    plain = engine.B200Model(p, w, max_batch=B)
    want = plain.forward(rows, want_logits=True)
    plain.close()
    os.environ.update(env)
    m = engine.B200Model(p, w, max_batch=B, library=jit)
    for k in env:
      os.environ.pop(k)
    n = runs if B < 1024 else max(20, runs // 4)
    t0 = time.time()
    bad = 0
    ms = []
    pk = m.pack_rows(rows)
    for i in range(n):
      # the one-kernel path alternates between float32 rows and packed rows (staged by bulk copy in the embedding kernel)
      got = m.forward_packed(pk, want_logits=True) if (not env and i % 2) else m.forward(rows, want_logits=True)
      ms.append(m.last_ms)
      if env:
        ok = np.array_equal(got["logits"], first["logits"]) if i else True     # per-layer path: self-consistency
        if i == 0:
          first = got
      else:
        ok = (np.array_equal(got["logits"], want["logits"]) and np.array_equal(got["bases"], want["bases"]) and
              np.array_equal(got["quals"], want["quals"]))
      bad += not ok
    m.close()
    rec = dict(path=tag, batch=B, runs=n, mismatching_runs=bad, jittered_ms_min=round(min(ms), 3),
               jittered_ms_max=round(max(ms), 3), seconds=round(time.time() - t0, 1))
    print(json.dumps(rec), flush=True)
    report.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/r02_jitter.json", "w") as f:
  json.dump(report, f, indent=1)
if any(r["mismatching_runs"] for r in report):
  sys.exit("JITTER RUNS DIFFER")
print("all jittered runs bit-identical")
