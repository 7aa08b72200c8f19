"""Thread sweep of the CPU arm (oracle, torch-CPU fp32) on the GPU box's host: which torch thread count scores the
bench workload fastest.  Writes gpurun_out/r02_cpu_sweep.json; commit it as profiles/r02_cpu_sweep.json -- bench.py's
cpu_threads() uses its `best_threads` for `--impl reference` and `cpu_baseline` (else: every available core).

  gpurun -- python scripts/cpu_sweep.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepconsensus_b200 import weights as W  # noqa: E402

p = bench.model_params()
w = W.init_weights(p, seed=1)
avail = len(os.sched_getaffinity(0))
counts = sorted({t for t in (8, 16, 32, 48, 64, 96, 128, avail) if t <= avail})
rows = []
for t in counts:
  v, secs = bench.cpu_reference_windows_per_sec(p, w, sample_windows=256, reps=2, threads=t)
  rows.append(dict(threads=t, windows_per_s=round(v, 1), sec_per_pass=round(secs, 2), sample_windows=256))
  print(rows[-1], flush=True)
best = max(rows, key=lambda r: r["windows_per_s"])
v, secs = bench.cpu_reference_windows_per_sec(p, w, sample_windows=1024, reps=1, threads=best["threads"])
out = dict(host_cores=avail, best_threads=best["threads"], sweep=rows,
           full_batch=dict(threads=best["threads"], windows_per_s=round(v, 1), sec_per_pass=round(secs, 2), sample_windows=1024),
           workload=bench.WORKLOAD)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/r02_cpu_sweep.json", "w") as f:
  json.dump(out, f, indent=1)
print(json.dumps(out))
