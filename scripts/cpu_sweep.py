import os, sys, time
sys.path.insert(0, os.getcwd())
import bench, torch
from deepconsensus_b200 import weights as W
p = bench.model_params(); w = W.init_weights(p, seed=1)
for t in (16, 32, 48, 64):
    v, secs = bench.cpu_reference_windows_per_sec(p, w, sample_windows=128, reps=3, threads=t)
    print("threads", t, "windows/s %.1f" % v, "sec/pass %.2f" % secs, flush=True)
for t, sw in ((64, 256), (32, 256)):
    v, secs = bench.cpu_reference_windows_per_sec(p, w, sample_windows=sw, reps=2, threads=t)
    print("threads", t, "sample", sw, "windows/s %.1f" % v, "sec/pass %.2f" % secs, flush=True)
