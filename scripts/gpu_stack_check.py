"""First-light / regression check of stack_pair_kernel against the per-layer kernels and the oracle (GPU, diagnostics)."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # layers, B, rezero, L, bq
    (1, 2, True, 120, False), (2, 5, True, 120, False), (2, 3, False, 100, True), (6, 300, True, 120, False),
    (5, 301, False, 100, True)]

def child(idx, stack):
  from deepconsensus_b200 import params as P, weights as W, synthetic, engine
  nl, B, rz, L, bq = CASES[idx]
  os.environ["DCB_STACK"] = str(stack)
  p = P.synthetic_params(20, L, num_hidden_layers=nl, rezero=rz, use_ccs_bq=bq)
  w = W.init_weights(p, seed=5 + idx)
  rows = synthetic.make_rows(p, B, seed=77 + idx)
  m = engine.B200Model(p, w, max_batch=B)
  out = m.forward(rows, want_logits=True, strict_input=False)
  out2 = out if os.environ.get("STACK_WATCHDOG") else m.forward(rows, want_logits=True, strict_input=False)
  print("case", idx, "stack", stack, "launches", m.last_launches, "ms %.3f" % m.last_ms,
        "deterministic", bool(np.array_equal(out["logits"], out2["logits"])), flush=True)
  if os.environ.get("STACK_WATCHDOG"):
    import ctypes
    lib = engine.load_library()
    buf = (ctypes.c_uint64 * (256 * 16))()
    lib.dcb_debug_trace(buf, 256 * 16)
    a = np.array(buf[:], dtype=np.uint64).reshape(256, 16)
    for b in range(4):
      for wp in range(12):
        v = int(a[b, wp])
        if v >> 63:
          print("   block %d warp %2d stuck: tag %4d parity %d after %d cycles" % (b, wp, v & 0xffff, (v >> 16) & 0xff, (v >> 24) & 0xffffffff))
  np.save("gpurun_out/stack_case%d_s%d.npy" % (idx, stack), out["logits"])
  m.close()

if __name__ == "__main__":
  os.makedirs("gpurun_out", exist_ok=True)
  if len(sys.argv) == 3:
    child(int(sys.argv[1]), int(sys.argv[2]))
    sys.exit(0)
  from deepconsensus_b200 import params as P, weights as W, synthetic
  from oracle import model as omodel
  for idx in range(len(CASES)):
    ok = True
    for stack in (0, 1):
      r = subprocess.run(["timeout", "60", sys.executable, __file__, str(idx), str(stack)], capture_output=True, text=True)
      print(r.stdout.strip(), ("| rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])) if r.returncode else "", flush=True)
      ok = ok and r.returncode == 0
    if not ok:
      print("case", idx, "FAILED to run; stopping"); break
    a = np.load("gpurun_out/stack_case%d_s0.npy" % idx); b = np.load("gpurun_out/stack_case%d_s1.npy" % idx)
    nl, B, rz, L, bq = CASES[idx]
    p = P.synthetic_params(20, L, num_hidden_layers=nl, rezero=rz, use_ccs_bq=bq)
    w = W.init_weights(p, seed=5 + idx)
    rows = synthetic.make_rows(p, min(B, 16), seed=77 + idx) if False else synthetic.make_rows(p, B, seed=77 + idx)[:16]
    ref = omodel.forward(rows, p, w)["logits"]
    print("  case %d: max|stack - per_layer| = %.4e   max|stack - oracle| = %.4e  max|per_layer - oracle| = %.4e  nan=%d" %
          (idx, np.abs(a - b).max(), np.abs(b[:16] - ref).max(), np.abs(a[:16] - ref).max(), int(np.isnan(b).sum())), flush=True)
    if not np.isfinite(b).all() or np.abs(a - b).max() > 0.1:
      d = np.abs(a - b).reshape(B, L, 5).max(-1)
      bad = np.argwhere(d > 0.1)
      print("   bad windows:", sorted(set(bad[:, 0].tolist()))[:20], "bad positions (first window):", bad[bad[:, 0] == bad[0, 0]][:, 1].tolist()[:40])
