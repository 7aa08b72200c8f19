"""GPU diagnostic: per-stage comparison of the CUDA engine against the oracle.

Run on a B200 (gpurun).  Test/diagnostic infrastructure: imports oracle/.
Writes a report to gpurun_out/diag.txt (and stdout).
"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import params as P, weights as W, synthetic, engine
from deepconsensus_b200 import calibration as C
from oracle import model as omodel, postprocess as opost

out_dir = "gpurun_out"
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "diag.txt"), "a")
def say(*a):
  s = " ".join(str(x) for x in a)
  print(s, flush=True); log.write(s + "\n"); log.flush()

def stage_names(nl):
  names = ["embedded"]
  for n in range(nl):
    names += ["attn_%d" % n, "ffn_%d" % n]
  return names

def run_case(tag, p, B, seed, calib="0,1.197654,-0.99781"):
  say("=== case", tag, "B=%d L=%d P=%d layers=%d rezero=%s bq=%s" % (B, p.max_length, p.max_passes, p.num_hidden_layers, p.rezero, p.use_ccs_bq))
  w = W.init_weights(p, seed=seed)
  rows = synthetic.make_rows(p, B, seed=seed + 100)
  cal = C.parse_calibration_string(calib)
  m = engine.B200Model(p, w, max_batch=max(B, 1), calibration=cal)
  nodebug = bool(os.environ.get('DIAG_NODEBUG'))
  if not nodebug:
    m.set_debug(True)
  t = time.time()
  out = m.forward(rows, want_probs=True, want_logits=True, strict_input=False)
  say("forward ok: wall %.3fs device %.3f ms launches %d" % (time.time() - t, m.last_ms, m.last_launches))
  ref = omodel.forward(rows, p, w, emulate="bf16", return_intermediates=True)
  ref32 = omodel.forward(rows, p, w, emulate=None)
  L = p.max_length
  # residual stream of the last chunk
  ntok = B * L
  try:
    for si, name in enumerate(stage_names(p.num_hidden_layers) if not nodebug else []):
      got = m.debug_residual(si, ntok) if ntok <= 200000 else None
      want = ref["intermediates"][name].reshape(-1, 280)[-got.shape[0]:]
      err = np.abs(got - want)
      say("  stage %-9s max|d|=%.4e mean|d|=%.3e  ref absmax=%.3f  nan=%d" % (name, err.max(), err.mean(), np.abs(want).max(), int(np.isnan(got).sum())))
      if err.max() > 0.25 and si <= 2:
        bad = np.argwhere(err > 0.25)
        say("    first bad (tok,col):", bad[:8].tolist(), "rows bad:", len(set(bad[:,0].tolist())), "cols bad:", len(set(bad[:,1].tolist())))
  except Exception as ex:
    say("  debug_residual failed:", repr(ex))
  dl = np.abs(out["logits"] - ref["logits"]).max()
  dl32 = np.abs(out["logits"] - ref32["logits"]).max()
  say("  logits max|gpu-oracle_bf16|=%.4e  max|gpu-oracle_fp32|=%.4e" % (dl, dl32))
  cal_t = (cal.threshold, cal.w, cal.b) if cal.enabled else None
  y, q = opost.quality_from_probs(ref["probs"], 93, cal_t)
  y32, q32 = opost.quality_from_probs(ref32["probs"], 93, cal_t)
  ob, oq = opost.to_ascii(y, q)
  ob32, oq32 = opost.to_ascii(y32, q32)
  say("  bases match vs bf16-oracle %.5f, vs fp32-oracle %.5f" % ((out["bases"] == ob).mean(), (out["bases"] == ob32).mean()))
  dq = np.abs(out["quals"].astype(int) - oq.astype(int)); dq32 = np.abs(out["quals"].astype(int) - oq32.astype(int))
  say("  quals exact vs bf16-oracle %.5f (max d %d), vs fp32-oracle %.5f (max d %d)" % ((dq == 0).mean(), dq.max(), (dq32 == 0).mean(), dq32.max()))
  # self-consistency of the device epilogue: recompute bases/quals from the device's own probs
  yy, qq = opost.quality_from_probs(out["probs"], 93, cal_t)
  sb, sq = opost.to_ascii(yy, qq)
  say("  epilogue self-check: bases %.6f quals %.6f (max d %d)" % ((sb == out["bases"]).mean(), (sq == out["quals"]).mean(), np.abs(sq.astype(int) - out["quals"].astype(int)).max()))
  m.close()
  return dict(tag=tag, logits_err=float(dl), logits_err32=float(dl32))

if __name__ == "__main__":
  which = sys.argv[1:] or ["small"]
  res = []
  if "small" in which:
    res.append(run_case("small-rezero", P.synthetic_params(20, 120), 3, 1))
  if "variants" in which:
    res.append(run_case("ln-bq-5L-L100", P.synthetic_params(20, 100, use_ccs_bq=True, num_hidden_layers=5, rezero=False), 5, 2))
    res.append(run_case("P32-L200", P.synthetic_params(32, 200), 4, 3, calib="skip"))
    res.append(run_case("multi-chunk", P.synthetic_params(20, 120), 700, 4))
  if "perf" in which:
    p = P.synthetic_params(20, 120)
    w = W.init_weights(p, seed=1)
    B = 1024
    rows = synthetic.make_rows(p, B, seed=7)
    m = engine.B200Model(p, w, max_batch=B)
    for i in range(5):
      m.forward(rows)
      say("perf B=%d: device %.3f ms -> %.0f windows/s (launches %d)" % (B, m.last_ms, B / m.last_ms * 1e3, m.last_launches))
    m.close()
  say(json.dumps(res))
