"""Small cases of every round-2 kernel path, meant to run under compute-sanitizer:
  compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py
(strict fp32 path, packed rows in the embedding kernel, device unpack, wide-window stack kernel incl. the cross-CTA halo,
post-model kernels: stitch_fastq / skip_mask / fill_skipped)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import calibration, engine, params as P, synthetic, weights as W

cal = calibration.parse_calibration_string("0,1.1,-0.5")
for (Pn, L, bq, layers, rezero) in ((20, 120, False, 2, True), (20, 100, True, 2, False), (8, 200, False, 2, True), (8, 136, True, 1, False)):
  p = P.synthetic_params(Pn, L, use_ccs_bq=bq, num_hidden_layers=layers, rezero=rezero)
  w = W.init_weights(p, seed=1)
  rows = synthetic.make_rows(p, 5, seed=2)
  m = engine.B200Model(p, w, max_batch=4, calibration=cal)
  a = m.forward(rows, want_logits=True)
  s = m.forward(rows, want_logits=True, strict=True)
  pk = m.pack_rows(rows)
  b = m.forward_packed(pk, want_logits=True)
  bs = m.forward_packed(pk, want_logits=True, strict=True)
  assert np.array_equal(a["logits"], b["logits"]) and np.array_equal(s["logits"], bs["logits"])
  print("P=%d L=%d bq=%s: launches %d, |default - strict| max %.4f" % (Pn, L, bq, m.last_launches, np.abs(a["logits"] - s["logits"]).max()))
  names = ["z/1/ccs"] * 2 + ["z/2/ccs"] * 3
  pos = [0, L, 0, L, 3 * L]
  fq, off, outc, avg = m.stitch_fastq(a["bases"], a["quals"], np.array([0, 2, 5], np.int32), pos, ["z/1/ccs", "z/2/ccs"], 0, 0, length=L)
  mask, av = m.skip_mask(np.random.default_rng(0).integers(-1, 94, size=(5, L)).astype(np.int16), 45)
  bb, qq = a["bases"].copy(), a["quals"].copy()
  m.fill_skipped(np.random.default_rng(1).integers(0, 5, size=(2, L)).astype(np.uint8), np.full((2, L), 30, np.int16), np.array([4, 1], np.int32), bb, qq, calibration=cal)
  print("  post-model: outcomes", outc.tolist(), "fastq bytes", len(fq), "skip mask", mask.tolist())
  m.close()
print("done")
