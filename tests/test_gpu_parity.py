"""Parity of the CUDA engine (through the C-ABI / ctypes binding) against the oracle.  -m gpu.

Tolerances (written here, per BASELINE.json north_star: "within fp32 logit tolerance (identical
argmax bases)"): the engine computes its tensor-core contractions with bf16 operands and fp32
accumulation, so
  * |logit - oracle_fp32|  <= LOGIT_TOL_FP32 (absolute) everywhere,
  * |logit - oracle_bf16|  <= LOGIT_TOL_EMU  where the oracle rounds at the same points,
  * bases identical at every position whose fp32-oracle top-2 logit margin exceeds MARGIN
    (positions inside the margin are near-ties that random-weight models produce in bulk),
  * quality characters within +-1 at those positions, and
  * the device epilogue is bit-exact given the device's own probabilities (integer/byte work).
"""
import ast
import os

import numpy as np
import pytest

from deepconsensus_b200 import calibration, params as params_lib, synthetic, weights as weights_lib
from oracle import model as omodel, postprocess as opost

pytestmark = pytest.mark.gpu

LOGIT_TOL_FP32 = 0.25
LOGIT_TOL_EMU = 0.20
MARGIN = 0.5
CAL = "0,1.197654,-0.99781"


@pytest.fixture(scope="module")
def engine_mod():
  from deepconsensus_b200 import engine
  engine.load_library()
  return engine


def _check(engine_mod, p, w, rows, cal_str=CAL, chunk_tiles=0, max_batch=None):
  cal = calibration.parse_calibration_string(cal_str)
  model = engine_mod.B200Model(p, w, max_batch=max_batch or rows.shape[0], calibration=cal, chunk_tiles=chunk_tiles)
  out = model.forward(rows, want_probs=True, want_logits=True, strict_input=False)
  launches = model.last_launches
  model.close()
  assert launches > 0
  cal_t = (cal.threshold, cal.w, cal.b) if cal.enabled else None
  ref = omodel.forward(rows, p, w)
  emu = omodel.forward(rows, p, w, emulate="bf16")
  assert np.isfinite(out["logits"]).all()
  assert np.abs(out["logits"] - ref["logits"]).max() <= LOGIT_TOL_FP32
  assert np.abs(out["logits"] - emu["logits"]).max() <= LOGIT_TOL_EMU
  assert np.abs(out["probs"].sum(-1) - 1).max() < 1e-5
  y, q = opost.quality_from_probs(ref["probs"], 93, cal_t)
  rb, rq = opost.to_ascii(y, q)
  srt = np.sort(ref["logits"], axis=-1)
  safe = (srt[..., -1] - srt[..., -2]) > MARGIN
  assert safe.mean() > 0.5
  assert (out["bases"][safe] == rb[safe]).all()
  assert np.abs(out["quals"].astype(int) - rq.astype(int))[safe].max() <= 1
  assert (out["bases"] == rb).mean() > 0.98
  # device epilogue is exact integer/byte work on the device's own probabilities
  yy, qq = opost.quality_from_probs(out["probs"], 93, cal_t)
  sb, sq = opost.to_ascii(yy, qq)
  assert np.array_equal(sb, out["bases"])
  assert (sq == out["quals"]).mean() > 0.999 and np.abs(sq.astype(int) - out["quals"].astype(int)).max() <= 1
  return out


def test_c2_shape_rezero(engine_mod):
  p = params_lib.synthetic_params(20, 120)
  _check(engine_mod, p, weights_lib.init_weights(p, seed=1), synthetic.make_rows(p, 9, seed=2))


def test_layernorm_bq_5_layers_L100(engine_mod):
  p = params_lib.synthetic_params(20, 100, use_ccs_bq=True, num_hidden_layers=5, rezero=False)
  _check(engine_mod, p, weights_lib.init_weights(p, seed=3), synthetic.make_rows(p, 7, seed=4), cal_str="10,0.9,1.5")


def test_c5_shape_P32_L200(engine_mod):
  p = params_lib.synthetic_params(32, 200)
  _check(engine_mod, p, weights_lib.init_weights(p, seed=5), synthetic.make_rows(p, 5, seed=6), cal_str="skip")


def test_full_attention_when_no_window(engine_mod):
  p = params_lib.synthetic_params(20, 100, attn_win_size=None, num_hidden_layers=2)
  _check(engine_mod, p, weights_lib.init_weights(p, seed=7), synthetic.make_rows(p, 3, seed=8))


def test_real_windows_from_reference_fixture(engine_mod, golden_dir):
  rows = np.load(os.path.join(golden_dir, "real_windows_human_1m.npz"))["rows"]
  p = params_lib.synthetic_params(20, 100)
  _check(engine_mod, p, weights_lib.init_weights(p, seed=9), rows)


def test_ragged_batches_chunks_and_determinism(engine_mod):
  p = params_lib.synthetic_params(20, 120, num_hidden_layers=2)
  w = weights_lib.init_weights(p, seed=10)
  rows = synthetic.make_rows(p, 37, seed=11)
  whole = _check(engine_mod, p, w, rows)
  model = engine_mod.B200Model(p, w, max_batch=16, chunk_tiles=3,     # 3 engine calls, several chunks each
                               calibration=calibration.parse_calibration_string(CAL))
  split = model.forward(rows, want_logits=True, strict_input=False)
  again = model.forward(rows, want_logits=True, strict_input=False)
  one = model.forward(rows[:1], want_logits=True, strict_input=False)
  model.close()
  assert np.array_equal(split["logits"], again["logits"])             # deterministic
  assert np.array_equal(split["bases"], whole["bases"]) and np.array_equal(split["quals"], whole["quals"])
  assert np.array_equal(split["logits"], whole["logits"])             # windows are independent units
  assert np.array_equal(one["logits"][0], whole["logits"][0])         # batch of 1


def test_out_of_range_input_is_flagged(engine_mod):
  p = params_lib.synthetic_params(20, 100, num_hidden_layers=1)
  w = weights_lib.init_weights(p, seed=12)
  rows = synthetic.make_rows(p, 2, seed=13)
  rows[0, 0, 5, 0] = 7.0                                               # base id 7 does not exist
  model = engine_mod.B200Model(p, w, max_batch=2)
  with pytest.raises(engine_mod.DcbError) as ei:
    model.forward(rows)
  assert ei.value.code == -5
  model.close()


def test_empty_batch_and_bad_shapes(engine_mod):
  p = params_lib.synthetic_params(20, 100, num_hidden_layers=1)
  model = engine_mod.B200Model(p, weights_lib.init_weights(p, seed=1), max_batch=4)
  out = model.forward(np.zeros((0, 85, 100, 1), np.float32))
  assert out["bases"].shape == (0, 100)
  with pytest.raises(ValueError):
    model.forward(np.zeros((1, 86, 100, 1), np.float32))
  model.close()


def test_run_model_on_examples_and_stitch(engine_mod, golden_dir):
  """The reference-facing call: feature dicts in, DCModelOutput list out, FASTQ via stitch_utils."""
  from deepconsensus_b200 import inference, stitch_utils
  z = np.load(os.path.join(golden_dir, "real_windows_human_1m.npz"))
  rows, names, pos = z["rows"], z["names"], z["window_pos"]
  p = params_lib.synthetic_params(20, 100)
  w = weights_lib.init_weights(p, seed=14)
  cal = calibration.parse_calibration_string(CAL)
  opts = inference.InferenceOptions(max_length=100, example_height=85, max_passes=20, min_quality=0, min_length=0,
                                    batch_size=24, use_ccs_bq=False, cpus=0, skip_windows_above=45,
                                    use_saved_model=False, max_base_quality=93, dc_calibration_values=cal,
                                    ccs_calibration_values=calibration.parse_calibration_string("skip"))
  model, p = inference.initialize_model("", p, opts, weights=w)
  fds = [dict(subreads=rows[i][..., None], **{"subreads/num_passes": 3}, window_pos=int(pos[i]), name=str(names[i]),
              ccs_base_quality_scores=np.zeros(100), ec=1.0, np_num_passes=3, rq=0.99, rg="rg") for i in range(len(rows))]
  preds = inference.run_model_on_examples(fds, model, p, opts)
  # fast path: the same windows, grouped by read and sorted by position, straight to FASTQ records with the byte work
  # on the device -- must equal stitch_to_fastq over the per-window objects, read for read
  order = sorted(range(len(fds)), key=lambda i: (fds[i]["name"], fds[i]["window_pos"]))
  cnt_fast = stitch_utils.OutcomeCounter()
  fast = inference.run_model_and_stitch([fds[i] for i in order], model, p, opts, cnt_fast)
  cnt_ref, slow, i = stitch_utils.OutcomeCounter(), [], 0
  while i < len(order):
    j = i
    while j < len(order) and fds[order[j]]["name"] == fds[order[i]]["name"]:
      j += 1
    slow.append(stitch_utils.stitch_to_fastq(fds[order[i]]["name"], [preds[k] for k in order[i:j]], 100, 0, 0, cnt_ref))
    i = j
  assert fast == slow and cnt_fast.__dict__ == cnt_ref.__dict__
  model.close()
  assert len(preds) == len(rows) and all(len(o.sequence) == 100 and len(o.quality_string) == 100 for o in preds)
  ref = omodel.forward(rows, p, w)
  y, q = opost.quality_from_probs(ref["probs"], 93, (cal.threshold, cal.w, cal.b))
  agree = np.mean([np.mean(np.frombuffer(o.sequence.encode(), np.uint8) == opost.to_ascii(y[i], q[i])[0]) for i, o in enumerate(preds)])
  assert agree > 0.98
  # windows of one ZMW, re-indexed contiguously, stitch into a FASTQ record
  first = str(names[0])
  mine = [o for o in preds if o.molecule_name == first]
  for k, o in enumerate(sorted(mine, key=lambda o: o.window_pos)):
    o.window_pos = k * 100
  cnt = stitch_utils.OutcomeCounter()
  fq = stitch_utils.stitch_to_fastq(first, sorted(mine, key=lambda o: o.window_pos), 100, 0, 0, cnt)
  assert fq is not None and fq.startswith("@" + first + "\n") and cnt.success == 1


def test_unfused_fallback_paths_agree_with_fused(engine_mod):
  """DCB_STACK / DCB_FUSE_HEAD / DCB_FUSE_OPROJ / DCB_FUSE_EMBED / DCB_FUSE_QA / DCB_ALIGN / DCB_FFN_PAIR select measured alternatives of the same math;
  they are read when an engine is created, so flip them around model construction."""
  p = params_lib.synthetic_params(20, 120, num_hidden_layers=2)
  w = weights_lib.init_weights(p, seed=21)
  rows = synthetic.make_rows(p, 5, seed=22)
  ref = omodel.forward(rows, p, w)["logits"]
  outs = {}
  for name, env in (("fused", {}),                                     # default: whole stack in one kernel
                    ("per_layer", {"DCB_STACK": "0"}),                  # QKV+attention and out-proj+FFN kernels per layer
                    ("fused_head", {"DCB_FUSE_HEAD": "1"}),             # head in the tail of the stack kernel
                    ("unfused", {"DCB_FUSE_OPROJ": "0", "DCB_FUSE_EMBED": "0", "DCB_FUSE_QA": "0"}),
                    ("packed", {"DCB_ALIGN": "0"}),                     # windows packed back to back, separate QKV / attention
                    ("single_cta", {"DCB_FFN_PAIR": "0", "DCB_FUSE_QA": "0"}),
                    ("qkv2", {"DCB_FUSE_QA": "0", "DCB_QKV2": "1"})):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
      model = engine_mod.B200Model(p, w, max_batch=8)
      outs[name] = model.forward(rows, want_logits=True, strict_input=False)["logits"]
      model.close()
    finally:
      for k, v in old.items():
        if v is None:
          os.environ.pop(k, None)
        else:
          os.environ[k] = v
    assert np.abs(outs[name] - ref).max() <= LOGIT_TOL_FP32, name
  assert np.abs(outs["fused"] - outs["per_layer"]).max() < 0.05
  assert np.abs(outs["fused"] - outs["fused_head"]).max() < 1e-3
  assert np.abs(outs["fused"] - outs["unfused"]).max() < 0.05
  assert np.abs(outs["fused"] - outs["packed"]).max() < 0.05
  assert np.abs(outs["qkv2"] - outs["unfused"]).max() < 0.05


@pytest.mark.parametrize("name", ["rezero_p20", "layernorm_p20", "rezero_p20_bq", "layernorm_p20_bq", "rezero_p5_win3"])
def test_engine_against_reference_code_goldens(engine_mod, golden_dir, name):
  """CUDA path vs outputs of the reference's OWN model code (tests/golden/ref_model_*.npz, generated by
  scripts/make_model_golden.py) -- no oracle in between."""
  z = np.load(os.path.join(golden_dir, "ref_model_%s.npz" % name))
  p = params_lib.get_config(str(z["config"]))
  for k, v in ast.literal_eval(str(z["overrides"])).items():
    p[k] = v
  params_lib.modify_params(p, max_length=int(z["max_length"]))
  w = weights_lib.init_weights(p, seed=int(z["seed"]))
  rows = z["rows"]
  model = engine_mod.B200Model(p, w, max_batch=rows.shape[0])
  out = model.forward(rows, want_probs=True, want_logits=True, strict_input=False)
  model.close()
  assert np.abs(out["logits"] - z["logits"]).max() <= LOGIT_TOL_FP32
  srt = np.sort(z["logits"], axis=-1)
  safe = (srt[..., -1] - srt[..., -2]) > MARGIN
  y, q = opost.quality_from_probs(z["probs"], 93, None)
  rb, rq = opost.to_ascii(y, q)
  assert (out["bases"][safe] == rb[safe]).all()
  assert (out["bases"] == rb).mean() > 0.98
  assert np.abs(out["probs"] - z["probs"]).max() < 0.05


def test_pipelined_submit_wait_matches_blocking_forward(engine_mod):
  """dcb_submit / dcb_wait (two batches in flight, H2D of batch i+1 under the kernels of batch i) returns exactly what
  dcb_forward returns, keeps order, reports per-ticket input errors and refuses a third outstanding submission."""
  p = params_lib.synthetic_params(20, 120)
  w = weights_lib.init_weights(p, seed=31)
  model = engine_mod.B200Model(p, w, max_batch=16)
  batches = [synthetic.make_rows(p, n, seed=40 + i) for i, n in enumerate((16, 7, 16, 1, 12))]
  blocking = [model.forward(b, want_probs=True) for b in batches]
  piped = list(model.forward_batches(batches, want_probs=True))
  assert len(piped) == len(blocking)
  for a, b in zip(blocking, piped):
    assert np.array_equal(a["bases"], b["bases"]) and np.array_equal(a["quals"], b["quals"])
    assert np.array_equal(a["probs"], b["probs"])
  # a bad batch between two good ones: only its own ticket reports the range error
  bad = batches[1].copy()
  bad[0, 0, 0] = 9.0
  h0 = model.submit(batches[0])
  h1 = model.submit(bad)
  with pytest.raises(engine_mod.DcbError):
    model.submit(batches[2])                      # two already in flight
  o0 = model.wait(h0)
  with pytest.raises(engine_mod.DcbError):
    model.wait(h1)
  with pytest.raises(engine_mod.DcbError):
    model.wait(h1)                                # not in flight any more
  h2 = model.submit(batches[2])
  o2 = model.wait(h2)
  assert np.array_equal(o0["bases"], blocking[0]["bases"]) and np.array_equal(o2["quals"], blocking[2]["quals"])
  model.close()


def test_full_size_properties_c2_batch_1024(engine_mod):
  """BASELINE configs[1] at its full size (1024 windows, 20 x 120, 6 layers): too big for the oracle in test time, so the
  CUDA path is checked through size-independent properties -- windows are independent units, hence
  (a) permuting the batch permutes the outputs bit-exactly, (b) any sub-batch reproduces its rows of the full batch
  bit-exactly (different tile -> SM assignment, different pairing), (c) probabilities sum to one, (d) the device epilogue
  (argmax / Phred / calibration / ASCII) is exact integer work on the device's own probabilities, (e) a 16-window slice
  agrees with the oracle."""
  p = params_lib.synthetic_params(20, 120)
  w = weights_lib.init_weights(p, seed=101)
  B = 1024
  rows = synthetic.make_rows(p, B, seed=102)
  cal = calibration.parse_calibration_string(CAL)
  model = engine_mod.B200Model(p, w, max_batch=B, calibration=cal)
  full = model.forward(rows, want_probs=True, strict_input=False)
  again = model.forward(rows, want_probs=True, strict_input=False)
  assert np.array_equal(full["probs"], again["probs"]) and np.array_equal(full["quals"], again["quals"])
  rng = np.random.default_rng(7)
  perm = rng.permutation(B)
  permuted = model.forward(rows[perm], want_probs=True, strict_input=False)
  assert np.array_equal(permuted["probs"], full["probs"][perm])
  assert np.array_equal(permuted["bases"], full["bases"][perm]) and np.array_equal(permuted["quals"], full["quals"][perm])
  for lo, hi in ((0, 1), (5, 12), (300, 811), (1023, 1024)):
    sub = model.forward(rows[lo:hi], want_probs=True, strict_input=False)
    assert np.array_equal(sub["probs"], full["probs"][lo:hi]), (lo, hi)
    assert np.array_equal(sub["bases"], full["bases"][lo:hi]) and np.array_equal(sub["quals"], full["quals"][lo:hi])
  assert np.isfinite(full["probs"]).all() and np.abs(full["probs"].sum(-1) - 1).max() < 1e-5
  y, q = opost.quality_from_probs(full["probs"], 93, (cal.threshold, cal.w, cal.b))
  sb, sq = opost.to_ascii(y, q)
  assert np.array_equal(sb, full["bases"])
  assert (sq == full["quals"]).mean() > 0.999 and np.abs(sq.astype(int) - full["quals"].astype(int)).max() <= 1
  ref = omodel.forward(rows[500:516], p, w)
  model.close()
  m2 = engine_mod.B200Model(p, w, max_batch=16, calibration=cal)
  o16 = m2.forward(rows[500:516], want_probs=True, want_logits=True, strict_input=False)
  m2.close()
  assert np.array_equal(o16["probs"], full["probs"][500:516])
  assert np.abs(o16["logits"] - ref["logits"]).max() <= LOGIT_TOL_FP32


@pytest.mark.parametrize("layers,ff,rezero,win,L,B", [
    (1, 128, True, 12, 120, 3),      # one layer, one hidden chunk: the FFN stage program never reaches the tail slots
    (3, 640, False, 16, 128, 4),     # pre-LN, band at the two-pass limit, window exactly one tile
    (8, 256, True, 1, 64, 5),        # deepest stack the one-kernel path takes, narrowest band, short windows
    (9, 256, True, 12, 100, 2),      # deeper than kMaxLayers: falls back to the per-layer kernels
])
def test_stack_kernel_corner_shapes(engine_mod, layers, ff, rezero, win, L, B):
  p = params_lib.synthetic_params(20, L, num_hidden_layers=layers, rezero=rezero, attn_win_size=win)
  p.filter_size = ff
  w = weights_lib.init_weights(p, seed=50 + layers)
  rows = synthetic.make_rows(p, B, seed=60 + layers)
  model = engine_mod.B200Model(p, w, max_batch=B)
  out = model.forward(rows, want_logits=True, strict_input=False)
  launches = model.last_launches
  model.close()
  assert launches == (3 if layers <= 8 else 2 + 2 * layers)
  ref = omodel.forward(rows, p, w)
  assert np.isfinite(out["logits"]).all()
  assert np.abs(out["logits"] - ref["logits"]).max() <= LOGIT_TOL_FP32


# ----------------------------------------------------------------------------------------------------------------
# SURVEY section 8(f) "next": stitching.  dcb_stitch does get_full_sequence + remove_gaps on the device (bytes: exact).
def _tiny_model(engine_mod, L=100):
  p = params_lib.synthetic_params(20, L, num_hidden_layers=1)
  return engine_mod.B200Model(p, weights_lib.init_weights(p, seed=1), max_batch=4), p


def test_device_stitch_against_executed_reference(engine_mod, golden_dir):
  """All 120 cases produced by executing the reference's stitch_utils (tests/golden/ref_stitch.json), stitched in
  batches on the device: FASTQ records and outcome counters must be identical, read for read."""
  import json
  from deepconsensus_b200 import stitch_gpu, stitch_utils
  with open(os.path.join(golden_dir, "ref_stitch.json")) as f:
    cases = json.load(f)["cases"]
  model, _ = _tiny_model(engine_mod)
  groups = {}
  for c in cases:
    groups.setdefault((c["max_length"], c["min_quality"], c["min_length"]), []).append(c)
  checked = 0
  for (L, min_q, min_len), cs in groups.items():
    bases, quals, names, pos = [], [], [], []
    for i, c in enumerate(cs):
      for w in c["windows"]:
        if w["dropped"]:
          continue
        assert len(w["sequence"]) == L and len(w["quality_string"]) == L
        bases.append(np.frombuffer(w["sequence"].encode("latin-1"), np.uint8))
        quals.append(np.frombuffer(w["quality_string"].encode("latin-1"), np.uint8))
        names.append("%s#%d" % (c["name"], i))          # reads with equal names in different cases stay separate
        pos.append(w["window_pos"])
    cnt = stitch_utils.OutcomeCounter()
    got = stitch_gpu.stitch_batch_to_fastq(model, np.stack(bases), np.stack(quals), names, pos, L, min_q, min_len, cnt)
    want_cnt = stitch_utils.OutcomeCounter()
    j = 0
    for i, c in enumerate(cs):
      if not any(not w["dropped"] for w in c["windows"]):
        continue                                          # a read with no windows at all never reaches the batch
      want = c["fastq"]
      if want is not None:
        want = want.replace("@" + c["name"] + "\n", "@%s#%d\n" % (c["name"], i), 1)
      assert got[j] == want, (c["name"], i)
      for k, v in c["counter"].items():
        setattr(want_cnt, k, getattr(want_cnt, k) + v)
      j += 1
      checked += 1
    assert j == len(got)
    assert cnt.__dict__ == want_cnt.__dict__
  assert checked >= 100
  model.close()


def test_device_stitch_chain_from_forward_outputs(engine_mod):
  """forward (outputs left on the device) -> dcb_stitch on those device buffers == forward to host -> Python mirror
  of stitch_to_fastq, for reads of ragged window counts, including one with a missing window and empty inputs."""
  from deepconsensus_b200 import stitch_gpu, stitch_utils
  model, p = _tiny_model(engine_mod, L=100)
  model.close()
  p = params_lib.synthetic_params(20, 100, num_hidden_layers=2)
  w = weights_lib.init_weights(p, seed=77)
  B, L = 37, 100
  model = engine_mod.B200Model(p, w, max_batch=B)
  rows = synthetic.make_rows(p, B, seed=78)
  host = model.forward(rows, strict_input=False)
  counts = [1, 5, 2, 9, 3, 7, 10]
  assert sum(counts) == B
  names, pos = [], []
  for z, n in enumerate(counts):
    for i in range(n):
      names.append("m/%d/ccs" % z)
      pos.append(i * L if not (z == 3 and i >= 4) else (i + 1) * L)     # read 3 misses its 5th window
  # device chain
  dev_rows = model.alloc_device(rows.nbytes)
  model.memcpy_h2d(dev_rows, rows[..., 0])
  db, dq = model.alloc_device(B * L), model.alloc_device(B * L)
  model.forward_raw(dev_rows, B, engine_mod.DCB_ROWS_ON_DEVICE | engine_mod.DCB_OUT_ON_DEVICE, db, dq)
  cnt = stitch_utils.OutcomeCounter()
  got = stitch_gpu.stitch_batch_to_fastq(model, db, dq, names, pos, L, 0, 0, cnt, n_windows=B, on_device=True)
  # Python mirror, read by read
  want, want_cnt, k = [], stitch_utils.OutcomeCounter(), 0
  for z, n in enumerate(counts):
    preds = []
    for i in range(n):
      o = stitch_utils.DCModelOutput(names[k], pos[k], 1.0, 3, 0.99, "rg")
      o.sequence = host["bases"][k].tobytes().decode("ascii")
      o.quality_string = host["quals"][k].tobytes().decode("ascii")
      preds.append(o)
      k += 1
    want.append(stitch_utils.stitch_to_fastq(names[k - 1], preds, L, 0, 0, want_cnt))
  assert got == want and cnt.__dict__ == want_cnt.__dict__
  assert want[3] is None and cnt.empty_sequence == 1 and cnt.success >= 5
  # degenerate inputs
  s, q, l = model.stitch(np.zeros((0, L), np.uint8), np.zeros((0, L), np.uint8), np.array([0], np.int32))
  assert l.shape == (0,)
  allgap = np.full((2, L), ord(" "), np.uint8)
  s, q, l = model.stitch(allgap, allgap, np.array([0, 2], np.int32))
  assert l.tolist() == [0]
  for d in (dev_rows, db, dq):
    model.free_device(d)
  model.close()
