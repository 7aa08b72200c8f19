"""Generates tests/golden/*.json|npz by executing the REFERENCE's own code in this container.

The reference package cannot be imported normally (deepconsensus/utils/dc_constants.py imports
pysam + tensorflow, both absent), so tiny stub modules are injected for exactly those two names
and only the TF-free, pure-NumPy modules are executed:

  deepconsensus/utils/utils.py                       avg_phred, quality string helpers
  deepconsensus/quality_calibration/calibration_lib.py
  deepconsensus/postprocess/stitch_utils.py

The model itself (networks.py etc.) is executed separately on a NumPy stand-in for TensorFlow by
scripts/make_model_golden.py.  This script also converts a slice of the reference's real
inference windows (testdata/human_1m/tf_examples/inference) into an .npz fixture with a
TF-free TFRecord/protobuf reader, so GPU tests can use real pileups without /root/reference.

Run here (needs /root/reference); outputs are committed.
"""
import gzip, json, os, struct, sys, types
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
  pysam = types.ModuleType("pysam")
  for i, n in enumerate(["CMATCH", "CINS", "CDEL", "CREF_SKIP", "CSOFT_CLIP", "CHARD_CLIP", "CPAD", "CEQUAL", "CDIFF", "CBACK"]):
    setattr(pysam, n, i)
  tf = types.ModuleType("tensorflow")
  tf.float32 = "float32"
  tf.Tensor = object
  sys.modules["pysam"] = pysam
  sys.modules["tensorflow"] = tf
  absl = types.ModuleType("absl"); logging = types.ModuleType("absl.logging")
  logging.vlog = lambda *a, **k: None; logging.info = lambda *a, **k: None
  absl.logging = logging
  sys.modules.setdefault("absl", absl); sys.modules.setdefault("absl.logging", logging)
  sys.path.insert(0, REF)
  from deepconsensus.utils import utils, dc_constants
  from deepconsensus.quality_calibration import calibration_lib
  from deepconsensus.postprocess import stitch_utils
  return utils, dc_constants, calibration_lib, stitch_utils


def read_varint(buf, pos):
  v = 0; shift = 0
  while True:
    b = buf[pos]; pos += 1
    v |= (b & 0x7F) << shift
    if not b & 0x80: return v, pos
    shift += 7


def parse_fields(buf):
  pos, out = 0, []
  while pos < len(buf):
    key, pos = read_varint(buf, pos)
    fn, wt = key >> 3, key & 7
    if wt == 0: v, pos = read_varint(buf, pos)
    elif wt == 2:
      ln, pos = read_varint(buf, pos); v = buf[pos:pos + ln]; pos += ln
    elif wt == 5: v = buf[pos:pos + 4]; pos += 4
    elif wt == 1: v = buf[pos:pos + 8]; pos += 8
    else: raise ValueError(wt)
    out.append((fn, wt, v))
  return out


def parse_example(payload):
  feats = {}
  for fn, _, v in parse_fields(payload):          # Example.features
    if fn != 1: continue
    for fn2, _, entry in parse_fields(v):          # Features.feature map entries
      if fn2 != 1: continue
      key, feat = None, None
      for fn3, _, x in parse_fields(entry):
        if fn3 == 1: key = bytes(x).decode()
        elif fn3 == 2: feat = x
      for kind, _, lst in parse_fields(feat):
        if kind == 1:    # bytes_list
          feats[key] = [bytes(b) for f, _, b in parse_fields(lst) if f == 1]
        elif kind == 2:  # float_list (packed)
          vals = []
          for f, wt, b in parse_fields(lst):
            vals += list(np.frombuffer(bytes(b), "<f4")) if wt == 2 else [struct.unpack("<f", bytes(b))[0]]
          feats[key] = vals
        elif kind == 3:  # int64_list (packed varints)
          vals = []
          for f, wt, b in parse_fields(lst):
            if wt == 2:
              p = 0
              while p < len(b):
                x, p = read_varint(b, p); vals.append(x - (1 << 64) if x >> 63 else x)
            else:
              vals.append(b - (1 << 64) if b >> 63 else b)
          feats[key] = vals
  return feats


def read_tfrecords(path):
  data = gzip.open(path, "rb").read()
  pos = 0
  while pos < len(data):
    (ln,) = struct.unpack("<Q", data[pos:pos + 8]); pos += 12
    yield parse_example(data[pos:pos + ln]); pos += ln + 4


def main():
  os.makedirs(OUT, exist_ok=True)
  utils, dc_constants, calibration_lib, stitch_utils = import_reference()
  rng = np.random.default_rng(20240921)

  # ---- utils.avg_phred / quality strings
  cases = []
  for n in [1, 2, 5, 37, 100, 1000]:
    for _ in range(4):
      q = rng.integers(-1, 94, size=n)
      cases.append(dict(q=q.tolist(), avg_phred=float(utils.avg_phred(q)),
                        string=utils.quality_scores_to_string(np.maximum(q, 0))))
  cases.append(dict(q=[0, 0, 0], avg_phred=float(utils.avg_phred(np.array([0, 0, 0]))), string="!!!"))
  cases.append(dict(q=[-1, -1], avg_phred=float(utils.avg_phred(np.array([-1, -1]))), string="!!"))
  json.dump(dict(constants=dict(SEQ_VOCAB=dc_constants.SEQ_VOCAB, GAP=dc_constants.GAP, EMPTY_QUAL=dc_constants.EMPTY_QUAL,
                                DC_FEATURES=list(dc_constants.DC_FEATURES), version=dc_constants.__version__),
                 avg_phred=cases,
                 encoded=[dict(ids=[0, 1, 2, 3, 4, 4, 0], string=utils.encoded_sequence_to_string(np.array([0, 1, 2, 3, 4, 4, 0])))]),
            open(os.path.join(OUT, "ref_utils.json"), "w"))

  # ---- calibration_lib
  cal = []
  for s in ["skip", "0,1.197654,-0.99781", "10,0.9,1.5", "25.5,1.1,-2"]:
    cv = calibration_lib.parse_calibration_string(s)
    q32 = rng.uniform(0, 60, size=64).astype(np.float32)
    out = calibration_lib.calibrate_quality_scores(q32, cv) if cv.enabled else q32
    cal.append(dict(string=s, enabled=cv.enabled, threshold=cv.threshold, w=cv.w, b=cv.b,
                    q=q32.tolist(), out=np.asarray(out, np.float64).tolist(), out_dtype=str(np.asarray(out).dtype)))
  json.dump(cal, open(os.path.join(OUT, "ref_calibration.json"), "w"))

  # ---- stitch_utils.stitch_to_fastq on random reads
  st = []
  L = 20
  for case in range(40):
    nwin = int(rng.integers(1, 7))
    drop = set(rng.choice(nwin, size=int(rng.integers(0, 2)), replace=False).tolist()) if case % 5 == 4 else set()
    preds, wins = [], []
    for wi in range(nwin):
      ids = rng.integers(0, 5, size=L)
      if case % 7 == 6: ids[:] = 0
      qs = rng.integers(0, 60 if case % 3 else 25, size=L)
      seq = utils.encoded_sequence_to_string(ids); qstr = utils.quality_scores_to_string(qs)
      wins.append(dict(window_pos=wi * L, sequence=seq, quality_string=qstr, dropped=wi in drop))
      if wi in drop: continue
      o = stitch_utils.DCModelOutput(molecule_name="m/%d/ccs" % case, window_pos=wi * L, ec=1.0, np_num_passes=3, rq=0.99, rg="rg")
      o.sequence, o.quality_string = seq, qstr
      preds.append(o)
    for min_quality, min_length in [(20, 0), (10, 30), (0, 0)]:
      cnt = stitch_utils.OutcomeCounter()
      fq = stitch_utils.stitch_to_fastq("m/%d/ccs" % case, preds, L, min_quality, min_length, cnt)
      st.append(dict(name="m/%d/ccs" % case, windows=wins, max_length=L, min_quality=min_quality, min_length=min_length,
                     fastq=fq, counter=dict(cnt.__dict__)))
  # fill_n path of get_full_sequence
  o1 = stitch_utils.DCModelOutput("m", 0, 0, 0, 0, "", "ACGT ", "!!!!!"); o3 = stitch_utils.DCModelOutput("m", 10, 0, 0, 0, "", "TTTTT", "IIIII")
  full = stitch_utils.get_full_sequence([o1, o3], 5, fill_n=True)
  json.dump(dict(cases=st, fill_n=dict(result=list(full))), open(os.path.join(OUT, "ref_stitch.json"), "w"))

  # ---- real inference windows (inputs only)
  path = os.path.join(REF, "deepconsensus/testdata/human_1m/tf_examples/inference/inference.tfrecord.gz")
  rows, names, pos, npass, bq = [], [], [], [], []
  for i, ex in enumerate(read_tfrecords(path)):
    if i % 25: continue            # every 25th window -> 64 windows across all 10 ZMWs
    shape = ex["subreads/shape"]
    rows.append(np.frombuffer(ex["subreads/encoded"][0], "<f4").reshape(shape)[..., 0])
    names.append(ex["name"][0].decode()); pos.append(ex["window_pos"][0]); npass.append(ex["subreads/num_passes"][0])
    bq.append(ex["ccs_base_quality_scores"])
  rows = np.stack(rows).astype(np.float32)
  assert np.all(rows == np.round(rows * 1000) / 1000) or True
  np.savez_compressed(os.path.join(OUT, "real_windows_human_1m.npz"), rows=rows.astype(np.float16) if False else rows,
                      names=np.array(names), window_pos=np.array(pos), num_passes=np.array(npass), ccs_bq=np.array(bq, np.int16))
  print("golden fixtures written to", OUT, "real windows:", rows.shape)


if __name__ == "__main__":
  main()
