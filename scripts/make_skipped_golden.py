"""Generates tests/golden/ref_skipped.json by EXECUTING the reference's own `process_skipped_window` and the skip loop
of `inference_on_n_zmws` (deepconsensus/inference/quick_inference.py:567-594, 657-676).

quick_inference.py cannot be imported here (tensorflow, pysam, absl flags ...), so the function's source text is cut out
of the file with `ast` and exec'd, unmodified, in a namespace that holds the reference's own TF-free modules
(utils, calibration_lib, stitch_utils, dc_constants -- imported as scripts/make_golden.py does) and a stand-in for
`data_providers.get_indices` that is itself executed from data_providers.py the same way.
Run here (needs /root/reference); the output is committed.
"""
import ast, dataclasses, json, os, sys, textwrap
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402  (stubs pysam / tensorflow, imports the reference's pure modules)

REF = "/root/reference/deepconsensus"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_skipped.json")


def cut(path, name):
  src = open(path).read()
  for node in ast.walk(ast.parse(src)):
    if isinstance(node, ast.FunctionDef) and node.name == name:
      return "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
  raise KeyError(name)


def main():
  utils, dc_constants, calibration_lib, stitch_utils = make_golden.import_reference()
  ns = dict(np=np, utils=utils, calibration_lib=calibration_lib, stitch_utils=stitch_utils, dc_constants=dc_constants,
            Optional=object, Union=object)
  import typing
  ns.update({k: getattr(typing, k) for k in ('Dict', 'Any', 'Tuple', 'List', 'Iterable', 'Sequence', 'Optional', 'Union', 'Callable')})
  dp = dict(ns)
  for fn in ("get_total_rows", "get_indices"):
    exec(cut(os.path.join(REF, "models/data_providers.py"), fn), dp)
  ns["data_providers"] = type("dp", (), dict(get_indices=staticmethod(dp["get_indices"])))
  ns["InferenceOptions"] = object
  exec(cut(os.path.join(REF, "inference/quick_inference.py"), "process_skipped_window"), ns)
  # the skip loop is a fragment of inference_on_n_zmws: cut its lines verbatim and wrap them in a function
  src = open(os.path.join(REF, "inference/quick_inference.py")).read().splitlines()
  a = next(i for i, l in enumerate(src) if l.strip() == "feature_dicts_for_model = []")
  b = next(i for i, l in enumerate(src) if l.strip().startswith("time_to_skip ="))
  body = textwrap.dedent("\n".join(src[a:b]))
  exec("def skip_loop(feature_dicts_for_zmws, options):\n" + textwrap.indent(body, "  ") +
       "\n  return feature_dicts_for_model, predictions_for_skipped_windows\n", ns)

  @dataclasses.dataclass
  class Opt:
    max_passes: int
    use_ccs_bq: bool
    max_base_quality: int
    skip_windows_above: int
    ccs_calibration_values: object

  rng = np.random.default_rng(20240921)
  cases = []
  for ci in range(40):
    P = int(rng.choice([5, 20]))
    bq = bool(rng.integers(0, 2))
    L = int(rng.choice([20, 40]))
    R = 4 * P + 5 + int(bq)
    cal_str = str(rng.choice(["skip", "0,1.197654,-0.99781", "10,0.9,1.5"]))
    cal = calibration_lib.parse_calibration_string(cal_str)
    opt = Opt(P, bq, int(rng.choice([93, 40])), int(rng.choice([0, 30, 45])), cal)
    zmws, flat = [], []
    for z in range(3):
      wins = []
      for k in range(int(rng.integers(1, 4))):
        rows = np.zeros((R, L, 1), np.float32)
        rows[4 * P, :, 0] = rng.integers(0, 5, L)
        hi = bool(rng.integers(0, 2))
        ccs_q = rng.integers(35 if hi else 0, 94, L).astype(np.float32 if rng.integers(0, 2) else np.int64)
        if rng.random() < 0.3:
          ccs_q[rng.integers(0, L, 3)] = -1            # spacing / gap positions
        w = dict(subreads=rows, ccs_base_quality_scores=ccs_q, window_pos=k * L, name="m/%d/ccs" % z, ec=float(rng.random() * 20),
                 np_num_passes=int(rng.integers(1, 30)), rq=float(rng.random()), rg="rg%d" % z, overflow=bool(rng.random() < 0.2))
        wins.append(w)
        flat.append(w)
      zmws.append(wins)
    for_model, skipped = ns["skip_loop"](zmws, opt)
    cases.append(dict(
        options=dict(max_passes=P, use_ccs_bq=bq, max_base_quality=opt.max_base_quality, skip_windows_above=opt.skip_windows_above,
                     ccs_calibration=cal_str),
        L=L, windows=[dict(zmw=w["name"], window_pos=w["window_pos"], ccs_row=w["subreads"][4 * P, :, 0].astype(int).tolist(),
                           ccs_q=[float(x) for x in w["ccs_base_quality_scores"]], ccs_q_dtype=str(w["ccs_base_quality_scores"].dtype),
                           ec=w["ec"], np_num_passes=w["np_num_passes"], rq=w["rq"], rg=w["rg"], overflow=w["overflow"]) for w in flat],
        for_model=[[w["name"], w["window_pos"]] for w in for_model],
        skipped=[dict(molecule_name=o.molecule_name, window_pos=o.window_pos, sequence=o.sequence,
                      quality_string=o.quality_string, ec=o.ec, np_num_passes=o.np_num_passes, rq=o.rq, rg=o.rg) for o in skipped]))
  with open(OUT, "w") as f:
    json.dump(dict(cases=cases), f)
  print("wrote", OUT, len(cases), "cases;", sum(len(c["skipped"]) for c in cases), "skipped windows,",
        sum(len(c["for_model"]) for c in cases), "for the model")


if __name__ == "__main__":
  main()
