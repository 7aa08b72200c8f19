"""Host-side statistics that travel with every throughput number (deepconsensus_b200/parity.py) and the bench line's
config contract (bench.py): no GPU."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepconsensus_b200 import parity  # noqa: E402


def _outputs(logits):
  b = logits.argmax(-1).astype(np.uint8)
  return dict(bases=b, quals=np.full(b.shape, 40, np.uint8), logits=logits.astype(np.float32))


def test_compare_counts_mismatches_margins_and_logit_errors():
  rng = np.random.default_rng(0)
  ref = rng.normal(size=(4, 50, 5)).astype(np.float32)
  test = ref.copy()
  test[0, 0] = np.roll(ref[0, 0], 1)                        # a gross error on one position
  st = parity.compare(_outputs(test), _outputs(ref), margin=1e-3)
  assert st["positions"] == 200 and st["base_mismatches"] == 1 and st["base_mismatches_outside_margin"] == 1
  assert st["bases_identical_pct"] == 99.5 and st["qv_exact_pct"] == 100.0 and st["max_dq"] == 0
  assert st["max_logit_err"] > 0 and st["largest_margin_of_a_mismatch"] == float(parity.top2_margin(ref)[0, 0])


def test_expected_flips_follows_the_margin_distribution():
  """expected_flips = sum over positions of P(N(0, 2 rms^2) > margin): near-ties flip, clear calls do not."""
  rng = np.random.default_rng(1)
  n = 4000
  ref = np.zeros((1, n, 5), np.float32)
  margins = np.concatenate([np.full(n // 2, 0.005, np.float32), np.full(n - n // 2, 5.0, np.float32)])
  ref[0, :, 1] = margins                                    # class 1 leads class 0 by `margin`
  ref[0, :, 2:] = -10.0
  noise = rng.normal(scale=0.01, size=ref.shape).astype(np.float32)
  test = ref + noise
  st = parity.compare(_outputs(test), _outputs(ref), margin=0.25)
  rms = st["rms_logit_err"]
  sd = math.sqrt(2.0) * rms
  want = (n // 2) * 0.5 * math.erfc(0.005 / (sd * math.sqrt(2.0)))
  assert abs(st["expected_flips"] - want) < 1e-3 * want + 1e-9
  # the observed count is what the model predicts (binomial spread), and none of the clear calls flipped
  assert abs(st["base_mismatches"] - st["expected_flips"]) < 5 * math.sqrt(st["expected_flips"])
  assert st["base_mismatches_outside_margin"] == 0 and st["largest_margin_of_a_mismatch"] <= 0.0051


def test_bench_config_is_the_same_for_both_arms_and_states_the_l2_policy():
  import bench
  a, b = bench.config_dict(2, 1024), bench.config_dict(2, 1024)
  assert a == b and a["global_batch"] == 2048 and a["batch_per_gpu"] == 1024
  assert "workload" in a and "L2" in a["l2"] and "model" not in a
