"""ORACLE -- test infrastructure.  Post-model arithmetic of `run_model_on_examples`.

Restates quick_inference.py:377-389 (argmax / error prob / Phred / calibration /
clip / round / int / floor) and :390-414 (string building) with NumPy, keeping the
reference's dtypes: softmax output float32, `1 - max` and `-10*log10` in float32,
calibration float32 when threshold == 0 and float64 otherwise
(calibration_lib.py:89-99), np.round half-to-even.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

SEQ_VOCAB = " ATCG"   # dc_constants.py:39-41


def quality_from_probs(probs: np.ndarray, max_base_quality: int = 93,
                       calibration: Optional[Tuple[float, float, float]] = None,
                       log10: str = "libm") -> Tuple[np.ndarray, np.ndarray]:
  """probs [B,L,5] float32 -> (y_preds int64 [B,L], quality int32 [B,L]).

  log10="libm": `np.log10` on the float32 array, as the reference writes it -- the platform's float32 log10 (glibc /
  SVML: <= 1 ulp, not always correctly rounded, so the value is platform-dependent in its last bit).
  log10="exact": the correctly rounded float32 log10 (float64 log10 rounded once) -- the platform-independent
  definition the device epilogue implements (csrc/head_finish.cuh); differs from "libm" only where the platform's
  float32 log10 is off by an ulp AND that ulp crosses a rounding boundary of the final integer.
  """
  probs = np.asarray(probs, dtype=np.float32)
  y_preds = np.argmax(probs, -1)                         # :377
  error_prob = 1 - np.max(probs, axis=-1)                # :378 (float32)
  with np.errstate(divide="ignore"):
    if log10 == "exact":
      q = np.float32(-10) * np.log10(error_prob.astype(np.float64)).astype(np.float32)
    else:
      q = -10 * np.log10(error_prob)                     # :379 (float32; inf when p == 1)
  if calibration is not None:                            # :380-383
    thr, w, b = calibration
    if thr == 0:
      q = q * w + b
    else:
      q = q * np.where(q > thr, w, 1.0) + np.where(q > thr, b, 0.0)
  q = np.minimum(q, max_base_quality)                    # :385
  q = np.round(q, decimals=0)                            # :386
  q = q.astype(dtype=np.int32)                           # :387
  q = np.maximum(q, 0)                                   # :389
  return y_preds, q


def to_strings(y_pred: np.ndarray, quality: np.ndarray) -> Tuple[str, str]:
  """One window: ids -> ' ATCG' string, scores -> Phred+33 string (:408-411, utils.py:60-62)."""
  seq = "".join(SEQ_VOCAB[int(i)] for i in y_pred)
  qual = "".join(chr(int(s) + 33) for s in quality)
  return seq, qual


def to_ascii(y_preds: np.ndarray, quality: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
  """Batch form of to_strings: uint8 [B,L] base characters and Phred+33 characters."""
  vocab = np.frombuffer(SEQ_VOCAB.encode(), dtype=np.uint8)
  return vocab[y_preds], (quality + 33).astype(np.uint8)
