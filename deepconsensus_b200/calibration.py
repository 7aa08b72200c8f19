"""Base-quality calibration (mirror of `quality_calibration/calibration_lib.py:35-99`).

The device epilogue applies the same linear map (csrc/head_finish.cuh, csrc/post_kernels.cu); this module
is the host-side parser plus a NumPy implementation for skipped windows and tests.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class QualityCalibrationValues:
  """enabled / threshold / w / b, as in calibration_lib.py:35-50."""
  enabled: bool
  threshold: float
  w: float
  b: float


def parse_calibration_string(calibration: str) -> QualityCalibrationValues:
  """'skip' or 'threshold,w,b' (calibration_lib.py:52-75)."""
  if calibration == "skip":
    return QualityCalibrationValues(enabled=False, threshold=0.0, w=1.0, b=0.0)
  fields = calibration.split(",")
  if len(fields) != 3:
    raise ValueError(
        ("Malformed calibration string. Expected 3 values (or set "
         'to "skip" to perform no quality calibration).'), calibration)
  t, w, b = (float(x) for x in fields)
  return QualityCalibrationValues(enabled=True, threshold=t, w=w, b=b)


def calibrate_quality_scores(quality_scores: np.ndarray,
                             calibration_values: QualityCalibrationValues) -> np.ndarray:
  """q*w+b for every score (threshold==0) or only where q > threshold (calibration_lib.py:77-99).

  dtype behaviour is the reference's: with threshold==0 a float32 input stays
  float32 (python-scalar multiply); otherwise np.where yields float64 factors.
  """
  cv = calibration_values
  if cv.threshold == 0:
    return quality_scores * cv.w + cv.b
  above = quality_scores > cv.threshold
  return quality_scores * np.where(above, cv.w, 1.0) + np.where(above, cv.b, 0.0)
