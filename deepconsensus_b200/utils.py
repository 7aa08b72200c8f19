"""Small host helpers on the output surface (mirror of `deepconsensus/utils/utils.py`).

Same names and results as the reference's pure-NumPy helpers
(utils.py:36-67, 88-106); byte-oriented instead of per-character Python loops.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np

from deepconsensus_b200 import constants


def encoded_sequence_to_string(encoded_sequence: np.ndarray) -> str:
  """Vocabulary ids (0..4) -> ' ATCG' string (utils.py:36-40)."""
  ids = np.asarray(encoded_sequence).astype(np.int64)
  return constants.SEQ_VOCAB_ASCII[ids].tobytes().decode("ascii")


def quality_score_to_string(score: int) -> str:
  """One Phred score -> its FASTQ character (utils.py:43-57)."""
  return chr(int(score) + constants.PHRED_OFFSET)


def quality_scores_to_string(scores: np.ndarray) -> str:
  """Array of Phred scores -> FASTQ quality string (utils.py:60-62)."""
  arr = np.asarray(scores).astype(np.int64) + constants.PHRED_OFFSET
  return "".join(map(chr, arr.tolist()))


def quality_string_to_array(quality_string: str) -> List[int]:
  """FASTQ quality string -> list of Phred ints (utils.py:65-67)."""
  return [c - constants.PHRED_OFFSET for c in quality_string.encode("latin-1")]


def avg_phred(base_qualities: Union[np.ndarray, Sequence[int]]) -> float:
  """Phred of the mean error probability, ignoring negative (spacing) entries.

  utils.py:88-106: entries < 0 are dropped; all-zero (or empty) input -> 0.0;
  otherwise -10*log10(mean(10**(-q/10))) in float64.
  """
  q = np.asarray(base_qualities)
  q = q[q >= 0]
  if not q.any():
    return 0.0
  err = np.power(10.0, q / -10.0)
  return -10 * np.log10(err.sum() / len(err))
