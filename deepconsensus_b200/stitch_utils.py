"""Output surface: windows -> read -> FASTQ (mirror of `postprocess/stitch_utils.py`).

Same dataclasses, function names, argument meaning and filter order as the
reference (stitch_utils.py:39-189) so `deepconsensus run`'s post-model code can
bind to it unchanged.  Gap stripping works on bytes instead of the reference's
per-character string appends (stitch_utils.py:84-98) -- same result.
"""
from __future__ import annotations

import dataclasses
import logging
from typing import Iterable, Optional, Tuple

import numpy as np

from deepconsensus_b200 import constants
from deepconsensus_b200 import utils


@dataclasses.dataclass
class DCModelOutput:
  """Per-window model output + read metadata (stitch_utils.py:39-48)."""
  molecule_name: str
  window_pos: int
  ec: float
  np_num_passes: int
  rq: float
  rg: str
  sequence: Optional[str] = None
  quality_string: Optional[str] = None


@dataclasses.dataclass
class OutcomeCounter:
  """Per-read outcome tallies (stitch_utils.py:122-128)."""
  empty_sequence: int = 0
  only_gaps: int = 0
  failed_quality_filter: int = 0
  failed_length_filter: int = 0
  success: int = 0


def get_full_sequence(deepconsensus_outputs: Iterable[DCModelOutput], max_length: int,
                      fill_n: bool = False) -> Tuple[Optional[str], str]:
  """Concatenate sorted windows; a missing window aborts (or is N-filled).

  Literal restatement of stitch_utils.py:51-81, including `start` advancing by
  `max_length` per emitted window.
  """
  seq_parts, qual_parts = [], []
  filler_q = utils.quality_scores_to_string(np.full(max_length, constants.EMPTY_QUAL))
  start = 0
  for out in deepconsensus_outputs:
    while out.window_pos > start:
      if not fill_n:
        return None, ""
      seq_parts.append("N" * max_length)
      qual_parts.append(filler_q)
      start += max_length
    seq_parts.append(out.sequence)
    qual_parts.append(out.quality_string)
    start += max_length
  return "".join(seq_parts), "".join(qual_parts)


def remove_gaps(sequence: str, quality_string: str) -> Tuple[str, str]:
  """Drop gap characters and the quality characters under them (stitch_utils.py:84-98).

  Like the reference's zip(), extra trailing characters of the longer input are ignored.
  """
  n = min(len(sequence), len(quality_string))
  seq = np.frombuffer(sequence[:n].encode("latin-1"), dtype=np.uint8)
  qual = np.frombuffer(quality_string[:n].encode("latin-1"), dtype=np.uint8)
  keep = seq != ord(constants.GAP)
  return seq[keep].tobytes().decode("latin-1"), qual[keep].tobytes().decode("latin-1")


def is_quality_above_threshold(quality_string: str, min_quality: int) -> bool:
  """round(avg_phred, 5) >= min_quality (stitch_utils.py:101-109)."""
  phred = round(utils.avg_phred(utils.quality_string_to_array(quality_string)), 5)
  return phred >= min_quality


def format_as_fastq(molecule_name: str, sequence: str, quality_string: str) -> str:
  """Four-line FASTQ record (stitch_utils.py:112-119)."""
  return "@%s\n%s\n+\n%s\n" % (molecule_name, sequence, quality_string)


def stitch_to_fastq(molecule_name: str, predictions: Iterable[DCModelOutput], max_length: int,
                    min_quality: int, min_length: int,
                    outcome_counter: OutcomeCounter) -> Optional[str]:
  """Stitch, strip gaps, filter (empty / only gaps / quality / length), format.

  Filter order and counters as stitch_utils.py:131-189.
  """
  full_seq, full_qual = get_full_sequence(predictions, max_length=max_length)
  if not full_seq:
    outcome_counter.empty_sequence += 1
    logging.debug("Filtered out read that was empty after stitching: %s", molecule_name)
    return None
  seq, qual = remove_gaps(full_seq, full_qual)
  if not seq:
    outcome_counter.only_gaps += 1
    logging.debug("Filtered out read that contained only gaps: %s", molecule_name)
    return None
  if not is_quality_above_threshold(qual, min_quality):
    outcome_counter.failed_quality_filter += 1
    logging.debug("Filtered out read below quality threshold: %s", molecule_name)
    return None
  if len(seq) < min_length:
    outcome_counter.failed_length_filter += 1
    logging.debug("Filtered out read below length threshold: %s", molecule_name)
    return None
  outcome_counter.success += 1
  return format_as_fastq(molecule_name, seq, qual)
