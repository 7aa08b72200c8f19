"""Batch stitching with the concatenation + gap compaction on the device (`dcb_stitch`).

`stitch_utils.stitch_to_fastq` (mirror of postprocess/stitch_utils.py:131-189) handles one read at a time from
per-window strings.  Here a whole batch of reads goes through at once, straight from the engine's per-window byte
arrays: the device does get_full_sequence + remove_gaps (stitch_utils.py:51-98); the host keeps what needs the window
positions and read names -- the missing-window check, the empty / only-gaps / quality / length filters (same order,
same counters, `round(avg_phred, 5)` evaluated by the same NumPy code as the reference) and the FASTQ formatting.
The result is identical, read for read, to calling stitch_utils.stitch_to_fastq on each read.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from deepconsensus_b200 import stitch_utils, utils


def group_reads(molecule_names: Sequence[str]) -> np.ndarray:
  """zmw_start for windows already grouped by read: index of the first window of every run of equal names, + end."""
  n = len(molecule_names)
  starts = [0] if n else []
  for i in range(1, n):
    if molecule_names[i] != molecule_names[i - 1]:
      starts.append(i)
  return np.asarray(starts + [n], dtype=np.int32)


def stitch_batch_to_fastq(model, bases, quals, molecule_names: Sequence[str], window_pos: Sequence[int],
                          max_length: int, min_quality: int, min_length: int,
                          outcome_counter: stitch_utils.OutcomeCounter,
                          n_windows: Optional[int] = None, on_device: bool = False) -> List[Optional[str]]:
  """One FASTQ record (or None) per read, for windows grouped by read and sorted by window position.

  `bases` / `quals`: uint8 [n_windows, max_length] arrays as `B200Model.forward` returns them, or device addresses
  of the same (`on_device=True`, e.g. the DCB_OUT_ON_DEVICE outputs of `forward_raw`).
  """
  zs = group_reads(molecule_names)
  nz = len(zs) - 1
  seq, qual, lens = model.stitch(bases, quals, zs, n_windows=n_windows, on_device=on_device, length=max_length)
  out: List[Optional[str]] = []
  for z in range(nz):
    w0, w1 = int(zs[z]), int(zs[z + 1])
    name = molecule_names[w0]
    # get_full_sequence (stitch_utils.py:51-81): a window further right than expected means one is missing
    missing = any(int(window_pos[w0 + i]) > i * max_length for i in range(w1 - w0))
    if missing or w1 == w0 or max_length == 0:
      outcome_counter.empty_sequence += 1
      out.append(None)
      continue
    n = int(lens[z])
    if n == 0:
      outcome_counter.only_gaps += 1
      out.append(None)
      continue
    o = w0 * max_length
    q = qual[o:o + n]
    phred = round(utils.avg_phred(q.astype(np.int64) - 33), 5)       # is_quality_above_threshold (:101-109)
    if not phred >= min_quality:
      outcome_counter.failed_quality_filter += 1
      out.append(None)
      continue
    if n < min_length:
      outcome_counter.failed_length_filter += 1
      out.append(None)
      continue
    outcome_counter.success += 1
    out.append(stitch_utils.format_as_fastq(name, seq[o:o + n].tobytes().decode("latin-1"),
                                            q.tobytes().decode("latin-1")))
  return out
