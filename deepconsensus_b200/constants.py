"""Constants of the hot path.

Values restate `deepconsensus/utils/dc_constants.py:36-45,114-127` of the
reference (which cannot be imported here: it pulls in pysam + tensorflow at
module import, dc_constants.py:32-33).
"""
import numpy as np

REFERENCE_VERSION = "1.2.0"            # dc_constants.py:36

GAP = " "                              # dc_constants.py:39
ALLOWED_BASES = "ATCG"                 # dc_constants.py:40
SEQ_VOCAB = GAP + ALLOWED_BASES        # dc_constants.py:41  -> ' ATCG'
SEQ_VOCAB_SIZE = len(SEQ_VOCAB)        # dc_constants.py:42  -> 5
GAP_INT = SEQ_VOCAB.index(GAP)         # dc_constants.py:45  -> 0

NP_DATA_TYPE = np.float32              # dc_constants.py:85

# Feature keys carried through batching (dc_constants.py:114-125).
DC_FEATURES = (
    "rows", "label", "num_passes", "window_pos", "name",
    "ccs_base_quality_scores", "ec", "np_num_passes", "rq", "rg",
)

EMPTY_QUAL = 0                         # dc_constants.py:127

# ASCII lookup used by the device epilogue and the host fast path.
SEQ_VOCAB_ASCII = np.frombuffer(SEQ_VOCAB.encode("ascii"), dtype=np.uint8)
PHRED_OFFSET = 33                      # utils.py:51,62
