"""TF-free tensor-bundle reader (deepconsensus_b200/tf_checkpoint.py).

Pinned against the reference's own fixtures: tests/golden/ckpt/{model,model_bq}/ are byte copies of
deepconsensus/testdata/model{,_bq}/{checkpoint, checkpoint-*.index, params.json} (data files written by real
TensorFlow; the reference ships them without their data shards).  They pin (a) the SSTable / snappy / protobuf parsing
-- block checksums verify -- and (b) `weights.variable_shapes`: every variable the engine expects for that
params.json is in the checkpoint under the same name with the same dtype and shape, and nothing else is.
"""
import json
import os

import numpy as np
import pytest

from deepconsensus_b200 import params as params_lib, tf_checkpoint as ckpt, weights as weights_lib


def _params_from_json(path):
  p = params_lib.read_params_from_json(path)
  params_lib.modify_params(p)
  return p


@pytest.mark.parametrize("name,prefix", [("model", "checkpoint-1"), ("model", "checkpoint-2"), ("model_bq", "checkpoint-1")])
def test_reference_fixture_lists_exactly_the_variables_the_engine_expects(golden_dir, name, prefix):
  d = os.path.join(golden_dir, "ckpt", name)
  header, entries = ckpt.read_index(os.path.join(d, prefix), verify_checksums=True)
  assert header.num_shards == 1 and header.little_endian
  got = ckpt.variable_entries(entries)
  p = _params_from_json(os.path.join(d, "params.json"))
  want = dict(weights_lib.variable_shapes(p))
  assert sorted(got) == sorted(want)
  for k, shape in want.items():
    assert got[k].dtype == ckpt.DT_FLOAT and tuple(got[k].shape) == tuple(shape), k
    assert got[k].size == 4 * int(np.prod(shape, dtype=np.int64)) and not got[k].sliced
  assert weights_lib.count_params(p) == sum(e.size // 4 for e in got.values())
  # what the reference's expect_partial() ignores is there too: optimizer slots, save_counter, the object graph
  assert "_CHECKPOINTABLE_OBJECT_GRAPH" in entries and "save_counter/.ATTRIBUTES/VARIABLE_VALUE" in entries
  assert any(".OPTIMIZER_SLOT" in k for k in entries)
  # entries of the data shard do not overlap
  spans = sorted((e.offset, e.offset + e.size) for e in entries.values())
  assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))


def test_latest_checkpoint_and_missing_shard(golden_dir):
  d = os.path.join(golden_dir, "ckpt", "model")
  assert ckpt.latest_checkpoint(d) == os.path.join(d, "checkpoint-1")
  assert ckpt.resolve_prefix(d) == os.path.join(d, "checkpoint-1")
  assert ckpt.resolve_prefix(os.path.join(d, "checkpoint-2.index")) == os.path.join(d, "checkpoint-2")
  with pytest.raises(ckpt.CheckpointError, match="data shard"):
    ckpt.load_variables(os.path.join(d, "checkpoint-1"))       # the reference ships the index only


@pytest.mark.parametrize("compress", [True, False])
def test_write_read_round_trip(tmp_path, compress):
  p = params_lib.synthetic_params(20, 100, use_ccs_bq=True, rezero=False, num_hidden_layers=2)
  w = weights_lib.init_weights(p, seed=3)
  prefix = str(tmp_path / "checkpoint-7")
  extra = {"save_counter/.ATTRIBUTES/VARIABLE_VALUE": np.array(7, np.int64),
           "model/fc1/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE": np.zeros((280, 5), np.float32)}
  ckpt.write_checkpoint(prefix, w, compress=compress, extra_keys=extra, block_entries=5)
  back = ckpt.load_variables(prefix, verify_tensor_crc=True)
  assert sorted(back) == sorted(w)
  for k in w:
    assert back[k].dtype == np.float32 and back[k].shape == np.shape(w[k])
    np.testing.assert_array_equal(back[k], np.asarray(w[k], np.float32))
  weights_lib.check_weights(p, back)
  assert ckpt.resolve_prefix(str(tmp_path)) == prefix
  # corruption is detected
  raw = bytearray(open(prefix + ".index", "rb").read())
  raw[10] ^= 0xFF
  open(prefix + ".index", "wb").write(bytes(raw))
  with pytest.raises(ckpt.CheckpointError):
    ckpt.read_index(prefix)


def test_snappy_copies_and_crc_known_answers():
  assert ckpt.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
  assert ckpt.crc32c(bytes(32)) == 0x8A9136AA                          # rfc3720 B.4: 32 bytes of zeros
  # literal "abcd" then a 1-byte-offset copy of length 8 from offset 4 (overlapping its own output)
  stream = bytes([12, (3 << 2) | 0]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
  assert ckpt.snappy_decompress(stream) == b"abcdabcdabcd"
  # 2-byte-offset copy
  stream = bytes([8, (3 << 2) | 0]) + b"wxyz" + bytes([((4 - 1) << 2) | 2, 4, 0])
  assert ckpt.snappy_decompress(stream) == b"wxyzwxyz"
  blob = bytes(range(256)) * 300
  assert ckpt.snappy_decompress(ckpt.snappy_compress_literal(blob)) == blob
  with pytest.raises(ckpt.CheckpointError):
    ckpt.snappy_decompress(bytes([4, (0 << 2) | 1, 9]))                # copy before any output


def test_initialize_model_reads_params_json_and_resolves_the_checkpoint(tmp_path, golden_dir):
  """inference.load_weights accepts what `--checkpoint` accepts: a prefix, a directory, an .npz export."""
  from deepconsensus_b200 import inference
  p = params_lib.synthetic_params(20, 100, num_hidden_layers=1)
  w = weights_lib.init_weights(p, seed=4)
  prefix = str(tmp_path / "checkpoint-3")
  ckpt.write_checkpoint(prefix, w)
  for path in (prefix, str(tmp_path), prefix + ".index"):
    back = inference.load_weights(path)
    np.testing.assert_array_equal(back["model/fc1/kernel"], w["model/fc1/kernel"])
  np.savez(str(tmp_path / "w.npz"), **w)
  back = inference.load_weights(str(tmp_path / "w.npz"))
  np.testing.assert_array_equal(back["model/fc1/bias"], w["model/fc1/bias"])
  p2 = inference.read_params_from_json(os.path.join(golden_dir, "ckpt", "model", "checkpoint-1"))
  assert p2.max_passes == 20 and p2.max_length == 100 and p2.num_hidden_layers == 6 and p2.rezero
  assert p2.dc_calibration == "0,1.197654,-0.99781"
