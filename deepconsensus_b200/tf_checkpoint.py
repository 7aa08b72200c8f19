"""TensorFlow-free reader (and writer) of TF2 object-graph checkpoints ("tensor bundles").

The reference restores its model with `tf.train.Checkpoint(model=model).restore(checkpoint_path)`
(quick_inference.py:515-529).  This module reads the same files without TensorFlow and hands the variables to
`B200Model` / `dcb_load_weights` under the names `deepconsensus_b200/weights.py` lists:

  <prefix>.index                 LevelDB-format table ("SSTable"): sorted string keys -> serialized protos, in
                                 prefix-compressed blocks, each block optionally snappy-compressed, followed by a
                                 5-byte trailer (type, masked crc32c); a 48-byte footer points at the index block.
                                 Key ""  -> BundleHeaderProto {num_shards=1, endianness=2, version=3}
                                 Key k   -> BundleEntryProto  {dtype=1, shape=2, shard_id=3, offset=4, size=5,
                                                               crc32c=6 (fixed32, masked), slices=7}
  <prefix>.data-SSSSS-of-NNNNN   the raw little-endian tensors, entry i at [offset, offset + size) of its shard.

Variables of `tf.train.Checkpoint(model=...)` are keyed `model/<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`;
optimizer slots (`.../.OPTIMIZER_SLOT/...`, `optimizer/...`) and bookkeeping (`save_counter`,
`_CHECKPOINTABLE_OBJECT_GRAPH`) are skipped, as the reference's `expect_partial()` does.

Formats implemented from their public specifications: LevelDB table_format.md, the snappy format description,
protobuf wire encoding, tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/hash/crc32c.h (mask).
Pinned against the reference's own fixtures `deepconsensus/testdata/model{,_bq}/checkpoint-1.index` (committed copies
under tests/golden/ckpt/): every variable `weights.variable_shapes` expects is there with the same dtype and shape.
"""
from __future__ import annotations

import dataclasses
import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

VARIABLE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT16, DT_INT8, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 4, 5, 6, 7, 9, 10
DT_BFLOAT16, DT_HALF = 14, 19
_NP_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_UINT8: np.dtype("u1"),
              DT_INT16: np.dtype("<i2"), DT_INT8: np.dtype("i1"), DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("?"),
              DT_HALF: np.dtype("<f2")}


class CheckpointError(ValueError):
  pass


# ----------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_crc_table() -> List[int]:
  table = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    table.append(c)
  return table


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
  c = crc ^ 0xFFFFFFFF
  tab = _CRC_TABLE
  for b in data:
    c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
  """crc32c::Mask: rotate right by 15 and add a constant (stored CRCs are masked)."""
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------- varints / protobuf
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
  result = shift = 0
  while True:
    if pos >= len(buf):
      raise CheckpointError("truncated varint")
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7
    if shift > 70:
      raise CheckpointError("varint too long")


def _put_varint(v: int) -> bytes:
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
  """(field number, wire type, value) of one serialized message; value = int or bytes."""
  pos = 0
  while pos < len(buf):
    key, pos = _varint(buf, pos)
    field, wt = key >> 3, key & 7
    if wt == 0:
      v, pos = _varint(buf, pos)
    elif wt == 1:
      v = struct.unpack_from("<Q", buf, pos)[0]
      pos += 8
    elif wt == 2:
      n, pos = _varint(buf, pos)
      v = bytes(buf[pos:pos + n])
      if len(v) != n:
        raise CheckpointError("truncated length-delimited field")
      pos += n
    elif wt == 5:
      v = struct.unpack_from("<I", buf, pos)[0]
      pos += 4
    else:
      raise CheckpointError("unsupported protobuf wire type %d" % wt)
    yield field, wt, v


def _signed64(v: int) -> int:
  return v - (1 << 64) if v >= 1 << 63 else v


# ----------------------------------------------------------------------------------------------- snappy
def snappy_decompress(data: bytes) -> bytes:
  n, pos = _varint(data, 0)
  out = bytearray()
  while pos < len(data):
    tag = data[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:                                    # literal
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(data[pos:pos + nb], "little")
        pos += nb
      ln += 1
      out += data[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:                                    # copy, 1-byte offset
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | data[pos]
      pos += 1
    elif kind == 2:                                  # copy, 2-byte offset
      ln = (tag >> 2) + 1
      off = data[pos] | (data[pos + 1] << 8)
      pos += 2
    else:                                            # copy, 4-byte offset
      ln = (tag >> 2) + 1
      off = int.from_bytes(data[pos:pos + 4], "little")
      pos += 4
    if off == 0 or off > len(out):
      raise CheckpointError("corrupt snappy stream (bad copy offset)")
    for _ in range(ln):                              # byte-wise: copies may overlap their own output
      out.append(out[-off])
  if len(out) != n:
    raise CheckpointError("corrupt snappy stream: %d bytes, header says %d" % (len(out), n))
  return bytes(out)


def snappy_compress_literal(data: bytes) -> bytes:
  """A valid snappy stream made of literals only (used by the test writer to exercise the compressed-block path)."""
  out = bytearray(_put_varint(len(data)))
  pos = 0
  while pos < len(data):
    chunk = data[pos:pos + 65536]
    ln = len(chunk) - 1
    if ln < 60:
      out.append(ln << 2)
    else:
      nb = (ln.bit_length() + 7) // 8
      out.append((59 + nb) << 2)
      out += ln.to_bytes(nb, "little")
    out += chunk
    pos += len(chunk)
  return bytes(out)


# ----------------------------------------------------------------------------------------------- LevelDB table
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
  raw = buf[offset:offset + size]
  trailer = buf[offset + size:offset + size + 5]
  if len(raw) != size or len(trailer) != 5:
    raise CheckpointError("block [%d, +%d) outside the index file" % (offset, size))
  if verify:
    want = struct.unpack("<I", trailer[1:])[0]
    if mask_crc(crc32c(raw + trailer[:1])) != want:
      raise CheckpointError("block checksum mismatch at offset %d" % offset)
  if trailer[0] == 0:
    return raw
  if trailer[0] == 1:
    return snappy_decompress(raw)
  raise CheckpointError("unknown block compression type %d" % trailer[0])


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
  if len(block) < 4:
    raise CheckpointError("block too small")
  num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * num_restarts
  if end < 0:
    raise CheckpointError("bad restart array")
  pos, key = 0, b""
  while pos < end:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    vlen, pos = _varint(block, pos)
    if shared > len(key):
      raise CheckpointError("bad key prefix length")
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    yield key, block[pos:pos + vlen]
    pos += vlen


def read_table(path: str, verify_checksums: bool = True) -> Dict[bytes, bytes]:
  """All (key, value) pairs of a LevelDB-format table file."""
  with open(path, "rb") as f:
    buf = f.read()
  if len(buf) < 48:
    raise CheckpointError("%s: too small for a table footer" % path)
  footer = buf[-48:]
  if struct.unpack("<Q", footer[40:])[0] != _TABLE_MAGIC:
    raise CheckpointError("%s: not a LevelDB-format table (bad magic)" % path)
  pos = 0
  _, pos = _varint(footer, pos)          # metaindex handle (unused)
  _, pos = _varint(footer, pos)
  idx_off, pos = _varint(footer, pos)
  idx_size, pos = _varint(footer, pos)
  out: Dict[bytes, bytes] = {}
  for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify_checksums)):
    off, p = _varint(handle, 0)
    size, p = _varint(handle, p)
    for k, v in _block_entries(_read_block(buf, off, size, verify_checksums)):
      out[bytes(k)] = bytes(v)
  return out


# ----------------------------------------------------------------------------------------------- tensor bundle
@dataclasses.dataclass
class BundleEntry:
  dtype: int
  shape: Tuple[int, ...]
  shard_id: int
  offset: int
  size: int
  crc32c: int
  sliced: bool = False


@dataclasses.dataclass
class BundleHeader:
  num_shards: int
  little_endian: bool
  version: Optional[Tuple[int, int]] = None


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
  dims = []
  for f, _, v in _proto_fields(buf):
    if f == 2:                                       # TensorShapeProto.dim
      size = 0
      for f2, _, v2 in _proto_fields(v):
        if f2 == 1:
          size = _signed64(v2)
      dims.append(size)
    elif f == 3 and v:
      raise CheckpointError("tensor of unknown rank in a checkpoint")
  return tuple(dims)


def _parse_entry(buf: bytes) -> BundleEntry:
  e = BundleEntry(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=0)
  for f, _, v in _proto_fields(buf):
    if f == 1:
      e.dtype = v
    elif f == 2:
      e.shape = _parse_shape(v)
    elif f == 3:
      e.shard_id = v
    elif f == 4:
      e.offset = v
    elif f == 5:
      e.size = v
    elif f == 6:
      e.crc32c = v
    elif f == 7:
      e.sliced = True
  return e


def _parse_header(buf: bytes) -> BundleHeader:
  h = BundleHeader(num_shards=1, little_endian=True)
  for f, _, v in _proto_fields(buf):
    if f == 1:
      h.num_shards = v
    elif f == 2:
      h.little_endian = v == 0
    elif f == 3:
      producer = min_consumer = 0
      for f2, _, v2 in _proto_fields(v):
        if f2 == 1:
          producer = v2
        elif f2 == 2:
          min_consumer = v2
      h.version = (producer, min_consumer)
  return h


def read_index(prefix: str, verify_checksums: bool = True) -> Tuple[BundleHeader, Dict[str, BundleEntry]]:
  """Header and entries of `<prefix>.index` (keys as stored, including the /.ATTRIBUTES/... suffix)."""
  table = read_table(prefix + ".index", verify_checksums)
  if b"" not in table:
    raise CheckpointError("%s.index has no bundle header" % prefix)
  header = _parse_header(table[b""])
  if not header.little_endian:
    raise CheckpointError("big-endian tensor bundles are not supported")
  entries = {k.decode("utf-8"): _parse_entry(v) for k, v in table.items() if k != b""}
  return header, entries


def variable_entries(entries: Dict[str, BundleEntry]) -> Dict[str, BundleEntry]:
  """The model variables: `model/...` keys with the VARIABLE_VALUE suffix stripped; optimizer slots dropped."""
  out = {}
  for key, e in entries.items():
    if not key.endswith(VARIABLE_SUFFIX) or ".OPTIMIZER_SLOT" in key or not key.startswith("model/"):
      continue
    out[key[:-len(VARIABLE_SUFFIX)]] = e
  return out


def shard_path(prefix: str, shard_id: int, num_shards: int) -> str:
  return "%s.data-%05d-of-%05d" % (prefix, shard_id, num_shards)


def load_variables(prefix: str, verify_tensor_crc: bool = False) -> Dict[str, np.ndarray]:
  """All model variables of the checkpoint at `prefix` as float32 arrays keyed like `weights.variable_shapes`."""
  header, entries = read_index(prefix)
  out: Dict[str, np.ndarray] = {}
  shards: Dict[int, np.memmap] = {}
  for name, e in sorted(variable_entries(entries).items()):
    if e.sliced:
      raise CheckpointError("%s: partitioned (sliced) variables are not supported" % name)
    if e.dtype not in _NP_DTYPES:
      raise CheckpointError("%s: unsupported dtype %d" % (name, e.dtype))
    dt = _NP_DTYPES[e.dtype]
    count = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
    if count * dt.itemsize != e.size:
      raise CheckpointError("%s: %d bytes stored, shape %s needs %d" % (name, e.size, e.shape, count * dt.itemsize))
    if e.shard_id not in shards:
      path = shard_path(prefix, e.shard_id, header.num_shards)
      if not os.path.exists(path):
        raise CheckpointError("%s is missing (the checkpoint's index is present but not its data shard)" % path)
      shards[e.shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
    raw = shards[e.shard_id][e.offset:e.offset + e.size]
    if raw.shape[0] != e.size:
      raise CheckpointError("%s: data shard too short" % name)
    if verify_tensor_crc and mask_crc(crc32c(raw.tobytes())) != e.crc32c:
      raise CheckpointError("%s: tensor checksum mismatch" % name)
    out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e.shape).astype(np.float32)
  return out


def latest_checkpoint(model_dir: str) -> Optional[str]:
  """tf.train.latest_checkpoint: the prefix named by `model_checkpoint_path` in <dir>/checkpoint."""
  state = os.path.join(model_dir, "checkpoint")
  if not os.path.exists(state):
    return None
  with open(state) as f:
    for line in f:
      line = line.strip()
      if line.startswith("model_checkpoint_path:"):
        name = line.split(":", 1)[1].strip().strip('"')
        return name if os.path.isabs(name) else os.path.join(model_dir, name)
  return None


def resolve_prefix(checkpoint_path: str) -> str:
  """`--checkpoint` is a prefix (".../checkpoint-50"); a directory means its latest checkpoint."""
  if os.path.isdir(checkpoint_path):
    latest = latest_checkpoint(checkpoint_path)
    if latest is None:
      raise CheckpointError("%s: no `checkpoint` state file" % checkpoint_path)
    return latest
  if checkpoint_path.endswith(".index"):
    return checkpoint_path[:-len(".index")]
  return checkpoint_path


# ----------------------------------------------------------------------------------------------- writer (tests, export)
def _block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
  out, restarts, prev = bytearray(), [], b""
  for i, (k, v) in enumerate(entries):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(out))
    else:
      while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
        shared += 1
    out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
    prev = k
  if not restarts:
    restarts = [0]
  for r in restarts:
    out += struct.pack("<I", r)
  out += struct.pack("<I", len(restarts))
  return bytes(out)


def _field(num: int, wt: int, payload) -> bytes:
  key = _put_varint((num << 3) | wt)
  if wt == 0:
    return key + _put_varint(payload)
  if wt == 2:
    return key + _put_varint(len(payload)) + payload
  if wt == 5:
    return key + struct.pack("<I", payload)
  raise ValueError(wt)


def write_checkpoint(prefix: str, variables: Dict[str, np.ndarray], compress: bool = True,
                     extra_keys: Optional[Dict[str, np.ndarray]] = None, block_entries: int = 24) -> None:
  """Writes `variables` (names as in weights.variable_shapes) as a single-shard TF2 tensor bundle at `prefix`.
  For tests and for exporting engine-ready checkpoints without TensorFlow."""
  items = {name + VARIABLE_SUFFIX: np.asarray(a) for name, a in variables.items()}
  for k, a in (extra_keys or {}).items():
    items[k] = np.asarray(a)
  data = bytearray()
  table: List[Tuple[bytes, bytes]] = []
  header = _field(1, 0, 1) + _field(2, 0, 0) + _field(3, 2, _field(1, 0, 1))
  table.append((b"", header))
  rev = {v: k for k, v in _NP_DTYPES.items()}
  for key in sorted(items, key=lambda s: s.encode("utf-8")):
    a = items[key]
    dt = np.dtype(a.dtype).newbyteorder("<") if a.dtype.byteorder == ">" else np.dtype(a.dtype)
    if dt not in rev:
      raise CheckpointError("cannot store dtype %s" % a.dtype)
    raw = np.ascontiguousarray(a, dtype=dt).tobytes()
    shape = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in a.shape)
    entry = _field(1, 0, rev[dt])
    if shape:
      entry += _field(2, 2, shape)
    entry += _field(4, 0, len(data)) + _field(5, 0, len(raw)) + _field(6, 5, mask_crc(crc32c(raw)))
    table.append((key.encode("utf-8"), entry))
    data += raw
  out = bytearray()
  index_entries: List[Tuple[bytes, bytes]] = []

  def emit(block: bytes) -> Tuple[int, int]:
    body, ctype = (snappy_compress_literal(block), 1) if compress else (block, 0)
    off = len(out)
    out.extend(body)
    out.extend(bytes([ctype]) + struct.pack("<I", mask_crc(crc32c(body + bytes([ctype])))))
    return off, len(body)

  for i in range(0, len(table), block_entries):
    chunk = table[i:i + block_entries]
    off, size = emit(_block(chunk))
    index_entries.append((chunk[-1][0], _put_varint(off) + _put_varint(size)))
  moff, msize = emit(_block([]))
  ioff, isize = emit(_block(index_entries, restart_interval=1))
  footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
  footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
  out += footer
  with open(prefix + ".index", "wb") as f:
    f.write(bytes(out))
  with open(shard_path(prefix, 0, 1), "wb") as f:
    f.write(bytes(data))
  state = os.path.join(os.path.dirname(prefix) or ".", "checkpoint")
  base = os.path.basename(prefix)
  with open(state, "w") as f:
    f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
