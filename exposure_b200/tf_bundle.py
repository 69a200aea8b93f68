"""TF-free reader of TensorFlow checkpoint bundles (V2: `<prefix>.index` + `<prefix>.data-00000-of-00001`),
enough to load the reference's shipped `models/example/pretrained/model.ckpt-20000`
(net.py:271 `tf.train.Saver`, evaluate.py:27-28 `net.restore(20000)`).

Format facts: the .index file is an uncompressed LevelDB-style table -- 48-byte
footer (metaindex handle, index handle, magic 0xdb4775248b80fb57), blocks of prefix-compressed
entries `varint shared | varint non_shared | varint value_len | key suffix | value` followed by a
restart array; values of data blocks are BundleEntryProto messages (dtype, shape, shard, offset,
size); tensor bytes are raw little-endian row-major in the .data shard."""
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
  out = shift = 0
  while True:
    b = buf[pos]
    pos += 1
    out |= (b & 0x7F) << shift
    if not b & 0x80:
      return out, pos
    shift += 7


def _block_entries(buf, offset, size):
  """Yield (key, value) of the block at [offset, offset+size) (compression type must be 0)."""
  if buf[offset + size] != 0:
    raise ValueError("compressed table blocks are not supported")
  block = buf[offset:offset + size]
  n_restarts = struct.unpack_from("<I", block, size - 4)[0]
  end = size - 4 - 4 * n_restarts
  pos, key = 0, b""
  while pos < end:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    vlen, pos = _varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + non_shared])
    pos += non_shared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def _parse_proto(msg):
  """Minimal protobuf wire decoder -> {field: [values]} (varint, fixed32/64, length-delimited)."""
  out, pos = {}, 0
  while pos < len(msg):
    tag, pos = _varint(msg, pos)
    field, wt = tag >> 3, tag & 7
    if wt == 0:
      v, pos = _varint(msg, pos)
    elif wt == 1:
      v = msg[pos:pos + 8]; pos += 8
    elif wt == 2:
      ln, pos = _varint(msg, pos)
      v = msg[pos:pos + ln]; pos += ln
    elif wt == 5:
      v = msg[pos:pos + 4]; pos += 4
    else:
      raise ValueError("unsupported wire type %d" % wt)
    out.setdefault(field, []).append(v)
  return out


def read_index(prefix):
  """{variable name: dict(dtype, shape, offset, size)} from `<prefix>.index`."""
  buf = open(prefix + ".index", "rb").read()
  footer = buf[-48:]
  if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
    raise ValueError("not a TF bundle index (bad magic)")
  pos = 0
  _, pos = _varint(footer, pos)          # metaindex offset
  _, pos = _varint(footer, pos)          # metaindex size
  idx_off, pos = _varint(footer, pos)
  idx_size, pos = _varint(footer, pos)
  entries = {}
  for _, handle in _block_entries(buf, idx_off, idx_size):
    off, p = _varint(handle, 0)
    size, p = _varint(handle, p)
    for key, value in _block_entries(buf, off, size):
      if not key:
        continue                         # BundleHeaderProto
      f = _parse_proto(value)
      shape = []
      if 2 in f:
        for dim in _parse_proto(f[2][0]).get(2, []):
          shape.append(_parse_proto(dim).get(1, [0])[0])
      entries[key.decode()] = dict(dtype=f.get(1, [1])[0], shape=tuple(shape), shard=f.get(3, [0])[0],
                                   offset=f.get(4, [0])[0], size=f.get(5, [0])[0])
  return entries


def load_bundle(prefix, include_optimizer_slots=False):
  """{variable name: numpy array} for every tensor of the single-shard bundle `prefix`."""
  entries = read_index(prefix)
  data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
  out = {}
  for name, e in entries.items():
    if not include_optimizer_slots and ("/Adam" in name or "OptimizeLoss" in name or "beta1_power" in name or "beta2_power" in name):
      continue
    dt = _DTYPES.get(e["dtype"])
    if dt is None:
      continue
    arr = np.frombuffer(data[e["offset"]:e["offset"] + e["size"]].tobytes(), dtype=dt)
    out[name] = arr.reshape(e["shape"]) if e["shape"] else arr.reshape(())
  return out


def load_pretrained_into(trainer, prefix):
  """Load a reference checkpoint into a Trainer (names/layouts map 1:1, see checkpoint.py)."""
  import torch
  from .checkpoint import import_named
  named = {k: torch.from_numpy(np.array(v)) for k, v in load_bundle(prefix).items() if v.dtype == np.float32 and v.ndim > 0}
  import_named(trainer, named)
  return named
