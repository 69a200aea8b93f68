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


# ------------------------------------------------------------------------------------------
# Writer: produces `<prefix>.index` + `<prefix>.data-00000-of-00001` in the same V2 bundle format
# (what tf.train.Saver.save writes at net.py:383-387), so new checkpoints stay readable by the
# reference's Saver.restore and by read_index / load_bundle above.
# ------------------------------------------------------------------------------------------
_CRC_TABLE = None


def crc32c(data, crc=0):
  """CRC-32C (Castagnoli), the checksum of LevelDB tables and BundleEntryProto.crc32c.  Table-driven,
  vectorised over 8 interleaved lanes with numpy for large buffers (a pure byte loop for small ones)."""
  global _CRC_TABLE
  if _CRC_TABLE is None:
    tab = np.zeros(256, dtype=np.uint32)
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
      tab[i] = c
    _CRC_TABLE = tab
  tab = _CRC_TABLE
  buf = np.frombuffer(bytes(data), dtype=np.uint8)
  c = np.uint32(crc ^ 0xFFFFFFFF)
  t = tab.tolist()
  c = int(c)
  for b in buf.tolist():
    c = t[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def _crc32c_fast(data):
  """crc32c through the C library when it is built (exp_crc32c, ~1 GB/s), else the Python fallback."""
  try:
    from . import _cabi
    import ctypes
    lib = _cabi.lib()
    b = bytes(data)
    return int(lib.exp_crc32c(0, ctypes.c_char_p(b), len(b))) & 0xFFFFFFFF
  except Exception:                       # library not built (CPU-only checkout): slow but correct
    return crc32c(data)


def _mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _put_varint(v):
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _field(num, wt, payload):
  return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype, shape, offset, size, crc):
  dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
  msg = _field(1, 0, _put_varint(dtype)) + _field(2, 2, _put_varint(len(dims)) + dims)
  if offset:
    msg += _field(4, 0, _put_varint(offset))
  msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", _mask_crc(crc)))
  return msg


def _build_block(items, restart_interval=16):
  """LevelDB block of (key, value) pairs (keys sorted), prefix-compressed, + restart array."""
  out, restarts, prev = bytearray(), [], b""
  for i, (k, v) in enumerate(items):
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


def _emit_block(fh, block):
  off = fh.tell()
  trailer = b"\x00"                                   # no compression
  fh.write(block + trailer + struct.pack("<I", _mask_crc(crc32c(block + trailer))))
  return off, len(block)


_NP2TF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def save_bundle(prefix, tensors, block_bytes=4096):
  """Write {name: numpy array} as a single-shard V2 bundle.  Names are stored sorted (bytewise), the
  tensor bytes back to back in that order, like tf.train.Saver."""
  names = sorted(tensors, key=lambda s: s.encode())
  entries, offset = [], 0
  with open(prefix + ".data-00000-of-00001", "wb") as fh:
    for n in names:
      a = np.asarray(tensors[n])          # (np.ascontiguousarray would turn 0-d scalars into shape (1,))
      raw = a.tobytes()                   # C order
      entries.append((n.encode(), _entry_proto(_NP2TF[a.dtype], a.shape, offset, len(raw), _crc32c_fast(raw))))
      fh.write(raw)
      offset += len(raw)
  header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # num_shards 1, version.producer 1
  items = [(b"", header)] + entries
  with open(prefix + ".index", "wb") as fh:
    index_items, cur, cur_bytes = [], [], 0
    for k, v in items:
      cur.append((k, v))
      cur_bytes += len(k) + len(v) + 3
      if cur_bytes >= block_bytes:
        off, size = _emit_block(fh, _build_block(cur))
        index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
        cur, cur_bytes = [], 0
    if cur:
      off, size = _emit_block(fh, _build_block(cur))
      index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
    meta_off, meta_size = _emit_block(fh, _build_block([]))
    idx_off, idx_size = _emit_block(fh, _build_block(index_items, restart_interval=1))
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    fh.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))


def verify_bundle(prefix):
  """Re-read a bundle and check every block CRC and every tensor CRC.  Returns the entry count."""
  buf = open(prefix + ".index", "rb").read()
  footer = buf[-48:]
  pos = 0
  handles = []
  for _ in range(4):
    v, pos = _varint(footer, pos)
    handles.append(v)

  def check(off, size):
    want = struct.unpack_from("<I", buf, off + size + 1)[0]
    if _mask_crc(crc32c(buf[off:off + size + 1])) != want:
      raise ValueError("block CRC mismatch at %d" % off)
  check(handles[0], handles[1]); check(handles[2], handles[3])
  data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
  n = 0
  for _, handle in _block_entries(buf, handles[2], handles[3]):
    off, p = _varint(handle, 0)
    size, p = _varint(handle, p)
    check(off, size)
    for key, value in _block_entries(buf, off, size):
      if not key:
        continue
      f = _parse_proto(value)
      o, s = f.get(4, [0])[0], f.get(5, [0])[0]
      want = struct.unpack("<I", f[6][0])[0]
      if _mask_crc(_crc32c_fast(data[o:o + s].tobytes())) != want:
        raise ValueError("tensor CRC mismatch for %s" % key.decode())
      n += 1
  return n


_OPT_SCOPE = {"rl_value": "OptimizeLoss", "generator": "OptimizeLoss_1", "critic": "OptimizeLoss_2"}


def export_checkpoint(trainer):
  """Every variable tf.train.Saver would save for this model (names as in the shipped checkpoint):
  parameters, Adam slots `<scope>/<var>/Adam{,_1}`, beta powers, the three global steps, the EMA."""
  from .checkpoint import export_named
  from .nets import FC1
  out = {}
  for key, store, counter in (("generator", trainer.gen, trainer.counter_g), ("rl_value", trainer.val, trainer.counter_v),
                              ("critic", trainer.cri, trainer.counter_c)):
    scope = _OPT_SCOPE[key]
    for name, (off, k) in store.offsets.items():
      shape = store.p[name].shape
      views = {"": store.flat, "/Adam": store.m, "/Adam_1": store.v}
      for suffix, flat in views.items():
        t = flat[off:off + k].view(shape).detach().cpu().numpy()
        if "/filter_fc1_all/" in name:
          kind = name.rsplit("/", 1)[1]
          for j in range(8):
            blk = t[:, j * FC1:(j + 1) * FC1] if kind == "weights" else t[j * FC1:(j + 1) * FC1]
            base = "generator/filter_%d/fc1/%s" % (j, kind)
            out[(scope + "/" + base + suffix) if suffix else base] = np.ascontiguousarray(blk)
        else:
          out[(scope + "/" + name + suffix) if suffix else name] = t
    b1, b2 = float(trainer.cfg.adam_beta1), float(trainer.cfg.adam_beta2)
    out[scope + "/beta1_power"] = np.float32(b1 ** (counter + 1)).reshape(())     # TF stores beta^(t+1) after t updates
    out[scope + "/beta2_power"] = np.float32(b2 ** (counter + 1)).reshape(())
  out["Variable"] = np.int32(trainer.counter_v).reshape(())
  out["Variable_1"] = np.int32(trainer.counter_g).reshape(())
  out["Variable_2"] = np.int32(trainer.counter_c).reshape(())
  ema = getattr(trainer, "ema", {"value": 0.0, "biased": 0.0, "local_step": 0.0})
  out["mul_8/ExponentialMovingAverage"] = np.float32(ema["value"]).reshape(())
  out["mul_8/ExponentialMovingAverage/biased"] = np.float32(ema["biased"]).reshape(())
  out["mul_8/ExponentialMovingAverage/local_step"] = np.float32(ema["local_step"]).reshape(())
  return out


def save_checkpoint(trainer, prefix):
  save_bundle(prefix, export_checkpoint(trainer))


def restore_checkpoint(trainer, prefix, optimizer_state=True):
  """Saver.restore: parameters and, when present, Adam slots and step counters."""
  import torch
  from .checkpoint import import_named
  from .nets import FC1
  allv = load_bundle(prefix, include_optimizer_slots=True)
  named = {k: torch.from_numpy(np.array(v)) for k, v in allv.items() if v.dtype == np.float32 and v.ndim > 0 and "/Adam" not in k}
  import_named(trainer, named)
  if not optimizer_state:
    return
  for key, store in (("generator", trainer.gen), ("rl_value", trainer.val), ("critic", trainer.cri)):
    scope = _OPT_SCOPE[key]
    for name, (off, k) in store.offsets.items():
      shape = store.p[name].shape
      for suffix, flat in (("/Adam", store.m), ("/Adam_1", store.v)):
        dst = flat[off:off + k].view(shape)
        if "/filter_fc1_all/" in name:
          kind = name.rsplit("/", 1)[1]
          for j in range(8):
            src = allv.get("%s/generator/filter_%d/fc1/%s%s" % (scope, j, kind, suffix))
            if src is None:
              continue
            s = torch.from_numpy(np.array(src)).to(dst.device)
            if kind == "weights":
              dst[:, j * FC1:(j + 1) * FC1].copy_(s)
            else:
              dst[j * FC1:(j + 1) * FC1].copy_(s)
        else:
          src = allv.get(scope + "/" + name + suffix)
          if src is not None:
            dst.copy_(torch.from_numpy(np.array(src)).to(dst.device).reshape(shape))
  if "Variable" in allv:
    trainer.counter_v = int(allv["Variable"]); trainer.counter_g = int(allv["Variable_1"]); trainer.counter_c = int(allv["Variable_2"])
  if "mul_8/ExponentialMovingAverage" in allv:
    trainer.ema = {"value": float(allv["mul_8/ExponentialMovingAverage"]), "biased": float(allv["mul_8/ExponentialMovingAverage/biased"]),
                   "local_step": float(allv["mul_8/ExponentialMovingAverage/local_step"])}
