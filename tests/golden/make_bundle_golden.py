#!/usr/bin/env python
"""Golden vectors for the TF-bundle writer's checksums, taken from the reference's shipped checkpoint
(models/example/pretrained/model.ckpt-20000, written by TF 1.6's tf.train.Saver): the bytes of one
table block of the .index file with the masked CRC-32C stored behind it, and the bytes of one small
tensor with the masked CRC-32C stored in its BundleEntryProto.  Run in the build container (the
reference tree is not available on the GPU box); output: tests/golden/bundle_golden.npz."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from exposure_b200 import tf_bundle as tb  # noqa: E402

prefix = "/root/reference/models/example/pretrained/model.ckpt-20000"
buf = open(prefix + ".index", "rb").read()
footer = buf[-48:]
pos, hs = 0, []
for _ in range(4):
  v, pos = tb._varint(footer, pos)
  hs.append(v)
idx_off, idx_size = hs[2], hs[3]
block = np.frombuffer(buf[idx_off:idx_off + idx_size + 1], dtype=np.uint8)          # contents + compression-type byte
block_crc = struct.unpack_from("<I", buf, idx_off + idx_size + 1)[0]
tensor = tensor_crc = header = None
for _, handle in tb._block_entries(buf, idx_off, idx_size):
  o, q = tb._varint(handle, 0)
  s, q = tb._varint(handle, q)
  for k, v in tb._block_entries(buf, o, s):
    if k == b"":
      header = np.frombuffer(v, dtype=np.uint8)
    if k == b"generator/Conv/biases":
      f = tb._parse_proto(v)
      data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
      tensor = np.array(data[f[4][0]:f[4][0] + f[5][0]])
      tensor_crc = struct.unpack("<I", f[6][0])[0]
np.savez(os.path.join(os.path.dirname(__file__), "bundle_golden.npz"), block=block, block_crc=np.uint32(block_crc), tensor=tensor,
         tensor_crc=np.uint32(tensor_crc), header=header, footer=np.frombuffer(footer, dtype=np.uint8))
print("block %d B crc %08x, tensor %d B crc %08x, header %s" % (block.size, block_crc, tensor.size, tensor_crc, bytes(header).hex()))
