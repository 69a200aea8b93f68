"""In-tree build of the C-ABI CUDA library (sm_100a only).

``python -m exposure_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``exposure_b200/_lib/libexposure_b200.so`` is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libexposure_b200.so")
STAMP = os.path.join(OUT_DIR, "build.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
OBJ_DIR = os.path.join(OUT_DIR, "obj")
# development builds only, e.g. EXPOSURE_NVCC_EXTRA=-DEXPO_TMA_TRACE (tools/tma_trace.py); part of the build digest
NVCC_FLAGS += os.environ.get("EXPOSURE_NVCC_EXTRA", "").split()


def _sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
  h = hashlib.sha256()
  for root in (CSRC, os.path.join(HERE, "..", "include")):
    for f in sorted(os.listdir(root)):
      if f.endswith((".cu", ".cuh", ".h")):
        with open(os.path.join(root, f), "rb") as fh:
          h.update(f.encode() + b"\0" + fh.read())
  h.update((" ".join(NVCC_FLAGS) + " link:gencode-sm_100a").encode())
  return h.hexdigest()


def build_lib(force=False, verbose=False):
  """Compile every .cu under csrc/ into one shared library.  Returns its path."""
  os.makedirs(OUT_DIR, exist_ok=True)
  digest = _digest()
  if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
    with open(STAMP) as fh:
      if fh.read().strip() == digest:
        return LIB_PATH
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  os.makedirs(OBJ_DIR, exist_ok=True)
  # one translation unit per .cu, compiled concurrently (nvcc is single-threaded per file), then linked
  jobs = []
  for src in _sources():
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    jobs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  failed = False
  for src, obj, proc in jobs:
    out, _ = proc.communicate()
    if proc.returncode != 0:
      failed = True
      sys.stderr.write(out)
    elif verbose:
      sys.stderr.write(out)
  if failed:
    raise RuntimeError("nvcc failed building %s" % LIB_PATH)
  res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [j[1] for j in jobs],
                       capture_output=True, text=True)
  if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    raise RuntimeError("nvcc failed linking %s" % LIB_PATH)
  with open(STAMP, "w") as fh:
    fh.write(digest)
  return LIB_PATH


if __name__ == "__main__":
  print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
