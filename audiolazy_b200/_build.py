"""Build the native CUDA library in-tree (``audiolazy_b200/_native/libalz_b200.so``).

``nvcc`` cross-compiles for sm_100a without a GPU; the built ``.so`` is git-ignored
but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
NATIVE_DIR = os.path.join(_PKG, "_native")
LIB_PATH = os.path.join(NATIVE_DIR, "libalz_b200.so")
INCLUDE = os.path.join(os.path.dirname(_PKG), "include")

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a",
  "-O3", "-lineinfo", "-std=c++17",
  "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
         [os.path.join(INCLUDE, "alz_b200.h")]


def is_stale() -> bool:
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > t for s in _sources())


def find_nvcc():
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  return nvcc if os.path.exists(nvcc) else None


def build_native(force: bool = False, verbose: bool = False) -> str:
  """Compile ``csrc/alz_capi.cu`` for sm_100a if the library is missing or stale."""
  if not force and not is_stale():
    return LIB_PATH
  nvcc = find_nvcc()
  if nvcc is None:
    raise RuntimeError("nvcc not found: cannot build audiolazy_b200's CUDA library")
  os.makedirs(NATIVE_DIR, exist_ok=True)
  tmp = LIB_PATH + ".tmp.%d" % os.getpid()
  cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "alz_capi.cu")]
  subprocess.check_call(cmd)
  os.replace(tmp, LIB_PATH)
  return LIB_PATH
