"""Build the native CUDA library in-tree (``audiolazy_b200/_native/libalz_b200.so``).

``nvcc`` cross-compiles for sm_100a without a GPU; the built ``.so`` is git-ignored
but travels to the GPU box with the repository snapshot. The translation units
(``csrc/*.cu``: the C ABI plus one unit of kernel instantiations per cascade length) are
compiled in parallel into ``_native/obj/`` and linked into ONE shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
NATIVE_DIR = os.path.join(_PKG, "_native")
OBJ_DIR = os.path.join(NATIVE_DIR, "obj")
LIB_PATH = os.path.join(NATIVE_DIR, "libalz_b200.so")
INCLUDE = os.path.join(os.path.dirname(_PKG), "include")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + [
  "-O3", "-lineinfo", "-std=c++17",
  "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",   # only the extern "C" ABI of include/alz_b200.h is exported
]


def _units():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
         [os.path.join(INCLUDE, "alz_b200.h")]


def _sources():
  return _units() + _headers()


def is_stale() -> bool:
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > t for s in _sources())


def find_nvcc():
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  return nvcc if os.path.exists(nvcc) else None


def build_native(force: bool = False, verbose: bool = False) -> str:
  """Compile ``csrc/*.cu`` for sm_100a (only the units that changed) and link the library."""
  if not force and not is_stale():
    return LIB_PATH
  nvcc = find_nvcc()
  if nvcc is None:
    raise RuntimeError("nvcc not found: cannot build audiolazy_b200's CUDA library")
  os.makedirs(OBJ_DIR, exist_ok=True)
  newest_header = max(os.path.getmtime(h) for h in _headers())
  jobs = []
  for src in _units():
    obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
    if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
      jobs.append((src, obj))

  def compile_one(job):
    src, obj = job
    tmp = obj + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", tmp, src]
    subprocess.check_call(cmd)
    os.replace(tmp, obj)

  with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
    list(pool.map(compile_one, jobs))
  objs = [os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(s))[0] + ".o") for s in _units()]
  tmp = LIB_PATH + ".tmp.%d" % os.getpid()
  subprocess.check_call([nvcc] + ARCH_FLAGS + ["-shared", "-o", tmp] + objs)
  os.replace(tmp, LIB_PATH)
  return LIB_PATH
