"""Builds libxmc.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc. No GPU needed to build."""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libxmc.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
  deps.append(os.path.join(PKG_DIR, "..", "include", "xmc.h"))
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
  if not force and not needs_build():
    return LIB_PATH
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  objs = []
  build_dir = os.path.join(PKG_DIR, "build")
  os.makedirs(build_dir, exist_ok=True)
  procs = []
  for src in sources():
    obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
    if verbose:
      cmd.insert(1, "-Xptxas=-v")
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  for src, pr in procs:
    out, _ = pr.communicate()
    if pr.returncode != 0:
      raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    if verbose and out:
      print(out)
  cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
  subprocess.check_call(cmd)
  return LIB_PATH


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
