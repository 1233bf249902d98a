"""Build libvq_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libvq_b200.so")
SOURCES = ["vq_store.cu", "vq_scan.cu", "vq_labelled.cu", "vq_batch.cu", "vq_exchange.cu", "vq_ingest.cu", "vq_hostx.cu", "vq_rng.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr"]


PYHOST = os.path.join(PKG, "lib", "libvq_pyhost.so")     # CPython-API glue (record unboxing), plain C, no CUDA


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(PYHOST):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(PYHOST))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "vq.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_pyhost():
    import sysconfig
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O3", "-fPIC", "-shared", "-pthread", "-I", sysconfig.get_paths()["include"], "-o", PYHOST,
           os.path.join(CSRC, "vq_pyhost.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvq_pyhost.so failed:\n" + r.stdout + r.stderr)
    return PYHOST


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", LIB,
           *[os.path.join(CSRC, f) for f in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    cmd[1:1] = os.environ.get("VQ_NVCC_EXTRA", "").split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    build_pyhost()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
