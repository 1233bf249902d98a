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
    tmp = PYHOST + ".tmp.%d" % os.getpid()
    cmd = [cc, "-O3", "-fPIC", "-shared", "-pthread", "-I", sysconfig.get_paths()["include"], "-o", tmp,
           os.path.join(CSRC, "vq_pyhost.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvq_pyhost.so failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, PYHOST)
    return PYHOST


def build(force=False, verbose=False):
    """Compile when a source is newer than the libraries.  Several processes may get here at once (torchrun starts one
    bench.py per GPU): one of them builds under a file lock, into temporary files that are renamed into place, and the
    others find the work done when they get the lock."""
    if not force and not needs_build():
        return LIB
    import fcntl
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    with open(os.path.join(os.path.dirname(LIB), ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", tmp,
                   *[os.path.join(CSRC, f) for f in SOURCES]]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            cmd[1:1] = os.environ.get("VQ_NVCC_EXTRA", "").split()
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            if verbose:
                print(r.stderr)
            os.replace(tmp, LIB)
            build_pyhost()
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
