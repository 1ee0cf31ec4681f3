"""Builds libctr_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libctr_b200.so")
SRC = os.path.join(HERE, "csrc")
INC = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(INC, f) for f in os.listdir(INC)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = SO + ".tmp.%d" % os.getpid()          # built aside and renamed: a snapshot of the tree never sees a half-written library
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, os.path.join(SRC, "engine.cu")]
    if verbose:
        print(" ".join(cmd))
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, SO)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
