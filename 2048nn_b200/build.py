"""Build lib2048nn_b200.so in-tree with nvcc for sm_100a (B200).  No GPU needed to compile."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("B2048_LIB") or os.path.join(HERE, "lib2048nn_b200.so")  # B2048_LIB: build a variant elsewhere
SOURCES = ["b2048_capi.cu", "b2048_train.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b2048.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    exp = [f"-D{d}" for d in os.environ.get("B2048_EXP", "").split(",") if d]
    cmd = [nvcc] + NVCC_FLAGS + exp + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)  # the image's CC wrapper lacks some specs; use nvcc's default host gcc
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building lib2048nn_b200.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
