"""Builds libccsim.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(_HERE, "csrc", "ccsim_engine.cu")]
DEPS = SRC + [os.path.join(_HERE, "csrc", "ccsim_device.cuh"), os.path.join(_HERE, "..", "include", "ccsim.h")]
OUT = os.path.join(_HERE, "libccsim.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-lcudart", "-ldl"]


def build(force=False, verbose=False):
    newest = max(os.path.getmtime(p) for p in DEPS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
