"""Builds tsp_gnn_b200/libtspgnn.so for sm_100a with nvcc (in-tree, so it travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtspgnn.so")
SOURCES = ["tspgnn.cu"]
HEADERS = ["common.cuh", "simt_kernels.cuh", "tc_ptx.cuh", "tc_kernels.cuh", "tc_fused.cuh", "train_kernels.cuh", "tc_train.cuh", "generic_kernels.cuh", "train_host.inc"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math=false",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "tspgnn.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("TSPGNN_NVCC_EXTRA", "").split()
    cmd = [nvcc] + flags + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libtspgnn.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
