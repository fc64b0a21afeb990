"""Build libb200_decode.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc. No JIT cache: the .so travels
with the repo snapshot to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200_decode.so")
SOURCES = ["c_api.cu"]
HEADERS = ["ptx.cuh", "paged_decode_attn.cuh", "wo_gemm.cuh", "aux_kernels.cuh", "../../include/b200_decode_ops.h"]
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-shared",
              "-Xcompiler", "-fPIC", "-DB200_BUILD"]
if os.environ.get("B200_DEV"):           # developer build: clock64 timeline + ablation switches in the GEMM (tools/gemm_trace.py)
    NVCC_FLAGS.append("-DB200_GEMM_DEV")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
