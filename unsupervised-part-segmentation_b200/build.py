"""In-tree build of the CUDA library (sm_100a) and of the host-side test helper."""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libups_b200.so")
HOST_LIB = os.path.join(CSRC, "libups_canon_host.so")
CU_SOURCES = ["cabi.cu", "tps.cu", "parts_ops.cu", "step_fused.cu", "step_decode_bwd_tma.cu", "stats_ops.cu", "priors_ops.cu", "ingest.cu",
              "inject_conv.cu", "parts_conv.cu"]
HEADERS = ["common.cuh", "canon_math.cuh", "pk_math.cuh", os.path.join("..", "..", "include", "ups_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps)


def build_cuda(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libups_b200.so"""
    if not force and not _stale(LIB, CU_SOURCES + HEADERS):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + CU_SOURCES + ["-o", LIB]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


def build_host_helper(force=False):
    """g++ build of canon_math.cuh for CPU bit-exactness tests (test support only)."""
    if not force and not _stale(HOST_LIB, ["canon_host.cpp", "canon_math.cuh"]):
        return HOST_LIB
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "canon_host.cpp",
                           "-o", HOST_LIB], cwd=CSRC)
    return HOST_LIB


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
    print(build_host_helper(force=True))
