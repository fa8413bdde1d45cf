"""In-tree build of the CUDA library (sm_100a) and of the host-side test helper.

Every translation unit is compiled to its own object (in parallel) and linked into
csrc/libups_b200.so.  The library carries a hash of the sources it was built from
(`ups_version()` -> "... src=<12 hex>"): `build_cuda()` rebuilds whenever that hash
differs from the hash of the sources in the tree, so a binary that does not correspond to
the committed sources cannot travel to the GPU box unnoticed (modification times play no part).
"""
import hashlib
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libups_b200.so")
HOST_LIB = os.path.join(CSRC, "libups_canon_host.so")
CU_SOURCES = ["cabi.cu", "tps.cu", "parts_ops.cu", "step_fused.cu", "step_decode_bwd_tma.cu", "stats_ops.cu", "priors_ops.cu", "ingest.cu",
              "inject_conv.cu", "parts_conv.cu", "dp_allreduce.cu", "standin.cu", "step_fwd_fused.cu", "parts_conv_bwd_tc.cu"]
HEADERS = ["common.cuh", "canon_math.cuh", "pk_math.cuh", "tps_warp_fwd.cuh", "step_decode_fwd.cuh", "tc_helpers.cuh", os.path.join("..", "..", "include", "ups_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _digest(names):
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(n.encode() + b"\0" + f.read() + b"\0")
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:12]


def source_hash():
    """Hash of every source and header the library is built from (and of the compiler flags)."""
    return _digest(sorted(CU_SOURCES) + sorted(HEADERS))


def built_hash(path=LIB):
    """The source hash recorded inside a built library (read from the file, not through dlopen), or None."""
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        m = re.search(rb"ups_b200 [0-9.]+ \(sm_100a, src=([0-9a-f]{12})\)", f.read())
    return m.group(1).decode() if m else None


def _compile_one(src, full_hash, verbose):
    """One translation unit -> build/<src>.o, skipped if its own inputs (source + headers) are unchanged."""
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, src + ".o")
    stamp = obj + ".hash"
    # cabi.cu embeds the hash of the whole tree, so it is rebuilt whenever anything changes
    want = _digest([src] + sorted(HEADERS)) + (full_hash if src == "cabi.cu" else "")
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [f"-DUPS_SRC_HASH=\"{full_hash}\"", "-c", src, "-o", obj]
    subprocess.check_call(cmd, cwd=CSRC)
    with open(stamp, "w") as f:
        f.write(want)
    return obj


def build_cuda(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> csrc/libups_b200.so"""
    full = source_hash()
    if not force and built_hash() == full:
        return LIB
    if force and os.path.isdir(OBJ):
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, full, verbose), CU_SOURCES))
    nvcc = os.environ.get("NVCC", "nvcc")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs +
                          ["-o", LIB], cwd=CSRC)
    assert built_hash() == full, (built_hash(), full)
    return LIB


def build_host_helper(force=False):
    """g++ build of canon_math.cuh for CPU bit-exactness tests (test support only)."""
    deps = ["canon_host.cpp", "canon_math.cuh"]
    stamp = HOST_LIB + ".hash"
    want = _digest(deps)
    if not force and os.path.exists(HOST_LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return HOST_LIB
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "canon_host.cpp",
                           "-o", HOST_LIB], cwd=CSRC)
    with open(stamp, "w") as f:
        f.write(want)
    return HOST_LIB


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_helper())
