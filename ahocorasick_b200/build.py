"""Build libacgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment knobs (tools/gpu_quick.sh A/B runs): extra nvcc flags and another output file; the product build uses neither
EXTRA = os.environ.get("ACGPU_NVCC_EXTRA", "").split()
LIB = os.environ.get("ACGPU_LIB_OUT") or os.path.join(HERE, "libacgpu.so")
SOURCES = ["engine.cu", "builder.cpp"]
TIER_KS = range(1, 9)  # tier_inst.cu is compiled once per K (-DTIER_K=k), in parallel
HEADERS = ["kernels.cuh", "kernel_tier.cuh", "kernel_mask.cuh", "kernel_pair.cuh", "kernel_emit.cuh", "kernel_fuse.cuh", "kernel_wide.cuh", "kernel_sel2.cuh", "kernel_ww.cuh", "kernel_ww3.cuh", "kernel_wwlit.cuh", "tier_launch.hpp", "tier_inst.cu", "device_tables.cuh", "builder.hpp", "trie_insert.hpp",
           "java_char_tables.h", os.path.join("..", "..", "include", "acgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall"]
OBJ_DIR = os.path.join(HERE, "build" + ("_" + os.path.basename(LIB) if os.environ.get("ACGPU_LIB_OUT") else ""))


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libacgpu.so cannot be built")
    return p


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(job):
    src, obj, extra, verbose = job
    cmd = [nvcc_path()] + NVCC_FLAGS + EXTRA + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit for sm_100a (in parallel) and link libacgpu.so in-tree."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = [(os.path.join(CSRC, s), os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o"), [], verbose) for s in SOURCES]
    jobs += [(os.path.join(CSRC, "tier_inst.cu"), os.path.join(OBJ_DIR, "tier_k%d.o" % k), ["-DTIER_K=%d" % k], verbose)
             for k in TIER_KS]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile, jobs))
    log = "".join(out for _, out in results)
    if any(rc != 0 for rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libacgpu.so")
    cmd = [nvcc_path(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [j[1] for j in jobs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(log + r.stdout + r.stderr)
        raise RuntimeError("linking libacgpu.so failed")
    if verbose:
        sys.stderr.write(log + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
