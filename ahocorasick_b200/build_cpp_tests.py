"""Build tests/cpp/reference_style_test (g++, C++17) against include/acgpu.hpp and the in-tree libacgpu.so."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "tests", "cpp", "reference_style_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "build")
EXE = os.path.join(OUT_DIR, "reference_style_test")
DEPS = [SRC, os.path.join(ROOT, "include", "acgpu.hpp"), os.path.join(ROOT, "include", "acgpu.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in DEPS):
        return EXE
    os.makedirs(OUT_DIR, exist_ok=True)
    # $ORIGIN-relative rpath: the binary travels with the repo snapshot to the GPU box
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), SRC,
           "-L" + HERE, "-lacgpu", "-Wl,-rpath,$ORIGIN/../../../ahocorasick_b200", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building tests/cpp/reference_style_test")
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
