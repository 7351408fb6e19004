"""Test infrastructure: build tests/cpp/reference_style_test (g++, C++17) against include/acgpu.hpp and the in-tree
libacgpu.so, and its oracle-mocked twin for CPU-only runs."""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "ahocorasick_b200")  # where libacgpu.so is built
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


def build_mock(force: bool = False) -> str:
    """tests/cpp/build/reference_style_test_mock: the same test program linked against libacgpu_mock_oracle.so (the C ABI
    answered by the CPU oracle; tests/cpp/mock_acgpu_oracle.cpp).  CPU-only test runs use it to exercise the C++ host
    logic; it is test infrastructure and never part of the product."""
    mock_src = os.path.join(ROOT, "tests", "cpp", "mock_acgpu_oracle.cpp")
    oracle_dir = os.path.join(ROOT, "oracle")
    exe = EXE + "_mock"
    lib = os.path.join(OUT_DIR, "libacgpu_mock_oracle.so")
    deps = DEPS + [mock_src, os.path.join(oracle_dir, "ac_oracle.h"), os.path.join(oracle_dir, "liboracle.so")]
    if not force and os.path.exists(exe) and os.path.exists(lib) and all(
            os.path.getmtime(d) <= min(os.path.getmtime(exe), os.path.getmtime(lib)) for d in deps):
        return exe
    os.makedirs(OUT_DIR, exist_ok=True)
    cmds = [
        ["g++", "-std=c++17", "-O1", "-Wall", "-shared", "-fPIC", mock_src, "-L" + oracle_dir, "-l:liboracle.so",
         "-Wl,-rpath,$ORIGIN/../../../oracle", "-o", lib],
        ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + OUT_DIR,
         "-lacgpu_mock_oracle", "-Wl,-rpath,$ORIGIN", "-o", exe],
    ]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building the mock-backed C++ test")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
