"""The C++ host-side mirror of the reference API (include/acgpu.hpp) above the C ABI.

tests/cpp/reference_style_test.cpp restates the reference's SetTest / MapTest / WholeWordMatchTest strategy in C++
(counting listener + brute-force count per family, plus the ordered streams of SURVEY.md section 8c).  Here: build it,
run its host-only part on the CPU, run all of it on the GPU, and cross-check its --dump output (the streams a C++ user's
listener sees) against the oracle on seeded random inputs for every family x {Set, Map, Readable}.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe():
    from ahocorasick_b200 import build as acbuild
    import cpp_build
    if not os.path.exists(acbuild.LIB):   # normally prebuilt by __graft_entry__.build(); never rebuilt behind the tests' back
        acbuild.build()
    return cpp_build.build()


@pytest.fixture(scope="module")
def mock_exe():
    """The same program linked against libacgpu_mock_oracle.so (tests/cpp/mock_acgpu_oracle.cpp: the C ABI answered by
    the oracle) - exercises the C++ host logic and the test expectations without a GPU.  Test infrastructure only."""
    from oracle import oracle as ora
    import cpp_build
    ora.build()
    return cpp_build.build_mock()


def _run(cmd, timeout=600):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, "exit %d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


def test_cpp_mirror_builds_and_host_checks_pass(exe):
    import torch
    out = _run([exe, "--host-only" if torch.cuda.is_available() else "--no-device"])
    assert " 0 failures" in out


def test_cpp_mirror_declares_every_reference_class():
    """Same class names as the reference's public API."""
    text = open(os.path.join(ROOT, "include", "acgpu.hpp")).read()
    for fam in ("AhoCorasick", "LongestMatch", "ShortestMatch", "WholeWordMatch", "WholeWordLongestMatch"):
        for kind in ("Set", "Map"):
            assert "%s%s" % (fam, kind) in text
    for name in ("StringSet", "StringMap", "SetMatchListener", "MapMatchListener", "ReadableMatchListener", "Readable",
                 "WordCharacters", "Thresholder", "RangeNodeThreshold", "IllegalArgumentException", "getWordChars",
                 "generateWordCharsFlags"):
        assert name in text, name
    assert "oracle" not in text.lower()


def test_cpp_host_logic_against_mocked_abi(mock_exe):
    """Every reference-style C++ test with the C ABI mocked by the oracle: checks include/acgpu.hpp's packing, zipping,
    replay and quirk logic and the expectations of the test program itself; the GPU run below checks the kernels."""
    out = _run([mock_exe], timeout=600)
    assert " 0 failures" in out, out


@pytest.mark.gpu
def test_cpp_reference_style_suite(exe):
    out = _run([exe], timeout=1200)
    assert " 0 failures" in out, out


FAMILIES = ["ahocorasick", "longest", "shortest", "wholeword", "wholewordlongest"]


def _random_case(seed, fam):
    rng = np.random.default_rng(seed)
    alpha = "abcdeXY" if fam < 3 else "abcXY"
    n_kw = int(rng.integers(5, 200))
    kws = sorted({"".join(rng.choice(list(alpha), size=int(rng.integers(1, 7)))) for _ in range(n_kw)})
    if fam == 4:
        kws += [a + " " + b for a, b in zip(kws[::7], kws[3::7])]
    rng.shuffle(kws)
    hay_alpha = list(alpha) + [" ", ",", "x"]
    hay = "".join(rng.choice(hay_alpha, size=int(rng.integers(0, 30000))))
    return list(kws), hay


def test_cpp_usage_example_compiles_and_runs_against_mocked_abi(mock_exe):
    """examples/cpp_usage.cpp = the snippet of INTEGRATION.md section 5; documentation that does not compile is a bug."""
    build_dir = os.path.dirname(mock_exe)
    exe = os.path.join(build_dir, "cpp_usage_mock")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cpp_usage.cpp"), "-L" + build_dir, "-lacgpu_mock_oracle",
                           "-Wl,-rpath," + build_dir, "-o", exe])
    assert _run([exe]).split("\n")[:7] == ["1 4 2", "2 4 1", "2 6 3", "v 2", "v 1", "v 3", "ww 1"]


def test_cpp_dump_plumbing_against_mocked_abi(mock_exe, tmp_path):
    for fam in range(5):
        _check_streams(mock_exe, tmp_path, fam, fam % 2, seeds=(0,))


@pytest.mark.gpu
@pytest.mark.parametrize("fam", range(5), ids=FAMILIES)
@pytest.mark.parametrize("cs", [1, 0], ids=["cs", "ci"])
def test_cpp_streams_equal_oracle(exe, tmp_path, fam, cs):
    _check_streams(exe, tmp_path, fam, cs, seeds=range(3))


def _check_streams(exe, tmp_path, fam, cs, seeds):
    from oracle import oracle as ora
    for seed in seeds:
        kws, hay = _random_case(1000 * fam + 10 * seed + cs, fam)
        kp, hp = str(tmp_path / "kw.u16"), str(tmp_path / "hay.u16")
        open(kp, "wb").write("\n".join(kws).encode("utf-16-le"))
        open(hp, "wb").write(hay.encode("utf-16-le"))
        want = ora.Matcher(FAMILIES[fam], kws, n_values=len(kws), case_sensitive=bool(cs)).match(hay)
        got = np.array([[int(x) for x in ln.split()] for ln in _run([exe, "--dump", str(fam), "set", str(cs), kp, hp]).splitlines()],
                       dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(got[:, 0], want["start"]) and np.array_equal(got[:, 1], want["end"]), (fam, seed)
        got = np.array([[int(x) for x in ln.split()] for ln in _run([exe, "--dump", str(fam), "map", str(cs), kp, hp]).splitlines()],
                       dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(got[:, 0], want["start"]) and np.array_equal(got[:, 1], want["end"])
        assert np.array_equal(got[:, 2], want["value"].astype(np.int64)), (fam, seed)
        wantr = ora.Matcher(FAMILIES[fam], kws, n_values=len(kws), case_sensitive=bool(cs)).match(hay, readable=True)
        gotr = np.array([int(ln) for ln in _run([exe, "--dump", str(fam), "readable", str(cs), kp, hp]).splitlines()], dtype=np.int64)
        assert np.array_equal(gotr, wantr["value"].astype(np.int64)), (fam, seed)
