"""CPU-only test of the Python host mirror (matchers.py, streaming.py): the C ABI is mocked by the oracle
(tests/cpp/mock_acgpu_oracle.cpp, loaded through the ACGPU_LIB override in a subprocess so this process keeps the real
library), the mirror's packing, zipping, replay quirks and Readable block logic run end to end and must reproduce the
oracle's literal listener-call sequences.  The kernels are not involved - tests/test_gpu_parity.py covers those."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_mirror_against_mocked_abi():
    from oracle import oracle as ora
    import cpp_build
    ora.build()
    cpp_build.build_mock()
    env = dict(os.environ, ACGPU_LIB=os.path.join(ROOT, "tests", "cpp", "build", "libacgpu_mock_oracle.so"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "python_host_mock_checks.py")], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0 and "python host mirror ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
