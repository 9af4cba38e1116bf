import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ref_possible():
    import pyoracle
    return pyoracle.ref_available() or os.path.isdir("/root/reference")


@pytest.fixture(scope="session", params=["port", "ref"])
def oracle(request):
    """the checker every parity test compares against: "port" = the restated CPU oracle (oracle/oracle_impl.hpp), "ref" = the
    REFERENCE'S OWN kernels compiled from /root/reference (oracle/_ref/libhexed_ref.so, built here, shipped to the GPU box);
    both expose the same interface (pyoracle.Oracle / pyoracle.RefOracle), so every test runs against both."""
    import pyoracle
    if request.param == "ref":
        if not _ref_possible():
            pytest.skip("oracle/_ref/libhexed_ref.so is absent and /root/reference is not here to build it from")
        return pyoracle.RefOracle()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def emu_lib():
    """host-thread emulation build of the CUDA sources (tests only; see tests/emu/cuda_emu.hpp)"""
    import subprocess
    import fcntl
    path = os.path.join(ROOT, "tests", "emu", "libhexed_b200_emu.so")
    with open(os.path.join(ROOT, "tests", "emu", ".build.lock"), "w") as lock:  # pytest-xdist workers must not rebuild it under each other
        fcntl.flock(lock, fcntl.LOCK_EX)
        subprocess.run(["make", "-C", os.path.join(ROOT, "hexed_b200", "csrc"), "-j8", "emu"], check=True, stdout=subprocess.DEVNULL)
    return path


@pytest.fixture(scope="session")
def gpu_lib():
    """the product library; the GPU tests fail loudly if it is missing or no device is present"""
    from hexed_b200.kernels import LIB_PATH, load_library
    import ctypes
    lib = load_library(LIB_PATH)
    n = ctypes.c_int(0)
    lib.hexed_b200_device_count(ctypes.byref(n))
    assert n.value > 0, "GPU test selected but no CUDA device is visible"
    return LIB_PATH
