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
def port_oracle():
    """the restated oracle only: for tests whose point does not depend on the checker (e.g. two kernel variants bit-identical to each other)"""
    import pyoracle
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


# The CPU suite runs every oracle-using test against both checkers ("port" = restated oracle, "ref" = the reference's own compiled kernels). For
# the tests that spend their time in the HOST-THREAD EMULATION of the CUDA kernels (tens of seconds each on a small container), the second
# checker only repeats the emulated run: tests/test_ref_oracle.py and tests/golden/ref_kernel_vectors.npz already tie the two checkers together at
# 1e-13. Those tests keep the "port" variant here unless HEXED_B200_ALL_ORACLES=1; the GPU suite (-m gpu) always runs both.
HEAVY_EMULATED = ("test_vortex_emulated", "test_adapter_multi_device_box_morton_emu", "test_update_loops_with_chebyshev_steps",
                  "test_max_dt_running_screen_is_exact", "test_box_3d_other_local_kernels", "test_navier_stokes_3d_line_kernel",
                  "test_naca_class_emulated", "test_adapter_multi_device_refined_box_emu", "test_adapter_multi_device_euler_emu",
                  "test_update_euler_device_time_step", "test_av_smoothness_pipeline", "test_adapter_async_boundary_traffic_emu",
                  "test_navier_stokes_2d_line_kernel", "test_box_2d_pipelined_local", "test_box_3d_pipelined_local", "test_cylinder_class_emulated",
                  "test_update_navier_stokes_device_time_step", "test_adapter_multi_device_navier_stokes_emu")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("HEXED_B200_ALL_ORACLES") == "1":
        return
    keep, dropped = [], []
    for item in items:
        name = item.name.split("[")[0]
        heavy = name in HEAVY_EMULATED and "[ref" in item.name and item.get_closest_marker("gpu") is None
        (dropped if heavy else keep).append(item)
    if dropped:
        config.hook.pytest_deselected(items=dropped)
        items[:] = keep
