import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        import t1k_b200._lib as L
        n = ctypes.c_int(0)
        L.lib().t1k_device_count(ctypes.byref(n))
        return n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def emu():
    src = os.path.join(ROOT, "tests", "host_emu.cpp")
    so = os.path.join(ROOT, "tests", "_build", "libhostemu.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "t1k_b200", "csrc", f) for f in ("t1k_core.cuh", "t1k_host.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-w", "-std=c++14", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_create.restype = C.c_void_p
    lib.emu_create.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]
    lib.emu_destroy.argtypes = [C.c_void_p]
    lib.emu_assign.restype = C.c_int32
    lib.emu_assign.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.emu_coverage.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.emu_set_fast.argtypes = [C.c_void_p, C.c_int32]
    lib.emu_counters.restype = C.POINTER(C.c_longlong)
    lib.emu_align.restype = C.c_int32
    lib.emu_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_align_info.restype = C.c_int32
    lib.emu_align_info.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    return lib
