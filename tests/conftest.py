import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

from vit_testlib import PKG_DIR, has_gpu, load_pkg  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def V():
    return load_pkg()


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def emu():
    """Host SIMT emulator of the product kernel source (tests/emu), built on demand."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(d, "libvitemu.so")
    srcs = [os.path.join(d, "vit_emu.cpp"), os.path.join(PKG_DIR, "csrc", "vit_kernel.cuh"), os.path.join(PKG_DIR, "csrc", "vit_kernel_map.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.vit_emu_decode.restype = C.c_int
    L.vit_emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_size_t]
    return L

