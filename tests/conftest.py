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


def _build_emu(name, flags):
    """tests/emu/vit_emu.cpp (the product kernel source compiled for the host) -> tests/emu/<name>, rebuilt when stale."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(d, name)
    srcs = [os.path.join(d, "vit_emu.cpp")] + [os.path.join(PKG_DIR, "csrc", f) for f in ("vit_kernel.cuh", "vit_kernel_map.inc", "vit_code.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O1", "-std=c++17", "-fPIC", "-shared"] + list(flags) + ["-o", so, srcs[0]])
    L = C.CDLL(so)
    L.vit_emu_decode.restype = C.c_int
    L.vit_emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_size_t]
    L.vit_emu_polynomials.restype, L.vit_emu_polynomials.argtypes = None, [C.POINTER(C.c_int)] * 2
    return L


@pytest.fixture(scope="session")
def emu():
    """Host SIMT emulator of the product kernel source (tests/emu), built on demand."""
    return _build_emu("libvitemu.so", [])


@pytest.fixture(scope="session")
def emu_for_polynomials():
    """emu_for_polynomials(p1, p2): the emulator compiled for other generator polynomials (csrc/vit_code.h), product lane
    geometry only."""
    cache = {}

    def get(p1, p2):
        if (p1, p2) not in cache:
            cache[(p1, p2)] = _build_emu("libvitemu_p%o_%o.so" % (p1, p2),
                                         ["-DVIT_EMU_L8_ONLY", "-DVIT_POLY1=0%o" % p1, "-DVIT_POLY2=0%o" % p2])
        return cache[(p1, p2)]
    return get

