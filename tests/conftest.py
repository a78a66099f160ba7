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


EMU_DIR = os.path.join(ROOT, "tests", "emu")
SIM_DIR = os.path.join(ROOT, "tests", "sim")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
KERNEL_SRCS = [os.path.join(PKG_DIR, "csrc", f) for f in ("vit_kernel.cuh", "vit_kernel_map.inc", "vit_kernel_l1.inc", "vit_code.h")]


def _emu_target(name, flags):
    """(output, sources it depends on, command): tests/emu/vit_emu.cpp = the product kernel source compiled for the host"""
    so = os.path.join(EMU_DIR, name)
    srcs = [os.path.join(EMU_DIR, "vit_emu.cpp")] + KERNEL_SRCS
    return so, srcs, [GXX, "-O1", "-std=c++17", "-fPIC", "-shared"] + list(flags) + ["-o", so, srcs[0]]


def _alt_emu_target(p1, p2):
    return _emu_target("libvitemu_p%o_%o.so" % (p1, p2), ["-DVIT_EMU_L8_ONLY", "-DVIT_POLY1=0%o" % p1, "-DVIT_POLY2=0%o" % p2])


def _sim_target():
    """tests/sim: the library's host code (csrc/vit_api.cu) compiled for the host against the stand-in CUDA runtime + emulator"""
    so = os.path.join(SIM_DIR, "libvitsim.so")
    api = os.path.join(PKG_DIR, "csrc", "vit_api.cu")
    srcs = [api, os.path.join(SIM_DIR, "sim_runtime.cpp"), os.path.join(EMU_DIR, "vit_emu.cpp"), os.path.join(SIM_DIR, "sim_stubs.cpp"),
            os.path.join(SIM_DIR, "cuda_runtime.h"), os.path.join(ROOT, "include", "vit_b200.h")] + KERNEL_SRCS
    srcs += [os.path.join(PKG_DIR, "csrc", f) for f in ("vit_stage_pool.h", "vit_launch.h", "vit_internal.h")]
    return so, srcs, [GXX, "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DVIT_EMU_L8_ONLY", "-DVIT_EMU_WITH_L1", "-I", SIM_DIR, "-x", "c++",
                      api, srcs[1], srcs[2], srcs[3], "-o", so]


def _stale(target):
    so, srcs, _ = target
    return not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)


_prebuilt = False


def _build_host_libs(wanted):
    """Build `wanted` if stale.  The first call also starts every OTHER stale host-side test library in parallel (the
    emulator takes 1.5 minutes to compile; in a fresh checkout the four of them would otherwise queue up)."""
    global _prebuilt
    from vit_testlib import ALT_EMU_PAIRS
    targets = [wanted]
    if not _prebuilt:
        _prebuilt = True
        targets += [t for t in [_emu_target("libvitemu.so", []), _sim_target()] + [_alt_emu_target(*p) for p in ALT_EMU_PAIRS]
                    if t[0] != wanted[0]]
    procs = [(t, subprocess.Popen(t[2])) for t in targets if _stale(t)]
    for t, p in procs:
        if p.wait() != 0:
            raise RuntimeError("build failed: " + " ".join(t[2]))


def _build_emu(name, flags):
    """tests/emu/<name>, rebuilt when stale, loaded with its entry points typed"""
    import ctypes as C
    target = _emu_target(name, flags)
    _build_host_libs(target)
    L = C.CDLL(target[0])
    L.vit_emu_decode.restype = C.c_int
    L.vit_emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_size_t]
    L.vit_emu_polynomials.restype, L.vit_emu_polynomials.argtypes = None, [C.POINTER(C.c_int)] * 2
    return L


@pytest.fixture(scope="session")
def sim_lib_path():
    """tests/sim/libvitsim.so, built on demand"""
    target = _sim_target()
    _build_host_libs(target)
    return target[0]


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

