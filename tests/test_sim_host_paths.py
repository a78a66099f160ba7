"""CPU run of the library's HOST code (csrc/vit_api.cu itself, compiled with a plain C++ compiler) against the kernel
SOURCE in the host emulator, through a stand-in CUDA runtime (tests/sim: test scaffolding, see sim_runtime.cpp).

The stand-in runtime schedules adversarially: "device" memory starts as garbage, a chunk-pipeline kernel runs the moment
its own chunk's bytes have been copied (later uploads are not even enqueued yet), and the gate-waiting kernel of the
time-sliced upload is re-run from scratch at EVERY gate opening with the memory as it is then -- whatever such a partial run
emits must already be final (sim_violations).  So vit_run's paths (sequential, time-sliced with pinned / pageable / mixed
buffers incl. the worker-pool staging, the segment-range chunk pipeline) and vit_stream_push are executed for real --
geometry, strided copies, gates, staging, downloads, carry arithmetic -- and compared with the golden model, over more
stream shapes than the GPU suite samples.  The product has no CPU path: this library exists only in the tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from vit_testlib import PKG_DIR, ROOT

SIM = os.path.join(ROOT, "tests", "sim")
SEQUENTIAL, CHUNKED, GATED = 1, 2, 3
W = 64                       # segments (vit_set_segments): the smallest count the overlapped paths accept


@pytest.fixture(scope="module")
def sim(sim_lib_path):
    so = sim_lib_path
    L = C.CDLL(so)
    vp, sz = C.c_void_p, C.c_size_t
    L.vit_create.restype, L.vit_create.argtypes = C.c_int, [C.POINTER(vp), C.c_int, C.c_int, sz]
    L.vit_destroy.restype, L.vit_destroy.argtypes = None, [vp]
    L.vit_run.restype, L.vit_run.argtypes = C.c_int, [vp, vp, vp, sz, C.POINTER(C.c_float)]
    L.vit_set_segments.restype, L.vit_set_segments.argtypes = C.c_int, [vp, C.c_uint]
    L.vit_set_upload_mode.restype, L.vit_set_upload_mode.argtypes = C.c_int, [vp, C.c_int]
    L.vit_upload_mode_in_effect.restype, L.vit_upload_mode_in_effect.argtypes = C.c_int, [vp]
    L.vit_launch_count.restype, L.vit_launch_count.argtypes = C.c_ulonglong, [vp]
    L.vit_last_error.restype = C.c_char_p
    L.vit_stream_reset.restype, L.vit_stream_reset.argtypes = C.c_int, [vp]
    L.vit_stream_push.restype, L.vit_stream_push.argtypes = C.c_int, [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.vit_stream_pending.restype, L.vit_stream_pending.argtypes = sz, [vp]
    L.sim_pinned_alloc.restype, L.sim_pinned_alloc.argtypes = vp, [sz]
    L.sim_pinned_free.restype, L.sim_pinned_free.argtypes = None, [vp]
    for f in ("sim_violations", "sim_kernel_runs", "sim_gated_attempts"):
        getattr(L, f).restype = C.c_ulonglong
    L.sim_set_table.argtypes = [C.c_int]
    L.sim_check_canaries.restype = None
    return L


@pytest.fixture(autouse=True)
def no_violations(sim):
    """every test: no partial run emitted a non-final word, and nothing wrote outside a device or pinned buffer (the
    stand-in runtime keeps canaries around every allocation and checks them when it is freed)"""
    sim.sim_reset_counters()
    yield
    sim.sim_check_canaries()
    assert sim.sim_violations() == 0


class Pinned:
    """a numpy view of page-locked memory as the stand-in runtime knows it (is_pinned() -> true)"""
    def __init__(self, L, nbytes):
        self.L, self.p = L, L.sim_pinned_alloc(nbytes)
        self.a = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.p))

    def free(self):
        self.L.sim_pinned_free(self.p)


def _make(L, opt):
    h = C.c_void_p()
    assert L.vit_create(C.byref(h), opt, 0, 0) == 0, L.vit_last_error()
    assert L.vit_set_segments(h, W) == 0
    return h


def _run(L, O, h, opt, n_bits, seed, mode, pin_in, pin_out, sigma=0.8):
    bits, packed, N = O.make_channel_det(n_bits, opt & 0xF, seed=seed, sigma=sigma)
    O.set_segments(W)
    try:
        exp = O.decode(opt, packed, N)
    finally:
        O.set_segments(0)
    in_bytes, out_bytes = O.input_size(opt, N), O.output_size(opt, N)
    raw = np.ascontiguousarray(packed).view(np.uint8)[:in_bytes]
    bi = Pinned(L, in_bytes) if pin_in else None
    bo = Pinned(L, out_bytes) if pin_out else None
    src = bi.a if pin_in else raw.copy()
    if pin_in:
        src[:] = raw
    dst = bo.a if pin_out else np.zeros(out_bytes, np.uint8)
    dst[:] = 0x5A
    assert L.vit_set_upload_mode(h, mode) == 0
    rc = L.vit_run(h, src.ctypes.data, dst.ctypes.data, N, None)
    assert rc == 0, L.vit_last_error()
    got = dst.view(exp.dtype).copy()
    for b in (bi, bo):
        if b:
            b.free()
    return got, exp


# (options, message bits): 20+ packs per segment -> >= 8 super-steps (the forced time-sliced mode needs them); ragged
# counts, 16-bit packs with odd tails, hard input (24 channel bytes per super-step), fp32
SHAPES = [
    (0x011, (W * 22 + 7) * 32 + 64 + 5),
    (0x011, (W * 31) * 32 + 64),
    (0x000, (W * 26 + 63) * 32 + 64 + 31),
    (0x112, (W * 45 + 3) * 16 + 64 + 9),
    (0x100, (W * 47 + 33) * 16 + 64),
    (0x004, (W * 21 + 1) * 32 + 64 + 2),
    (0x022, (W * 24 + 40) * 32 + 64),
    (0x2003, (W * 23 + 11) * 32 + 64),
]


@pytest.mark.parametrize("opt,n_bits", SHAPES)
def test_time_sliced_upload_host_code_on_the_cpu(sim, O, opt, n_bits):
    """vit_run in the forced time-sliced mode: one launch per call, every pinned / pageable mix (the pageable sides go
    through the worker pool and the pinned staging buffers), output equal to the golden model, and no partial run of the
    gate-waiting kernel ever emitted a word that differed from the final one."""
    L = sim
    h = _make(L, opt)
    L.sim_reset_counters()
    launches = L.vit_launch_count(h)
    for k, (pi, po) in enumerate(((True, True), (False, False), (True, False), (False, True))):
        got, exp = _run(L, O, h, opt, n_bits, 100 + k, GATED, pi, po)
        assert np.array_equal(got, exp), (hex(opt), pi, po)
    assert L.vit_launch_count(h) - launches == 4                  # ONE launch per call: no fallback happened
    assert L.vit_upload_mode_in_effect(h) == GATED
    L.sim_check_canaries()                                          # no write outside any device / pinned buffer
    assert L.sim_violations() == 0
    assert L.sim_gated_attempts() >= 4 * 2                        # the kernel was re-run at the gate openings (2-4 gates per call)
    L.vit_destroy(h)


@pytest.mark.parametrize("opt,n_bits", SHAPES[:4])
def test_sequential_path_and_small_table_build(sim, O, opt, n_bits):
    L = sim
    h = _make(L, opt)
    for tbl in (96, 32):
        L.sim_set_table(tbl)
        try:
            got, exp = _run(L, O, h, opt, n_bits, 7 + tbl, SEQUENTIAL, False, False)
            assert np.array_equal(got, exp), (hex(opt), tbl)
            got, exp = _run(L, O, h, opt, n_bits // 3, 9 + tbl, GATED, True, True)     # growing then shrinking inputs on one handle
            assert np.array_equal(got, exp), (hex(opt), tbl)
        finally:
            L.sim_set_table(96)
    L.vit_destroy(h)


@pytest.mark.parametrize("opt", [0x011, 0x100, 0x004])
def test_stream_push_host_code_on_the_cpu(sim, O, opt):
    """vit_stream_push: windows = carried symbols ++ chunk, decoded like a run() of the window; the carry arithmetic and the
    device-side tail copies are the library's own code here.  Compared push by push with the oracle's restatement."""
    L = sim
    it = opt & 0xF
    spw = {0: 32, 1: 8, 2: 4, 3: 2, 4: 1}[it]
    h = _make(L, opt)
    n_bits = 40_000
    bits, packed, N = O.make_channel_det(n_bits, it, seed=5, sigma=0.7)
    words = np.ascontiguousarray(packed).view(np.uint32)
    rng = np.random.default_rng(opt)
    cuts = sorted(set(int(x) for x in rng.integers(1, words.size, 9)) | {3, words.size})
    chunks, prev = [], 0
    for c in cuts:
        chunks.append((prev, c))
        prev = c
    O.set_segments(W)
    try:
        exp, pending = O.decode_chunked(opt, packed, [(b - a) * spw for a, b in chunks])
    finally:
        O.set_segments(0)
    assert L.vit_stream_reset(h) == 0
    itemsize = 2 if opt & 0x100 else 4
    for (a, b), e in zip(chunks, exp):
        buf = np.zeros(e.size * itemsize + 64, np.uint8)
        nout = C.c_size_t(0)
        chunk = words[a:b].copy()
        assert L.vit_stream_push(h, chunk.ctypes.data, (b - a) * spw, buf.ctypes.data, buf.size, C.byref(nout)) == 0, L.vit_last_error()
        assert nout.value == e.size * itemsize
        assert np.array_equal(buf[:nout.value].view(e.dtype), e), (a, b)
    assert L.vit_stream_pending(h) == pending
    L.vit_destroy(h)


def test_chunk_pipeline_host_code_on_the_cpu(sim, O):
    """vit_run's segment-range chunk pipeline (pinned buffers where gates cannot be used): 8.4 MB of fp32 input = two
    chunks.  The stand-in runtime runs chunk 0's kernel the moment chunk 0's bytes are copied -- chunk 1's upload has not
    even been enqueued and its part of the device buffer is still garbage -- so the byte range vit_run uploads per chunk
    must cover everything that chunk's segments read."""
    L = sim
    opt, n_bits = 0x004, (W * 512 + 17) * 32 + 64 + 3
    h = _make(L, opt)
    launches = L.vit_launch_count(h)
    got, exp = _run(L, O, h, opt, n_bits, 41, CHUNKED, True, True)
    assert np.array_equal(got, exp)
    assert L.vit_launch_count(h) - launches == 2
    L.vit_destroy(h)


def test_host_paths_fuzz(sim, O):
    """seeded fuzz of vit_run over stream shapes, option values, upload modes and buffer kinds, one handle per option value
    kept across calls (device / staging buffers grow and are reused); 12 s"""
    import time
    L = sim
    rng = np.random.default_rng(20261019)
    opts = [0x011, 0x000, 0x112, 0x101, 0x002, 0x021, 0x2000]
    handles = {o: _make(L, o) for o in opts}
    L.sim_reset_counters()
    t0, cases = time.perf_counter(), 0
    while time.perf_counter() - t0 < 12.0:
        opt = opts[int(rng.integers(len(opts)))]
        bpp = 16 if opt & 0x100 else 32
        packs_per_seg = int(rng.integers(20, 34)) * (32 // bpp)
        n_bits = (W * packs_per_seg + int(rng.integers(0, W))) * bpp + 64 + int(rng.integers(0, bpp))
        mode = GATED if rng.random() < 0.8 else SEQUENTIAL
        pi, po = bool(rng.integers(2)), bool(rng.integers(2))
        got, exp = _run(L, O, handles[opt], opt, n_bits, int(rng.integers(1, 1 << 30)), mode, pi, po, sigma=float(rng.choice([0.3, 0.8, 1.5])))
        assert np.array_equal(got, exp), (hex(opt), n_bits, mode, pi, po)
        cases += 1
    L.sim_check_canaries()
    assert L.sim_violations() == 0 and cases >= 5, cases
    for h in handles.values():
        assert L.vit_upload_mode_in_effect(h) in (GATED, SEQUENTIAL)
        L.vit_destroy(h)


def test_argument_checks_of_the_host_code(sim, O):
    """the C ABI's argument validation (csrc/vit_api.cu launch / vit_run / vit_create / vit_stream_push), which returns codes
    and messages instead of faulting in a kernel"""
    L = sim
    L.vit_run_device_batch.restype = C.c_int
    L.vit_run_device_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.POINTER(C.c_float)]
    L.sim_device_alloc.restype, L.sim_device_alloc.argtypes = C.c_void_p, [C.c_size_t]
    L.sim_device_free.restype, L.sim_device_free.argtypes = None, [C.c_void_p]
    h = C.c_void_p()
    assert L.vit_create(C.byref(h), 0x013, 0, 0) == 1 and b"unsupported" in L.vit_last_error()        # b16 x s16
    assert L.vit_create(C.byref(h), 0x2021, 0, 0) == 1                                                   # no half2 DPX code
    assert L.vit_create(None, 0x011, 0, 0) == 3
    h = _make(L, 0x111)
    N = 2 * ((W * 4) * 16 + 64)
    d_in, d_out = L.sim_device_alloc(N // 2 + 64), L.sim_device_alloc(4096)
    assert L.vit_run_device_batch(h, d_in + 4, d_out, N, 1, 0, 0, None, None) == 3 and b"16-byte aligned" in L.vit_last_error()
    assert L.vit_run_device_batch(h, d_in, d_out, N, 2, 100, 512, None, None) == 3 and b"16-byte aligned" in L.vit_last_error()
    assert L.vit_run_device_batch(h, d_in, d_out + 1, N, 1, 0, 0, None, None) == 3 and b"2-byte packs" in L.vit_last_error()
    assert L.vit_run_device_batch(h, d_in, d_out, N, 70000, 0, 0, None, None) == 3 and b"65535" in L.vit_last_error()
    assert L.vit_run_device_batch(h, None, d_out, N, 1, 0, 0, None, None) == 3
    assert L.vit_run_device_batch(h, d_in, d_out, 100, 1, 0, 0, None, None) == 0                          # shorter than the window: nothing to decode
    assert L.vit_run(h, None, None, N, None) == 3
    assert L.vit_set_upload_mode(h, 7) == 3 and b"unknown upload mode" in L.vit_last_error()
    out = np.zeros(8, np.uint8)
    nout = C.c_size_t(99)
    words = np.zeros(3, np.uint32)
    assert L.vit_stream_push(h, words.ctypes.data, 7, out.ctypes.data, 8, C.byref(nout)) == 3 and b"whole 32-bit channel packs" in L.vit_last_error()
    big = np.zeros(4000, np.uint32)
    assert L.vit_stream_push(h, big.ctypes.data, 4000 * 8, out.ctypes.data, 8, C.byref(nout)) == 3 and b"output buffer holds" in L.vit_last_error()
    assert nout.value == 0
    L.sim_device_free(d_in)
    L.sim_device_free(d_out)
    L.vit_destroy(h)
    L.vit_destroy(None)


_LOST_GATE = r'''
import ctypes as C, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
from oracle import oracle as O
L = C.CDLL(%(so)r)
vp, sz = C.c_void_p, C.c_size_t
L.vit_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, sz]
L.vit_run.argtypes = [vp, vp, vp, sz, C.POINTER(C.c_float)]
L.vit_set_segments.argtypes = [vp, C.c_uint]; L.vit_set_upload_mode.argtypes = [vp, C.c_int]
L.vit_upload_mode_in_effect.argtypes = [vp]; L.vit_launch_count.argtypes = [vp]; L.vit_launch_count.restype = C.c_ulonglong
L.vit_last_error.restype = C.c_char_p; L.vit_destroy.argtypes = [vp]
L.sim_violations.restype = C.c_ulonglong
opt, W = 0x004, 64
n_bits = (W * 130 + 5) * 32 + 64          # fp32: 2.1 MB of input, 44 super-steps -> the automatic mode takes the time-sliced upload
bits, packed, N = O.make_channel_det(n_bits, 4, seed=3, sigma=0.8)
O.set_segments(W); exp = O.decode(opt, packed, N); O.set_segments(0)
raw = np.ascontiguousarray(packed).view(np.uint8)[:O.input_size(opt, N)].copy()
h = vp(); assert L.vit_create(C.byref(h), opt, 0, 0) == 0; L.vit_set_segments(h, W)
assert L.vit_upload_mode_in_effect(h) == 3                      # AUTO resolves to the time-sliced upload
out = np.zeros(O.output_size(opt, N), np.uint8)
assert L.vit_run(h, raw.ctypes.data, out.ctypes.data, N, None) == 0, L.vit_last_error()
assert np.array_equal(out.view(exp.dtype), exp)                # the last gate never opened: decoded again through the fallback
assert L.vit_launch_count(h) == 2 and L.vit_upload_mode_in_effect(h) == 2       # ... and the handle stops using gates
out[:] = 0
assert L.vit_run(h, raw.ctypes.data, out.ctypes.data, N, None) == 0 and np.array_equal(out.view(exp.dtype), exp)
assert L.vit_launch_count(h) == 3
L.vit_set_upload_mode(h, 3)                                      # insisting on the time-sliced upload reports the time-out instead
assert L.vit_run(h, raw.ctypes.data, out.ctypes.data, N, None) == 2 and b"upload gate timed out" in L.vit_last_error()
L.vit_destroy(h)
assert L.sim_violations() == 0
print("OK")
'''


def test_lost_gate_falls_back_and_disables_gates(sim_lib_path):
    """VIT_TEST_LOSE_GATE (the last gate's flag copy is skipped): the kernel's warps give up at the closed gate, vit_run sees
    it, decodes again without gates and stops using them on that handle; with the time-sliced mode forced it is an error."""
    import sys
    code = _LOST_GATE % dict(root=ROOT, tests=os.path.join(ROOT, "tests"), so=sim_lib_path)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, VIT_TEST_LOSE_GATE="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
