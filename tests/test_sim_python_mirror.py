"""CPU run of the PYTHON mirror of the host class (gpu-accelerated-viterbi-decoder_b200/__init__.py: ViterbiCUDA, its buffer
checks, run / run_device / stream_push wrappers) over the host-path simulation (tests/sim/libvitsim.so selected with
VIT_B200_LIB: the library's own host code + the kernel source in the emulator + a stand-in CUDA runtime; test scaffolding).
The same class runs against the real library in the -m gpu suite."""
import numpy as np
import pytest

from vit_testlib import load_pkg_variant

W = 64


@pytest.fixture(scope="module")
def VS(sim_lib_path):
    mod = load_pkg_variant(sim_lib_path, "gpu_accelerated_viterbi_decoder_b200_sim")
    assert mod.LIB_PATH == sim_lib_path and mod.code_parameters() == (7, 0o171, 0o133)
    return mod


def _oracle(O, opt, packed, N):
    O.set_segments(W)
    try:
        return O.decode(opt, packed, N)
    finally:
        O.set_segments(0)


@pytest.mark.parametrize("opt", [0x011, 0x100, 0x004, 0x123, 0x2002])
def test_run_host_buffers(VS, O, opt):
    bits, packed, N = O.make_channel_det((W * 23 + 9) * 32 + 64 + 3, opt & 0xF, seed=61, sigma=0.8)
    dec = VS.ViterbiCUDA(opt, N)
    dec.set_segments(W)
    exp = _oracle(O, opt, packed, N)
    assert dec.getInputSize(N) == O.input_size(opt, N) and dec.getMessageLen(N) == O.message_len(opt, N)
    out, ms = dec.run(packed, N, want_kernel_time=True)                 # the reference's copy -> launch -> copy sequence
    assert np.array_equal(out, exp) and out.dtype == dec.decPack_t and ms > 0
    dec.set_upload_mode(VS.UPLOAD_GATED)
    out2 = np.zeros_like(exp)
    assert dec.run(packed, N, output_h=out2) is out2 and np.array_equal(out2, exp)      # caller's output buffer, overlapped path
    assert dec.launch_count() == 2 and dec.upload_mode_in_effect() == VS.UPLOAD_GATED
    dec.close()
    dec.close()                                                          # idempotent


def test_run_rejects_bad_buffers(VS, O):
    opt = 0x011
    bits, packed, N = O.make_channel_det((W * 6) * 32 + 64, 1, seed=62, sigma=0.5)
    dec = VS.ViterbiCUDA(opt)
    dec.set_segments(W)
    nw = dec.getOutputSize(N) // 4
    with pytest.raises(VS.ViterbiError):
        dec.run(packed[: packed.size // 2], N)                          # input shorter than getInputSize
    with pytest.raises(VS.ViterbiError):
        dec.run(packed, N, output_h=np.zeros(nw - 1, np.uint32))        # output too short
    with pytest.raises(VS.ViterbiError):
        dec.run(packed, N, output_h=np.zeros(2 * nw, np.uint32)[::2])   # not contiguous
    ro = np.zeros(nw, np.uint32)
    ro.setflags(write=False)
    with pytest.raises(VS.ViterbiError):
        dec.run(packed, N, output_h=ro)                                 # not writable
    with pytest.raises(VS.ViterbiError):
        dec.run(packed, N, output_h=[0] * nw)                           # not an array
    with pytest.raises(VS.ViterbiError):
        dec.set_upload_mode(9)
    with pytest.raises(VS.ViterbiError):
        VS.ViterbiCUDA(0x013)                                           # b16 x s16: rejected like the reference
    assert np.array_equal(dec.run(packed, N), _oracle(O, opt, packed, N))                # the handle is still good
    assert dec.run(packed[:40], 100).size == 0                          # fewer than 64 stages: nothing to decode
    dec.close()


def test_run_device_and_batches(VS, O):
    """device-resident entry points over the stand-in runtime's "device" memory: one stream, then three streams with strides"""
    import ctypes as C
    opt, ns = 0x112, 3
    L = VS.lib()
    streams = [O.make_channel_det((W * 11 + 5) * 16 + 64 + 7, 2, seed=70 + s, sigma=0.8) for s in range(ns)]
    N = streams[0][2]
    dec = VS.ViterbiCUDA(opt)
    dec.set_segments(W)
    in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
    in_stride, out_stride = (in_bytes + 255) // 256 * 256, (out_bytes + 255) // 256 * 256
    d_in, d_out = C.c_void_p(), C.c_void_p()
    assert L.vit_dev_alloc(C.byref(d_in), ns * in_stride) == 0 and L.vit_dev_alloc(C.byref(d_out), ns * out_stride) == 0
    for s, (_, packed, _) in enumerate(streams):
        raw = np.ascontiguousarray(packed).view(np.uint8)[:in_bytes].copy()
        assert L.vit_dev_copy_from_host(d_in.value + s * in_stride, raw.ctypes.data, in_bytes) == 0
    ms = dec.run_device(d_in.value, d_out.value, N, want_kernel_time=True)
    assert ms > 0 and np.array_equal(VS.dev_to_host(d_out.value, out_bytes).view(np.uint16), _oracle(O, opt, streams[0][1], N))
    dec.run_device(d_in.value, d_out.value, N, nstreams=ns, in_stride=in_stride, out_stride=out_stride)
    assert L.vit_dev_sync() == 0
    for s, (_, packed, _) in enumerate(streams):
        got = VS.dev_to_host(d_out.value + s * out_stride, out_bytes).view(np.uint16)
        assert np.array_equal(got, _oracle(O, opt, packed, N)), s
    with pytest.raises(VS.ViterbiError):
        dec.run_device(d_in.value + 2, d_out.value, N)
    L.vit_dev_free(d_in)
    L.vit_dev_free(d_out)
    dec.close()


def test_stream_push_wrapper(VS, O):
    opt = 0x011
    bits, packed, N = O.make_channel_det(30_000, 1, seed=63, sigma=0.6)
    words = np.ascontiguousarray(packed).view(np.uint32)
    cuts = [(0, 700), (700, 703), (703, 2600), (2600, words.size)]
    O.set_segments(W)
    try:
        exp, pending = O.decode_chunked(opt, packed, [(b - a) * 8 for a, b in cuts])
    finally:
        O.set_segments(0)
    dec = VS.ViterbiCUDA(opt)
    dec.set_segments(W)
    dec.stream_reset()
    total = 0
    for (a, b), e in zip(cuts, exp):
        got = dec.stream_push(words[a:b], (b - a) * 8)
        assert np.array_equal(got, e), (a, b)
        total += e.size * 32
    assert dec.stream_pending() == pending and dec.stream_bits() == total
    with pytest.raises(VS.ViterbiError):
        dec.stream_push(words[:2], 15)                                   # not whole 32-bit channel packs
    dec.stream_reset()
    assert dec.stream_pending() == 0 and dec.stream_bits() == 0
    dec.close()


def test_entry_points_that_need_a_gpu_say_so(VS):
    """the simulation has no device source / multi-GPU code: the mirror turns the library's error codes into exceptions"""
    with pytest.raises(VS.ViterbiError, match="not available in the host-path simulation"):
        VS.synth_device(1, 1000, 0)
    assert VS.options_valid(0x122) and not VS.options_valid(0x013) and VS.options_valid_ref(0x011) and not VS.options_valid_ref(0x122)
    assert VS.parse_options("s8", "f16", "b16") == 0x122


@pytest.mark.parametrize("opt", [0x012, 0x111, 0x024])
def test_one_lane_geometry_through_the_library(VS, O, opt):
    """vit_set_geometry(L1) (experimental, opt-in): gate-free launches of the packed cores run the one-lane-per-segment kernel
    -- same words -- while vit_run's time-sliced upload keeps the 8-lane kernel (it needs the upload gates), and so does a
    core the geometry is not built for."""
    import ctypes as C
    L = VS.lib()
    ns = 3
    bpp = 16 if opt & 0x100 else 32
    streams = [O.make_channel_det((W * 9 + 5) * bpp + 64 + 3, opt & 0xF, seed=80 + s, sigma=0.8) for s in range(ns)]
    N = streams[0][2]
    dec = VS.ViterbiCUDA(opt)
    dec.set_segments(W)
    assert dec.last_launch_geometry() == VS.GEOMETRY_L8
    dec.set_geometry(VS.GEOMETRY_L1)
    in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
    in_stride, out_stride = (in_bytes + 255) // 256 * 256, (out_bytes + 255) // 256 * 256
    d_in, d_out = C.c_void_p(), C.c_void_p()
    assert L.vit_dev_alloc(C.byref(d_in), ns * in_stride) == 0 and L.vit_dev_alloc(C.byref(d_out), ns * out_stride) == 0
    for s, (_, packed, _) in enumerate(streams):
        raw = np.ascontiguousarray(packed).view(np.uint8)[:in_bytes].copy()
        assert L.vit_dev_copy_from_host(d_in.value + s * in_stride, raw.ctypes.data, in_bytes) == 0
    dec.run_device(d_in.value, d_out.value, N, nstreams=ns, in_stride=in_stride, out_stride=out_stride)
    assert L.vit_dev_sync() == 0 and dec.last_launch_geometry() == VS.GEOMETRY_L1
    for s, (_, packed, _) in enumerate(streams):
        got = VS.dev_to_host(d_out.value + s * out_stride, out_bytes).view(dec.decPack_t)
        assert np.array_equal(got, _oracle(O, opt, packed, N)), s
    # host buffers, plain sequence: L1 as well; time-sliced upload: the 8-lane kernel
    big = O.make_channel_det((W * 24) * 32 + 64, opt & 0xF, seed=90, sigma=0.8)        # >= 8 super-steps per segment
    out = dec.run(big[1], big[2], want_kernel_time=True)[0]
    assert np.array_equal(out, _oracle(O, opt, big[1], big[2])) and dec.last_launch_geometry() == VS.GEOMETRY_L1
    dec.set_upload_mode(VS.UPLOAD_GATED)
    assert np.array_equal(dec.run(big[1], big[2]), out) and dec.last_launch_geometry() == VS.GEOMETRY_L8
    with pytest.raises(VS.ViterbiError):
        dec.set_geometry(5)
    L.vit_dev_free(d_in)
    L.vit_dev_free(d_out)
    dec.close()
    d32 = VS.ViterbiCUDA(0x001)                                       # int32 core: the geometry is not built for it
    d32.set_segments(W)
    d32.set_geometry(VS.GEOMETRY_L1)
    small = O.make_channel_det((W * 3) * 32 + 64, 1, seed=91, sigma=0.8)
    assert np.array_equal(d32.run(small[1], small[2], want_kernel_time=True)[0], _oracle(O, 0x001, small[1], small[2]))
    assert d32.last_launch_geometry() == VS.GEOMETRY_L8
    d32.close()
