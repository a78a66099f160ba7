"""CPU tests of the product kernel SOURCE: vit_kernel.cuh compiled for the host and run as 32
lock-step fibers per warp (tests/emu), compared word for word with the golden model.  This checks
the trellis mapping, operand tables, survivor fields, ring and traceback logic without a GPU; the
real sm_100a binary is checked by the -m gpu tests."""
import numpy as np
import pytest

from vit_testlib import ALL_OPTS



def emu_decode(emu, O, opt, packed, N, W):
    packed = np.ascontiguousarray(packed)
    buf = np.zeros((packed.nbytes + 31) // 16 * 16 + 16, np.uint8)
    off = (-buf.ctypes.data) % 16
    buf[off:off + packed.nbytes] = packed.view(np.uint8)
    nw = O.output_size(opt, N) // np.dtype(O.out_dtype(opt)).itemsize
    out = np.full(nw + 4, 0xDEAD, O.out_dtype(opt))
    assert emu.vit_emu_decode(opt, buf[off:].ctypes.data, out.ctypes.data, N, W, 1, 0, 0) == 0
    assert np.all(out[nw:] == 0xDEAD), "wrote past the end of the output"
    return out[:nw]


def run_case(emu, O, opt, n, W, **kw):
    zero = kw.pop("zero", False)
    bits, packed, N = O.make_channel_det(n, opt & 0xF, zero=zero, **kw)
    O.set_segments(W)
    try:
        ref = O.decode(opt, packed, N)
    finally:
        O.set_segments(0)
    got = emu_decode(emu, O, opt, packed, N, W)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("opt", ALL_OPTS)
def test_kernel_source_matches_oracle(emu, O, opt):
    run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)        # ragged segments, noisy
    run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)             # every compare a tie
    run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)      # fewer packs than segments


@pytest.mark.parametrize("opt", [0x011, 0x000, 0x121, 0x112, 0x004, 0x023])
def test_kernel_source_small_table_variant(emu, O, opt):
    """The TBL=32 build of the kernel (operand table rebuilt before every 32-stage slide; used for launches whose
    warps do not all fit with the 96-stage table) decodes the same words."""
    emu.vit_emu_set_table(32)
    try:
        run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
        run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)
        run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)
        amp = {0: 64, 1: 7, 2: 127, 3: 32767, 4: 128}[opt & 0xF]
        run_case(emu, O, opt, 12000 + 64, 4, seed=11, sigma=1.0, amp=amp)
    finally:
        emu.vit_emu_set_table(96)


@pytest.mark.parametrize("opt", [0x011, 0x000, 0x121, 0x112, 0x004, 0x023, 0x002])
def test_kernel_source_sixteen_lane_geometry(emu, O, opt):
    emu.vit_emu_set_lanes(16)
    try:
        for tbl in (96, 32):
            emu.vit_emu_set_table(tbl)
            run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
            run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)
            run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)
    finally:
        emu.vit_emu_set_lanes(8)
        emu.vit_emu_set_table(96)


@pytest.mark.parametrize("tbl", [96, 32])
@pytest.mark.parametrize("opt", ALL_OPTS)
def test_kernel_source_four_lane_geometry(emu, O, opt, tbl):
    """The l4 instantiation of the kernel (4 lanes per segment, 8 segments per warp, 16 states per lane, three half
    exchanges per 6 stages; used for multi-stream launches) decodes the same words."""
    emu.vit_emu_set_lanes(4)
    emu.vit_emu_set_table(tbl)
    try:
        run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
        run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)
        run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)
        if tbl == 32:
            amp = {0: 64, 1: 7, 2: 127, 3: 32767, 4: 128}[opt & 0xF]
            run_case(emu, O, opt, 12000 + 64, 4, seed=11, sigma=1.0, amp=amp)
    finally:
        emu.vit_emu_set_lanes(8)
        emu.vit_emu_set_table(96)


@pytest.mark.parametrize("opt", [0x012, 0x112, 0x011, 0x024, 0x022, 0x003])
def test_kernel_source_long_segments(emu, O, opt):
    """Several 96-stage super-steps per segment, saturated symbols: exercises the metric range
    management (offset int16 operands, normalisation) and the double-buffered staging."""
    amp = {0: 64, 1: 7, 2: 127, 3: 32767, 4: 128}[opt & 0xF]
    run_case(emu, O, opt, 12000 + 64, 4, seed=11, sigma=1.0, amp=amp)
    run_case(emu, O, opt, 12000 + 64, 4, seed=12, sigma=0.2, amp=amp)


@pytest.mark.parametrize("opt", [0x2000, 0x2100, 0x2001, 0x2002, 0x2003, 0x2004, 0x2011, 0x2112])
def test_kernel_source_dpx_tie_rule(emu, O, opt):
    """CompMode value 2 (extension): the tie rule of the reference's dormant DPX code paths (viterbiACS.cuh:123-134) --
    int32 core: the partner wins ties in every phase; int16x2 core: identical to REG.  Kernel source vs oracle table."""
    run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
    run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)             # every compare a tie: output = the tie table
    emu.vit_emu_set_table(32)
    try:
        run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)
    finally:
        emu.vit_emu_set_table(96)


def test_dpx_tie_rule_differs_from_reg_only_for_int32(O):
    bits, packed, N = O.make_channel_det(1500 + 64, O.SOFT4, seed=1, zero=True)             # all-zero soft symbols: every compare ties
    assert not np.array_equal(O.decode(0x2001, packed, N), O.decode(0x0001, packed, N))      # int32: phase-0 ties differ
    assert np.array_equal(O.decode(0x2011, packed, N), O.decode(0x0011, packed, N))          # int16x2: same table
    with pytest.raises(ValueError):
        O.decode(0x2021, packed, N)                                                          # no half2 DPX code


@pytest.mark.parametrize("lanes", [8, 4, 16])
@pytest.mark.parametrize("opt", [0x011, 0x112, 0x100, 0x004, 0x121])
def test_kernel_source_staged_output_stores(emu, O, opt, lanes):
    """KParams::stage_out (output buffer in a peer GPU's memory): packs are staged in shared memory and stored 8 slides at
    a time; same words, for both pack widths (16-bit packs: aligned and unaligned segment starts, odd tails)."""
    emu.vit_emu_set_stage_out(1)
    emu.vit_emu_set_lanes(lanes)
    try:
        run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
        run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)
        run_case(emu, O, opt, 12000 + 64 + 16, 5, seed=11, sigma=1.0)          # > 8 slides per segment, ragged
        run_case(emu, O, opt, 9000 + 64, 7, seed=12, sigma=0.3)
    finally:
        emu.vit_emu_set_stage_out(0)
        emu.vit_emu_set_lanes(8)


@pytest.mark.parametrize("opt", [0x012, 0x111, 0x004])
def test_kernel_source_multi_stream_strides(emu, O, opt):
    """Several independent streams per launch (grid.y = stream; BASELINE configs[4] shape): stream s is read at
    in + s * in_stride and decoded to out + s * out_stride, nothing is written between or after the streams."""
    W, ns, n = 9, 3, 2000 + 64 + 5
    _multi_stream_case(emu, O, opt, W, ns, n, (96, 32))


def test_kernel_source_multi_stream_strides_one_lane_geometry(emu, O):
    emu.vit_emu_set_lanes(1)
    try:
        _multi_stream_case(emu, O, 0x012, 41, 3, 41 * 32 * 9 + 64 + 5, (96,))
    finally:
        emu.vit_emu_set_lanes(8)


def _multi_stream_case(emu, O, opt, W, ns, n, tbls):
    streams = [O.make_channel_det(n, opt & 0xF, seed=40 + s, sigma=0.8) for s in range(ns)]
    N = streams[0][2]
    in_bytes, out_bytes = O.input_size(opt, N), O.output_size(opt, N)
    in_stride, out_stride = (in_bytes + 4 + 255) // 256 * 256, (out_bytes + 255) // 256 * 256
    buf = np.zeros(ns * in_stride + 64, np.uint8)
    off = (-buf.ctypes.data) % 16
    for s, (_, packed, _) in enumerate(streams):
        buf[off + s * in_stride: off + s * in_stride + in_bytes] = packed.view(np.uint8)[:in_bytes]
    out = np.full(ns * out_stride + 16, 0xEE, np.uint8)
    O.set_segments(W)
    try:
        refs = [O.decode(opt, packed, N) for _, packed, _ in streams]
    finally:
        O.set_segments(0)
    for tbl in tbls:
        emu.vit_emu_set_table(tbl)
        try:
            out[:] = 0xEE
            assert emu.vit_emu_decode(opt, buf[off:].ctypes.data, out.ctypes.data, N, W, ns, in_stride, out_stride) == 0
        finally:
            emu.vit_emu_set_table(96)
        for s in range(ns):
            got = out[s * out_stride: s * out_stride + out_bytes].view(O.out_dtype(opt))
            assert np.array_equal(got, refs[s]), (tbl, s)
            assert np.all(out[s * out_stride + out_bytes: (s + 1) * out_stride] == 0xEE), (tbl, s)
        assert np.all(out[ns * out_stride:] == 0xEE)


@pytest.mark.parametrize("opt", ALL_OPTS + [0x2001, 0x2104, 0x2011])
def test_kernel_source_one_lane_per_segment_geometry(emu, O, opt):
    """vitk::l1 (csrc/vit_kernel_l1.inc): one lane per segment, 32 segments per warp, all 64 states of a segment in one
    thread's registers -- no exchanges, no operand table, a one-slot ring, ping-pong register sets.  The many-stream
    geometry identified as the next step; same output words as the golden model for every option value:
    noisy, all-tie, short, long saturated segments, several warps with a partly filled last one, several streams."""
    emu.vit_emu_set_lanes(1)
    try:
        run_case(emu, O, opt, 3000 + 64 + 7, 12, seed=5, sigma=0.9)
        run_case(emu, O, opt, 1500 + 64, 5, seed=1, zero=True)
        run_case(emu, O, opt, 64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)
        amp = {0: 64, 1: 7, 2: 127, 3: 32767, 4: 128}[opt & 0xF]
        run_case(emu, O, opt, 12000 + 64, 4, seed=11, sigma=1.0, amp=amp)
        run_case(emu, O, opt, 40 * 32 * 70 + 64 + 5, 70, seed=3, sigma=0.8)
    finally:
        emu.vit_emu_set_lanes(8)


def _gate_case(emu, O, opt, n, W, sup, open_count):
    """decode with upload gates `sup` of which the first open_count are open; returns (out, reference, reads, gate error)"""
    import ctypes as C
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=17, sigma=0.8)
    O.set_segments(W)
    try:
        ref = O.decode(opt, packed, N)
    finally:
        O.set_segments(0)
    arr = (C.c_uint * len(sup))(*sup)
    emu.vit_emu_set_gates(len(sup), arr, open_count)
    emu.vit_emu_record_reads(1 << 20)
    try:
        packed = np.ascontiguousarray(packed)
        buf = np.zeros((packed.nbytes + 31) // 16 * 16 + 16, np.uint8)
        off = (-buf.ctypes.data) % 16
        buf[off:off + packed.nbytes] = packed.view(np.uint8)
        nw = O.output_size(opt, N) // np.dtype(O.out_dtype(opt)).itemsize
        out = np.full(nw + 4, 0xDEAD, O.out_dtype(opt))
        assert emu.vit_emu_decode(opt, buf[off:].ctypes.data, out.ctypes.data, N, W, 1, 0, 0) == 0
        rec = np.zeros(3 << 20, np.uint64)
        cnt = emu.vit_emu_fetch_reads(rec.ctypes.data, 1 << 20)
        assert cnt <= 1 << 20
        return out[:nw], ref, rec[:3 * cnt].reshape(-1, 3).astype(np.int64), emu.vit_emu_gate_error(), N
    finally:
        emu.vit_emu_set_gates(0, (C.c_uint * 1)(0), 0)
        emu.vit_emu_record_reads(0)


@pytest.mark.parametrize("opt", [0x011, 0x000, 0x112, 0x004])
def test_kernel_source_respects_upload_gates(emu, O, opt):
    """The kernel side of vit_run's time-sliced upload (the host side: tests/test_upload_plan_model.py).  A warp passes gate g
    before it requests the channel words of any super-step x >= gate_super[g].  With every gate open the output is the
    golden model's; with gates 1.. closed the warps stop (and report it), and NO input byte at or beyond the column the
    host uploads with the closed gate's block has been requested: per segment, every requested byte lies in columns
    [-15, gate_super[k] * b96 + 48) of its row -- exactly what blocks 0..k-1 of run_gated have delivered."""
    import ctypes as C
    emu.vit_emu_record_reads.argtypes = [C.c_size_t]
    emu.vit_emu_fetch_reads.restype, emu.vit_emu_fetch_reads.argtypes = C.c_size_t, [C.c_void_p, C.c_size_t]
    emu.vit_emu_set_gates.argtypes = [C.c_uint, C.c_void_p, C.c_uint]
    emu.vit_emu_gate_error.restype = C.c_uint
    it, bpp = opt & 0xF, (16 if opt & 0x100 else 32)
    b96 = {0: 24, 1: 96, 2: 192, 3: 384, 4: 768}[it]
    W, packs = 12, 12 * 40 + 5                                    # 40-41 packs per segment: 14-15 super-steps
    n = packs * bpp + 64 + 3
    nsuper = (64 + 32 * ((41 * bpp + 31) // 32) + 95) // 96
    sup = [0, nsuper // 2, nsuper * 7 // 8]
    out, ref, reads, err, N = _gate_case(emu, O, opt, n, W, sup, 3)
    assert err == 0 and np.array_equal(out, ref)                  # all gates open: nothing changes
    P = O.message_len(opt, N) // bpp
    q, r = divmod(P, W)
    pack_bytes = bpp * b96 // 96
    rows = np.array([(q * w + min(w, r)) * pack_bytes for w in range(W)], np.int64)
    for open_count in (1, 2):
        out, ref, reads, err, N = _gate_case(emu, O, opt, n, W, sup, open_count)
        assert err == 1                                           # the warps gave up at the closed gate and said so
        limit = sup[open_count] * b96 + 48
        col = reads[:, 1] - rows[reads[:, 0]]
        assert len(reads) > 0 and col.min() >= -15
        assert (col + reads[:, 2]).max() <= limit, (open_count, int((col + reads[:, 2]).max()), limit)
        assert (col + reads[:, 2]).max() > (sup[open_count] - 1) * b96     # ... and they did get as far as the gate allows
        written = out != 0xDEAD
        assert np.array_equal(out[written], ref[written])        # whatever was emitted before the stop is correct


@pytest.mark.parametrize("opt", [0x011, 0x112, 0x000])
def test_kernel_source_segment_range_launches(emu, O, opt):
    """vit_run's chunk pipeline decodes a stream with several launches over segment ranges [seg_first, seg_limit) (cut at
    multiples of 8 segments).  Each launch must write exactly the packs its segments own, and together they must give the
    one-launch output."""
    import ctypes as C
    emu.vit_emu_set_segment_range.argtypes = [C.c_uint, C.c_uint]
    bpp = 16 if opt & 0x100 else 32
    W, n = 40, (40 * 9 + 13) * bpp + 64 + 5              # the first 13 segments are one pack longer
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=23, sigma=0.8)
    O.set_segments(W)
    try:
        ref = O.decode(opt, packed, N)
    finally:
        O.set_segments(0)
    P = O.message_len(opt, N) // bpp
    q, r = divmod(P, W)
    start = lambda w: q * w + min(w, r)
    packed = np.ascontiguousarray(packed)
    buf = np.zeros((packed.nbytes + 31) // 16 * 16 + 16, np.uint8)
    off = (-buf.ctypes.data) % 16
    buf[off:off + packed.nbytes] = packed.view(np.uint8)
    total = np.full(ref.size, 0xDEAD, ref.dtype)
    try:
        for a, b in ((0, 8), (8, 24), (24, 40)):
            out = np.full(ref.size + 4, 0xDEAD, ref.dtype)
            emu.vit_emu_set_segment_range(a, b)
            assert emu.vit_emu_decode(opt, buf[off:].ctypes.data, out.ctypes.data, N, W, 1, 0, 0) == 0
            lo, hi = start(a), (start(b) if b < W else P)
            assert np.array_equal(out[lo:hi], ref[lo:hi]), (a, b)
            assert np.all(out[:lo] == 0xDEAD) and np.all(out[hi:] == 0xDEAD), (a, b)     # nothing outside its own range
            total[lo:hi] = out[lo:hi]
    finally:
        emu.vit_emu_set_segment_range(0, 0)
    assert np.array_equal(total, ref)
