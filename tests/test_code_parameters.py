"""Code parameters (reference viterbi.h:61-63): K = 7 is fixed, the generator polynomials are
compile-time parameters of the library (csrc/vit_code.h).  CPU tests: the golden model with other polynomials decodes what
its encoder produced, and the product kernel SOURCE compiled for those polynomials (host emulator, tests/emu) equals it word
for word -- noisy inputs, all-tie inputs, both operand-table builds.  The reference itself hard-codes 0171 / 0133, so for
other polynomials the golden model is the only checker; what pins it is that the same model with the default polynomials
reproduces the reference's vectors (tests/test_oracle.py)."""
import ctypes as C

import numpy as np
import pytest

from test_emu_kernel import emu_decode

# (polyn1, polyn2): the bit-reversed standard code; the standard pair in the other order; another pair tapping both ends
from vit_testlib import ALT_EMU_PAIRS as PAIRS  # noqa: E402
OPTS = [0x011, 0x000, 0x121, 0x112, 0x004, 0x022, 0x2001]


@pytest.fixture
def polys(O):
    """sets the golden model's polynomials for one test and restores the reference's afterwards"""
    yield O.set_polynomials
    O.set_polynomials(0, 0)


def test_defaults_are_the_reference_code(V, O, emu):
    assert O.get_polynomials() == (0o171, 0o133)
    p1, p2 = C.c_int(0), C.c_int(0)
    emu.vit_emu_polynomials(C.byref(p1), C.byref(p2))
    assert (p1.value, p2.value) == (0o171, 0o133)
    assert V.code_parameters() == (7, 0o171, 0o133)          # the built product library (no GPU needed for this call)
    assert (V.constLen, V.polyn1, V.polyn2) == (7, 0o171, 0o133)


@pytest.mark.parametrize("bad", [(0o170, 0o133), (0o171, 0o033), (0o371, 0o133)])
def test_polynomials_must_tap_both_ends(O, bad):
    with pytest.raises(ValueError):
        O.set_polynomials(*bad)
    assert O.get_polynomials() == (0o171, 0o133)


@pytest.mark.parametrize("pair", PAIRS + [(0o165, 0o127)])
@pytest.mark.parametrize("opt", [0x000, 0x011, 0x122, 0x104])
def test_golden_model_other_polynomials_round_trip(O, polys, pair, opt):
    """noiseless: decode(encode(bits)) returns the message (out bit j = message bit j + 26); the default decoder does
    not decode this code (so the polynomials really are in effect)."""
    polys(*pair)
    n = 6400 * 32 + 64 + 32 * 5
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=21, sigma=0.0)
    out = O.decode(opt, packed, N)
    M = O.message_len(opt, N)
    assert O.count_errors(opt, out, M, bits) == 0
    polys(0, 0)
    assert O.count_errors(opt, O.decode(opt, packed, N), M, bits) > M // 8


@pytest.mark.parametrize("pair", PAIRS)
@pytest.mark.parametrize("opt", OPTS)
def test_kernel_source_other_polynomials(emu_for_polynomials, O, polys, pair, opt):
    emu = emu_for_polynomials(*pair)
    p1, p2 = C.c_int(0), C.c_int(0)
    emu.vit_emu_polynomials(C.byref(p1), C.byref(p2))
    assert (p1.value, p2.value) == pair
    polys(*pair)

    def case(n, W, **kw):
        zero = kw.pop("zero", False)
        bits, packed, N = O.make_channel_det(n, opt & 0xF, zero=zero, **kw)
        O.set_segments(W)
        try:
            ref = O.decode(opt, packed, N)
        finally:
            O.set_segments(0)
        assert np.array_equal(emu_decode(emu, O, opt, packed, N, W), ref)
        return bits, ref, N
    for tbl in (96, 32):
        emu.vit_emu_set_table(tbl)
        try:
            case(3000 + 64 + 7, 12, seed=5, sigma=0.9)                # ragged segments, noisy
            case(1500 + 64, 5, seed=1, zero=True)                     # every compare a tie
            case(64 + 32 * 3 + 16, 8, seed=9, sigma=0.5)              # fewer packs than segments
            bits, ref, N = case(12000 + 64, 4, seed=11, sigma=0.3)    # several super-steps per segment ...
            if (opt & 0xF) != 0:
                M = O.message_len(opt, N)
                assert O.count_errors(opt, ref, M, bits) == 0         # ... and it is the message that comes out
        finally:
            emu.vit_emu_set_table(96)


def test_variant_library_reports_its_polynomials():
    """libvitb200_p117_155.so (built by __graft_entry__.build()): same C ABI, other code; loads without a GPU."""
    import os
    import subprocess
    from vit_testlib import ALT_LIB, ALT_POLYS, PKG_DIR, load_pkg_variant
    if not os.path.exists(ALT_LIB):
        pytest.skip("variant library not built")
    VA = load_pkg_variant(ALT_LIB, "gpu_accelerated_viterbi_decoder_b200_p117_155")
    assert VA.LIB_PATH == ALT_LIB
    assert VA.code_parameters() == (7,) + ALT_POLYS
    exported = lambda so: {l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", so], text=True).splitlines() if " T " in l}
    assert exported(ALT_LIB) == exported(os.path.join(PKG_DIR, "libvitb200.so"))
