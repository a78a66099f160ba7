"""GPU test of a library build for other generator polynomials (csrc/vit_code.h):
libvitb200_p117_155.so (built by __graft_entry__.build()) decodes the K=7 (0117, 0155) code.  Its sm_100a kernels are
compared word for word with the golden model set to the same polynomials -- noisy and all-tie inputs, every core --, its
device source with the CPU twin, and the encode -> decode round trip with the message.  The CPU suite checks the same
kernel source in the host emulator (tests/test_code_parameters.py)."""
import os

import numpy as np
import pytest

from vit_testlib import ALT_LIB, ALT_POLYS, load_pkg_variant

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def VA():
    if not os.path.exists(ALT_LIB):
        pytest.skip("variant library not built (python -c 'import __graft_entry__ as g; g.build()')")
    mod = load_pkg_variant(ALT_LIB, "gpu_accelerated_viterbi_decoder_b200_p117_155")
    assert mod.code_parameters() == (7,) + ALT_POLYS
    return mod


@pytest.fixture
def alt_oracle(O):
    O.set_polynomials(*ALT_POLYS)
    yield O
    O.set_polynomials(0, 0)


@pytest.mark.parametrize("opt", [0x011, 0x000, 0x121, 0x112, 0x004, 0x022, 0x2001])
def test_variant_library_matches_golden_model(VA, alt_oracle, opt):
    O = alt_oracle
    dec = VA.ViterbiCUDA(opt)
    for n, kw in ((6400 * 32 * 3 + 64 + 32 * 77, dict(seed=31, sigma=0.9)),       # noisy, ragged
                  (6400 * 32 * 2 + 64, dict(seed=32, zero=True)),                 # every compare a tie
                  (6400 * 32 * 12 + 64, dict(seed=33, sigma=0.3))):               # several super-steps per segment
        bits, packed, N = O.make_channel_det(n, opt & 0xF, **kw)
        out = dec.run(packed, N, want_kernel_time=True)[0]
        assert np.array_equal(out, O.decode(opt, packed, N)), (hex(opt), n)
    if (opt & 0xF) != 0:
        assert O.count_errors(opt, out, dec.getMessageLen(N), bits) == 0          # sigma 0.3: the message comes out
    dec.close()


def test_variant_library_device_source_and_round_trip(VA, alt_oracle):
    """the variant's device source encodes with ITS polynomials (bit-identical to the CPU twin under the same
    polynomials) and its decoder returns the message"""
    import ctypes as C
    O = alt_oracle
    L = VA.lib()
    opt, n = 0x011, 6400 * 32 * 8 + 64
    N = 2 * n
    dec = VA.ViterbiCUDA(opt)
    in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
    ptrs = [C.c_void_p() for _ in range(3)]
    for p, nbytes in zip(ptrs, (in_bytes + 64, n, out_bytes)):
        assert L.vit_dev_alloc(C.byref(p), nbytes) == 0
    d_in, d_bits, d_out = [p.value for p in ptrs]
    try:
        VA.synth_device(opt & 0xF, n, d_in, d_bits, seed=7, amp=0, sigma=0.4)
        assert L.vit_dev_sync() == 0
        bits, packed, _ = O.make_channel_det(n, opt & 0xF, seed=7, sigma=0.4, bits_source="hash")
        assert np.array_equal(VA.dev_to_host(d_bits, n), bits)
        assert np.array_equal(VA.dev_to_host(d_in, in_bytes), packed.view(np.uint8)[:in_bytes])
        dec.run_device(d_in, d_out, N)
        assert L.vit_dev_sync() == 0
        assert VA.count_errors_device(opt, d_out, d_bits, dec.getMessageLen(N)) == 0
        assert np.array_equal(VA.dev_to_host(d_out, out_bytes).view(np.uint32), O.decode(opt, packed, N))
    finally:
        for p in ptrs:
            L.vit_dev_free(p)
        dec.close()
