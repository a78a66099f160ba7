"""Generate tests/golden/ref_vectors_dpx.npz: outputs of the reference's DORMANT DPX code paths, made live by
oracle/_ref/libvitref_dpx.so (the unmodified reference sources compiled with compMode forwarded to forwardACS,
oracle/ref_dpx_shim.cu), on a B200.  They pin the oracle's VO_DPX_TIES tie table without a GPU (tests/test_oracle.py).

Run on the GPU box:   python tests/golden/make_golden_dpx.py gpurun_out/ref_vectors_dpx.npz
then copy the file to tests/golden/ref_vectors_dpx.npz and commit it.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
from vit_golden_cases import DPX_CASES  # noqa: E402


def main(dst):
    if O.ref_dpx_lib() is None:
        raise SystemExit("needs oracle/_ref/libvitref_dpx.so and a GPU")
    out = {}
    for name, opt, n, seed, sigma, zero in DPX_CASES:
        bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=seed, sigma=sigma, zero=zero)
        ref, _ = O.ref_decode_dpx((opt & 0xFFF) | O.DPX, packed, N)        # the reference's own option value for DPX
        out[name + "/out"] = ref
        out[name + "/sha"] = np.frombuffer(hashlib.sha256(packed.tobytes()).digest(), np.uint8)
        orc = O.decode((opt & 0xFFF) | O.DPX_TIES, packed, N)
        reg = O.decode(opt & 0xFFF, packed, N)
        ov = O.overrun_words(opt & 0xFFF, N).astype(np.int64)
        m = np.ones(ref.size, bool)
        m[ov] = False
        print("%-16s opt=%#06x n=%d  dpx ref == oracle(DPX_TIES) on owned words: %s   == oracle(REG): %s"
              % (name, opt, n, np.array_equal(ref[m], orc[m]), np.array_equal(ref[m], reg[m])))
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_vectors_dpx.npz"))
