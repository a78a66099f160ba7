"""Generate tests/golden/ref_vectors.npz by running the UNMODIFIED reference decoder
(oracle/_ref/libvitref.so, built from /root/reference by oracle/Makefile) on a B200.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/ref_vectors.npz
then copy the file to tests/golden/ref_vectors.npz and commit it.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
from vit_golden_cases import CASES  # noqa: E402


def main(dst):
    if O.ref_lib() is None or O.ref_lib().ref_device_count() < 1:
        raise SystemExit("needs oracle/_ref/libvitref.so and a GPU")
    out = {}
    for name, opt, n, seed, sigma, zero in CASES:
        bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=seed, sigma=sigma, zero=zero)
        ref, ms = O.ref_decode(opt, packed, N)
        ref2, _ = O.ref_decode(opt, packed, N)
        out[name + "/out"] = ref
        out[name + "/stable"] = np.array([int(np.array_equal(ref, ref2))])
        out[name + "/sha"] = np.frombuffer(hashlib.sha256(packed.tobytes()).digest(), np.uint8)
        orc = O.decode(opt, packed, N)
        orc_ov = O.decode(opt, packed, N, flags=O.FLAG_REF_OVERRUN)
        ov = O.overrun_words(opt, N)
        mask = np.ones(ref.size, bool)
        mask[ov.astype(np.int64)] = False
        print("%-18s opt=%#06x n=%d words=%d  ref==oracle(owned words): %s  ref==oracle(overrun emu): %s  "
              "overrun words=%d differing there=%d  ref stable=%s  BER=%.3e"
              % (name, opt, n, ref.size, np.array_equal(ref[mask], orc[mask]), np.array_equal(ref, orc_ov),
                 ov.size, int(np.count_nonzero(ref[~mask] != orc[~mask])), np.array_equal(ref, ref2),
                 O.count_errors(opt, ref, O.message_len(opt, N), bits) / max(1, O.message_len(opt, N))))
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_vectors.npz"))
