"""Small decodes of every core / table build, for compute-sanitizer (memcheck, racecheck, initcheck).
usage: compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402

V = bench.load_pkg()
bad = 0
for opt in (0x011, 0x000, 0x121, 0x112, 0x004, 0x023, 0x002):
    n = 64 + 32 * 40 * 9 + 16 * 7
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=7, sigma=0.8)
    dec = V.ViterbiCUDA(opt, N)
    dec.set_segments(40)
    O.set_segments(40)
    try:
        exp = O.decode(opt, packed, N)
    finally:
        O.set_segments(0)
    out = dec.run(packed, N)
    ok = np.array_equal(out, exp)
    bad += not ok
    print("options %#05x table %s: %s" % (opt, os.environ.get("VIT_TBL", "auto"), "equal to the golden model" if ok else "MISMATCH"))
    dec.close()
sys.exit(1 if bad else 0)
