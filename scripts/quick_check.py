"""Torch-free end-of-round check for a short GPU slot: smoke(), then one handle fed growing pageable inputs through the
overlapped host-buffer path, a mixed pinned/pageable run and a chunked stream push, each against the golden model."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

t0 = time.perf_counter()
g.smoke()
print("smoke ok  %.1f s" % (time.perf_counter() - t0), flush=True)
from oracle import oracle as O  # noqa: E402  (checker only)
V = g._load_pkg()
opt = 0x011
dec = V.ViterbiCUDA(opt, 0, device=0)
for packs in (20, 47, 99):
    bits, packed, N = O.make_channel_det(6400 * 32 * packs + 64 + 32 * 311, opt & 0xF, seed=packs, sigma=0.8)
    for rep in range(2):
        out = dec.run(packed, N)
        assert np.array_equal(out, O.decode(opt, packed, N)), (packs, rep)
    print("growing pageable input: %d bits ok (mode %d, launches %d)  %.1f s"
          % (dec.getMessageLen(N), dec.upload_mode_in_effect(), dec.launch_count(), time.perf_counter() - t0), flush=True)
dec.stream_reset()
got = [dec.stream_push(packed[a:b], (b - a) * 8) for a, b in ((0, 100000), (100000, 100003), (100003, 700000))]
exp, _pending = O.decode_chunked(opt, packed, [100000 * 8, 3 * 8, 599997 * 8])
assert all(np.array_equal(x, y) for x, y in zip(got, exp))
print("stream push ok  %.1f s" % (time.perf_counter() - t0), flush=True)
dec.close()
print("QUICK CHECK PASSED")
