"""Stress of vit_run's host-buffer paths on a B200: one handle per option value, random stream lengths around the thresholds of
the time-sliced upload and the chunk pipeline, pinned and pageable buffers alternating, every output compared with the
device-resident decode of the same bytes.  usage: host_run_stress.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from oracle import oracle as O  # noqa: E402

V = bench.load_pkg()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(7)
opts = [0x011, 0x000, 0x010, 0x112, 0x021, 0x002, 0x004, 0x123]
decs = {o: V.ViterbiCUDA(o) for o in opts}
t0, cases, bad, gated = time.time(), 0, 0, 0
while time.time() - t0 < budget:
    opt = int(rng.choice(opts))
    it = opt & 0xF
    bytes_per_bit = {0: 0.25, 1: 1, 2: 2, 3: 4, 4: 8}[it]
    target = int(rng.choice([1.5e6, 2.2e6, 4.1e6, 8.5e6, 17e6, 40e6]))                 # input bytes
    n = int(target / bytes_per_bit) + int(rng.integers(0, 5000))
    bits, packed, N = O.make_channel_det(n, it, seed=int(rng.integers(1, 1 << 30)), sigma=0.6)
    dec = decs[opt]
    in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
    src = torch.from_numpy(packed.view(np.uint8)[:in_bytes].copy())
    pinned = bool(rng.integers(0, 3) != 0)
    h_in = src.pin_memory() if pinned else src
    h_out = torch.zeros(out_bytes, dtype=torch.uint8)
    h_out = h_out.pin_memory() if pinned else h_out
    l0 = dec.launch_count()
    dec.run(h_in.numpy(), N, output_h=h_out.numpy().view(dec.decPack_t))
    gated += (dec.launch_count() - l0 == 1 and in_bytes >= (8 << 20))
    d_in = src.cuda()
    d_out = torch.zeros(out_bytes + 256, dtype=torch.uint8, device="cuda")
    dec.run_device(d_in.data_ptr(), d_out.data_ptr(), N)
    torch.cuda.synchronize()
    ok = torch.equal(d_out[:out_bytes].cpu(), h_out)
    cases += 1
    if not ok:
        bad += 1
        print("MISMATCH options %#x n=%d pinned=%s" % (opt, n, pinned))
print("host-run stress: %d cases, %d single-launch (time-sliced) runs of >= 8 MB among them, mismatches vs device-resident decode: %d" % (cases, gated, bad))
sys.exit(1 if bad else 0)
