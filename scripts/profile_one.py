"""Run a few decodes of one configuration (for ncu).  usage: profile_one.py <options hex> <message bits> [reps]"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

opt, n, reps = int(sys.argv[1], 16), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
V = bench.load_pkg()
dev = torch.device("cuda", 0)
packed, N = bench.make_stream_device(V, torch, n, opt & 0xF, 15.0, 1, dev)
dec = V.ViterbiCUDA(opt, N)
out = torch.zeros(dec.getOutputSize(N) + 256, dtype=torch.uint8, device=dev)
ms = [dec.run_device(packed.data_ptr(), out.data_ptr(), N, want_kernel_time=True) for _ in range(reps)]
errs = V.count_errors_synth_device(opt, out.data_ptr(), dec.getMessageLen(N), seed=1, source=V.SOURCE_PRBS31)
print("options %#x  n=%d  kernel ms %s  -> %.1f Gb/s  bit errors %d" % (opt, n, ["%.3f" % m for m in ms], dec.getMessageLen(N) / min(ms) / 1e6, errs))
sys.exit(1 if errs else 0)
