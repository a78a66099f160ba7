import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import bench, torch
V = bench.load_pkg()
opt, n = 0x011, 32000000
dev = torch.device("cuda", 0)
bits, packed, N = bench.make_stream_device(torch, n, opt & 0xF, 15.0, 1, dev)
dec = V.ViterbiCUDA(opt, N)
in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
h_in = packed[:in_bytes].cpu().numpy().copy()          # pageable
h_out = np.zeros(out_bytes // 4, np.uint32)
for _ in range(3): dec.run(h_in, N, output_h=h_out)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); dec.run(h_in, N, output_h=h_out); ts.append(time.perf_counter() - t0)
print("pageable host buffers: median %.3f ms -> %.1f Gb/s" % (sorted(ts)[10] * 1e3, dec.getMessageLen(N) / sorted(ts)[10] / 1e9))
ts = []
for _ in range(20):
    t0 = time.perf_counter(); out, ms = dec.run(h_in, N, output_h=h_out, want_kernel_time=True); ts.append(time.perf_counter() - t0)
print("pageable + kernel time (reference sequence): median %.3f ms -> %.1f Gb/s" % (sorted(ts)[10] * 1e3, dec.getMessageLen(N) / sorted(ts)[10] / 1e9))
