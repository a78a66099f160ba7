"""e2e probe: time vit_run (host pinned buffers) for one configuration, plus raw pinned H2D/D2H bandwidth.
usage: e2e_probe.py <options hex> <message bits> [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

opt, n, reps = int(sys.argv[1], 16), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 20
V = bench.load_pkg()
dev = torch.device("cuda", 0)
packed, N = bench.make_stream_device(V, torch, n, opt & 0xF, 15.0, 1, dev)
dec = V.ViterbiCUDA(opt, N)
in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)
h_in = packed[:in_bytes].cpu().pin_memory()
h_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
for _ in range(3):
    d_in.copy_(h_in, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d_in.copy_(h_in, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("pinned H2D %d bytes: %.3f ms = %.1f GB/s" % (in_bytes, dt * 1e3, in_bytes / dt / 1e9))
h_in_np, h_out_np = h_in.numpy(), h_out.numpy().view(dec.decPack_t)
for _ in range(3):
    dec.run(h_in_np, N, output_h=h_out_np)
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); dec.run(h_in_np, N, output_h=h_out_np); ts.append(time.perf_counter() - t0)
M = dec.getMessageLen(N)
ref = torch.empty(out_bytes + 256, dtype=torch.uint8, device=dev)
dec.run_device(packed.data_ptr(), ref.data_ptr(), N)
torch.cuda.synchronize()
same = bool((ref[:out_bytes].cpu() == h_out).all())
print("mode %s options %#x: e2e best %.3f ms median %.3f ms -> %.1f Gb/s (median); output equals device-resident decode: %s"
      % (os.environ.get("VIT_RUN_MODE", "0"), opt, min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3, M / sorted(ts)[len(ts) // 2] / 1e9, same))
