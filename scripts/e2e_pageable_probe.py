"""vit_run from pageable buffers: timeline (VIT_RUN_DEBUG) and throughput by staging threads / column blocks."""
import os, sys, time, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    V = bench.load_pkg()
    opt, n = int(sys.argv[2], 16), int(sys.argv[3])
    bits, packed, N = O.make_channel(n, opt & 0xF, snr_db=15.0, seed=5)
    dec = V.ViterbiCUDA(opt, N)
    M = dec.getMessageLen(N)
    ins = [np.array(packed, copy=True) for _ in range(3)]
    out = np.empty(dec.getOutputSize(N) // np.dtype(dec.decPack_t).itemsize, dec.decPack_t)
    for k in range(4):
        dec.run(ins[k % 3], N, output_h=out)
    os.environ.pop("VIT_RUN_DEBUG", None)
    t0 = time.perf_counter()
    K = 40
    for k in range(K):
        dec.run(ins[k % 3], N, output_h=out)
    dt = time.perf_counter() - t0
    print("opt %#x n %d threads %s blocks %s: %.3f ms per call = %.1f Gb/s" % (opt, n, os.environ.get("VIT_STAGE_THREADS", "def"),
          os.environ.get("VIT_STAGE_BLOCKS", "def"), dt / K * 1e3, M * K / dt / 1e9), flush=True)
    sys.exit(0)
for opt, n in (("011", 32_000_000), ("000", 32_000_000)):
    for th in ("2", "4", "8", "12", "16"):
        for bl in ("4", "8"):
            env = dict(os.environ, VIT_STAGE_THREADS=th, VIT_STAGE_BLOCKS=bl)
            if th == "8":
                env["VIT_RUN_DEBUG"] = "1"
            r = subprocess.run([sys.executable, __file__, "child", opt, str(n)], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
            if th == "8":
                print("   ", "\n    ".join(r.stderr.strip().splitlines()[-2:]))
