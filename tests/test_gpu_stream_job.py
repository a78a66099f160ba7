"""GPU tests (-m gpu) of the sharded stream job (C ABI vit_job_* / vit_comm_*, csrc/vit_mg.cu): BASELINE.json configs[4]
at a reduced stream count and length, through the same code the 1024-stream run uses.  One-GPU cases always run; the
2-GPU cases (NCCL send/recv, copy-engine and direct-store gathers; one process with threads, and one process per GPU
under torchrun) need a box with two GPUs."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from vit_testlib import PKG_DIR, ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _expected(V, O, torch, opt, n_bits, seed, sigma, source):
    per = {0: 32, 1: 8, 2: 4, 3: 2, 4: 1}[opt & 0xF]
    nsym = 2 * n_bits
    nbytes = ((nsym + per - 1) // per) * 4 if (opt & 0xF) != 4 else nsym * 4
    buf = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
    V.synth_device(opt & 0xF, n_bits, buf.data_ptr(), None, seed=seed, sigma=sigma, source=source)
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    return O.decode(opt, host, nsym)


@pytest.mark.parametrize("gather", ["none", "copy", "nccl", "direct"])
def test_stream_job_one_gpu(V, O, gather):
    """7 streams, waves of 2, batches of 4 (ragged last wave and batch): every stream of the gathered buffer equals the
    golden model's decode of the regenerated stream; per-stream bit errors are reported."""
    import torch
    opt, n_bits, ns = 0x112, 64 + 16 * 6400 * 5 + 16 * 3, 7
    job = V.StreamJob(None, 0, options=opt, n_bits=n_bits, nstreams=ns, wave=2, batch=4, seed=40, source=V.SOURCE_PRBS31,
                      amp=0, sigma=0.9, gather=V.GATHER_MODES[gather], root=0)
    r = job.run()
    assert r["streams"] == ns and r["launches"] == 4 and r["decode_ms"] > 0 and r["job_ms"] >= r["decode_ms"]
    errs = job.stream_errors()
    assert len(errs) == ns and sum(errs) == r["bit_errors"] and max(errs) == r["max_stream_errors"] and min(errs) > 0
    gptr, stride = job.gathered()
    if gather == "none":
        assert not gptr
    else:
        out_bytes = V.lib().vit_output_size(opt, 2 * n_bits)
        for s in range(ns):
            got = V.dev_to_host(gptr + s * stride, out_bytes).view(np.uint16)
            assert np.array_equal(got, _expected(V, O, torch, opt, n_bits, 40 + s, 0.9, V.SOURCE_PRBS31)), s
    r2 = job.run()                                     # a job can be run again (warm-up + timed passes)
    assert r2["bit_errors"] == r["bit_errors"]
    job.close()


def test_prbs_source_matches_cpu_twin(V, O):
    """PRBS-31 message source on the device (jump-ahead per thread) == the sequential CPU generator (vo_prbs31)."""
    import torch
    n = 300_000 + 7
    for seed in (1, 0x7FFFFFFF, 123456789):
        d_p = torch.zeros(n // 2 + 64, dtype=torch.uint8, device="cuda")           # hard input: 4 bytes per 16 bits
        d_b = torch.zeros(n, dtype=torch.uint8, device="cuda")
        V.synth_device(0, n, d_p.data_ptr(), d_b.data_ptr(), seed=seed, source=V.SOURCE_PRBS31)
        torch.cuda.synchronize()
        assert np.array_equal(d_b.cpu().numpy(), O.prbs31(seed, n)), seed
    # ... and the regenerating error counter agrees with the oracle's counter on a noisy decode
    opt, n_bits = 0x011, 64 + 32 * 9000
    per_bytes = n_bits           # s4: 1 byte per message bit
    d_p = torch.zeros(per_bytes + 64, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n_bits, dtype=torch.uint8, device="cuda")
    V.synth_device(1, n_bits, d_p.data_ptr(), d_b.data_ptr(), seed=77, sigma=1.1, source=V.SOURCE_PRBS31)
    dec = V.ViterbiCUDA(opt)
    N = 2 * n_bits
    d_o = torch.zeros(dec.getOutputSize(N), dtype=torch.uint8, device="cuda")
    dec.run_device(d_p.data_ptr(), d_o.data_ptr(), N)
    torch.cuda.synchronize()
    M = dec.getMessageLen(N)
    want = O.count_errors(opt, d_o.cpu().numpy().view(np.uint32), M, d_b.cpu().numpy())
    assert want > 0
    assert V.count_errors_synth_device(opt, d_o.data_ptr(), M, seed=77, source=V.SOURCE_PRBS31) == want
    assert V.count_errors_device(opt, d_o.data_ptr(), d_b.data_ptr(), M) == want
    dec.close()


def test_harness_stream_job_one_gpu():
    exe = os.path.join(PKG_DIR, "host", "main")
    out = subprocess.run([exe, "--streams", "6", "-n", "8000000", "-i", "s8", "-m", "b16", "-o", "b16", "--wave", "4", "--batch", "4",
                          "--seed", "3", "--prbs", "--gather", "copy"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if "box time" in l]
    assert len(lines) == 2 and all(int(re.search(r"BEN: (\d+)", l).group(1)) == 0 for l in lines), out.stdout[-800:]
    assert all("Gb/s" in l for l in lines)


@pytest.mark.parametrize("gather", ["nccl", "copy", "direct"])
def test_harness_stream_job_two_gpus_one_process(gather):
    """host/main.cpp --gpus 2: one process, one thread + decoder + communicator per GPU (vit_comm_init_all)."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(PKG_DIR, "host", "main")
    out = subprocess.run([exe, "--streams", "7", "--gpus", "2", "-n", "8000000", "-i", "s8", "-m", "b16", "--wave", "2", "--batch", "2",
                          "--seed", "5", "--gather", gather], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if "box time" in l]
    assert len(lines) == 2 and all(int(re.search(r"BEN: (\d+)", l).group(1)) == 0 for l in lines), out.stdout[-800:]


@pytest.mark.parametrize("gather", ["nccl", "copy", "direct"])
def test_bench_config5_two_gpus_torchrun(gather):
    """bench.py --workload config5 under torchrun (one process per GPU, vit_comm_init_rank): 6 streams of 256 Mbit, oracle
    slices of the first and last stream of every rank read from the gathered buffer on rank 0."""
    import json
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29551", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "config5", "--streams", "6",
           "--wave", "2", "--batch", "2", "--gather", gather]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["check"]["bit_errors"] == 0
    assert line["check"]["oracle_slices_equal"] is True and len(line["check"]["streams_checked"]) == 4
    assert line["value"] > 0 and line["value_without_gather"] > 0


@pytest.mark.parametrize("gather", ["copy", "nccl", "direct"])
def test_bench_default_two_gpus_torchrun(gather):
    """The per-step bench (weak scaling, config 2 per rank) with the library's gather: rank 0 checks a slice of what the
    last rank sent against the golden model."""
    import json
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29552", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "20", "--warmup", "3", "--gather", gather,
           "--no-e2e"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["check"]["bit_errors"] == 0 and line["check"]["oracle_slice_equal"]
    assert line["check"]["gathered_slice_of_last_rank_equal"] is True
