#!/usr/bin/env python
"""bench.py -- decoded Gb/s of the B200-native Viterbi decoder on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--gather MODE]

One "step" = one decode of one synthetic codeword stream through the hot path.  The stream comes from the library's own
device source (vit_synth_device_ex): PRBS-31 message bits -> K=7 0171/0133 encoder -> BPSK + noise, sd = 10^(-snr/5) ->
x40000 -> saturating quantiser -> MSB-first packing, i.e. the reference harness's channel (reference src/main.cpp:131-138;
at that scale every soft symbol saturates, as in ./main).  Default workload = BASELINE.json configs[1]: 32,000,000 message
bits, 4-bit soft input, int16x2 core, 32-bit packs.

Printed JSON (rank 0, one line):
  value     whole-job decoded Gb/s with inputs resident in HBM, CUDA events on the launch stream, max over ranks.  For
            N > 1 every rank decodes its own streams (weak scaling, no stream is split) and the packed output bits reach
            rank 0 inside the timed region: by default the decode kernel stores them straight into rank 0's buffer over
            NVLink (--gather direct), alternatively through copy engines (copy) or NCCL send/recv (nccl) -- all three are
            the library's own C code (vit_comm_gatherv, csrc/vit_mg.cu); value_without_gather is measured in the same run.
  e2e       the same metric through the reference-facing call ViterbiCUDA.run(host_in, host_out) (C ABI vit_run): H2D
            copy, kernel, D2H copy every step; e2e.value from pinned buffers, e2e.pageable from pageable numpy buffers (the
            reference's calling convention); both outputs are checked against the device-resident decode.
  roofline  the kernel is issue-bound, not HBM-bound: achieved decoded Gb/s against the ACS-op issue
            roofline N_SM*4*f_SM / (3|6 warp-instructions per decoded bit); roofline_hbm is the algorithmic-bytes view
            against MEASURED_PEAKS.json.
  cpu_baseline  the scalar C golden model (oracle/, "port") on the host cores, bounded sample.
  reference_cuda  the reference's own CUDA decoder (oracle/_ref, built for sm_100) timed in the same run on the same GPU (N = 1).
`--workload config5` runs BASELINE.json configs[4] (1024 x 256-Mbit s8 streams, strong scaling) through vit_job_run.
`--impl reference` times the reference's own CUDA decoder (oracle/_ref/libvitref.so, built from the unmodified reference
sources for sm_100) through its own run(); if that library is absent it times the C golden model on the host cores.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (options, message bits, snr dB)                      BASELINE.json configs[...]
    "s4_b16_o32_32M": (0x011, 32_000_000, 15.0),   # configs[1] (the headline: metric is quoted on this)
    "hard_b32_o32_32M": (0x000, 32_000_000, 15.0),  # ./main with no flags
    "hard_b32_o32_1M": (0x000, 1_000_000, 5.5),     # configs[0]
    "s8_f16_o16_256M": (0x122, 256_000_000, 3.0),   # configs[2]
    "s16_f16_o16_256M": (0x123, 256_000_000, 3.0),  # configs[2]
    "f_b32_o32_4G": (0x004, 4_000_000_000, 15.0),   # configs[3]
    "s8_b16_o32_256M": (0x012, 256_000_000, 15.0),  # configs[4] per-stream shape
    "config5": (0x012, 256_000_000, 15.0),          # configs[4]: the 1024-stream job over 1/2/4/8 GPUs (--streams to scale it down)
}
BYTES_PER_BIT_IN = {0: 0.25, 1: 1.0, 2: 2.0, 3: 4.0, 4: 8.0}
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the decode
# kernel on that workload), with the committed capture it was read from; workloads without a capture report null.
# The channel words are read exactly once; the decoded packs are still in L2 when the kernel ends (bytes written = 0).
NCU_DRAM_BYTES = {
    ("s4_b16_o32_32M", 1): (32.03e6, "profiles/r1_v7_ncu_core_0x011.txt"),
}
N_SM = 148


def load_pkg():
    name = "gpu_accelerated_viterbi_decoder_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------
# synthetic channel: generated on the device by the library's own source (vit_synth_device_ex, csrc/vit_synth.cu)
# ------------------------------------------------------------------------------------------------
HARNESS_SCALE = 40000          # the reference harness's quantiser scale (main.cpp:137)


def make_stream_device(V, torch, n_bits, input_type, snr_db, seed, device, pad_to=None):
    """PRBS-31 message bits (state = seed) -> K=7 0171/0133 encoder -> BPSK +-amp + noise (sd = 10^(-snr/5) * amp) ->
    the reference's saturating quantiser and MSB-first packing, all on the device.  amp = 40000 quantiser units, the
    reference harness's scale (main.cpp:135-137): every soft symbol saturates (4-bit soft input is -8 / +7), exactly as
    `./main` feeds its decoder.  Returns (packed uint8 tensor on the device, inputNum)."""
    per = {0: 32, 1: 8, 2: 4, 3: 2, 4: 1}[input_type]
    nsym = 2 * n_bits
    nbytes = ((nsym + per - 1) // per) * 4 if input_type != 4 else nsym * 4
    size = max(nbytes + 64, pad_to or 0)
    packed = torch.zeros(size, dtype=torch.uint8, device=device)
    V.synth_device(input_type, n_bits, packed.data_ptr(), None, seed=seed, amp=HARNESS_SCALE, sigma=10.0 ** (-snr_db / 5.0),
                   source=V.SOURCE_PRBS31)
    return packed, nsym


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  A thread polls NVML
    (pynvml, ~0.1 ms per query) every period_ms; if NVML is unavailable it falls back to `nvidia-smi -lms`.
    mark()/unmark() bracket the timed region so that the summary can tell its samples from the ones taken during
    the warm-up and kernel-timing loops (same kernels, same load)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=2):
        self.index, self.rows, self.proc, self.period_ms = index, [], None, period_ms
        self.t_mark = self.t_unmark = None
        self.stop, self.t, self.nv, self.max_mhz, self.source, self.nq = False, None, None, None, None, 0

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except ValueError:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv, self.source = pynvml, "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(max(self.period_ms, 5))],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.15)               # let the first samples arrive before load starts
        except Exception:
            self.proc = None
        return self

    def mark(self):
        self.t_mark = time.perf_counter()

    def unmark(self):
        self.t_unmark = time.perf_counter()

    def _poll_nvml(self):
        nv = self.nv
        bits = ((getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4), "sw_power_cap"))
        while not self.stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                flags = ["Active" if mask & b else "Not Active" for b, _ in bits]
                self.rows.append((time.perf_counter(), [str(self.index), str(mhz), str(self.max_mhz), str(pw), hex(mask)] + flags))
                self.nq += 1
            except Exception:
                pass
            time.sleep(self.period_ms / 1000.0)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        self.stop = True
        if self.nv is not None:
            if self.t:
                self.t.join(timeout=1)
            try:
                self.nv.nvmlShutdown()
            except Exception:
                pass
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        def num(v):
            try:
                return float(v)
            except (ValueError, TypeError):
                return None
        rows = [(t, r) for t, r in self.rows if len(r) >= 9 and num(r[1]) is not None]
        inside = [(t, r) for t, r in rows if self.t_mark is not None and self.t_unmark is not None and self.t_mark <= t <= self.t_unmark + 0.002]
        use = inside if inside else rows
        sm = [num(r[1]) for _, r in use]
        mx = [num(r[2]) for _, r in rows if num(r[2]) is not None]
        pw = [num(r[3]) for _, r in use if num(r[3]) is not None]
        reasons = set()
        for _, r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons),
                "samples_in_timed_region": len(inside), "samples_under_load": len(rows), "period_ms": self.period_ms,
                "source": self.source}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


def workload_config(name, options, n_bits, snr, gpus=1):
    """`config`: identical in both arms (the driver compares the two dicts); run-specific detail goes to `run_info`."""
    return {"workload": name, "message_bits": n_bits, "options": "0x%03x" % options, "snr_db": snr, "segments": 6400,
            "source": "PRBS-31 message bits, K=7 0171/0133, BPSK + noise sd 10^(-snr/5), x40000 saturating quantiser "
                      "(the reference harness's scale: soft symbols saturate, e.g. -8/+7 for 4-bit input)",
            "l2": "every step reads its input from HBM: inputs rotate over distinct device buffers larger than the 126 MB L2 "
                  "(reference arm: the stream is uploaded from the host again every step)",
            "parallelism": "single GPU" if gpus <= 1 else "independent streams sharded over %d GPUs, none split; packed output bits gathered to rank 0" % gpus}


def cpu_baseline(options, n_bits, snr, budget_s=12.0):
    """Golden model on the host cores: whole stream if it is small, else a 32-Mbit stream of the same options."""
    from oracle import oracle as O
    n_cpu = min(n_bits, 32_000_000)
    bits, packed, N = O.make_channel(n_cpu, options & 0xF, snr_db=snr, seed=5, prbs=True)
    M = O.message_len(options, N)
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.decode(options, packed, N)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or reps >= 20:
            break
    return {"value": M * reps / dt / 1e9, "unit": "Gb/s", "cores": O.num_threads(), "kind": "port",
            "sample": "%d x full decode of a %d-bit stream (same options), OpenMP over the 6400 segments" % (reps, n_cpu)}


def reference_cuda_baseline(workload, reps=10):
    """The reference's own CUDA decoder (oracle/_ref/libvitref.so, the unmodified sources built for sm_100) timed in THIS run
    on the same GPU, beside the CPU baseline (BASELINE.json: "three baselines timed on the same box in the same run"): the
    reference arm of this script in a child process -- its run() answers a CUDA error with exit(), which must not be able to
    cost this line.  A reported baseline; the driver's ratio comes from its own `--impl reference` run.  Never raises."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                            "--steps", str(reps), "--warmup", "3"], capture_output=True, text=True, timeout=300,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = json.loads(r.stdout.strip().splitlines()[-1])
        if line.get("cpu_baseline", {}).get("kind") != "reference":
            return {"unavailable": line.get("note", "the reference's CUDA decoder did not run")}
        return {"value": line["value"], "unit": "Gb/s", "kernel_ms": line["ms_per_step"], "e2e": line["e2e"]["value"],
                "kind": "reference CUDA decoder, -arch=sm_100", "sample": line["cpu_baseline"]["sample"]}
    except Exception as e:          # a baseline must never cost the line
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args, options, n_bits, snr):
    """--impl reference: the reference's own CUDA decoder through its own run() (pageable host buffers)."""
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"metric": "decoded Gb/s", "unit": "Gb/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference",
            "config": workload_config(args.workload, options, n_bits, snr, args.gpus)}
    have_gpu = O.ref_lib() is not None and O.ref_lib().ref_device_count() > 0
    if not have_gpu or not O.lib().vo_options_valid_ref(options):
        cb = cpu_baseline(options, n_bits, snr, budget_s=20.0)
        base.update({"value": cb["value"], "ms_per_step": None, "dtype": "int64", "cpu_baseline": cb, "gpu_launches": 0,
                     "e2e": {"value": cb["value"], "unit": "Gb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": "reference CUDA decoder unavailable for this combination: C golden model (port) on host cores"})
        print(json.dumps(base))
        return
    bits, packed, N = O.make_channel(min(n_bits, 256_000_000), options & 0xF, snr_db=snr, seed=5, prbs=True)
    M = O.message_len(options, N)
    for _ in range(args.warmup):
        O.ref_decode(options, packed, N)
    kms, t0 = [], time.perf_counter()
    with ClockSampler(0) as cs:
        for _ in range(args.steps):
            _, ms = O.ref_decode(options, packed, N)
            kms.append(ms)
    wall = time.perf_counter() - t0
    base.update({
        "value": M / (statistics.mean(kms) * 1e6), "ms_per_step": statistics.mean(kms),
        "dtype": {0x00: "int32", 0x10: "int16x2", 0x20: "f16x2"}[options & 0xF0],
        "gpu_launches": args.steps, "clocks": cs.summary(),
        "e2e": {"value": M * args.steps / wall / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": int(O.input_size(options, N)),
                "d2h_bytes_per_step": int(O.output_size(options, N)), "host_memory": "pageable (numpy), the reference's calling convention",
                "note": "wall clock of ViterbiCUDA::run incl. its per-call cudaMalloc / pageable cudaMemcpy / cudaFree (viterbi.cu:217-237); "
                        "compare with our e2e.pageable (same buffers), not only with our pinned e2e.value"},
        "cpu_baseline": {"value": M / (statistics.mean(kms) * 1e6), "unit": "Gb/s", "cores": 0, "kind": "reference",
                         "sample": "reference CUDA decoder (oracle/_ref/libvitref.so, -arch=sm_100) on GPU 0; value = its own cudaEvent kernel time, "
                                   "e2e = wall clock of ViterbiCUDA::run incl. its cudaMalloc/cudaMemcpy/cudaFree"},
    })
    print(json.dumps(base))


def make_comm(V, torch, dist, world, rank, local, dev):
    """The library's own communicator (NCCL loaded by libvitb200.so); torch.distributed only carries the 128-byte id."""
    idt = torch.zeros(V.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(V.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    return V.Comm(world, rank, bytes(idt.cpu().numpy().tobytes()), local)


def reduce_max(torch, dist, dev, x, world):
    if world == 1:
        return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(torch, dist, dev, x, world):
    if world == 1:
        return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_config5(args, V, torch, dist, rank, world, local, dev):
    """BASELINE.json configs[4]: --streams (default 1024) independent 256-Mbit s8 streams, int16x2 core, sharded over the
    ranks in contiguous blocks, generated per batch on the device, decoded --wave streams per launch, the packed output
    bits gathered to rank 0 per finished wave (vit_job_run, csrc/vit_mg.cu).  STRONG scaling: the job is fixed."""
    options, n_bits, snr = WORKLOADS[args.workload]
    nstreams = args.streams if args.streams > 1 else 1024
    comm = make_comm(V, torch, dist, world, rank, local, dev) if world > 1 else None
    mode = V.GATHER_MODES[args.gather]
    passes = {}
    checks = {}
    for name, m in (("with_gather", mode), ("without_gather", V.GATHER_NONE)):
        if name == "with_gather" and mode == V.GATHER_NONE:
            continue
        job = V.StreamJob(comm, local, options=options, n_bits=n_bits, nstreams=nstreams, wave=args.wave, batch=args.batch, seed=1,
                          source=V.SOURCE_PRBS31, amp=HARNESS_SCALE, sigma=10.0 ** (-snr / 5.0), gather=m, root=0)
        with ClockSampler(local) as cs:
            cs.mark()
            r = job.run()
            cs.unmark()
        errs = job.stream_errors()
        r["job_ms_max"] = reduce_max(torch, dist, dev, r["job_ms"], world)
        r["decode_ms_max"] = reduce_max(torch, dist, dev, r["decode_ms"], world)
        r["bits_total"] = reduce_sum(torch, dist, dev, float(r["decoded_bits"]), world)
        r["errors_total"] = reduce_sum(torch, dist, dev, float(r["bit_errors"]), world)
        r["worst_stream_errors"] = reduce_max(torch, dist, dev, float(max(errs) if errs else 0), world)
        r["launches_total"] = reduce_sum(torch, dist, dev, float(r["launches"]), world)
        r["clocks"] = cs.summary()
        passes[name] = r
        if name == "with_gather" and rank == 0:
            # oracle slices of two streams of EVERY rank's block, read from the gathered buffer on the root
            from oracle import oracle as O
            gptr, stride = job.gathered()
            N = 2 * n_bits
            bpp = 16 if options & 0x100 else 32
            ok, checked = True, []
            for p in range(world):
                first, count = V.shard_range(nstreams, world, p)
                for s in sorted({first, first + count - 1}):
                    buf, _ = make_stream_device(V, torch, n_bits, options & 0xF, snr, 1 + s, dev)
                    for sa, sb in ((0, 2), (6398, 6400)):
                        b0, nb, w0, nw = O.segment_window(options, N, sa, sb)
                        _, exp = O.decode_window(options, buf[b0:b0 + nb].cpu().numpy(), N, sa, sb)
                        got = V.dev_to_host(gptr + s * stride + w0 * (bpp // 8), nw * (bpp // 8)).view(exp.dtype)
                        ok = ok and bool(np.array_equal(got, exp))
                    checked.append(s)
                    del buf
            checks = {"oracle_slices_equal": ok, "streams_checked": checked}
        job.close()
    if rank == 0:
        main_pass = passes.get("with_gather", passes["without_gather"])
        wo = passes["without_gather"]
        bits = main_pass["bits_total"]
        _, sm_max_mhz, _ = measured_peaks()
        peak = N_SM * 4 * sm_max_mhz * 1e6 / 3 / 1e9
        k_gbps = bits / world / (wo["decode_ms_max"] * 1e6)       # per GPU, decode launches only
        line = {
            "metric": "decoded Gb/s", "value": bits / (main_pass["job_ms_max"] * 1e6), "unit": "Gb/s", "n_gpus": world, "steps": 1, "warmup": 0,
            "ms_per_step": main_pass["job_ms_max"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16x2",
            "data": "synthetic",
            "value_without_gather": wo["bits_total"] / (wo["job_ms_max"] * 1e6),
            "config": dict(workload_config(args.workload, options, n_bits, snr, world), streams=nstreams, wave=args.wave, batch=args.batch,
                           gather=args.gather, nccl_version=V.lib().vit_comm_nccl_version() if world > 1 else None,
                           sharding="contiguous blocks of %d streams per GPU, one decoder per GPU, packed output bits gathered to rank 0 per finished wave" % (nstreams // world),
                           timing="device events per batch round: first decode launch -> end of the round's last gather, summed; max over ranks; generation untimed"),
            "gpu_launches": int(main_pass["launches_total"]), "clocks": main_pass["clocks"],
            "check": dict(checks, bit_errors=int(main_pass["errors_total"]), ber=main_pass["errors_total"] / bits,
                          worst_stream_bit_errors=int(main_pass["worst_stream_errors"])),
            "passes": {k: {kk: v[kk] for kk in ("job_ms_max", "decode_ms_max", "synth_ms", "bits_total", "errors_total")} for k, v in passes.items()},
            "roofline": {"bound": "issue", "achieved": k_gbps, "peak": peak, "unit": "Gb/s decoded per GPU", "frac": k_gbps / peak,
                         "how": "per-GPU decode rate of the multi-stream launches against the ACS-op issue roofline (3 warp-instructions per decoded bit)", "traffic": None},
            "e2e": None,
        }
        print(json.dumps(line))
    if comm is not None:
        comm.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s4_b16_o32_32M", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=1, help="independent codeword streams decoded per step by ONE launch; config5: streams of the whole job (default 1024)")
    ap.add_argument("--gather", default=None, choices=["nccl", "copy", "direct", "none"],
                    help="N > 1: how the packed output bits reach rank 0 (default: direct)")
    ap.add_argument("--wave", type=int, default=8, help="config5: streams per decode launch (the last wave's gather is the exposed tail of a round)")
    ap.add_argument("--batch", type=int, default=64, help="config5: streams generated ahead of each timed decode phase (one gather tail per round)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    options, n_bits, snr = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, options, n_bits, snr)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    V = load_pkg()
    if args.workload == "config5":
        args.gather = args.gather or "direct"
        run_config5(args, V, torch, dist, rank, world, local, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    # direct stores: nothing runs but the decode kernel, which stages 8 slides per segment and stores 32 bytes at a time into
    # rank 0's buffer over NVLink: 99.0-99.7 % of N x one GPU at N = 2, 4, 8 (copy engines: 95 %, NCCL send/recv: 89 %;
    # profiles/bench_r2/scale8_staged, scale8)
    args.gather = args.gather or "direct"
    it, bpp = options & 0xF, (16 if options & 0x100 else 32)
    dec = V.ViterbiCUDA(options, 2 * n_bits, device=local)
    N = 2 * n_bits
    M = dec.getMessageLen(N)
    in_bytes, out_bytes = dec.getInputSize(N), dec.getOutputSize(N)

    # distinct input streams rotated step to step so that the working set exceeds L2 (126 MB)
    S = max(1, args.streams)
    in_stride = (in_bytes + 4 + 255) // 256 * 256
    nbuf = max(2, min(8, int(300e6 // (in_bytes * S)) + 1)) if in_bytes * S < 300e6 else 1
    seed0 = 1 + 1000 * rank
    streams = []
    for k in range(nbuf):
        buf = torch.zeros(S * in_stride, dtype=torch.uint8, device=dev)
        for j in range(S):
            packed, _ = make_stream_device(V, torch, n_bits, it, snr, seed0 + k * S + j, dev)
            buf[j * in_stride:j * in_stride + in_bytes] = packed[:in_bytes]
            del packed
        streams.append(buf)
    torch.cuda.synchronize()

    # Outputs go to a ring of 2 x GB slots per rank.  For N > 1 the packed output bits are gathered to rank 0 with the
    # library's own gather (vit_comm_gatherv, csrc/vit_mg.cu) on its own high-priority stream while the next decodes run:
    #   copy   one device-to-device copy per step into rank 0's buffer (copy engines over NVLink, no SM used)
    #   nccl   one grouped ncclSend/ncclRecv per GB steps (NCCL kernels beside the decode kernel)
    #   direct the decode kernel stores its packs straight into rank 0's buffer over NVLink
    GB = 8 if world > 1 else 1
    CG = 2                                                  # steps per copy-engine transfer (divides GB)
    out_stride1 = (out_bytes + 255) // 256 * 256          # per stream
    out_stride = out_stride1 * S                            # per step
    ring_bytes = 2 * GB * out_stride
    ring = torch.zeros(ring_bytes, dtype=torch.uint8, device=dev)
    comm = make_comm(V, torch, dist, world, rank, local, dev) if world > 1 else None
    root_buf = comm.shared_alloc(world * ring_bytes, 0) if comm is not None else None
    st = torch.cuda.current_stream()
    state = {"gather": V.GATHER_MODES[args.gather] if world > 1 else V.GATHER_NONE}

    def out_ptr(slot):
        if state["gather"] == V.GATHER_DIRECT:
            return root_buf + rank * ring_bytes + slot * out_stride      # rank 0's buffer, mapped into this process
        return ring.data_ptr() + slot * out_stride

    arrays = {}

    def gather_slots(slot0, nslots):
        key = (slot0, nslots)
        if key not in arrays:
            arrays[key] = (comm.size_array([p * ring_bytes + slot0 * out_stride for p in range(world)]), comm.size_array([nslots * out_stride] * world))
        offs, sizes = arrays[key]
        comm.gatherv(state["gather"], ring.data_ptr() + slot0 * out_stride, root_buf, offs, sizes, 0, st.cuda_stream)

    def step(k):
        slot = k % (2 * GB)
        h = slot // GB
        if comm is not None and slot % GB == 0 and state["gather"] in (V.GATHER_NCCL, V.GATHER_COPY):
            comm.stream_wait_mark(h, st.cuda_stream)   # the gathers that read this half of the ring (issued GB..2*GB steps ago) are done
        dec.run_device(streams[k % nbuf].data_ptr(), out_ptr(slot), N, stream=st.cuda_stream,
                       nstreams=S, in_stride=in_stride, out_stride=out_stride1)
        if state["gather"] == V.GATHER_COPY:
            if slot % CG == CG - 1:
                gather_slots(slot - CG + 1, CG)        # one copy-engine transfer per CG steps
        elif state["gather"] == V.GATHER_NCCL and slot % GB == GB - 1:
            gather_slots(slot - GB + 1, GB)
        if comm is not None and slot % GB == GB - 1 and state["gather"] in (V.GATHER_NCCL, V.GATHER_COPY):
            comm.mark(h)

    def flush_partial(last_k):
        g = GB if state["gather"] == V.GATHER_NCCL else CG if state["gather"] == V.GATHER_COPY else 0
        if g and (last_k % g) != g - 1:
            slot = last_k % (2 * GB)
            gather_slots(slot - slot % g, slot % g + 1)                  # partial group at the end of a run

    def drain(last_k=None):
        if last_k is not None:
            flush_partial(last_k)
        if comm is not None:
            st.synchronize()
            comm.barrier()                     # every rank's gathers have landed in rank 0's buffer

    def timed(steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for k in range(steps):
            step(k)
        flush_partial(steps - 1)
        if comm is not None and state["gather"] in (V.GATHER_NCCL, V.GATHER_COPY):
            comm.stream_wait(st.cuda_stream)   # the timed region ends when this rank's gathers have completed
        e1.record(st)
        drain()                                # host-waits for this rank, then meets the other ranks
        wall_ms = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), wall_ms

    # correctness gate inside the bench: BER == 0 at this SNR, and a slice equals the golden model
    step(0)
    drain(0)
    torch.cuda.synchronize()
    if comm is not None and state["gather"] == V.GATHER_DIRECT:
        # a sender's kernel must have recognised rank 0's buffer as remote (32-byte staged stores); rank 0 stores locally
        assert dec.last_launch_staged_output() == (rank != 0), "direct-store gather: remote output buffer not recognised"
    gathering = comm is not None and state["gather"] != V.GATHER_NONE
    # rank 0 checks what arrived in the gathered buffer; a sender checks its local packs (or, with direct stores, its
    # block of rank 0's buffer through the mapping)
    first_out = ring.data_ptr()
    if gathering and (rank == 0 or state["gather"] == V.GATHER_DIRECT):
        first_out = root_buf + rank * ring_bytes
    errs = V.count_errors_synth_device(options, first_out, M, seed=seed0, source=V.SOURCE_PRBS31)
    check = {"bit_errors": errs, "ber": errs / M}
    if rank == 0:
        from oracle import oracle as O
        host_in = streams[0][:in_bytes].cpu().numpy()   # stream 0 of the step
        seg = O.decode(options, host_in, N, segs=(1000, 1016))
        P = M // bpp
        q, r = divmod(P, 6400)
        a, b = q * 1000 + min(1000, r), q * 1016 + min(1016, r)
        got = V.dev_to_host(first_out, out_bytes).view(np.uint16 if bpp == 16 else np.uint32)
        check["oracle_slice_equal"] = bool(np.array_equal(got[a:b], seg[a:b]))
        if comm is not None and state["gather"] != V.GATHER_NONE:
            # what the other ranks sent: the last rank's first stream, regenerated here and decoded by the golden model
            pl, _ = make_stream_device(V, torch, n_bits, it, snr, 1 + 1000 * (world - 1), dev)
            seg = O.decode(options, pl[:in_bytes].cpu().numpy(), N, segs=(1000, 1016))
            got = V.dev_to_host(root_buf + (world - 1) * ring_bytes, out_bytes).view(np.uint16 if bpp == 16 else np.uint32)
            check["gathered_slice_of_last_rank_equal"] = bool(np.array_equal(got[a:b], seg[a:b]))
            del pl

    for k in range(args.warmup):
        step(k)
    drain(args.warmup - 1)
    with ClockSampler(local) as cs:
        # same kernels again, for about 0.3 s: NVML answers a clock query in milliseconds to tens of milliseconds,
        # so a short timed region alone may hold few samples; the sampler sees this identical load as well.
        # The step count is derived from the workload size only, NOT from a clock: every rank must run the same number
        # of steps, because step() issues the gathers.
        est_step_s = M * S / 80e9
        n_load = max(args.warmup, min(2000, int(0.3 / est_step_s) + 1))
        for k in range(n_load):
            step(k)
            if (k + 1) % 64 == 0:
                st.synchronize()
        drain(n_load - 1)
        launches0 = dec.launch_count()
        cs.mark()
        ms_dev, ms_wall = timed(args.steps)
        cs.unmark()
        launches = dec.launch_count() - launches0
        ms_total = reduce_max(torch, dist, dev, ms_dev, world)
        ms_wall_max = reduce_max(torch, dist, dev, ms_wall, world)
        ms_nogather = None
        if world > 1 and state["gather"] != V.GATHER_NONE:
            keep = state["gather"]
            state["gather"] = V.GATHER_NONE
            nd, _ = timed(args.steps)
            ms_nogather = reduce_max(torch, dist, dev, nd, world)
            state["gather"] = keep
        # kernel-only duration for the roofline: events bracketing the launch on the launch stream
        kms = []
        for k in range(min(args.steps, 20)):
            kms.append(dec.run_device(streams[k % nbuf].data_ptr(), ring.data_ptr(), N, stream=st.cuda_stream, want_kernel_time=True,
                                      nstreams=S, in_stride=in_stride, out_stride=out_stride1))
    ms_step = ms_total / args.steps
    value = M * S * world / (ms_step * 1e6)                   # whole-job decoded Gb/s
    kernel_ms = statistics.mean(kms)

    # e2e through the public host-buffer call, copies inside the timed region: pinned buffers, and pageable ones
    e2e = None
    if not args.no_e2e and in_bytes < 8e9 and S == 1:
        h_in = [s[:in_bytes].cpu().pin_memory() for s in streams[:min(nbuf, 4)]]
        h_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
        h_out_np = h_out.numpy().view(dec.decPack_t)
        h_in_np = [h.numpy() for h in h_in]

        def e2e_leg(bufs_in, buf_out, steps):
            for k in range(3):
                dec.run(bufs_in[k % len(bufs_in)], N, output_h=buf_out)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                dec.run(bufs_in[k % len(bufs_in)], N, output_h=buf_out)
            return reduce_max(torch, dist, dev, time.perf_counter() - t0, world)
        dt = e2e_leg(h_in_np, h_out_np, args.steps)
        e2e = {"value": M * world * args.steps / dt / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": int(in_bytes),
               "d2h_bytes_per_step": int(out_bytes), "host_memory": "pinned", "timer": "host wall clock around the synchronous call"}
        # the same call from PAGEABLE numpy buffers -- the reference's calling convention (std::vector storage,
        # viterbiDF.h:188-193) and what the reference arm's e2e is measured with
        p_in = [np.array(h, copy=True) for h in h_in_np[:3]]
        p_out = np.empty(out_bytes // np.dtype(dec.decPack_t).itemsize, dec.decPack_t)
        n_pg = max(3, args.steps // 2)
        dt = e2e_leg(p_in, p_out, n_pg)
        e2e["pageable"] = {"value": M * world * n_pg / dt / 1e9, "unit": "Gb/s", "steps": n_pg, "host_memory": "pageable (numpy)",
                           "how": "staged through the handle's pinned buffers by worker threads, time-sliced upload (csrc/vit_api.cu run_gated)"}
        dec.run(p_in[0], N, output_h=p_out)
        dec.run(h_in_np[0], N, output_h=h_out_np)
        e2e["pageable"]["equals_pinned_output"] = bool(np.array_equal(p_out, h_out_np))
        # ... and both equal the device-resident decode of the same stream (the host-buffer paths overlap copies and kernel:
        # a result that only matched itself would prove nothing)
        chk = torch.zeros(out_stride1 + 256, dtype=torch.uint8, device=dev)
        dec.run_device(streams[0].data_ptr(), chk.data_ptr(), N, stream=st.cuda_stream)
        torch.cuda.synchronize()
        e2e["equals_device_decode"] = bool(np.array_equal(chk[:out_bytes].cpu().numpy(), h_out.numpy()))
        assert e2e["equals_device_decode"] and e2e["pageable"]["equals_pinned_output"], "host-buffer decode differs from the device-resident decode"

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_kind = measured_peaks()
        clocks = cs.summary()
        f_mhz = clocks["sm_mhz"] or sm_max_mhz
        wi_per_bit = 6 if (options & 0xF0) == 0 else 3          # int32 core 6, packed cores 3
        issue_peak_nominal = N_SM * 4 * sm_max_mhz * 1e6 / wi_per_bit / 1e9
        issue_peak_at_clock = N_SM * 4 * f_mhz * 1e6 / wi_per_bit / 1e9
        k_gbps = M * S / (kernel_ms * 1e6)
        alg_bytes = M * S * (BYTES_PER_BIT_IN[it] + 0.125)
        info = dec.kernel_info()
        traffic = NCU_DRAM_BYTES.get((args.workload, S))
        cfg = workload_config(args.workload, options, n_bits, snr, world)
        run_info = {"streams_per_step_per_gpu": S, "l2": "inputs rotate over %d distinct device buffers (%.0f MB > 126 MB L2)" % (nbuf, nbuf * in_bytes * S / 1e6),
                    "timed_region_ms": ms_total,
                    "note": "a step is one 32-Mbit decode (~0.33 ms): K steps are a short timed region; the NVML sampler also covers an identical ~0.3 s load phase before it",
                    "gather": ("vit_comm_gatherv mode '%s' on its own stream, overlapped with the following decodes; the timed region ends when this rank's gathers have completed" % args.gather) if world > 1 else None}
        line = {
            "metric": "decoded Gb/s", "value": value, "unit": "Gb/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0x00: "int32", 0x10: "int16x2", 0x20: "f16x2"}[options & 0xF0], "data": "synthetic",
            "config": cfg, "run_info": run_info,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "check": check,
            "kernel": {"ms": kernel_ms, "gbps": k_gbps, "regs": info["regs"], "smem_bytes": info["smem_bytes"],
                       "grid": [1600, S, 1], "block": 32},
            "roofline": {"bound": "issue", "achieved": k_gbps, "peak": issue_peak_nominal, "unit": "Gb/s decoded", "frac": k_gbps / issue_peak_nominal,
                         "peak_at_measured_clock": issue_peak_at_clock, "frac_at_measured_clock": k_gbps / issue_peak_at_clock,
                         "achieved_acs_warp_inst_per_s": k_gbps * 1e9 * wi_per_bit, "peak_warp_inst_per_s": N_SM * 4 * sm_max_mhz * 1e6,
                         "how": "ACS-op roofline: 192 add/compare-select ops per decoded bit = %d warp-instructions; peak = 148 SM x 4 issue/clk x f_SM / that" % wi_per_bit,
                         "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (kernel_ms * 1e6), "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (kernel_ms * 1e6) / hbm_peak, "peak_kind": peak_kind, "traffic": traffic[0] if traffic else None,
                             "algorithmic_bytes_per_launch": alg_bytes},
        }
        if world > 1:
            line["value_without_gather"] = M * S * world * args.steps / (ms_nogather * 1e6) if ms_nogather else value
            line["gather"] = {"mode": args.gather, "wall_ms_incl_gather_tail": ms_wall_max,
                              "value_by_wall_clock": M * S * world * args.steps / (ms_wall_max * 1e6)}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(options, n_bits, snr)
            line["reference_cuda"] = reference_cuda_baseline(args.workload)
        print(json.dumps(line))
    dec.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
