"""CPU tests of bench.py's contract pieces that need no GPU: the reference arm's line (here, without a GPU, it times the C
golden model on the host cores and says so), the `config` dict shared by both arms, and the clock-sample summary."""
import importlib.util
import json
import os
import subprocess
import sys

from vit_testlib import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("vit_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_line_without_gpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "hard_b32_o32_1M",
                        "--steps", "3", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "decoded Gb/s" and line["unit"] == "Gb/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 3
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "1000000-bit" in cb["sample"]
    assert line["gpu_launches"] == 0
    # the same config dict our arm prints for this workload (the driver compares them)
    B = _bench()
    opt, n_bits, snr = B.WORKLOADS["hard_b32_o32_1M"]
    assert line["config"] == B.workload_config("hard_b32_o32_1M", opt, n_bits, snr, 1)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                        "hard_b32_o32_1M"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workloads_are_the_baseline_configs():
    B = _bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    assert B.WORKLOADS["s4_b16_o32_32M"] == (0x011, 32_000_000, 15.0)        # configs[1]: the metric's configuration, the default
    assert B.WORKLOADS["hard_b32_o32_1M"] == (0x000, 1_000_000, 5.5)          # configs[0]
    assert B.WORKLOADS["s8_f16_o16_256M"][:2] == (0x122, 256_000_000) and B.WORKLOADS["s16_f16_o16_256M"][0] == 0x123   # configs[2]
    assert B.WORKLOADS["f_b32_o32_4G"][:2] == (0x004, 4_000_000_000)          # configs[3]
    assert B.WORKLOADS["config5"][:2] == (0x012, 256_000_000)                 # configs[4]
    a = B.workload_config("s4_b16_o32_32M", 0x011, 32_000_000, 15.0, 1)
    assert a["workload"] == "s4_b16_o32_32M" and "model" not in a and a["segments"] == 6400


def test_clock_summary_flags_throttle_reasons():
    B = _bench()
    cs = B.ClockSampler(0)
    cs.source, cs.t_mark, cs.t_unmark = "nvml", 10.0, 11.0
    row = lambda mhz, sw_cap: ["0", str(mhz), "1965.0", "300.0", "0x0", "Not Active", "Not Active", "Not Active", "Active" if sw_cap else "Not Active"]
    cs.rows = [(9.5, row(1500, False)), (10.2, row(1965, False)), (10.5, row(1950, True)), (10.8, row(1965, False)), (12.0, row(1200, False))]
    s = cs.summary()
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0           # median of the samples INSIDE the timed region
    assert s["samples_in_timed_region"] == 3 and s["samples_under_load"] == 5
    assert s["reasons"] == ["sw_power_cap"]


def test_same_run_reference_baseline_never_raises():
    """bench.py's `reference_cuda` entry (the reference's CUDA decoder timed in the same run, in a child process): without
    a device it reports why it is unavailable instead of costing the line"""
    B = _bench()
    old = os.environ.get("CUDA_VISIBLE_DEVICES")
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    try:
        r = B.reference_cuda_baseline("hard_b32_o32_1M", reps=3)
    finally:
        if old is None:
            del os.environ["CUDA_VISIBLE_DEVICES"]
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = old
    assert "unavailable" in r and "golden model" in r["unavailable"]
    assert "unavailable" in B.reference_cuda_baseline("no_such_workload")
