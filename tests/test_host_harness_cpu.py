"""CPU end-to-end run of the host-side C++ (host/main.cpp, viterbiDF.h, dataflow.h, viterbi.h): the ./main-style harness is
run with a stand-in for the C ABI (tests/host/fake_cabi.c, test scaffolding: decodes with the golden model) preloaded in
front of libvitb200.so.  What is under test is everything ABOVE the boundary: flag parsing, the source | encoder | noise |
packer | decoder pipeline of the re-authored dataflow runtime, probes, status printing, the BER loop and the repeat path.
The product itself has no CPU path; the same harness runs against the real library in the -m gpu suite."""
import os
import re
import subprocess

import pytest

from vit_testlib import PKG_DIR, ROOT

HOST = os.path.join(PKG_DIR, "host")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


@pytest.fixture(scope="module")
def fake(tmp_path_factory, O):
    so = tmp_path_factory.mktemp("fake") / "libfake_cabi.so"
    subprocess.check_call([GCC, "-O1", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "host", "fake_cabi.c"),
                           "-L", os.path.join(ROOT, "oracle"), "-lvitoracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    exe = os.path.join(HOST, "main")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST])

    def run(*args):
        env = dict(os.environ, LD_PRELOAD=str(so), CUDA_VISIBLE_DEVICES="")
        return subprocess.run([exe] + list(args), capture_output=True, text=True, env=env, timeout=300)
    return run


@pytest.mark.parametrize("args", [
    ["-n", "200000", "-s", "5.5", "-m", "b32", "-i", "h", "--seed", "1"],                       # BASELINE configs[0] shape
    ["-n", "300000", "-i", "s4", "-m", "b16", "-o", "b32", "--seed", "2"],
    ["-n", "300000", "-i", "s8", "-m", "f16", "-o", "b16", "-s", "3", "--seed", "3", "--prbs", "--reps", "2"],
    ["-n", "250000", "-i", "f", "-m", "b32", "-c", "dpx", "--seed", "4"],
    ["-n", "250000", "-i", "s16", "-m", "b32", "-o", "b16", "-c", "dpxt", "--seed", "5"],
])
def test_harness_end_to_end_on_the_cpu(fake, args):
    r = fake(*args)
    assert r.returncode == 0, r.stderr[-800:]
    assert "Pipeline executed." in r.stdout
    assert int(re.search(r"BEN: (\d+)", r.stdout).group(1)) == 0, r.stdout[-400:]
    n = int(args[1])
    bpp = 16 if "-o" in args and args[args.index("-o") + 1] == "b16" else 32
    assert int(re.search(r"decoded: (\d+) bits", r.stdout).group(1)) == (n - 64) // bpp * bpp


def test_harness_verbose_status_block(fake):
    r = fake("-n", "150000", "-i", "s4", "-m", "b16", "--seed", "7", "-v")
    assert r.returncode == 0, r.stderr[-800:]
    out = r.stdout
    assert "Input Channel Type: 4-bit Soft Decision" in out and "Metric Type: 16-bit" in out
    assert "--- Pipeline Status ---" in out and "--- End of Status ---" in out
    assert len(re.findall(r"^Element \d+ \(type: ", out, flags=re.M)) == 5          # source, encoder, noise, packer, decoder
    assert len(re.findall(r"^  - Elapsed run time: [\d.]+ (us|ms|s)$", out, flags=re.M)) == 5
    assert re.search(r"^  - GPU kernel time: ", out, flags=re.M) and re.search(r"^  - Decoded Gb/s: ", out, flags=re.M)


def test_harness_at_a_noisy_point_counts_errors(fake):
    """1 dB: the decoder must make errors, and the harness must count them (BER in a plausible range), not report zero."""
    r = fake("-n", "200000", "-s", "1", "-i", "s8", "-m", "b16", "--seed", "9")
    assert r.returncode == 0
    ben = int(re.search(r"BEN: (\d+)", r.stdout).group(1))
    assert 0 < ben < 200000 // 10


@pytest.mark.parametrize("args", [
    ["-n", "1000000", "-s", "5.5", "-m", "b32", "-i", "h", "--seed", "1"],            # BASELINE.json configs[0], as ./main runs it
    ["-n", "250000", "-i", "s8", "-m", "f16", "-o", "b16", "-s", "3", "--seed", "3", "--prbs"],
])
def test_harness_over_the_real_host_code_and_kernel_source(sim_lib_path, args):
    """The same harness over the library's OWN host code and kernel source: tests/sim/libvitsim.so (csrc/vit_api.cu compiled
    for the host + stand-in CUDA runtime + the kernel source in the emulator) preloaded in front of libvitb200.so.  Everything
    but the CUDA runtime and the GPU is the product: flags, pipeline, C++ shim, C ABI, vit_run, the 6400-segment kernel."""
    exe = os.path.join(HOST, "main")
    env = dict(os.environ, LD_PRELOAD=sim_lib_path, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-800:]
    assert int(re.search(r"BEN: (\d+)", r.stdout).group(1)) == 0, r.stdout[-400:]
    n = int(args[1])
    bpp = 16 if "-o" in args and args[args.index("-o") + 1] == "b16" else 32
    assert int(re.search(r"decoded: (\d+) bits", r.stdout).group(1)) == (n - 64) // bpp * bpp
