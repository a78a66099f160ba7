"""CPU tests of the host-side C++ mirror (gpu-accelerated-viterbi-decoder_b200/host): the shim header exposes
the same option enums, masks, static constants and typedefs as the reference's viterbi.h (compiled side by
side when /root/reference is present), and a caller written against the reference interface builds with a
plain host compiler."""
import os
import subprocess
import sys

import pytest

from vit_testlib import PKG_DIR, ROOT

HOST = os.path.join(PKG_DIR, "host")
REF = "/root/reference/src/viterbi"
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
INC = os.path.join(ROOT, "tests", "host", "print_constants.inc")


def build_and_run(tmp_path, name, header_dir, valid_macro, compiler):
    src = tmp_path / (name + ".cu" if compiler[0] == "nvcc" else name + ".cpp")
    src.write_text('#include "viterbi.h"\n#define VIT_TEST_VALID(O) (%s)\n#include "%s"\n' % (valid_macro, INC))
    exe = tmp_path / name
    cmd = compiler + ["-std=c++17", "-I", header_dir, "-o", str(exe), str(src)]
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    return subprocess.check_output([str(exe)], text=True)


def test_shim_constants_equal_reference_header(tmp_path):
    if not os.path.isdir(REF):
        pytest.skip("reference sources not present on this machine")
    ours = build_and_run(tmp_path, "ours", HOST, "OptionsValid<O>::reference_value", [GXX])
    # the reference header includes cuda_fp16.h -> needs nvcc (host code only, no kernels)
    ref = build_and_run(tmp_path, "ref", REF, "OptionsValid<O>::value", ["nvcc", "-arch=sm_100"])
    o_lines, r_lines = ours.strip().splitlines(), ref.strip().splitlines()
    assert len(o_lines) == len(r_lines) == 61
    for a, b in zip(o_lines, r_lines):
        # metric_t is __half (2 bytes) in the reference and uint16_t in the CUDA-free shim: same size
        assert a == b, (a, b)


def test_shim_accepts_superset_of_reference_options(tmp_path):
    out = build_and_run(tmp_path, "ours2", HOST, "OptionsValid<O>::value", [GXX])
    valid = {int(l.split()[0]): int(l.split()[1].split("=")[1]) for l in out.strip().splitlines()[:60]}
    for o, v in valid.items():
        it, mt = o & 0xF, o & 0xF0
        assert v == (0 if (mt == 0x10 and it == 3) else 1), hex(o)


def test_reference_style_caller_builds_with_host_compiler(tmp_path):
    """A translation unit written against the reference's interface (ViterbiDecoder in a Pipeline, as
    reference src/main.cpp:131-142) compiles against our headers with g++ only."""
    src = tmp_path / "caller.cpp"
    src.write_text('''
#include "viterbiDF.h"
int main(int argc, char**) {
    constexpr int options = ChannelIn::SOFT4 | Metric::M_B16 | DecodeOut::O_B32 | CompMode::REG;
    static_assert(OptionsValid<options>::value, "");
    using decVec_t = typename ViterbiDecoder<options>::decVec_t;
    if (argc > 100) {   // never executed here (no GPU): this test is about the interface
        RandBitGen randGen(1000, 0);
        ConvolutionalEncoder convEnc(ViterbiCUDA<options>::constLen, ViterbiCUDA<options>::polyn1, ViterbiCUDA<options>::polyn2);
        AddNoise noise(0.1f, 1);
        SoftDecisionPacker packer(ViterbiCUDA<options>::inputType, 40000.0);
        ViterbiDecoder<options> viterbi;
        Pipeline pipe = randGen.probe() | convEnc | noise | packer | viterbi;
        PipelineResult result = pipe.run();
        pipe.printStatus();
        decVec_t out = std::any_cast<decVec_t>(result.final_output);
        Bits gen = std::any_cast<Bits>(result.probed_outputs[0]);
        ViterbiCUDA<options> dec(2000);
        float ms;
        std::vector<typename ViterbiCUDA<options>::encPack_t> in(dec.getInputSize(2000) / 4);
        dec.run(in.data(), out.data(), 2000, &ms);
        return (int)(out.size() + gen.size() + dec.getMessageLen(2000) + dec.getOutputSize(2000));
    }
    return 0;
}
''')
    exe = tmp_path / "caller"
    subprocess.check_call([GXX, "-std=c++17", "-Wall", "-I", HOST, "-o", str(exe), str(src), "-L", PKG_DIR, "-lvitb200",
                           "-Wl,-rpath," + PKG_DIR])
    assert subprocess.run([str(exe)]).returncode == 0


def test_harness_flags_without_gpu():
    exe = os.path.join(HOST, "main")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST])
    out = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("-n, --num", "-s, --snr", "-i, --input", "-m, --metric", "-o, --output", "-c, --compMode", "-v, --verbose"):
        assert flag in out.stdout
    bad = subprocess.run([exe, "-i", "nope"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Invalid" in bad.stderr
    bad = subprocess.run([exe, "--bogus"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Unknown or incomplete argument" in bad.stderr
    bad = subprocess.run([exe, "-i", "s16", "-m", "b16"], capture_output=True, text=True)
    assert bad.returncode != 0 and "16-bit metric does not support 16-bit soft decision input" in bad.stderr


def test_shim_refuses_a_library_built_for_another_code(tmp_path):
    """ViterbiCUDA<>::polyn1/polyn2 follow -DVIT_POLY1/-DVIT_POLY2 (csrc/vit_code.h), and constructing a decoder checks the
    linked library's vit_code_parameters against them BEFORE any CUDA call: a caller compiled for the reference's code
    refuses the (0117, 0155) library in the reference's print-and-exit style; compiled with the matching polynomials it
    gets past the check (and then fails on the missing GPU, here)."""
    from vit_testlib import ALT_LIB, ALT_POLYS
    if not os.path.exists(ALT_LIB):
        pytest.skip("variant library not built")
    src = tmp_path / "code.cpp"
    src.write_text('''
#include "viterbi.h"
int main() {
    std::printf("%d %o %o\\n", ViterbiCUDA<0>::constLen, ViterbiCUDA<0>::polyn1, ViterbiCUDA<0>::polyn2);
    std::fflush(stdout);
    ViterbiCUDA<0> dec;
    return 0;
}
''')
    libname = os.path.basename(ALT_LIB)[3:-3]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")          # no device for either binary: only the check itself differs
    for flags, constants, message in (([], "7 171 133", "built for other code parameters"),
                                      (["-DVIT_POLY1=0%o" % ALT_POLYS[0], "-DVIT_POLY2=0%o" % ALT_POLYS[1]], "7 117 155", None)):
        exe = tmp_path / ("code%d" % len(flags))
        subprocess.check_call([GXX, "-std=c++17", "-Wall", "-I", HOST] + flags + ["-o", str(exe), str(src), "-L", PKG_DIR,
                               "-l" + libname, "-Wl,-rpath," + PKG_DIR])
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
        assert r.stdout.strip() == constants
        assert r.returncode == 1                              # EXIT_FAILURE either way on a machine without a GPU
        if message:
            assert message in r.stderr
        else:
            assert "built for other code parameters" not in r.stderr and " at line " in r.stderr
