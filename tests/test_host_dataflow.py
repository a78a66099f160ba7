"""CPU tests of the host-side dataflow runtime and element library (gpu-accelerated-viterbi-decoder_b200/host/dataflow.h,
viterbiDF.h; reference src/dataflow/dataflow.h, src/viterbiDF.h:20-167): pipeline semantics (sources, probes, status,
errors, printStatus text) and the PRBS | encoder | noise | packer chain against the golden model's twins, word for word,
for every input type.  Host C++ only -- no decoder element, no GPU."""
import os
import subprocess

import numpy as np
import pytest

from vit_testlib import PKG_DIR

HOST = os.path.join(PKG_DIR, "host")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

PROGRAM = r'''
#include <cstring>
#include <fstream>
#include "viterbiDF.h"

struct Doubler : ComputeElement {           // a user element: status of its own, printable
    std::any process(const OptData& in) override {
        if (!in) throw std::runtime_error("Doubler expects input");
        std::vector<int> v = std::any_cast<std::vector<int>>(*in);
        for (int& x : v) x *= 2;
        setStatus("items", (int)v.size());
        return v;
    }
    std::string getStatusString(const std::string& key) const override {
        return key == "items" ? std::to_string(std::any_cast<int>(getStatus(key))) : ComputeElement::getStatusString(key);
    }
};
struct Counter : ComputeElement {           // a source
    std::any process(const OptData& in) override {
        if (in) throw std::runtime_error("Counter is a source");
        setStatus("opaque", 3.5);
        return std::vector<int>{1, 2, 3};
    }
};

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "semantics")) {
        Counter src; Doubler d1, d2;
        Pipeline p = src | d1.probe() | d2;
        PipelineResult r = p.run();
        auto out = std::any_cast<std::vector<int>>(r.final_output);
        auto mid = std::any_cast<std::vector<int>>(r.probed_outputs.at(0));
        std::printf("final %d %d %d probes %zu mid %d %d %d\n", out[0], out[1], out[2], r.probed_outputs.size(), mid[0], mid[1], mid[2]);
        std::printf("elapsed-is-us %d\n", (int)(src.getStatus("Elapsed run time").type() == typeid(std::chrono::microseconds)));
        p.printStatus();
        Doubler lone; Pipeline bad; bad.add(lone);
        try { bad.run(); } catch (const std::runtime_error& e) { std::printf("error: %s\n", e.what()); }
        Pipeline empty;
        try { empty.run(); } catch (const std::runtime_error& e) { std::printf("error: %s\n", e.what()); }
        try { src.getStatus("nope"); } catch (const std::out_of_range&) { std::printf("missing status throws\n"); }
        Counter c2; Doubler d3;
        c2.setStatus("Elapsed run time", std::chrono::microseconds(1234567)); std::printf("%s|", c2.getStatusStringAll("Elapsed run time").c_str());
        c2.setStatus("Elapsed run time", std::chrono::microseconds(3456)); std::printf("%s|", c2.getStatusStringAll("Elapsed run time").c_str());
        c2.setStatus("Elapsed run time", std::chrono::microseconds(12)); std::printf("%s\n", c2.getStatusStringAll("Elapsed run time").c_str());
        return 0;
    }
    // chain <input type 0..4> <n bits> <seed> <out file>: PRBS-31 | encoder | noiseless BPSK | packer (scale 3)
    const int it = std::atoi(argv[2]); const size_t n = std::strtoull(argv[3], nullptr, 10); const unsigned seed = std::atoi(argv[4]);
    PrbsBitGen bits(n, seed);
    ConvolutionalEncoder enc(ViterbiCUDA<0>::constLen, ViterbiCUDA<0>::polyn1, ViterbiCUDA<0>::polyn2);
    AddNoise channel;                        // default: no noise
    SoftDecisionPacker packer(static_cast<ChannelIn>(it), 3.0f);
    Pipeline p = bits.probe() | enc | channel | packer;
    PipelineResult r = p.run();
    std::ofstream f(argv[5], std::ios::binary);
    if (it == 4) { auto v = std::any_cast<Reals>(r.final_output); f.write((const char*)v.data(), v.size() * 4); }
    else { auto v = std::any_cast<Soft>(r.final_output); f.write((const char*)v.data(), v.size() * 4); }
    Bits b = std::any_cast<Bits>(r.probed_outputs.at(0));
    std::ofstream g(std::string(argv[5]) + ".bits", std::ios::binary);
    g.write((const char*)b.data(), b.size());
    return 0;
}
'''


@pytest.fixture(scope="module")
def prog(tmp_path_factory):
    d = tmp_path_factory.mktemp("dataflow")
    src = d / "prog.cpp"
    src.write_text(PROGRAM)
    exe = d / "prog"
    # the decoder element is a template that is never instantiated here, but the shim header declares the C ABI: link the library
    subprocess.check_call([GXX, "-O1", "-std=c++17", "-Wall", "-I", HOST, "-o", str(exe), str(src), "-L", PKG_DIR, "-lvitb200",
                           "-Wl,-rpath," + PKG_DIR])
    return str(exe)


def test_pipeline_semantics_and_status_text(prog):
    out = subprocess.check_output([prog, "semantics"], text=True).splitlines()
    assert out[0] == "final 4 8 12 probes 1 mid 2 4 6"
    assert out[1] == "elapsed-is-us 1"
    assert out[2] == "--- Pipeline Status ---"
    assert out[3].startswith("Element 0 (type: ") and "Counter" in out[3]
    # std::map order: "Elapsed run time" < "opaque"; a status without a printer shows the reference's placeholder
    assert out[4].startswith("  - Elapsed run time: ") and out[4].rstrip().endswith(("us", "ms", " s"))
    assert out[5] == "  - opaque: (Not printable)"
    assert out[6].startswith("Element 1 (type: ") and "Doubler" in out[6]
    assert out[8] == "  - items: 3"
    assert out[-5] == "--- End of Status ---"
    assert out[-4] == "error: Doubler expects input"
    assert out[-3] == "error: Pipeline produced no output"
    assert out[-2] == "missing status throws"
    assert out[-1] == "1.23 s|3.46 ms|12 us"


@pytest.mark.parametrize("it", range(5))
def test_host_chain_equals_golden_model_twins(prog, O, tmp_path, it):
    n, seed = 4096 + 37, 5
    f = tmp_path / ("chain%d.bin" % it)
    subprocess.check_call([prog, "chain", str(it), str(n), str(seed), str(f)])
    bits = np.fromfile(str(f) + ".bits", np.uint8)
    exp_bits = O.prbs31(0x7FFFFFFF ^ seed, n)
    assert np.array_equal(bits, exp_bits)
    soft = O.encode(exp_bits).astype(np.float32) * 2 - 1
    if it == 4:
        assert np.array_equal(np.fromfile(str(f), np.float32), soft * np.float32(3.0))
    else:
        per = {0: 32, 1: 8, 2: 4, 3: 2}[it]
        whole = soft[: soft.size // per * per]                # the reference's packer drops a trailing partial word
        assert np.array_equal(np.fromfile(str(f), np.int32), O.pack(it, whole, 3.0))
