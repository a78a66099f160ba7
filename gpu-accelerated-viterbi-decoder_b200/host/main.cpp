// main.cpp -- bench harness with the reference ./main's flags and semantics
// (reference src/main.cpp:14-264: -n/-s/-i/-m/-o/-c/-v/-h; sigma = 10^(-snr/5); quantiser scale 40000;
// BER = mismatches / messageLen with decoded bit i compared to generated bit i+extraL).
// Additions: reproducible seeds (--seed, --prbs), size_t message lengths, decoded Gb/s from device-side
// kernel time, repeated timed runs (--reps), a device-side channel source (--device-source) and the multi-GPU
// stream job (--streams S --gpus G [--gather nccl|copy|direct|none] [--wave W] [--batch B]): S independent
// codeword streams of -n bits each, generated on the GPUs, sharded over G GPUs in contiguous blocks (one host
// thread, one decoder and one communicator per GPU, no stream is ever split), the packed output bits gathered
// to GPU 0 over NCCL / NVLink, timed on the device, with and without the gather (BASELINE.json configs[4]).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "viterbiDF.h"

struct Args {
    size_t messageLen = 32000000;   // main.cpp:176
    float snr = 15.0f;              // main.cpp:177
    int options = 0;                // HARD | B32 | O_B32 | REG, main.cpp:178
    bool verbose = false;
    bool prbs = false;
    long seed = -1;                 // < 0: std::random_device, as the reference (main.cpp:131-135)
    int reps = 1;
    int streams = 1;
    int gpus = 1;
    bool deviceSource = false;      // generate, decode and count errors on the device (no host pipeline)
    std::string gather = "direct";  // how the stream job brings the packed output bits to GPU 0
    int wave = 8;                   // streams per decode launch
    int batch = 64;                 // streams generated ahead of each timed decode phase
};

static void usage(const char* prog) {
    std::cout << "Usage: " << prog << " [options]\n"
              << "Options:\n"
              << "  -n, --num <integer>      Set the message length.\n"
              << "  -s, --snr <float>        Set the Signal-to-Noise Ratio (SNR).\n"
              << "  -i, --input <type>       Set the input channel type (HARD|h, SOFT4|s4, SOFT8|s8, SOFT16|s16, FP32|f).\n"
              << "  -m, --metric <type>      Set the metric type (b16, b32, f16).\n"
              << "  -o, --output <type>      Set the output type (b16, b32).\n"
              << "  -c, --compMode <type>    Set the computation mode (REG|reg, DPX|dpx; as in the reference both run the\n"
              << "                           same core.  DPXT|dpxt: the tie rule of the reference's dormant DPX code).\n"
              << "  -v, --verbose            Enable verbose output.\n"
              << "      --seed <integer>     Fixed seed for bits (noise uses seed+1); default: random_device.\n"
              << "      --prbs               PRBS-31 message bits instead of mt19937.\n"
              << "      --reps <integer>     Timed decoder runs (best kernel time is reported).\n"
              << "      --streams <integer>  Stream job: independent streams of -n bits, generated on the device,\n"
              << "                           sharded over --gpus in contiguous blocks, outputs gathered to GPU 0.\n"
              << "      --gpus <integer>     Number of GPUs (one decoder, one communicator and one host thread per GPU).\n"
              << "      --gather <mode>      direct (the decode kernel stores into GPU 0's buffer over NVLink, default) |\n"
              << "                           copy (copy engines per finished wave) | nccl (ncclSend/ncclRecv) | none.\n"
              << "      --wave <integer>     Streams per decode launch (default 8).\n"
              << "      --batch <integer>    Streams generated ahead of each timed decode phase (default 64).\n"
              << "      --device-source      Generate the channel, decode and count bit errors on the device\n"
              << "                           (counter-based source; needed for multi-Gbit -n, e.g. -n 4000000000 -i f).\n"
              << "  -h, --help               Display this help message.\n";
}

static int lookup(const std::string& flag, const std::string& v, std::initializer_list<std::pair<const char*, int>> table) {
    for (const auto& kv : table) if (v == kv.first) return kv.second;
    std::cerr << "Error: Invalid value '" << v << "' for " << flag << "." << std::endl;
    std::exit(1);
}

static Args parseArg(int argc, char* argv[]) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string f = argv[i];
        auto value = [&]() -> std::string {
            if (i + 1 >= argc) { std::cerr << "Error: Unknown or incomplete argument: " << f << std::endl; std::exit(1); }
            return argv[++i];
        };
        auto number = [&](auto conv) {
            const std::string v = value();
            try { return conv(v); }
            catch (const std::exception&) { std::cerr << "Error: Invalid argument for " << f << "." << std::endl; std::exit(1); }
        };
        if (f == "-h" || f == "--help") { usage(argv[0]); std::exit(0); }
        else if (f == "-n" || f == "--num") a.messageLen = number([](const std::string& s) { return static_cast<size_t>(std::stoull(s)); });
        else if (f == "-s" || f == "--snr") a.snr = number([](const std::string& s) { return std::stof(s); });
        // each field REPLACES its bits (the reference ORs, so repeating a flag corrupts options: main.cpp:224-232)
        else if (f == "-m" || f == "--metric")
            a.options = (a.options & ~METRIC_MASK) | lookup(f, value(), {{"b16", M_B16}, {"b32", M_B32}, {"f16", M_FP16}});
        else if (f == "-i" || f == "--input")
            a.options = (a.options & ~CHANNEL_MASK) | lookup(f, value(), {{"HARD", HARD}, {"h", HARD}, {"SOFT4", SOFT4}, {"s4", SOFT4},
                        {"SOFT8", SOFT8}, {"s8", SOFT8}, {"SOFT16", SOFT16}, {"s16", SOFT16}, {"FP32", FP32}, {"f", FP32}});
        else if (f == "-o" || f == "--output")
            a.options = (a.options & ~DECODE_MASK) | lookup(f, value(), {{"b16", O_B16}, {"b32", O_B32}});
        else if (f == "-c" || f == "--compMode")
            a.options = (a.options & ~COMP_MASK) | lookup(f, value(), {{"REG", REG}, {"reg", REG}, {"DPX", DPX}, {"dpx", DPX}, {"DPXT", DPX_TIES}, {"dpxt", DPX_TIES}});
        else if (f == "-v" || f == "--verbose") a.verbose = true;
        else if (f == "--prbs") a.prbs = true;
        else if (f == "--device-source") a.deviceSource = true;
        else if (f == "--seed") a.seed = number([](const std::string& s) { return std::stol(s); });
        else if (f == "--reps") a.reps = std::max(1, number([](const std::string& s) { return std::stoi(s); }));
        else if (f == "--streams") a.streams = std::max(1, number([](const std::string& s) { return std::stoi(s); }));
        else if (f == "--gpus") a.gpus = std::max(1, number([](const std::string& s) { return std::stoi(s); }));
        else if (f == "--wave") a.wave = std::max(1, number([](const std::string& s) { return std::stoi(s); }));
        else if (f == "--batch") a.batch = std::max(1, number([](const std::string& s) { return std::stoi(s); }));
        else if (f == "--gather") { a.gather = value(); lookup(f, a.gather, {{"nccl", 1}, {"copy", 2}, {"direct", 3}, {"none", 0}}); }
        else { std::cerr << "Error: Unknown or incomplete argument: " << f << std::endl; std::exit(1); }
    }
    return a;
}

struct Outcome { size_t ben = 0, decoded = 0; double best_ms = 0; };

template <int options>
Outcome runPipeline(const Args& a) {
    using Dec = ViterbiCUDA<options>;
    using decVec_t = typename ViterbiDecoder<options>::decVec_t;
    constexpr int bitsPerPack = Dec::bitsPerPack;

    std::random_device rd;
    const unsigned seedBits = a.seed < 0 ? rd() : static_cast<unsigned>(a.seed);
    const unsigned seedNoise = a.seed < 0 ? rd() : static_cast<unsigned>(a.seed + 1);
    RandBitGen randGen(a.messageLen, seedBits);
    PrbsBitGen prbsGen(a.messageLen, seedBits);
    ConvolutionalEncoder convEnc(Dec::constLen, Dec::polyn1, Dec::polyn2);
    AddNoise noise(static_cast<float>(std::pow(10.0, -a.snr / 5.0)), seedNoise);   // main.cpp:135
    SoftDecisionPacker packer(Dec::inputType, 40000.0f);                            // main.cpp:137
    ViterbiDecoder<options> viterbi;

    ComputeElement& source = a.prbs ? static_cast<ComputeElement&>(prbsGen) : static_cast<ComputeElement&>(randGen);
    Pipeline front = source.probe() | convEnc | noise | packer;
    Pipeline pipe = front;
    pipe.add(viterbi.probe());
    PipelineResult result = pipe.run();
    if (a.verbose) { std::cout << std::endl; pipe.printStatus(); std::cout << std::endl; }

    Outcome out;
    const decVec_t& decoded = std::any_cast<const decVec_t&>(result.final_output);
    const Bits& gen = std::any_cast<const Bits&>(result.probed_outputs[0]);
    out.decoded = decoded.size() * bitsPerPack;
    for (size_t i = 0; i < out.decoded; ++i) {                                       // main.cpp:153-169
        const bool d = (decoded[i / bitsPerPack] >> (bitsPerPack - 1 - i % bitsPerPack)) & 1u;
        out.ben += d != (gen[i + Dec::extraL] == Bit::ON);
    }
    out.best_ms = std::any_cast<float>(viterbi.getStatus("GPU kernel time"));

    // timed repeats on the packed channel words of this stream
    if (a.reps > 1) {
        using encPack_t = typename Dec::encPack_t;
        Pipeline regen = front;   // same seeds -> same stream
        const auto packed = std::any_cast<std::vector<encPack_t>>(regen.run().final_output);
        const size_t inputNum = packed.size() * Dec::encDataPerPack;
        Dec dec(inputNum, 0);
        decVec_t o(dec.getOutputSize(inputNum) / sizeof(typename Dec::decPack_t));
        for (int r = 0; r < a.reps; ++r) {
            float ms = 0.f;
            dec.run(const_cast<encPack_t*>(packed.data()), o.data(), inputNum, &ms);
            out.best_ms = std::min<double>(out.best_ms, ms);
        }
    }
    return out;
}

// Device-resident variant of runPipeline: vit_synth_device -> ViterbiCUDA::runDevice -> vit_count_errors_device.
// Same channel model shape as the host chain (BPSK +- amp, additive noise with sd = 10^(-snr/5) * amp, the
// reference's saturating quantiser and packing) but counter based, so nothing of size n ever lives on the host.
template <int options>
Outcome runDevicePipeline(const Args& a) {
    using Dec = ViterbiCUDA<options>;
    Dec dec;
    const size_t per = Dec::encDataPerPack;
    const size_t inputNum = 2 * a.messageLen;
    const size_t inBytes = (dec.getInputSize(inputNum) + 4 * per + 255) / 256 * 256;
    const size_t outBytes = dec.getOutputSize(inputNum);
    void *in_d = nullptr, *bits_d = nullptr, *out_d = nullptr;
    VIT_HANDLE_ERROR(vit_dev_alloc(&in_d, inBytes));
    VIT_HANDLE_ERROR(vit_dev_alloc(&bits_d, a.messageLen + 64));
    VIT_HANDLE_ERROR(vit_dev_alloc(&out_d, outBytes + 256));
    const unsigned seed = a.seed < 0 ? std::random_device{}() : static_cast<unsigned>(a.seed);
    const double sigma = std::pow(10.0, -a.snr / 5.0);
    VIT_HANDLE_ERROR(vit_synth_device(Dec::inputType >> CHANNEL_SHIFT, a.messageLen, seed, 0, sigma, 0, in_d, bits_d, nullptr));
    VIT_HANDLE_ERROR(vit_dev_sync());
    Outcome out;
    out.best_ms = 1e30;
    for (int r = 0; r < std::max(1, a.reps); ++r) {
        float ms = 0.f;
        dec.runDevice(in_d, out_d, inputNum, nullptr, &ms);
        out.best_ms = std::min<double>(out.best_ms, ms);
    }
    out.decoded = dec.getMessageLen(inputNum);
    unsigned long long errs = 0;
    VIT_HANDLE_ERROR(vit_count_errors_device(options, out_d, bits_d, out.decoded, &errs, nullptr));
    out.ben = errs;
    vit_dev_free(in_d); vit_dev_free(bits_d); vit_dev_free(out_d);
    return out;
}

// The multi-GPU stream job (BASELINE.json configs[4]): everything below the flags is the C ABI's vit_comm_* / vit_job_*
// (csrc/vit_mg.cu); this harness only spawns one host thread per GPU and reduces the per-GPU device times.
struct JobPass { double box_ms = 0, decode_ms = 0, synth_ms = 0; unsigned long long errors = 0, worst = 0, bits = 0, launches = 0; bool ok = true; std::string err; };

static JobPass runJobPass(const Args& a, int gatherMode, std::vector<vit_comm*>& comms) {
    JobPass pass;
    std::vector<vit_job_result> res(a.gpus);
    std::vector<std::string> errs(a.gpus);
    std::vector<std::thread> workers;
    const unsigned seed = a.seed < 0 ? 1u : static_cast<unsigned>(a.seed);
    for (int g = 0; g < a.gpus; ++g)
        workers.emplace_back([&, g] {
            vit_job_config cfg{};
            cfg.options = a.options; cfg.n_bits = a.messageLen; cfg.nstreams = static_cast<unsigned>(a.streams);
            cfg.wave = static_cast<unsigned>(a.wave); cfg.batch = static_cast<unsigned>(a.batch); cfg.seed = seed;
            cfg.source = a.prbs ? 1 : 0; cfg.amp = 0; cfg.sigma = std::pow(10.0, -a.snr / 5.0); cfg.gather = gatherMode; cfg.root = 0;
            vit_job* job = nullptr;
            int rc = vit_job_create(&job, a.gpus > 1 ? comms[g] : nullptr, g, &cfg);
            if (rc == VIT_OK) rc = vit_job_run(job, &res[g]);
            if (rc != VIT_OK) errs[g] = vit_last_error();
            vit_job_destroy(job);
        });
    for (auto& w : workers) w.join();
    for (int g = 0; g < a.gpus; ++g) {
        if (!errs[g].empty()) { pass.ok = false; pass.err = "GPU " + std::to_string(g) + ": " + errs[g]; return pass; }
        pass.box_ms = std::max(pass.box_ms, res[g].job_ms);
        pass.decode_ms = std::max(pass.decode_ms, res[g].decode_ms);
        pass.synth_ms = std::max(pass.synth_ms, res[g].synth_ms);
        pass.errors += res[g].bit_errors; pass.worst = std::max(pass.worst, res[g].max_stream_errors);
        pass.bits += res[g].decoded_bits; pass.launches += res[g].launches;
    }
    return pass;
}

static int runStreamJob(const Args& a) {
    if (!vit_options_valid(a.options)) { std::cerr << "Error: unsupported option combination." << std::endl; return -1; }
    if (vit_dev_count() < a.gpus) { std::cerr << "Error: " << a.gpus << " GPUs requested, " << vit_dev_count() << " present." << std::endl; return -1; }
    const int mode = a.gather == "nccl" ? VIT_GATHER_NCCL : a.gather == "copy" ? VIT_GATHER_COPY : a.gather == "direct" ? VIT_GATHER_DIRECT : VIT_GATHER_NONE;
    std::vector<vit_comm*> comms(a.gpus, nullptr);
    if (a.gpus > 1) VIT_HANDLE_ERROR(vit_comm_init_all(comms.data(), a.gpus, nullptr));
    std::cout << "Stream job: " << a.streams << " streams x " << a.messageLen << " bits, options 0x" << std::hex << a.options << std::dec
              << ", " << a.gpus << " GPU(s), wave " << a.wave << ", batch " << a.batch << ", gather " << a.gather
              << (a.gpus > 1 ? ", NCCL " + std::to_string(vit_comm_nccl_version()) : std::string()) << std::endl;
    int rc = 0;
    for (int which = 0; which < (mode == VIT_GATHER_NONE ? 1 : 2); ++which) {
        const int m = which == 0 ? mode : VIT_GATHER_NONE;
        const JobPass p = runJobPass(a, m, comms);
        if (!p.ok) { std::cerr << "Error: " << p.err << std::endl; rc = -1; break; }
        std::cout << (m == VIT_GATHER_NONE ? "without gather" : "with gather (" + a.gather + ")") << " -> box time: " << p.box_ms
                  << " ms (decode " << p.decode_ms << " ms, max over GPUs; generation " << p.synth_ms << " ms not counted)   decoded: "
                  << p.bits << " bits   " << p.bits / (p.box_ms * 1e6) << " Gb/s   launches: " << p.launches << "   BEN: " << p.errors
                  << "   BER: " << static_cast<double>(p.errors) / static_cast<double>(p.bits) << "   worst stream BEN: " << p.worst << std::endl;
    }
    for (auto* c : comms) vit_comm_destroy(c);
    return rc;
}

// runtime options -> template instantiation (the reference: 60 nested-macro cases, main.cpp:79-104)
template <int options>
bool tryRun(const Args& a, Outcome& out) {
    if constexpr (OptionsValid<options>::value) {
        if (a.options == options) { out = a.deviceSource ? runDevicePipeline<options>(a) : runPipeline<options>(a); return true; }
    }
    return false;
}
template <int... I>
bool dispatch(const Args& a, Outcome& out, std::integer_sequence<int, I...>) {
    // index = in + 5*(met + 3*(out + 2*cmp))
    return (tryRun<((I % 5) << CHANNEL_SHIFT) | (((I / 5) % 3) << METRIC_SHIFT) | (((I / 15) % 2) << DECODE_SHIFT) |
                   ((I / 30) << COMP_SHIFT)>(a, out) || ...);      // I / 30: REG, DPX, DPX_TIES
}

int main(int argc, char* argv[]) {
    const Args a = parseArg(argc, argv);
    const int in = a.options & CHANNEL_MASK, met = a.options & METRIC_MASK;
    if (met == M_B16 && in == SOFT16) {                                               // main.cpp:26-29
        std::cerr << "Error: 16-bit metric does not support 16-bit soft decision input." << std::endl;
        return -1;
    }
    if (a.verbose) {
        static const char* inNames[] = {"Hard Decision", "4-bit Soft Decision", "8-bit Soft Decision", "16-bit Soft Decision", "32-bit Floating Point"};
        std::cout << "Message Length: " << a.messageLen << "\nSNR: " << a.snr << " dB\n"
                  << "Input Channel Type: " << inNames[in >> CHANNEL_SHIFT] << "\n"
                  << "Metric Type: " << (met == M_B16 ? "16-bit" : met == M_B32 ? "32-bit" : "FP16") << "\n"
                  << "Output Type: " << ((a.options & DECODE_MASK) == O_B16 ? "16-bit" : "32-bit") << "\n"
                  << "Computation Mode: " << ((a.options & COMP_MASK) == REG ? "Regular" : (a.options & COMP_MASK) == DPX ? "DPX" : "DPX tie rule") << "\n" << std::endl;
    }
    if (a.streams > 1 || a.gpus > 1) return runStreamJob(a);
    Outcome out;
    if (!dispatch(a, out, std::make_integer_sequence<int, 90>{})) {
        std::cerr << "Error: unsupported option combination." << std::endl;
        return -1;
    }
    std::cout << "Pipeline executed." << std::endl;
    if (a.deviceSource)
        std::cout << "(device source: counter-hash bits, noise = sum of four uniforms (Irwin-Hall), truncated at +-3.46 sigma -- a\n"
                     " bit-reproducible load generator; its BER is a decode check, not a BER-vs-SNR measurement)" << std::endl;
    std::cout << "Final results -> BEN: " << out.ben << "   BER: " << static_cast<double>(out.ben) / a.messageLen << std::endl;  // main.cpp:107-110
    std::cout << "Decoder -> kernel time: " << out.best_ms << " ms   decoded: " << out.decoded << " bits   "
              << out.decoded / (out.best_ms * 1e6) << " Gb/s";
    std::cout << std::endl;
    return 0;
}
