// dataflow.h -- the synchronous element/pipeline runtime behind the reference's dataflow interface
// (reference src/dataflow/dataflow.h:12-133).  What a reference caller sees is kept: ComputeElement with process / probe /
// status accessors, PipelineResult, Pipeline::add / run / printStatus, `a | b | c` chaining, the "Elapsed run time" status
// and the text printStatus writes.  Behind that interface this is a different small runtime: a stage hands its value to the
// next one by move (the reference copies the std::any between elements, and again into the result), the timing and the
// duration formatting live in helpers of their own, and every element can describe itself to a stream.
#pragma once

#include <any>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <map>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

using OptData = std::optional<std::any>;

namespace dataflow_detail {

// "12 us" / "3.46 ms" / "1.20 s": the three ranges of the reference's status line (dataflow.h:54-66)
inline std::string human_time(long long microseconds) {
    char text[48];
    if (microseconds > 1000000) std::snprintf(text, sizeof text, "%.2f s", (double)microseconds / 1e6);
    else if (microseconds > 1000) std::snprintf(text, sizeof text, "%.2f ms", (double)microseconds / 1e3);
    else std::snprintf(text, sizeof text, "%lld us", microseconds);
    return text;
}

// wall time of one stage, in the unit the status map stores
class Stopwatch {
public:
    Stopwatch() : begin_(clock::now()) {}
    std::chrono::microseconds lap() const { return std::chrono::duration_cast<std::chrono::microseconds>(clock::now() - begin_); }

private:
    using clock = std::chrono::high_resolution_clock;
    clock::time_point begin_;
};

inline const char* const kElapsedKey = "Elapsed run time";

}  // namespace dataflow_detail

// One processing stage.  A source ignores its (empty) input and produces data; every other stage transforms the value
// it is given.  Subclasses publish measurements through the status map and say how to print them.
class ComputeElement {
public:
    virtual ~ComputeElement() = default;

    virtual std::any process(const OptData& in) = 0;                                   // reference dataflow.h:27
    virtual std::string getStatusString(const std::string& /*key*/) const { return "(Not printable)"; }

    // keep this stage's output in PipelineResult::probed_outputs
    ComputeElement& probe() { keep_output_ = true; return *this; }
    bool isProbed() const { return keep_output_; }

    // status map: name -> value of any type (the runtime itself files the stage's wall time under "Elapsed run time")
    void setStatus(const std::string& key, std::any value) { status.insert_or_assign(key, std::move(value)); }
    std::any getStatus(const std::string& key) const { return status.at(key); }       // throws std::out_of_range
    const std::map<std::string, std::any>& getStatusMap() const { return status; }
    std::string getStatusStringAll(const std::string& key) const {
        if (key == dataflow_detail::kElapsedKey)
            return dataflow_detail::human_time(std::any_cast<std::chrono::microseconds>(status.at(key)).count());
        return getStatusString(key);
    }

    // the block printStatus shows for this stage
    void describe(std::ostream& os, int position) const {
        os << "Element " << position << " (type: " << typeid(*this).name() << "):\n";
        if (status.empty()) os << "  - No status information.\n";
        for (const auto& entry : status) os << "  - " << entry.first << ": " << getStatusStringAll(entry.first) << "\n";
    }

protected:
    std::map<std::string, std::any> status;

private:
    bool keep_output_ = false;
};

struct PipelineResult {
    std::any final_output;
    std::vector<std::any> probed_outputs;      // one per probed stage, in pipeline order
};

// An ordered chain of stages owned by the caller (the pipeline stores pointers, as the reference's does).
class Pipeline {
public:
    Pipeline& add(ComputeElement& stage) { chain_.push_back(&stage); return *this; }

    PipelineResult run() {
        PipelineResult result;
        OptData value;                          // empty: the first stage is a source
        for (ComputeElement* stage : chain_) {
            const dataflow_detail::Stopwatch watch;
            std::any produced = stage->process(value);
            stage->setStatus(dataflow_detail::kElapsedKey, watch.lap());
            if (stage->isProbed()) result.probed_outputs.push_back(produced);          // a probe is the only copy made
            value.emplace(std::move(produced));
        }
        if (!value) throw std::runtime_error("Pipeline produced no output");
        result.final_output = std::move(*value);
        return result;
    }

    void printStatus() const {
        std::cout << "--- Pipeline Status ---\n";
        for (size_t i = 0; i < chain_.size(); i++) chain_[i]->describe(std::cout, (int)i);
        std::cout << "--- End of Status ---\n";
    }

private:
    std::vector<ComputeElement*> chain_;
};

// `source | encoder | channel | decoder`
inline Pipeline operator|(ComputeElement& first, ComputeElement& second) {
    Pipeline p;
    p.add(first).add(second);
    return p;
}
inline Pipeline operator|(Pipeline chain, ComputeElement& next) { return std::move(chain.add(next)); }
