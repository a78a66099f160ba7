// dataflow.h -- minimal synchronous dataflow runtime with the reference's interface
// (reference src/dataflow/dataflow.h:1-133): ComputeElement / Pipeline / PipelineResult / operator|.
// Re-authored; values move through the pipeline instead of being copied between elements.
#pragma once

#include <any>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <map>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

using OptData = std::optional<std::any>;

class ComputeElement {
public:
    virtual ~ComputeElement() = default;

    // in == nullopt: the element is a source and makes its own data
    virtual std::any process(const OptData& in) = 0;

    ComputeElement& probe() { probed_ = true; return *this; }
    bool isProbed() const { return probed_; }

    void setStatus(const std::string& key, std::any value) { status[key] = std::move(value); }
    std::any getStatus(const std::string& key) const { return status.at(key); }
    const std::map<std::string, std::any>& getStatusMap() const { return status; }

    virtual std::string getStatusString(const std::string&) const { return "(Not printable)"; }

    std::string getStatusStringAll(const std::string& key) const {
        if (key != "Elapsed run time") return getStatusString(key);
        const auto us = std::any_cast<std::chrono::microseconds>(getStatus(key)).count();
        std::ostringstream os;
        if (us > 1000000) os << std::fixed << std::setprecision(2) << us / 1e6 << " s";
        else if (us > 1000) os << std::fixed << std::setprecision(2) << us / 1e3 << " ms";
        else os << us << " us";
        return os.str();
    }

protected:
    std::map<std::string, std::any> status;

private:
    bool probed_ = false;
};

struct PipelineResult {
    std::any final_output;
    std::vector<std::any> probed_outputs;
};

class Pipeline {
public:
    Pipeline& add(ComputeElement& e) { stages_.push_back(&e); return *this; }

    PipelineResult run() {
        PipelineResult res;
        OptData cur;
        for (ComputeElement* e : stages_) {
            const auto t0 = std::chrono::high_resolution_clock::now();
            cur = e->process(cur);
            const auto t1 = std::chrono::high_resolution_clock::now();
            e->setStatus("Elapsed run time", std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0));
            if (e->isProbed()) res.probed_outputs.push_back(*cur);
        }
        if (!cur.has_value()) throw std::runtime_error("Pipeline produced no output");
        res.final_output = std::move(*cur);
        return res;
    }

    void printStatus() const {
        std::cout << "--- Pipeline Status ---\n";
        int idx = 0;
        for (const ComputeElement* e : stages_) {
            std::cout << "Element " << idx++ << " (type: " << typeid(*e).name() << "):\n";
            if (e->getStatusMap().empty()) std::cout << "  - No status information.\n";
            for (const auto& kv : e->getStatusMap())
                std::cout << "  - " << kv.first << ": " << e->getStatusStringAll(kv.first) << "\n";
        }
        std::cout << "--- End of Status ---\n";
    }

private:
    std::vector<ComputeElement*> stages_;
};

inline Pipeline operator|(ComputeElement& a, ComputeElement& b) { Pipeline p; p.add(a).add(b); return p; }
inline Pipeline operator|(Pipeline p, ComputeElement& b) { p.add(b); return p; }
