// profiler.cuh — opt-in per-kernel timing with CUDA events on the launching stream, and an
// always-on launch counter.  bench.py uses it to report which kernel dominates a step and
// its achieved throughput without running under a profiler; off by default (two event
// records per launch when on).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace nrms {

struct Profiler {
    struct Rec {
        std::string name;
        cudaEvent_t a, b;
    };
    // suffix appended to the names recorded by this thread ("" or "@user": the user encoder's launches
    // of the shared kernels are reported separately from the news encoder's)
    static const char*& scope() {
        static thread_local const char* s = "";
        return s;
    }
    std::mutex mu;
    bool on = false;
    long long launches = 0;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;

    static Profiler& get() {
        static Profiler p;
        return p;
    }
    cudaEvent_t ev() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(const char* name, cudaStream_t s) {
        std::lock_guard<std::mutex> g(mu);
        ++launches;
        if (!on) return;
        Rec r{std::string(name) + scope(), ev(), ev()};
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    }
    void end(cudaStream_t s) {
        std::lock_guard<std::mutex> g(mu);
        if (!on || recs.empty()) return;
        cudaEventRecord(recs.back().b, s);
    }
    // synchronises; returns "name count total_ms\n" lines sorted by name and clears the log
    std::string collect() {
        std::lock_guard<std::mutex> g(mu);
        std::map<std::string, std::pair<long long, double>> agg;
        for (auto& r : recs) {
            cudaEventSynchronize(r.b);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            auto& a = agg[r.name];
            a.first += 1;
            a.second += ms;
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
        std::string out;
        char line[256];
        for (auto& kv : agg) {
            snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first,
                     kv.second.second);
            out += line;
        }
        return out;
    }
};

struct ProfileScope {
    const char* prev;
    explicit ProfileScope(const char* s) : prev(Profiler::scope()) { Profiler::scope() = s; }
    ~ProfileScope() { Profiler::scope() = prev; }
};

#define NRMS_LAUNCH(name, stream, ...)                \
    do {                                              \
        ::nrms::Profiler::get().begin(name, stream);  \
        __VA_ARGS__;                                  \
        ::nrms::Profiler::get().end(stream);          \
    } while (0)

}  // namespace nrms
