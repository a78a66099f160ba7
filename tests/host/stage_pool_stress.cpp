// stage_pool_stress.cpp -- CPU stress of csrc/vit_stage_pool.h (the worker pool behind vit_run's pageable-buffer path), meant
// to be built with -fsanitize=thread: many back-to-back parallel_for calls of varying size and varying function objects (as
// run_gated issues them: one per column block, then one for the output), inside begin_call / end_call brackets with sleeps
// between them; every item of every call must run exactly once, inside its call, with that call's function.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../gpu-accelerated-viterbi-decoder_b200/csrc/vit_stage_pool.h"

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 6;
    const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
    vit_host::StagePool pool(threads - 1);
    std::vector<std::atomic<int>> hits(4096);
    std::atomic<long long> current{-1};
    unsigned x = 12345;
    long long calls = 0, items = 0, errors = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
        pool.begin_call();
        const int jobs = 1 + (int)((x = x * 1664525u + 1013904223u) >> 29);
        for (int j = 0; j < jobs; j++) {
            x = x * 1664525u + 1013904223u;
            const size_t n = (x >> 8) % 7 == 0 ? (x >> 20) % 4 : 1 + (x >> 20) % 4095;       // tiny jobs too, and n == 0
            for (size_t i = 0; i < n; i++) hits[i].store(0, std::memory_order_relaxed);
            const long long id = calls;
            current.store(id);
            const std::function<void(size_t)> job = [&, id, n](size_t i) {
                if (current.load() != id || i >= n) __atomic_fetch_add(&errors, 1, __ATOMIC_RELAXED);   // wrong call or index
                hits[i].fetch_add(1, std::memory_order_relaxed);
            };
            pool.parallel_for(n, job);
            for (size_t i = 0; i < n; i++)
                if (hits[i].load(std::memory_order_relaxed) != 1) errors++;
            calls++; items += (long long)n;
        }
        pool.end_call();
        if ((calls & 63) == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));      // let the workers go to sleep
    }
    // the copy helpers: every length and alignment around the vector widths
    std::vector<char> src(5000), dst(5000);
    for (size_t i = 0; i < src.size(); i++) src[i] = (char)(i * 31 + 7);
    for (size_t off = 0; off < 40; off++)
        for (size_t n : {0u, 1u, 31u, 32u, 33u, 127u, 128u, 255u, 256u, 257u, 1000u, 4096u}) {
            memset(dst.data(), 0, dst.size());
            vit_host::stage_copy(dst.data() + off, src.data() + (off * 7) % 13, n);
            vit_host::stage_fence();
            if (memcmp(dst.data() + off, src.data() + (off * 7) % 13, n) != 0 || dst[off + n] != 0 || (off && dst[off - 1] != 0)) errors++;
        }
    printf("calls %lld items %lld errors %lld threads %d\n", calls, items, errors, pool.size());
    return errors ? 1 : 0;
}
