// vit_stage_pool.h -- host-only helpers of vit_run's pageable-buffer path (csrc/vit_api.cu): a pool of worker threads that
// copy pageable host memory into the pinned staging buffers block by block, and the non-temporal copy they use.
// No CUDA here: tests/host/stage_pool_stress.cpp builds this header alone (under ThreadSanitizer) and hammers it.
#pragma once

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace vit_host {

// Copy into a pinned staging buffer with non-temporal stores: the destination is read next by the GPU's DMA engine, not
// by this core, so write-allocating its cache lines (what a plain memcpy of a few KB does) only adds memory traffic.
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(char* dst, const char* src, size_t n) {
    const size_t head = std::min(n, (size_t)((32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31));
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    for (; i + 32 <= n; i += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
    if (i < n) memcpy(dst + i, src + i, n - i);
}
#endif
inline void stage_copy(char* dst, const char* src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("VIT_STAGE_PLAIN_MEMCPY") == nullptr;
    if (avx2 && n >= 256) { stream_copy_avx2(dst, src, n); return; }
#endif
    memcpy(dst, src, n);
}
inline void stage_fence() {
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

// The workers sleep between calls (begin_call / end_call bracket one vit_run) and spin between the jobs of one call: a
// condition-variable wake-up per column block would cost more than the block's copy.
//
// parallel_for publishes a job as ONE atomic ticket word (generation << 32 | next item) plus a limit word stamped with the
// same generation; a worker claims item k by CAS on the ticket and only for the generation it has seen in both words, so a
// worker that is still looking at an exhausted earlier job can neither claim an item of the next one early nor read its
// function or count before they are published.
class StagePool {
public:
    explicit StagePool(int nthreads) {
        for (int i = 0; i < nthreads; i++) th_.emplace_back([this] { worker(); });
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    StagePool(const StagePool&) = delete;
    StagePool& operator=(const StagePool&) = delete;
    int size() const { return (int)th_.size() + 1; }
    void begin_call() {
        { std::lock_guard<std::mutex> lk(m_); active_.store(true); }
        cv_.notify_all();
    }
    void end_call() { active_.store(false); }
    // fn(i) for i in [0, n), n < 2^32; the caller takes part; returns when all items are done.  One caller at a time.
    void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
        if (n == 0) return;
        const unsigned long long g = gen_.load(std::memory_order_relaxed) + 1;
        fn_.store(&fn, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        limit_.store((g << 32) | (unsigned long long)n, std::memory_order_release);
        next_.store(g << 32, std::memory_order_release);
        gen_.store(g, std::memory_order_release);
        drain(g);
        while (done_.load(std::memory_order_acquire) < n) spin_pause();
    }

private:
    static void spin_pause() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void drain(unsigned long long g) {
        for (;;) {
            unsigned long long v = next_.load(std::memory_order_acquire);
            if ((v >> 32) != g) return;                                  // a later job has been published: not ours
            const unsigned long long lim = limit_.load(std::memory_order_acquire);
            if ((lim >> 32) != g || v >= lim) return;                    // exhausted (or the limit already belongs to a later job)
            if (!next_.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) continue;
            (*fn_.load(std::memory_order_relaxed))((size_t)(v & 0xffffffffull));     // job g is still running: it waits for this item
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || active_.load(); });
                if (stop_) return;
            }
            unsigned idle = 0;
            for (;;) {
                const unsigned long long g = gen_.load(std::memory_order_acquire);
                if (g != seen) { seen = g; drain(g); idle = 0; continue; }
                if (!active_.load(std::memory_order_acquire)) break;
                for (int k = 0; k < 64 && gen_.load(std::memory_order_acquire) == seen; k++) spin_pause();
                if (++idle > 4096) { std::this_thread::yield(); idle = 0; }     // be polite if the host is oversubscribed
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
    std::atomic<bool> active_{false};
    std::atomic<const std::function<void(size_t)>*> fn_{nullptr};
    std::atomic<unsigned long long> gen_{0}, next_{0}, limit_{0};
    std::atomic<size_t> done_{0};
};

}  // namespace vit_host
