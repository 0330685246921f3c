// rm_pool.h -- the library's host threads (pure C++).
//
// The reference reassembles a frame single-threaded (engine/src/renderer.rs:92-108: one copy per pixel into
// frame.buffer[y][x]).  Behind the C ABI the same step -- busy tiles from the pinned staging buffer into the caller's
// rows, black pixels zero-filled, optionally widened to the reference's f64 -- is memory-bound host work that is spread
// over a persistent pool: workers sleep on a condition variable between frames, a frame wakes them once, items are
// handed out from an atomic counter (the same scheme the reference's Rayon loop uses for its patches).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rm {

// Threads for a piece of host work: RM_B200_HOST_THREADS when set, else the hardware's, at most `cap`.
inline int host_thread_count(int cap) {
    int n = (int)std::thread::hardware_concurrency();
    if (const char* env = std::getenv("RM_B200_HOST_THREADS")) n = std::atoi(env);
    return n < 1 ? 1 : (n > cap ? cap : n);
}

class HostPool {
public:
    explicit HostPool(int n_threads) {
        const int n = n_threads < 1 ? 1 : n_threads;
        for (int i = 0; i + 1 < n; i++) workers_.emplace_back([this, i] { loop(i + 1); });   // the caller is thread 0
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
            epoch_++;
            hot_.store(epoch_, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int threads() const { return (int)workers_.size() + 1; }

    // fn(i) for every i in [0, n), on the caller's thread and the workers (at most `limit` threads in all: a light job --
    // scattering a few megabytes -- gains nothing beyond eight, and threads that are not needed should sleep: on a box
    // whose every core spins, the caller's own next steps wait for a core); returns when every item is done.
    // Not re-entrant (one frame at a time: callers hold the library mutex).
    void run(int n, const std::function<void(int)>& fn, int limit = 1 << 30) {
        if (n <= 0) return;
        limit_ = limit;
        if (workers_.empty() || n == 1 || limit <= 1) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn;
            n_ = n;
            next_.store(0, std::memory_order_relaxed);
            pending_.store((int)workers_.size(), std::memory_order_relaxed);
            epoch_++;
            hot_.store(epoch_, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        // the workers are at most one item behind (tens of microseconds): spin -- going to sleep here costs more than that
        // to wake up from -- and only sleep if something holds a worker up for milliseconds
        const auto t0 = std::chrono::steady_clock::now();
        for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; spins++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(5)) {
                std::unique_lock<std::mutex> l(mu_);
                done_cv_.wait(l, [this] { return pending_.load(std::memory_order_acquire) == 0; });
                break;
            }
        }
    }

private:
    void work() {
        for (;;) {
            const int i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_) break;
            (*fn_)(i);
        }
    }
    void loop(const int id) {
        unsigned long long seen = 0;
        bool worked = false;
        for (;;) {
            // A frame hands the pool several jobs microseconds apart (clear, then scatter chunk by chunk): a thread that
            // worked on the last one stays awake for a while -- a sleeping thread takes tens of microseconds to come back --
            // then sleeps until the next frame.
            if (worked) {
                const auto t0 = std::chrono::steady_clock::now();
                for (int spins = 0; hot_.load(std::memory_order_acquire) == seen; spins++) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                    if ((spins & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
                }
            }
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            worked = id < limit_;
            if (worked) work();
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> l(mu_);
                done_cv_.notify_one();
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, limit_ = 1 << 30;
    std::atomic<int> next_{0}, pending_{0};
    std::atomic<unsigned long long> hot_{0};   // copy of epoch_ the workers poll without the mutex
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};

}  // namespace rm
