// pool_check.cpp -- exercises rusty_marcher_b200/csrc/rm_pool.h (the library's host threads) under ThreadSanitizer:
// thousands of back-to-back jobs of varying size and thread limit, every item must run exactly once, and runs separated by
// pauses long enough for the workers to go back to sleep.  Built and run by tests/test_host_pool.py.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "../../rusty_marcher_b200/csrc/rm_pool.h"

int main() {
    rm::HostPool pool(6);
    if (pool.threads() != 6) return 2;
    std::vector<int> hits(4096);
    long long total = 0, expect = 0;
    for (int round = 0; round < 3000; round++) {
        const int n = 1 + (round * 37) % 4096;
        const int limit = (round % 5 == 0) ? 1 : (round % 5 == 1) ? 2 : (round % 5 == 2) ? 4 : 1 << 30;
        for (int i = 0; i < n; i++) hits[i] = 0;
        std::atomic<long long> sum{0};
        pool.run(n, [&](int i) {
            hits[i]++;                               // exactly one thread may touch item i
            sum.fetch_add(i, std::memory_order_relaxed);
        }, limit);
        for (int i = 0; i < n; i++)
            if (hits[i] != 1) { std::printf("round %d: item %d ran %d times\n", round, i, hits[i]); return 1; }
        if (sum.load() != (long long)n * (n - 1) / 2) { std::printf("round %d: wrong sum\n", round); return 1; }
        total += sum.load();
        expect += (long long)n * (n - 1) / 2;
        if (round % 500 == 499) std::this_thread::sleep_for(std::chrono::milliseconds(2));   // workers fall asleep in between
    }
    {
        rm::HostPool one(1);                         // no workers: the caller does everything
        int c = 0;
        one.run(100, [&](int) { c++; });
        if (c != 100) return 1;
    }
    std::printf("pool ok: %lld\n", total);
    return total == expect ? 0 : 1;
}
