// Minimal stand-in for tbb/parallel_for.h -- ORACLE BUILD ONLY (test infrastructure).
// Hands out sub-ranges of `grainsize` items to a std::thread pool through one
// atomic cursor; the reference calls it once, over image rows with grain 1
// (source/main.cpp:329-331).  Results do not depend on the thread count because
// every row seeds its own RNG (source/main.cpp:204).
#pragma once
#include <atomic>
#include <thread>
#include <vector>
#include "blocked_range.h"
#include "task_scheduler_init.h"
namespace tbb {
template <typename T, typename Body>
void parallel_for(const blocked_range<T>& range, const Body& body) {
    const T grain = static_cast<T>(range.grainsize() ? range.grainsize() : 1);
    std::atomic<long long> cursor(static_cast<long long>(range.begin()));
    const long long end = static_cast<long long>(range.end());
    auto worker = [&]() {
        for (;;) {
            long long b = cursor.fetch_add(static_cast<long long>(grain));
            if (b >= end) break;
            long long e = b + static_cast<long long>(grain);
            if (e > end) e = end;
            body(blocked_range<T>(static_cast<T>(b), static_cast<T>(e), static_cast<std::size_t>(grain)));
        }
    };
    int n = task_scheduler_init::active_threads();
    if (n <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    pool.reserve(n - 1);
    for (int i = 1; i < n; ++i) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}
}  // namespace tbb
