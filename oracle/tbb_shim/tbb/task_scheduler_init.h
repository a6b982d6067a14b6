// Minimal stand-in for tbb/task_scheduler_init.h -- ORACLE BUILD ONLY.
// source/main.cpp:250-251 asks for default_num_threads() and constructs one.
// TBB_SHIM_THREADS overrides the thread count (used for 1-thread CPU baselines).
#pragma once
#include <cstdlib>
#include <thread>
namespace tbb {
class task_scheduler_init {
public:
    static int default_num_threads() {
        if (const char* e = std::getenv("TBB_SHIM_THREADS")) {
            int v = std::atoi(e);
            if (v > 0) return v;
        }
        unsigned n = std::thread::hardware_concurrency();
        return n ? static_cast<int>(n) : 1;
    }
    explicit task_scheduler_init(int n = -1) { active_threads() = n > 0 ? n : default_num_threads(); }
    static int& active_threads() {
        static int n = default_num_threads();
        return n;
    }
};
}  // namespace tbb
