// Shim for BS::thread_pool 4.1.0 (only what src/simulation.cpp:721,740-746 uses):
// thread_pool(n), detach_loop<T>(first, last, f) = static block partition, wait().
#pragma once
#include <cstddef>
#include <thread>
#include <vector>
namespace BS {
class thread_pool {
public:
    explicit thread_pool(std::size_t n) : n_(n ? n : 1) {}
    template <typename T, typename F>
    void detach_loop(T first, T last, F &&f) {
        const T total = last - first;
        const T nblk = static_cast<T>(n_) < total ? static_cast<T>(n_) : (total ? total : 1);
        for (T b = 0; b < nblk; ++b) {
            const T lo = first + total * b / nblk, hi = first + total * (b + 1) / nblk;
            workers_.emplace_back([lo, hi, &f] { for (T i = lo; i < hi; ++i) f(i); });
        }
    }
    void wait() { for (auto &t : workers_) t.join(); workers_.clear(); }
    ~thread_pool() { wait(); }
private:
    std::size_t n_;
    std::vector<std::thread> workers_;
};
}  // namespace BS
