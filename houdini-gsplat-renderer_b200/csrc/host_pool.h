// host_pool.h — a small persistent host thread pool for the cold path of libgsplat_b200 (registration uploads).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace gsb {

inline int host_threads()
{
    static int n = 0;
    if (!n) { unsigned h = std::thread::hardware_concurrency(); n = (int)std::min(16u, std::max(1u, h)); }
    return n;
}

// a small persistent pool (threads are created once per process, parked on a condition variable): the cold path issues
// ~100 slot copies per 2.6 GB registration and must not pay a thread spawn for each
class HostPool {
public:
    static HostPool& get() { static HostPool p; return p; }
    // f(begin, end) on disjoint ranges of [0, n), on the pool's threads and the caller; returns when all are done
    template <class F> void ranges(size_t n, size_t grain, F&& f)
    {
        const int parts = (int)std::min<size_t>((size_t)workers_.size() + 1, (n + grain - 1) / std::max<size_t>(grain, 1));
        if (parts <= 1) { f((size_t)0, n); return; }
        std::function<void(int)> job = [&](int k) { f(n * (size_t)k / (size_t)parts, n * (size_t)(k + 1) / (size_t)parts); };
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &job; parts_ = parts; next_ = 1; pending_ = parts - 1; ++generation_;
        }
        cv_.notify_all();
        job(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    HostPool()
    {
        const int n = host_threads() - 1;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void loop()
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> l(m_);
        for (;;) {
            cv_.wait(l, [&] { return stop_ || (generation_ != seen && job_ && next_ < parts_); });
            if (stop_) return;
            while (job_ && next_ < parts_) {
                const int k = next_++;
                std::function<void(int)>* j = job_;
                l.unlock();
                (*j)(k);
                l.lock();
                if (--pending_ == 0) done_.notify_all();
            }
            seen = generation_;
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(int)>* job_ = nullptr;
    int parts_ = 0, next_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};


}  // namespace gsb
