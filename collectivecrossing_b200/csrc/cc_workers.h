// cc_workers.h — the host threads of a handle (cc_set_host_expand): a fixed crew that sleeps between calls, so that a
// cc_step_host of a few hundred microseconds does not pay for creating threads.  Host C++ only; not part of the ABI.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

class cc_worker_pool {
  public:
    cc_worker_pool() = default;
    cc_worker_pool(const cc_worker_pool &) = delete;
    cc_worker_pool &operator=(const cc_worker_pool &) = delete;
    ~cc_worker_pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++round_;
        }
        wake_.notify_all();
        for (std::thread &t : threads_) t.join();
    }
    // fn(0) on the calling thread, fn(1) .. fn(n-1) on pool threads; returns when all have returned
    void run(int n, const std::function<void(int)> &fn) {
        if (n <= 1) { fn(0); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            while ((int)threads_.size() < n - 1) {
                const int idx = (int)threads_.size() + 1;
                threads_.emplace_back(&cc_worker_pool::loop, this, idx, round_);
            }
            fn_ = &fn;
            n_ = n;
            pending_ = n - 1;
            ++round_;
        }
        wake_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void loop(int idx, uint64_t seen) {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            wake_.wait(lk, [&] { return round_ != seen; });
            seen = round_;
            if (stop_) return;
            if (idx >= n_) continue;   // a smaller job than the crew
            const std::function<void(int)> *fn = fn_;
            lk.unlock();
            (*fn)(idx);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex m_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, pending_ = 0;
    uint64_t round_ = 0;
    bool stop_ = false;
};
