// host_pool.h -- a small persistent pool of host threads for the data-movement paths (the gather into
// pinned staging of apop_b200_pin, the text reader).  Spawning 16 threads per 64 MB chunk cost a
// third of the pin time of a 26 GB data set; the workers here are created once and parked on a
// condition variable between jobs.
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace apb {

class HostPool {
public:
    static HostPool& get() {
        static HostPool* p = new HostPool();   // lives for the process: workers must not be joined at exit
        return *p;
    }
    int size() const { return n_; }
    // runs fn(i) for i in [0, n), a contiguous share per worker; returns when all are done.  Not reentrant.
    void run(size_t n, const std::function<void(size_t)>& fn) {
        if (n <= 1 || n_ <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; total_ = n; next_ = 0; gen_++;
        }
        cv_.notify_all();
        work();                                // the caller takes items too
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return next_ >= total_ && busy_items_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        const char* e = getenv("APOP_B200_PIN_THREADS");
        n_ = e ? atoi(e) : 16;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && n_ > hw) n_ = hw;
        if (n_ < 1) n_ = 1;
        for (int t = 0; t < n_ - 1; t++) std::thread([this] { loop(); }).detach();
    }
    void work() {
        for (;;) {
            size_t i;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (!fn_ || next_ >= total_) return;
                i = next_++; busy_items_++;
            }
            (*fn_)(i);
            std::lock_guard<std::mutex> lk(m_);
            busy_items_--;
            if (next_ >= total_ && busy_items_ == 0) done_.notify_all();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
            }
            work();
        }
    }
    int n_ = 1;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t total_ = 0, next_ = 0, busy_items_ = 0;
    unsigned long long gen_ = 0;
};

}  // namespace apb
