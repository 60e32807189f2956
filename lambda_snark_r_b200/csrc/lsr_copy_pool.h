// lsr_copy_pool.h -- a few host threads that copy between caller memory and page-locked staging buffers.
//
// The batched host-pointer entry points accept any host memory.  From ordinary (pageable) memory -- a Rust Vec<u64>,
// a numpy array -- cudaMemcpyAsync is staged by the driver on the calling thread at ~10 GB/s, five times below
// PCIe 5 x16, and it blocks the caller meanwhile.  The staged path of lwe_commit_host copies chunk i+1 in and chunk i-3
// out with this pool while the GPU works on the chunks in between.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace lsr {

class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    // blocking parallel memcpy; calls from different threads are serialised
    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < (1u << 20) || workers_.empty()) { std::memcpy(dst, src, bytes); return; }
        std::lock_guard<std::mutex> call(call_mu_);
        const size_t parts = workers_.size() + 1;
        const size_t slice = (((bytes + parts - 1) / parts) + 4095) & ~(size_t)4095;     // parts * slice >= bytes
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = static_cast<char*>(dst); src_ = static_cast<const char*>(src); bytes_ = bytes; slice_ = slice;
            pending_ = workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        run_slice(parts - 1);                                   // the caller takes the last slice
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 4 ? (hw / 2 > 8 ? 8 : hw / 2) : 0;   // 0 workers on tiny hosts: plain memcpy
        if (n) n -= 1;                                          // the caller is one of the copiers
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run_slice(size_t idx) {
        const size_t lo = idx * slice_;
        if (lo >= bytes_) return;
        const size_t len = bytes_ - lo < slice_ ? bytes_ - lo : slice_;
        std::memcpy(dst_ + lo, src_ + lo, len);
    }
    void loop(unsigned idx) {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            run_slice(idx);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0, pending_ = 0, generation_ = 0;
    bool stop_ = false;
};

}  // namespace lsr
