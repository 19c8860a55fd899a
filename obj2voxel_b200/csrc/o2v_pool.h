// Host thread pool of the job runner (o2v_job.cpp): persistent threads that run index-parallel loops — expanding
// downloaded occupancy bitmaps into Voxel32 records, staging pageable triangle arrays into pinned memory.  The GPU does
// the voxelization; these threads only move bytes the PCIe link would otherwise carry less compactly.
//
// Reference counterpart: the worker threads of src/obj2voxel.cpp:957-1003, which pull 64^3-chunk commands from a ring
// buffer.  Here the chunks are voxelized on the device and the host threads take the part that is host work by nature.
#ifndef O2V_POOL_H
#define O2V_POOL_H

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace o2v {

inline void spinPause()
{
#if defined(__SSE2__)
    _mm_pause();
#endif
}

/// Polls `ready` for up to `microseconds` before the caller falls back to blocking: the loops of a job follow each other
/// within microseconds, a condition variable wakes a thread in tens of them.
template <typename Ready>
inline bool spinUntil(Ready &&ready, int microseconds)
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(microseconds);
    for (int round = 0;; ++round) {
        if (ready()) {
            return true;
        }
        if ((round & 63) == 63 && std::chrono::steady_clock::now() >= deadline) {
            return false;
        }
        spinPause();
    }
}

/// One index-parallel loop in flight: fn(i) for i in [0, count), indices handed out by an atomic counter.
class ParallelLoop {
public:
    ParallelLoop(size_t count, std::function<void(size_t)> fn) : count_(count), fn_(std::move(fn)) {}

    /// Runs indices until none is left; returns when this caller has no more to take (others may still be running).
    void help()
    {
        for (;;) {
            const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= count_) {
                return;
            }
            fn_(i);
            if (done_.fetch_add(1, std::memory_order_acq_rel) + 1 == count_) {
                std::lock_guard<std::mutex> lock{mutex_};
                finished_ = true;
                wake_.notify_all();
            }
        }
    }

    /// Helps, then waits until every index has run (polling first: the stragglers are microseconds away).
    void wait()
    {
        help();
        if (count_ == 0 || spinUntil([this] { return done_.load(std::memory_order_acquire) == count_; }, 200)) {
            return;
        }
        std::unique_lock<std::mutex> lock{mutex_};
        wake_.wait(lock, [this] { return finished_; });
    }

    bool exhausted() const { return next_.load(std::memory_order_relaxed) >= count_; }

private:
    const size_t count_;
    const std::function<void(size_t)> fn_;
    std::atomic<size_t> next_{0}, done_{0};
    std::mutex mutex_;
    std::condition_variable wake_;
    bool finished_ = false;
};

class HostPool {
public:
    explicit HostPool(unsigned threads)
    {
        for (unsigned t = 0; t < threads; ++t) {
            threads_.emplace_back([this] { run(); });
        }
    }

    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lock{mutex_};
            stopping_ = true;
        }
        wake_.notify_all();
        for (std::thread &t : threads_) {
            t.join();
        }
    }

    unsigned size() const { return (unsigned) threads_.size(); }

    /// Starts fn(i), i in [0, count), on the pool and returns at once; wait() on the result (the waiter helps).
    std::shared_ptr<ParallelLoop> start(size_t count, std::function<void(size_t)> fn)
    {
        auto loop = std::make_shared<ParallelLoop>(count, std::move(fn));
        if (count != 0) {
            {
                std::lock_guard<std::mutex> lock{mutex_};
                loops_.push_back(loop);
                epoch_.fetch_add(1, std::memory_order_release);
            }
            wake_.notify_all();
        }
        return loop;
    }

    /// fn(i) for i in [0, count) on the pool and the calling thread; returns when all are done.
    void parallelFor(size_t count, std::function<void(size_t)> fn) { start(count, std::move(fn))->wait(); }

private:
    /// The first loop that still has indices to hand out, or null.
    std::shared_ptr<ParallelLoop> take()
    {
        std::lock_guard<std::mutex> lock{mutex_};
        while (!loops_.empty() && loops_.front()->exhausted()) {
            loops_.erase(loops_.begin());
        }
        return loops_.empty() ? nullptr : loops_.front();
    }

    void run()
    {
        for (;;) {
            const unsigned long long seen = epoch_.load(std::memory_order_acquire);
            if (std::shared_ptr<ParallelLoop> loop = take()) {
                loop->help();
                continue;
            }
            // nothing to do: poll for the next loop for a while (a job issues them back to back), then sleep
            if (spinUntil([&] { return epoch_.load(std::memory_order_acquire) != seen; }, 300)) {
                continue;
            }
            std::unique_lock<std::mutex> lock{mutex_};
            wake_.wait(lock, [&] { return stopping_ || epoch_.load(std::memory_order_acquire) != seen; });
            if (stopping_) {
                return;
            }
        }
    }

    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::condition_variable wake_;
    std::vector<std::shared_ptr<ParallelLoop>> loops_;
    std::atomic<unsigned long long> epoch_{0};  // bumped by every start()
    bool stopping_ = false;
};

}  // namespace o2v

#endif  // O2V_POOL_H
