// Host thread pool of the job runner (o2v_job.cpp): persistent threads that run index-parallel loops — expanding
// downloaded occupancy bitmaps into Voxel32 records, staging pageable triangle arrays into pinned memory.  The GPU does
// the voxelization; these threads only move bytes the PCIe link would otherwise carry less compactly.
//
// Reference counterpart: the worker threads of src/obj2voxel.cpp:957-1003, which pull 64^3-chunk commands from a ring
// buffer.  Here the chunks are voxelized on the device and the host threads take the part that is host work by nature.
#ifndef O2V_POOL_H
#define O2V_POOL_H

#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace o2v {

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

    /// Helps, then blocks until every index has run.
    void wait()
    {
        help();
        if (count_ == 0) {
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
            std::lock_guard<std::mutex> lock{mutex_};
            loops_.push_back(loop);
        }
        wake_.notify_all();
        return loop;
    }

    /// fn(i) for i in [0, count) on the pool and the calling thread; returns when all are done.
    void parallelFor(size_t count, std::function<void(size_t)> fn) { start(count, std::move(fn))->wait(); }

private:
    void run()
    {
        for (;;) {
            std::shared_ptr<ParallelLoop> loop;
            {
                std::unique_lock<std::mutex> lock{mutex_};
                wake_.wait(lock, [this] {
                    while (!loops_.empty() && loops_.front()->exhausted()) {
                        loops_.erase(loops_.begin());
                    }
                    return stopping_ || !loops_.empty();
                });
                if (stopping_) {
                    return;
                }
                loop = loops_.front();
            }
            loop->help();
        }
    }

    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::condition_variable wake_;
    std::vector<std::shared_ptr<ParallelLoop>> loops_;
    bool stopping_ = false;
};

}  // namespace o2v

#endif  // O2V_POOL_H
