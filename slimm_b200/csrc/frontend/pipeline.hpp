// pipeline.hpp - the two building blocks of the host decode pipeline: a bounded queue and an ordered
// parallel stage (N worker threads, results handed on in submission order).
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace slimm_fe {

template <class T>
class BoundedQueue {
public:
    explicit BoundedQueue(size_t cap) : cap_(cap ? cap : 1) {}
    void push(T &&v)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    bool pop(T &out)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close()
    {
        std::lock_guard<std::mutex> lk(m_);
        closed_ = true;
        not_empty_.notify_all();
        not_full_.notify_all();
    }

private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// push(In) from one thread, pop(Out) from one thread; fn runs on n_threads workers.  At most max_inflight
// items are between push and pop, which bounds the memory the stage holds.
template <class In, class Out>
class OrderedStage {
public:
    OrderedStage(int n_threads, size_t max_inflight, std::function<void(In &, Out &)> fn)
        : fn_(std::move(fn)), max_inflight_(max_inflight ? max_inflight : 1)
    {
        for (int i = 0; i < (n_threads > 0 ? n_threads : 1); ++i) workers_.emplace_back([this] { work(); });
    }
    ~OrderedStage()
    {
        abort();
        for (auto &t : workers_) t.join();
    }
    void push(In &&in)
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_space_.wait(lk, [&] { return pushed_ - popped_ < max_inflight_ || aborted_; });
        if (aborted_) return;
        todo_.emplace_back(pushed_++, std::move(in));
        cv_work_.notify_one();
    }
    void close()
    {
        std::lock_guard<std::mutex> lk(m_);
        closed_ = true;
        cv_work_.notify_all();
        cv_done_.notify_all();
    }
    void abort()
    {
        std::lock_guard<std::mutex> lk(m_);
        aborted_ = closed_ = true;
        cv_work_.notify_all();
        cv_done_.notify_all();
        cv_space_.notify_all();
    }
    // hands a consumed result back: its buffers go to the next job instead of to the allocator (a result of this pipeline is
    // megabytes of arrays; fresh ones cost a page fault per 4 KB, and the threads' faults serialise in the kernel)
    void recycle(Out &&out)
    {
        std::lock_guard<std::mutex> lk(m_);
        if (free_.size() < max_inflight_ + workers_.size()) free_.push_back(std::move(out));
    }
    bool pop(Out &out)
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return done_.count(popped_) || (closed_ && popped_ == pushed_) || aborted_; });
        auto it = done_.find(popped_);
        if (it == done_.end()) return false;
        out = std::move(it->second);
        done_.erase(it);
        ++popped_;
        cv_space_.notify_one();
        return true;
    }

private:
    void work()
    {
        for (;;) {
            std::pair<uint64_t, In> job;
            Out out;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return !todo_.empty() || closed_; });
                if (todo_.empty() || aborted_) return;
                job = std::move(todo_.front());
                todo_.pop_front();
                if (!free_.empty()) { out = std::move(free_.back()); free_.pop_back(); }   // a recycled result: fn_ resets it
            }
            fn_(job.second, out);
            {
                std::lock_guard<std::mutex> lk(m_);
                done_.emplace(job.first, std::move(out));
                cv_done_.notify_all();
            }
        }
    }
    std::function<void(In &, Out &)> fn_;
    size_t max_inflight_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_, cv_space_;
    std::deque<std::pair<uint64_t, In>> todo_;
    std::map<uint64_t, Out> done_;
    std::vector<Out> free_;
    uint64_t pushed_ = 0, popped_ = 0;
    bool closed_ = false, aborted_ = false;
    std::vector<std::thread> workers_;
};

}  // namespace slimm_fe
