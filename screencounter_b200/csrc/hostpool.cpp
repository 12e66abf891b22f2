#include "hostpool.hpp"

#include <unistd.h>

#include <algorithm>

namespace scg {

namespace {
HostPool* g_pool = nullptr;
pid_t g_owner = 0;
std::mutex g_guard;
} // namespace

HostPool& HostPool::instance() {
    std::lock_guard<std::mutex> lock(g_guard);
    const pid_t me = getpid();
    if (!g_pool || g_owner != me) {
        // first use in this process (the parent's pool, if any, has no threads here: it is abandoned, not destroyed)
        g_pool = new HostPool();
        g_owner = me;
    }
    return *g_pool;
}

void HostPool::ensure(int workers) {
    while ((int)workers_.size() < workers) {
        workers_.emplace_back([this] { worker_loop(); });
        workers_.back().detach();
    }
}

void HostPool::worker_loop() {
    unsigned long long seen = 0;
    std::unique_lock<std::mutex> lock(mutex_);
    for (;;) {
        wake_.wait(lock, [&] { return generation_ != seen && job_ && next_ < total_ && running_ < allowed_; });
        seen = generation_;
        ++running_;
        while (job_ && next_ < total_) {
            const int k = next_++;
            const std::function<void(int)>* job = job_;
            lock.unlock();
            (*job)(k);
            lock.lock();
        }
        --running_;
        if (running_ == 0) done_.notify_all();
    }
}

void HostPool::parallel_for(int n, int threads, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) {
        for (int k = 0; k < n; ++k) fn(k);
        return;
    }
    std::lock_guard<std::mutex> entry(entry_);
    std::unique_lock<std::mutex> lock(mutex_);
    ensure(threads - 1);
    job_ = &fn;
    next_ = 0;
    total_ = n;
    allowed_ = threads - 1;
    ++generation_;
    wake_.notify_all();
    // the caller works too
    while (next_ < total_) {
        const int k = next_++;
        lock.unlock();
        fn(k);
        lock.lock();
    }
    done_.wait(lock, [&] { return running_ == 0; });
    job_ = nullptr;
}

} // namespace scg
