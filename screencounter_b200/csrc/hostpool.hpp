// A small persistent pool of host threads for the FASTQ splitter and the packer: parallel_for(n, fn) runs
// fn(0) .. fn(n-1), the caller taking part.  Threads are created on first use and kept (creating 2 x 16 threads
// per million reads costs more than the work they do on a virtualised host).  The pool belongs to the process that
// made it: after a fork (BiocParallel::MulticoreParam) the child starts its own.
#pragma once

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace scg {

class HostPool {
public:
    static HostPool& instance();
    // Runs fn(k) for k in [0, n) on up to `threads` threads (including the calling one); returns when all are done.
    // fn must not throw.
    void parallel_for(int n, int threads, const std::function<void(int)>& fn);

private:
    HostPool() {}
    void ensure(int workers);
    void worker_loop();

    std::mutex mutex_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)>* job_ = nullptr;
    int next_ = 0, total_ = 0, running_ = 0, allowed_ = 0;
    unsigned long long generation_ = 0;
    std::mutex entry_;   // one parallel_for at a time
};

} // namespace scg
