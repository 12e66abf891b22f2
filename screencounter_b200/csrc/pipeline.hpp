// FASTQ -> pinned staging -> HBM, double-buffered.  While the GPU works on batch k the host parses
// and packs batch k+1 into the other staging slot.
#pragma once

#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "engine.hpp"
#include "ingest.hpp"

namespace scg {

class ReadPipeline {
public:
    struct Batch {
        long long first_read = 0;   // global index of the batch's first read (pair)
        long long n = 0;
        ReadsDev reads1, reads2;    // reads2 only for paired input
        const uint8_t* odd1 = nullptr;   // device flags: read holds characters other than ACGTN (when asked for)
        const Record* recs1 = nullptr;   // host records of mate 1 (valid until the next call)
        void* slot = nullptr;
    };

    ReadPipeline(Context& ctx, FastqReader* r1, FastqReader* r2, int nthreads, bool want_odd);
    ~ReadPipeline();

    // Stages the next batch on the context's stream.  false = input exhausted.
    bool next(Batch& out);
    // Call after the kernels that read the batch have been enqueued.
    void submitted(Batch& b);

    // Reads of the batch flagged as holding characters other than ACGTN (callers that asked for the flags), or -1 when only
    // the device knows (batches of the device-side reader).
    long long count_odd(const Batch& b) const;
    // Raw characters of read `index` of the batch (mate 1), newlines of wrapped records removed.
    void raw_read(const Batch& b, long long index, std::string& seq);

private:
    static constexpr int kSlots = Context::kStagingSlots;
    static constexpr size_t kMaxBatchReads = 1u << 20;
    static constexpr size_t kMaxBatchBytes = 64u << 20;
    using Slot = StagingSlot;

    void stage(Slot& slot, int mate, const Record* recs, size_t count, uint32_t minlen, uint32_t maxlen);

    Context& ctx_;
    FastqReader* r1_;
    FastqReader* r2_;
    int nthreads_;
    bool want_odd_;
    Slot* slots_;   // the context's staging slots
    int next_slot_ = 0;
    const Record* recs1_ = nullptr;
    const Record* recs2_ = nullptr;
    size_t n1_ = 0, cur1_ = 0, n2_ = 0, cur2_ = 0;
    long long consumed_ = 0;
    std::unique_ptr<DeviceIngest> ingest_[2];   // set while the device-side reader is feeding the batches (one per mate)
    bool handover_pending_ = false;
    bool next_device(Batch& out, bool& ended);   // false = the device readers have handed over to the host readers
};

} // namespace scg
