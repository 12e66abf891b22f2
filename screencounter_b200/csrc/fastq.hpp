// Host FASTQ reader + packer.  Replaces the reference's byte-at-a-time kaori::FastqReader
// (inst/include/kaori/FastqReader.hpp:42-110) over byteme::PerByte with a chunked,
// memchr-driven record splitter and a multi-threaded packer that writes the tile-planar
// 2-bit + N-mask layout (layout.hpp) straight into pinned memory.  Grammar, line numbering
// and error texts follow the reference.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "common.hpp"
#include "layout.hpp"

namespace scg {

struct Record {
    const char* seq;     // first sequence byte (inside the reader's buffer)
    uint32_t span;       // bytes from seq up to (not including) the '+' that ends the sequence
    uint32_t len;        // bases = span minus embedded newlines
};

class FastqInput;  // raw file (mmap), gzip stream or caller memory

// Scratch of one chunk of the multi-threaded record splitter.
struct ParseChunk {
    std::vector<Record> recs;
    uint32_t min_len = 0xFFFFFFFFu, max_len = 0;
    size_t end = 0;
    bool okay_after = true, failed = false;
};

class FastqReader {
public:
    FastqReader(const char* path, const char* data, size_t size);
    ~FastqReader();

    // Parses up to max_records further records.  The returned records point into an internal
    // buffer that stays valid until the next call.  Empty result = end of input.
    const std::vector<Record>& next(size_t max_records);

    // Threads the record splitter may use on inputs that are entirely in memory (caller's buffer, mmap'd file).
    void set_threads(int n);

    // shortest / longest read of the batch returned by the last next()
    uint32_t batch_min_len() const { return batch_min_len_; }
    uint32_t batch_max_len() const { return batch_max_len_; }

    // The whole input as one span of host memory, for inputs that are (caller's buffer, mmap'd raw file) and have not
    // been read from yet: the device-side reader (ingest.hpp) takes the text from there.
    bool memory_text(const char** data, size_t* size) const;
    // A block-gzip input that has not been read from yet: its member index (the device-side reader inflates the members itself).
    // [*text_begin, *text_end): the part of the text this reader delivers (the whole of it unless set_text_range() narrowed it).
    bool bgzf_image(const struct BgzfIndex** index, size_t* text_begin = nullptr, size_t* text_end = nullptr) const;
    // Block-gzip inputs only, before the first read: deliver the bytes [begin, end) of the text, both record boundaries.
    void set_text_range(size_t begin, size_t end);
    // Continues the host parse at byte `offset` of such an input -- the start of record number `nrecords` (0-based),
    // everything before it having been consumed elsewhere as four-line records.
    void resume_at(size_t offset, long long nrecords);

    long long records_seen() const { return nrecords_; }
    double parse_seconds() const { return parse_s_; }

private:
    bool parse_one(Record& out);  // false = needs more data (or clean end of input)
    bool next_parallel(size_t max_records);
    void refill();

    std::unique_ptr<FastqInput> in_;
    const char* base_ = nullptr;  // current window
    size_t avail_ = 0;            // bytes in the window
    size_t pos_ = 0;              // parse position inside the window
    bool final_ = false;          // window reaches the end of the input
    bool okay_ = true;            // reference's `okay` flag (FastqReader.hpp:44,93-99)
    bool started_ = false;
    long long nrecords_ = 0;
    std::vector<Record> batch_;
    double parse_s_ = 0;
    int threads_ = 1;
    std::vector<ParseChunk> chunks_;
    uint32_t batch_min_len_ = 0xFFFFFFFFu, batch_max_len_ = 0;
};

// A likely record start at or after byte `from` of a FASTQ text, or (size_t)-1: a guess that the caller verifies by parsing
// what precedes it (fastq.cpp).
size_t guess_fastq_record_start(const char* text, size_t size, size_t from);

// Packs records [first, first+count) into `out` (tile-planar, W words per plane, count padded
// up to a multiple of 32 with empty reads).  lens receives the read lengths.  odd (nullable)
// receives 1 for reads holding a character other than upper-case A, C, G, T, N.
void pack_records(const Record* recs, size_t count, int W, uint32_t* out, uint16_t* lens, uint8_t* odd, int nthreads);

// Scalar reference implementation of the packer for one read (also used for pool rows).
void pack_read_scalar(const char* seq, uint32_t span, int W, uint32_t* h, uint32_t* l, uint32_t* n, size_t stride);

} // namespace scg
