#include "fastq.hpp"

#include "bgzf.hpp"
#include "hostpool.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace scg {

// ---------------------------------------------------------------------------------------
// Inputs.  A window is a contiguous span of not-yet-consumed FASTQ text; `final` says that
// nothing follows it.
// ---------------------------------------------------------------------------------------
class FastqInput {
public:
    virtual ~FastqInput() {}
    // Makes [*base, *base + *avail) hold the unconsumed tail starting at `keep_from` (an
    // offset into the previous window) plus as much further data as is convenient.
    virtual void window(size_t keep_from, const char** base, size_t* avail, bool* final) = 0;
    // Inputs that are entirely in host memory (caller's buffer, mmap'd raw file) say where.
    virtual bool memory(const char** data, size_t* size) const {
        (void)data;
        (void)size;
        return false;
    }
    // Host threads the input may use for its own work (block-gzip inflates its blocks in parallel).
    virtual void set_threads(int n) { (void)n; }
    // Block-gzip inputs say where their members are (the device can inflate them)...
    virtual const BgzfIndex* bgzf() const { return nullptr; }
    // ... and can be positioned at a byte of their TEXT: the next window() starts there.  false = not supported.
    virtual bool seek(size_t text_offset) {
        (void)text_offset;
        return false;
    }
    // ... and can be told to end at a byte of their text (a part of a file that several devices share)
    virtual void set_text_range(size_t begin, size_t end) {
        (void)begin;
        (void)end;
    }
    virtual void text_range(size_t* begin, size_t* end) const {
        *begin = 0;
        *end = 0;
    }
};

namespace {

class MemoryInput : public FastqInput {
public:
    MemoryInput(const char* data, size_t size) : data_(data), size_(size) {}
    void window(size_t keep_from, const char** base, size_t* avail, bool* final) override {
        offset_ += keep_from;
        *base = data_ + offset_;
        *avail = size_ - offset_;
        *final = true;
    }
    bool memory(const char** data, size_t* size) const override {
        *data = data_;
        *size = size_;
        return true;
    }

protected:
    const char* data_;
    size_t size_;
    size_t offset_ = 0;
};

class MmapInput : public MemoryInput {
public:
    MmapInput(int fd, void* map, size_t size) : MemoryInput(static_cast<const char*>(map), size), fd_(fd), map_(map), mapped_(size) {}
    ~MmapInput() override {
        if (map_ && mapped_) munmap(map_, mapped_);
        if (fd_ >= 0) close(fd_);
    }

private:
    int fd_;
    void* map_;
    size_t mapped_;
};

// gzip stream (byteme::GzipFileReader, inst/include/byteme/GzipFileReader.hpp:39-51), inflated in
// large chunks; the unconsumed tail of a chunk is carried to the front of the next one.
class GzInput : public FastqInput {
public:
    explicit GzInput(const char* path) {
        gz_ = gzopen(path, "rb");
        if (!gz_) throw Error(std::string("failed to open file at '") + path + "'");
        gzbuffer(gz_, 1 << 20);
        buf_.resize(kChunk);
    }
    ~GzInput() override {
        if (gz_) gzclose(gz_);
    }
    void window(size_t keep_from, const char** base, size_t* avail, bool* final) override {
        size_t tail = filled_ - keep_from;
        if (keep_from > 0 && tail > 0) std::memmove(buf_.data(), buf_.data() + keep_from, tail);
        filled_ = tail;
        if (!eof_) {
            if (buf_.size() - filled_ < kChunk / 2) buf_.resize(buf_.size() * 2);  // a record longer than the chunk
            while (filled_ < buf_.size()) {
                size_t want = std::min<size_t>(buf_.size() - filled_, 1u << 30);
                int got = gzread(gz_, buf_.data() + filled_, (unsigned)want);
                if (got < 0) {
                    int dummy;
                    throw Error(gzerror(gz_, &dummy));
                }
                if (got == 0) {
                    eof_ = true;
                    break;
                }
                filled_ += (size_t)got;
            }
        }
        *base = buf_.data();
        *avail = filled_;
        *final = eof_;
    }

private:
    static constexpr size_t kChunk = 64u << 20;
    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t filled_ = 0;
    bool eof_ = false;
};

// Block gzip (bgzf.hpp): the member chain is indexed without inflating; the members are inflated in parallel on the host
// thread pool here (kaori goes member by member on one thread), or on the device when the device-side reader takes the
// input (ingest.cu + inflate.cu).  The text is the same; plain gzip streams keep GzInput.
class BgzfInput : public FastqInput {
public:
    // nullptr when the file is not a complete chain of BGZF members
    static std::unique_ptr<FastqInput> open(int fd, size_t size) {
        if (size < 28) return nullptr;
        void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) return nullptr;
        std::unique_ptr<BgzfInput> in(new BgzfInput(fd, map, size));
        if (!bgzf_index(static_cast<const unsigned char*>(map), size, in->index_)) {
            in->fd_ = -1;   // the caller keeps the descriptor
            return nullptr;
        }
        madvise(map, size, MADV_SEQUENTIAL);
        return std::unique_ptr<FastqInput>(in.release());
    }
    // the same over an image in the caller's memory
    static std::unique_ptr<FastqInput> open_memory(const char* data, size_t size) {
        std::unique_ptr<BgzfInput> in(new BgzfInput(-1, nullptr, 0));
        if (!bgzf_index(reinterpret_cast<const unsigned char*>(data), size, in->index_)) return nullptr;
        return std::unique_ptr<FastqInput>(in.release());
    }
    ~BgzfInput() override {
        if (map_) munmap(map_, size_);
        if (fd_ >= 0) close(fd_);
    }
    void set_threads(int n) override { threads_ = n < 1 ? 1 : n; }
    const BgzfIndex* bgzf() const override { return &index_; }
    bool seek(size_t text_offset) override {
        const size_t b = std::min(index_.block_of(text_offset), index_.blocks.size());
        next_ = b;
        filled_ = 0;
        skip_ = b < index_.blocks.size() ? text_offset - index_.text_off[b] : 0;
        return true;
    }
    void set_text_range(size_t begin, size_t end) override {
        begin_ = std::min(begin, index_.text_size());
        limit_ = std::min(std::max(end, begin_), index_.text_size());
        seek(begin_);
    }
    void text_range(size_t* begin, size_t* end) const override {
        *begin = begin_;
        *end = limit_ == (size_t)-1 ? index_.text_size() : limit_;
    }
    void window(size_t keep_from, const char** base, size_t* avail, bool* final) override {
        const std::vector<BgzfBlock>& blocks_ = index_.blocks;
        const size_t tail = filled_ - keep_from;
        if (keep_from > 0 && tail > 0) std::memmove(buf_.data(), buf_.data() + keep_from, tail);
        filled_ = tail;
        const size_t limit = limit_ == (size_t)-1 ? index_.text_size() : limit_;
        if (next_ < blocks_.size() && index_.text_off[next_] < limit) {
            // the next batch of members: about kBatch bytes of text (never beyond the member that holds the end of the range)
            size_t last = next_, bytes = 0;
            while (last < blocks_.size() && index_.text_off[last] < limit && (bytes == 0 || bytes + blocks_[last].isize <= kBatch)) {
                bytes += blocks_[last++].isize;
            }
            if (buf_.size() < filled_ + bytes) buf_.resize(filled_ + bytes);
            std::vector<size_t> at(last - next_ + 1, filled_);
            for (size_t k = next_; k < last; ++k) at[k - next_ + 1] = at[k - next_] + blocks_[k].isize;
            const int nblocks = (int)(last - next_);
            const int pieces = std::max(1, std::min(nblocks, threads_ * 4));
            std::vector<int> failed((size_t)pieces, 0);
            char* out = buf_.data();
            const size_t first = next_;
            HostPool::instance().parallel_for(pieces, std::min(pieces, threads_), [&](int p) {
                const int b = (int)((long long)nblocks * p / pieces), e = (int)((long long)nblocks * (p + 1) / pieces);
                for (int k = b; k < e; ++k) {
                    if (!bgzf_inflate_block(index_, first + (size_t)k, out + at[(size_t)k])) {
                        failed[(size_t)p] = 1;
                        break;
                    }
                }
            });
            for (int f : failed) {
                if (f) throw Error("failed to inflate the block-gzip file (corrupt member)");
            }
            filled_ = at.back();
            next_ = last;
            // text beyond the end of the range is not this reader's
            if (index_.text_off[last] > limit) filled_ -= std::min(filled_, index_.text_off[last] - limit);
            if (skip_) {
                // a reader resumed inside a member: drop the text in front of that byte
                const size_t drop = std::min(skip_, filled_);
                std::memmove(buf_.data(), buf_.data() + drop, filled_ - drop);
                filled_ -= drop;
                skip_ = 0;
            }
        }
        *base = buf_.data();
        *avail = filled_;
        *final = next_ >= blocks_.size() || index_.text_off[next_] >= limit;
    }

private:
    static constexpr size_t kBatch = 64u << 20;

    BgzfInput(int fd, void* map, size_t size) : fd_(fd), map_(map), size_(size) {}

    size_t begin_ = 0, limit_ = (size_t)-1;   // the part of the text this reader delivers

    int fd_;
    void* map_;
    size_t size_;
    int threads_ = 1;
    BgzfIndex index_;
    size_t next_ = 0;
    size_t skip_ = 0;
    std::vector<char> buf_;
    size_t filled_ = 0;
};

std::unique_ptr<FastqInput> open_input(const char* path, const char* data, size_t size) {
    if (!path) {
        // a block-gzip image in the caller's memory is read like the file it would be on disk; anything else is FASTQ text
        if (size >= 28 && (unsigned char)data[0] == 0x1f && (unsigned char)data[1] == 0x8b && !std::getenv("SCG_NO_BGZF")) {
            std::unique_ptr<FastqInput> bgzf = BgzfInput::open_memory(data, size);
            if (bgzf) return bgzf;
        }
        return std::unique_ptr<FastqInput>(new MemoryInput(data, size));
    }
    // byteme::SomeFileReader (inst/include/byteme/SomeFileReader.hpp:31-44): sniff the gzip magic.
    int fd = open(path, O_RDONLY);
    if (fd < 0) throw Error(std::string("failed to open file at '") + path + "'");
    unsigned char header[2];
    ssize_t got = read(fd, header, 2);
    if (got == 2 && header[0] == 0x1f && header[1] == 0x8b) {
        struct stat gst;
        if (fstat(fd, &gst) == 0 && gst.st_size > 0 && !std::getenv("SCG_NO_BGZF")) {
            std::unique_ptr<FastqInput> bgzf = BgzfInput::open(fd, (size_t)gst.st_size);
            if (bgzf) return bgzf;   // owns the descriptor now
        }
        close(fd);
        return std::unique_ptr<FastqInput>(new GzInput(path));
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        throw Error(std::string("failed to open file at '") + path + "'");
    }
    if (st.st_size == 0) {
        close(fd);
        return std::unique_ptr<FastqInput>(new MemoryInput(nullptr, 0));
    }
    void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) {
        close(fd);
        throw Error(std::string("failed to read raw binary file (mmap of '") + path + "' failed)");
    }
    madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
    return std::unique_ptr<FastqInput>(new MmapInput(fd, map, (size_t)st.st_size));
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

// ---------------------------------------------------------------------------------------
// Record splitter
// ---------------------------------------------------------------------------------------

FastqReader::FastqReader(const char* path, const char* data, size_t size) : in_(open_input(path, data, size)) {}

FastqReader::~FastqReader() {}

void FastqReader::set_threads(int n) {
    threads_ = n < 1 ? 1 : n;
    in_->set_threads(threads_);
}

bool FastqReader::memory_text(const char** data, size_t* size) const {
    return !started_ && in_->memory(data, size);
}

bool FastqReader::bgzf_image(const BgzfIndex** index, size_t* text_begin, size_t* text_end) const {
    if (started_ || !in_->bgzf()) return false;
    *index = in_->bgzf();
    size_t b = 0, e = 0;
    in_->text_range(&b, &e);
    if (text_begin) *text_begin = b;
    if (text_end) *text_end = e;
    return true;
}

void FastqReader::set_text_range(size_t begin, size_t end) { in_->set_text_range(begin, end); }

void FastqReader::resume_at(size_t offset, long long nrecords) {
    if (in_->seek(offset)) {
        // (block gzip: the input inflates from the member that holds the byte)
        in_->window(0, &base_, &avail_, &final_);
        started_ = true;
        pos_ = 0;
    } else {
        if (!started_) refill();
        pos_ = offset;
    }
    nrecords_ = nrecords;
    okay_ = pos_ < avail_;
}

void FastqReader::refill() {
    in_->window(started_ ? pos_ : 0, &base_, &avail_, &final_);
    started_ = true;
    pos_ = 0;
}

// One record starting at b[p], following FastqReader::operator() (FastqReader.hpp:42-110).  Returns false when
// the window [b, b + n) ends before the record is complete (the caller knows whether more data can follow;
// with `final` set a truncated record throws like the reference).  Line numbers in the messages: the reference
// counts exactly four lines per record whatever the wrapping, so record k (0-based) "starts" at line 4k+1.
static bool parse_record(const char* b, size_t n, bool final, size_t& pos, long long rec_index, Record& out, bool& okay_after) {
    const long long init_line = 4 * rec_index;
    size_t p = pos;
    if (p >= n) return false;  // needs more data; the caller knows whether that is the end

    auto need_more = [&](int line_offset) -> bool {
        if (final) throw Error("premature end of the file at line " + std::to_string(init_line + line_offset));
        return false;
    };

    if (b[p] != '@') {  // :54-57
        throw Error("read name should start with '@' (starting line " + std::to_string(init_line + 1) + ")");
    }
    // name line: everything up to the first newline after '@' (:58-68)
    const char* e1 = (p + 1 < n) ? static_cast<const char*>(std::memchr(b + p + 1, '\n', n - p - 1)) : nullptr;
    if (!e1) return need_more(1);
    // sequence: up to the first '+', newlines dropped, every other byte kept (:70-78)
    size_t s0 = (size_t)(e1 - b) + 1;
    const char* plus = (s0 < n) ? static_cast<const char*>(std::memchr(b + s0, '+', n - s0)) : nullptr;
    if (!plus) return need_more(2);
    size_t span = (size_t)(plus - (b + s0));
    size_t len = span;
    if (span > 0) {
        // common case: one line, i.e. the only newline is the last byte of the span
        const char* nl = static_cast<const char*>(std::memchr(b + s0, '\n', span));
        if (nl == plus - 1) {
            len = span - 1;
        } else {
            len = 0;
            for (size_t i = 0; i < span; ++i) len += (b[s0 + i] != '\n');
        }
    }
    // '+' line (:81-85): the reference advances past the '+' before looking for the newline
    size_t q0 = (size_t)(plus - b) + 1;
    const char* e3 = (q0 < n) ? static_cast<const char*>(std::memchr(b + q0, '\n', n - q0)) : nullptr;
    if (!e3) return need_more(3);
    // qualities (:91-105): lines are consumed until at least `len` characters were seen
    size_t q = (size_t)(e3 - b) + 1;
    size_t qual = 0;
    bool ended = false, next_okay = false;
    while (q < n) {
        const char* e4 = static_cast<const char*>(std::memchr(b + q, '\n', n - q));
        if (!e4) {
            if (!final) return false;  // quality line not complete yet
            qual += n - q;
            q = n;
            break;
        }
        qual += (size_t)(e4 - (b + q));
        q = (size_t)(e4 - b) + 1;
        if (qual >= len) {
            ended = true;
            next_okay = q < n;
            break;
        }
    }
    if (!ended) {
        if (!final) return false;
        next_okay = false;  // ran off the end of the file inside the quality string
    } else if (!next_okay && !final) {
        // the record ended exactly at the window's edge: whether more input follows is not known yet
        return false;
    }
    if (qual != len) {
        throw Error("non-equal lengths for quality and sequence strings (starting line " + std::to_string(init_line + 1) + ")");
    }
    if (len > (size_t)MAX_READ_LEN) {
        throw Error("reads longer than " + std::to_string(MAX_READ_LEN) + " bases are not supported by this engine (read " +
                    std::to_string(rec_index + 1) + ")");
    }
    out.seq = b + s0;
    out.span = (uint32_t)span;
    out.len = (uint32_t)len;
    pos = q;
    okay_after = next_okay;
    return true;
}

bool FastqReader::parse_one(Record& out) {
    bool next_okay = false;
    if (!parse_record(base_, avail_, final_, pos_, nrecords_, out, next_okay)) return false;
    okay_ = next_okay;
    ++nrecords_;
    return true;
}

// A likely record start at or after `from`: a line starting with '@' whose line after next starts with '+'.
// Only a guess (the grammar allows wrapped sequences and '@' is a legal quality character): the caller checks it
// against where the preceding chunk's exact parse really ends.
static size_t guess_record_start(const char* b, size_t n, size_t from) {
    const char* nl = from < n ? static_cast<const char*>(std::memchr(b + from, '\n', n - from)) : nullptr;
    for (int tries = 0; nl && tries < 64; ++tries) {
        const size_t q = (size_t)(nl - b) + 1;
        if (q >= n) break;
        const char* e1 = static_cast<const char*>(std::memchr(b + q, '\n', n - q));
        if (!e1) break;
        if (b[q] == '@') {
            const size_t s = (size_t)(e1 - b) + 1;
            const char* e2 = s < n ? static_cast<const char*>(std::memchr(b + s, '\n', n - s)) : nullptr;
            if (e2 && (size_t)(e2 - b) + 1 < n && e2[1] == '+') return q;
        }
        nl = e1;
    }
    return (size_t)-1;
}

size_t guess_fastq_record_start(const char* text, size_t size, size_t from) { return guess_record_start(text, size, from); }

// Splits the next stretch of an in-memory input into chunks, parses them concurrently with the exact grammar and
// keeps the longest prefix of chunks whose parses chain up (each ends exactly where the next one started).  Whatever
// does not chain -- a wrong guess, a malformed record -- is left for the serial path, which then reproduces the
// reference's behaviour (and error text) at exactly that record.
bool FastqReader::next_parallel(size_t max_records) {
    const char* b = base_;
    const size_t n = avail_;
    size_t probe = pos_;
    Record first;
    bool dummy = false;
    if (!parse_record(b, n, true, probe, nrecords_, first, dummy)) return false;
    const size_t rec_bytes = std::max<size_t>(probe - pos_, 8);
    const size_t region_end = std::min(n, pos_ + std::max<size_t>(max_records, 1) * rec_bytes);
    const size_t region = region_end - pos_;
    int K = (int)std::min<size_t>((size_t)threads_, region / (1u << 20));
    if (K < 2) return false;
    std::vector<size_t> starts;
    starts.push_back(pos_);
    for (int k = 1; k < K; ++k) {
        const size_t g = guess_record_start(b, n, pos_ + region / K * k);
        if (g == (size_t)-1 || g <= starts.back() || g >= region_end) continue;
        starts.push_back(g);
    }
    K = (int)starts.size();
    if (K < 2) return false;
    // chunk scratch lives in the reader: after the first call no memory is allocated (or page-faulted) here
    if ((int)chunks_.size() < K) chunks_.resize(K);
    std::vector<ParseChunk>& chunks = chunks_;
    auto work = [&](int k) {
        ParseChunk& c = chunks[k];
        c.recs.clear();
        c.min_len = 0xFFFFFFFFu;
        c.max_len = 0;
        c.failed = false;
        c.okay_after = true;
        size_t p = starts[k];
        const size_t stop = k + 1 < K ? starts[k + 1] : region_end;
        try {
            Record r;
            bool ok = true;
            while (ok && p < stop) {
                if (!parse_record(b, n, true, p, 0, r, ok)) break;
                c.recs.push_back(r);
                c.min_len = std::min(c.min_len, r.len);
                c.max_len = std::max(c.max_len, r.len);
            }
            c.okay_after = ok;
        } catch (const std::exception&) {
            c.failed = true;  // records before the bad one stay valid; the serial path raises the error
        }
        c.end = p;
    };
    HostPool::instance().parallel_for(K, K, work);
    size_t total = 0;
    int accepted = 0;
    std::vector<size_t> offset(K + 1, 0);
    for (int k = 0; k < K; ++k) {
        total += chunks[k].recs.size();
        offset[k + 1] = total;
        ++accepted;
        if (chunks[k].failed || !chunks[k].okay_after) break;
        if (k + 1 < K && chunks[k].end != starts[k + 1]) break;
    }
    batch_.resize(total);
    HostPool::instance().parallel_for(accepted, accepted, [&](int k) {
        if (!chunks[k].recs.empty()) std::memcpy(batch_.data() + offset[k], chunks[k].recs.data(), chunks[k].recs.size() * sizeof(Record));
    });
    for (int k = 0; k < accepted; ++k) {
        batch_min_len_ = std::min(batch_min_len_, chunks[k].min_len);
        batch_max_len_ = std::max(batch_max_len_, chunks[k].max_len);
    }
    const ParseChunk& last = chunks[accepted - 1];
    pos_ = last.end;
    okay_ = last.okay_after;
    nrecords_ += (long long)batch_.size();
    return !batch_.empty();
}

const std::vector<Record>& FastqReader::next(size_t max_records) {
    double t0 = now_s();
    batch_.clear();
    batch_min_len_ = 0xFFFFFFFFu;
    batch_max_len_ = 0;
    if (!started_ || (pos_ >= avail_ && !final_)) refill();
    if (nrecords_ == 0 && avail_ == 0 && final_) okay_ = false;  // empty input: zero reads (FastqReader.hpp:30)
    // the whole remaining input is in memory (caller's buffer or mmap): parse it on all the threads we were given
    if (threads_ > 1 && final_ && okay_ && avail_ - pos_ >= (4u << 20)) {
        bool done = false;
        try {
            done = next_parallel(max_records);
        } catch (const std::exception&) {
            done = false;  // the serial path below raises the same error with the right line number
            batch_.clear();
        }
        if (done) {
            parse_s_ += now_s() - t0;
            return batch_;
        }
    }
    Record r;
    while (okay_ && batch_.size() < max_records) {
        if (parse_one(r)) {
            batch_.push_back(r);
            batch_min_len_ = std::min(batch_min_len_, r.len);
            batch_max_len_ = std::max(batch_max_len_, r.len);
            continue;
        }
        // needs more data
        if (final_) break;
        if (!batch_.empty()) break;  // hand out what is complete; records point into the current window
        size_t before = avail_ - pos_;
        refill();
        if (avail_ == before && final_ && avail_ == 0) break;
    }
    parse_s_ += now_s() - t0;
    return batch_;
}

// ---------------------------------------------------------------------------------------
// Packer
// ---------------------------------------------------------------------------------------

namespace {

struct Lut {
    uint8_t code[256];  // 0..3 = ACGT (any case), 4 = anything else
    uint8_t plain[256]; // 1 for upper-case A, C, G, T, N
    Lut() {
        std::memset(code, 4, sizeof code);
        std::memset(plain, 0, sizeof plain);
        const char* up = "ACGT";
        const char* lo = "acgt";
        for (int i = 0; i < 4; ++i) {
            code[(unsigned char)up[i]] = (uint8_t)i;
            code[(unsigned char)lo[i]] = (uint8_t)i;
            plain[(unsigned char)up[i]] = 1;
        }
        plain[(unsigned char)'N'] = 1;
    }
};
const Lut g_lut;

// Packs up to 32 bases starting at s into one word of each plane; returns "all plain" flag.
inline bool pack_word_scalar(const char* s, int count, uint32_t& h, uint32_t& l, uint32_t& n) {
    uint32_t hh = 0, ll = 0, nn = 0;
    uint8_t plain = 1;
    for (int i = 0; i < count; ++i) {
        unsigned char c = (unsigned char)s[i];
        uint32_t code = g_lut.code[c];
        plain &= g_lut.plain[c];
        hh |= ((code >> 1) & 1u) << i;
        ll |= (code & 1u) << i;
        nn |= (code >> 2) << i;
    }
    h = hh;
    l = ll;
    n = nn;
    return plain != 0;
}

#if defined(__x86_64__)
// 32 bases at once.  ASCII trick: bit 2 of A/C/G/T (either case) is 0/0/1/1 = the H plane, and
// bit 1 is 0/1/1/0, so L = bit1 ^ bit2.
__attribute__((target("avx2"))) inline bool pack_word_avx2(const char* s, uint32_t& h, uint32_t& l, uint32_t& n) {
    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    __m256i up = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));  // fold case
    __m256i isA = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('A'));
    __m256i isC = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'));
    __m256i isG = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('G'));
    __m256i isT = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('T'));
    // folding with 0xDF also maps a few non-letters onto letters (e.g. 0x61^0x20); require an alphabetic source byte
    __m256i alpha = _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8((char)0xC0)), _mm256_set1_epi8(0x40));
    __m256i valid = _mm256_and_si256(_mm256_or_si256(_mm256_or_si256(isA, isC), _mm256_or_si256(isG, isT)), alpha);
    uint32_t vmask = (uint32_t)_mm256_movemask_epi8(valid);
    uint32_t b2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
    uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
    h = b2 & vmask;
    l = (b1 ^ b2) & vmask;
    n = ~vmask;
    __m256i plainv = _mm256_or_si256(
        _mm256_or_si256(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(v, _mm256_set1_epi8('C'))),
        _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(v, _mm256_set1_epi8('T'))),
                        _mm256_cmpeq_epi8(v, _mm256_set1_epi8('N'))));
    return (uint32_t)_mm256_movemask_epi8(plainv) == 0xFFFFFFFFu;
}
#endif

bool have_avx2() {
#if defined(__x86_64__)
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
#else
    return false;
#endif
}

// One read whose bases are contiguous (no embedded newline).  Returns the "plain" flag.
#if defined(__x86_64__)
__attribute__((target("avx2")))
#endif
bool pack_contiguous(const char* s, uint32_t len, int W, uint32_t* h, uint32_t* l, uint32_t* n, size_t stride, bool avx2) {
    bool plain = true;
    uint32_t full = len / 32, w = 0;
    for (; w < full; ++w) {
        uint32_t hh, ll, nn;
#if defined(__x86_64__)
        if (avx2) {
            plain &= pack_word_avx2(s + 32 * w, hh, ll, nn);
        } else
#endif
        {
            plain &= pack_word_scalar(s + 32 * w, 32, hh, ll, nn);
        }
        h[w * stride] = hh;
        l[w * stride] = ll;
        n[w * stride] = nn;
    }
    uint32_t rem = len - 32 * full;
    if (rem) {
        uint32_t hh, ll, nn;
        plain &= pack_word_scalar(s + 32 * w, (int)rem, hh, ll, nn);
        h[w * stride] = hh;
        l[w * stride] = ll;
        n[w * stride] = nn;
        ++w;
    }
    for (; w < (uint32_t)W; ++w) {
        h[w * stride] = 0;
        l[w * stride] = 0;
        n[w * stride] = 0;
    }
    return plain;
}

} // namespace

void pack_read_scalar(const char* seq, uint32_t span, int W, uint32_t* h, uint32_t* l, uint32_t* n, size_t stride) {
    for (int w = 0; w < W; ++w) h[w * stride] = l[w * stride] = n[w * stride] = 0;
    uint32_t i = 0;
    for (uint32_t k = 0; k < span; ++k) {
        char c = seq[k];
        if (c == '\n') continue;
        uint32_t code = g_lut.code[(unsigned char)c];
        uint32_t bit = 1u << (i & 31);
        if (code & 2) h[(i >> 5) * stride] |= bit;
        if (code & 1) l[(i >> 5) * stride] |= bit;
        if (code & 4) n[(i >> 5) * stride] |= bit;
        ++i;
    }
}

void pack_records(const Record* recs, size_t count, int W, uint32_t* out, uint16_t* lens, uint8_t* odd, int nthreads) {
    const size_t ntiles = (count + TILE - 1) / TILE;
    const size_t tw = tile_words(W);
    const bool avx2 = have_avx2();
    auto work = [&](size_t tile_begin, size_t tile_end) {
        std::vector<char> scratch;
        for (size_t t = tile_begin; t < tile_end; ++t) {
            uint32_t* base = out + t * tw;
            for (int lane = 0; lane < TILE; ++lane) {
                size_t i = t * TILE + lane;
                uint32_t* h = base + (size_t)(PLANE_H * W) * TILE + lane;
                uint32_t* l = base + (size_t)(PLANE_L * W) * TILE + lane;
                uint32_t* n = base + (size_t)(PLANE_N * W) * TILE + lane;
                if (i >= count) {
                    for (int w = 0; w < W; ++w) h[w * TILE] = l[w * TILE] = n[w * TILE] = 0;
                    if (lens) lens[i] = 0;
                    if (odd) odd[i] = 0;
                    continue;
                }
                const Record& r = recs[i];
                const char* s = r.seq;
                bool plain;
                if (r.span == r.len || (r.span == r.len + 1 && r.seq[r.len] == '\n')) {
                    plain = pack_contiguous(s, r.len, W, h, l, n, TILE, avx2);
                } else {
                    scratch.clear();
                    for (uint32_t k = 0; k < r.span; ++k) {
                        if (s[k] != '\n') scratch.push_back(s[k]);
                    }
                    plain = pack_contiguous(scratch.data(), r.len, W, h, l, n, TILE, avx2);
                }
                if (lens) lens[i] = (uint16_t)r.len;
                if (odd) odd[i] = plain ? 0 : 1;
            }
        }
    };
    int nt = std::max(1, std::min<int>(nthreads, (int)ntiles));
    if (nt == 1 || ntiles < 64) {
        work(0, ntiles);
        return;
    }
    // more pieces than threads, so that a slow core does not hold everyone up
    const int pieces = nt * 4;
    const size_t per = (ntiles + pieces - 1) / pieces;
    HostPool::instance().parallel_for(pieces, nt, [&](int k) {
        const size_t b = (size_t)k * per, e = std::min(ntiles, b + per);
        if (b < e) work(b, e);
    });
}

} // namespace scg
