#include "bgzf.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <sched.h>

#include <zlib.h>

#include "common.hpp"
#include "hostpool.hpp"

namespace scg {

size_t BgzfIndex::block_of(size_t text_offset) const {
    // last block whose text starts at or before the offset and that is not empty there
    const auto it = std::upper_bound(text_off.begin(), text_off.end(), text_offset);
    if (it == text_off.begin()) return 0;
    return std::min<size_t>((size_t)(it - text_off.begin()) - 1, blocks.size());
}

namespace {

// One member at offset p: its block record and its whole size.  false = not a well-formed BGZF member there.
bool parse_member(const unsigned char* f, size_t size, size_t p, BgzfBlock& b, size_t& bsize) {
    if (p >= size || size - p < 18) return false;
    if (f[p] != 0x1f || f[p + 1] != 0x8b || f[p + 2] != 8 || f[p + 3] != 4) return false;   // FLG = FEXTRA only
    const size_t xlen = f[p + 10] | ((size_t)f[p + 11] << 8);
    if (size - p < 12 + xlen + 8) return false;
    bsize = 0;
    for (size_t q = p + 12; q + 4 <= p + 12 + xlen;) {
        const size_t slen = f[q + 2] | ((size_t)f[q + 3] << 8);
        if (f[q] == 'B' && f[q + 1] == 'C' && slen == 2 && q + 6 <= p + 12 + xlen) bsize = (f[q + 4] | ((size_t)f[q + 5] << 8)) + 1;
        q += 4 + slen;
    }
    if (bsize < 12 + xlen + 8 || bsize > size - p) return false;
    b.data = p + 12 + xlen;
    b.csize = (uint32_t)(bsize - 12 - xlen - 8);
    const unsigned char* t = f + p + bsize - 8;
    b.crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    b.isize = t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    return b.isize <= (1u << 16);   // BGZF members hold at most 64 KiB of text
}

// Members from offset p up to `stop` (the chain must land exactly there); false = broken chain.
bool walk_members(const unsigned char* f, size_t size, size_t p, size_t stop, std::vector<BgzfBlock>& blocks) {
    while (p < stop) {
        BgzfBlock b;
        size_t bsize = 0;
        if (!parse_member(f, size, p, b, bsize)) return false;
        blocks.push_back(b);
        p += bsize;
    }
    return p == stop;
}

int index_threads() {
    static const int n = [] {
        if (const char* env = std::getenv("SCG_BGZF_INDEX_THREADS")) return std::max(1, std::atoi(env));
        cpu_set_t set;
        CPU_ZERO(&set);
        const int usable = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : 1;
        return std::max(1, std::min(usable, 16));
    }();
    return n;
}

} // namespace

// The chain is sequential (every member announces only its own size) and a walk over a large image is one cache miss per
// member, 2 ms for the 19 000 members of a gigabyte of text.  Large images are therefore walked in pieces on the host pool:
// every piece starts at the first offset from which three well-formed members follow each other, and the pieces are only
// accepted when each one's chain lands exactly on the start of the next -- otherwise (it never happened) the serial walk decides.
bool bgzf_index(const unsigned char* f, size_t size, BgzfIndex& out) {
    out = BgzfIndex();
    if (size < 28) return false;
    std::vector<BgzfBlock> blocks;
    bool done = false;
    const int pieces = (int)std::min<size_t>((size_t)index_threads(), size >> 22);   // 4 MiB of image per piece at least
    if (pieces > 1) {
        std::vector<size_t> start((size_t)pieces + 1, 0);
        std::vector<std::vector<BgzfBlock>> part((size_t)pieces);
        std::vector<int> ok((size_t)pieces, 1);
        start[(size_t)pieces] = size;
        HostPool::instance().parallel_for(pieces, pieces, [&](int k) {
            if (k == 0) return;
            const size_t from = size / (size_t)pieces * (size_t)k, to = std::min(size, from + (1u << 17));
            ok[(size_t)k] = 0;
            for (size_t p = from; p < to; ++p) {
                if (f[p] != 0x1f) continue;
                BgzfBlock b;
                size_t q = p, bsize = 0;
                int chain = 0;
                while (chain < 3 && q < size && parse_member(f, size, q, b, bsize)) {
                    q += bsize;
                    ++chain;
                }
                if (chain == 3 || (chain > 0 && q == size)) {
                    start[(size_t)k] = p;
                    ok[(size_t)k] = 1;
                    break;
                }
            }
        });
        bool all = true;
        for (int k = 0; k < pieces; ++k) all = all && ok[(size_t)k] && start[(size_t)k] < start[(size_t)k + 1];
        if (all) {
            HostPool::instance().parallel_for(pieces, pieces, [&](int k) {
                part[(size_t)k].reserve((start[(size_t)k + 1] - start[(size_t)k]) / 4096 + 16);
                ok[(size_t)k] = walk_members(f, size, start[(size_t)k], start[(size_t)k + 1], part[(size_t)k]) ? 1 : 0;
            });
            for (int k = 0; k < pieces; ++k) all = all && ok[(size_t)k];
        }
        if (all) {
            size_t n = 0;
            for (const auto& v : part) n += v.size();
            blocks.reserve(n);
            for (const auto& v : part) blocks.insert(blocks.end(), v.begin(), v.end());
            done = true;
        }
    }
    if (!done) {
        blocks.clear();
        blocks.reserve(size / 16384 + 16);
        if (!walk_members(f, size, 0, size, blocks)) return false;
    }
    if (blocks.empty()) return false;
    out.text_off.resize(blocks.size() + 1);
    out.text_off[0] = 0;
    for (size_t i = 0; i < blocks.size(); ++i) out.text_off[i + 1] = out.text_off[i] + blocks[i].isize;
    out.blocks.swap(blocks);
    out.image = f;
    out.image_size = size;
    return true;
}

bool bgzf_inflate_block(const BgzfIndex& index, size_t block, char* out) {
    const BgzfBlock& blk = index.blocks[block];
    if (blk.isize == 0) return true;
    z_stream z;
    std::memset(&z, 0, sizeof z);
    if (inflateInit2(&z, -15) != Z_OK) return false;
    z.next_in = const_cast<unsigned char*>(index.image + blk.data);
    z.avail_in = blk.csize;
    z.next_out = reinterpret_cast<unsigned char*>(out);
    z.avail_out = blk.isize;
    const int rc = inflate(&z, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && z.avail_out == 0 && crc32(0L, reinterpret_cast<const unsigned char*>(out), blk.isize) == blk.crc;
    inflateEnd(&z);
    return ok;
}

size_t bgzf_compress(const char* text, size_t size, int level, size_t block_text, int nthreads, unsigned char* out, size_t capacity) {
    block_text = std::max<size_t>(1, std::min<size_t>(block_text, 0xff00));
    const size_t nblocks = (size + block_text - 1) / block_text;
    const size_t per_block = 18 + compressBound((uLong)block_text) + 8;
    const size_t bound = nblocks * per_block + 28;
    if (!out) return bound;
    if (capacity < bound) throw Error("bgzf_compress: output buffer too small");
    // every member is compressed into its own slot of `per_block` bytes, then the slots are closed up
    std::vector<uint32_t> sizes(nblocks, 0);
    std::vector<int> failed(nblocks, 0);
    auto member = [&](size_t k, const char* src, size_t n, unsigned char* dst) -> size_t {
        z_stream z;
        std::memset(&z, 0, sizeof z);
        if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
            if (k < nblocks) failed[k] = 1;
            return 0;
        }
        z.next_in = reinterpret_cast<unsigned char*>(const_cast<char*>(src));
        z.avail_in = (uInt)n;
        z.next_out = dst + 18;
        z.avail_out = (uInt)(per_block - 26);
        const int rc = deflate(&z, Z_FINISH);
        const size_t raw = z.total_out;
        deflateEnd(&z);
        if (rc != Z_STREAM_END) {
            if (k < nblocks) failed[k] = 1;
            return 0;
        }
        const size_t bsize = 18 + raw + 8;
        const unsigned char head[18] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)((bsize - 1) & 0xff),
                                         (unsigned char)((bsize - 1) >> 8) };
        std::memcpy(dst, head, 18);
        const uint32_t crc = (uint32_t)crc32(0L, reinterpret_cast<const unsigned char*>(src), (uInt)n), isize = (uint32_t)n;
        unsigned char* t = dst + 18 + raw;
        for (int b = 0; b < 4; ++b) {
            t[b] = (unsigned char)(crc >> (8 * b));
            t[4 + b] = (unsigned char)(isize >> (8 * b));
        }
        return bsize;
    };
    HostPool::instance().parallel_for((int)nblocks, std::max(1, nthreads), [&](int k) {
        const size_t from = (size_t)k * block_text;
        sizes[(size_t)k] = (uint32_t)member((size_t)k, text + from, std::min(block_text, size - from), out + (size_t)k * per_block);
    });
    for (int f : failed) {
        if (f) throw Error("bgzf_compress: deflate failed");
    }
    size_t at = 0;
    for (size_t k = 0; k < nblocks; ++k) {
        if (at != k * per_block) std::memmove(out + at, out + k * per_block, sizes[k]);
        at += sizes[k];
    }
    // the empty member that marks the end of a BGZF file: bgzip's fixed 28 bytes, whatever the level
    static const unsigned char eof_marker[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    std::memcpy(out + at, eof_marker, sizeof eof_marker);
    return at + sizeof eof_marker;
}

} // namespace scg
