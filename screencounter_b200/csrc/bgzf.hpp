// Block gzip (BGZF: what bgzip and Illumina's bcl-convert / bcl2fastq write): a chain of gzip members of at most 64 KiB of
// text, each announcing its compressed size in a 'BC' extra field, so the members can be found without inflating and
// inflated independently -- on the device (inflate.cuh) or on host threads (fastq.cpp).  The reference reads such files
// like any gzip stream, member by member on one thread (byteme::GzipFileReader, inst/include/byteme/GzipFileReader.hpp:39-51).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace scg {

struct BgzfBlock {
    size_t data;       // offset of the raw deflate stream in the image
    uint32_t csize;    // its length
    uint32_t isize;    // text bytes
    uint32_t crc;      // CRC-32 of the text
};

struct BgzfIndex {
    const unsigned char* image = nullptr;
    size_t image_size = 0;
    std::vector<BgzfBlock> blocks;
    std::vector<size_t> text_off;   // blocks.size() + 1 entries: where each block's text starts
    size_t text_size() const { return text_off.empty() ? 0 : text_off.back(); }
    size_t block_of(size_t text_offset) const;   // the block holding that byte of text (blocks.size() at the end)
};

// Walks the member chain; false unless every byte of the image belongs to a well-formed BGZF member.
bool bgzf_index(const unsigned char* image, size_t size, BgzfIndex& out);

// Inflates one block on the host into out (isize bytes), CRC checked.  false = corrupt member.
bool bgzf_inflate_block(const BgzfIndex& index, size_t block, char* out);

// Compresses text into a BGZF image (members of `block_text` bytes of text, zlib level `level`, closed by the empty member),
// on `nthreads` host threads.  Returns the image size; with out == nullptr only the bound needed is returned.
size_t bgzf_compress(const char* text, size_t size, int level, size_t block_text, int nthreads, unsigned char* out, size_t capacity);

} // namespace scg
