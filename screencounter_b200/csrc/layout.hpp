// Data layout shared by the host packer and the CUDA kernels ("Data layout in HBM").
//
// READS ("tile-planar", TP32).  Reads are grouped in tiles of 32.  A read is three bit planes
// of W 32-bit words each (W = ceil(longest read of the batch / 32)):
//     plane H = high bit of the 2-bit base code, plane L = low bit, plane N = "not ACGT" mask.
//     A = 00, C = 01, G = 10, T = 11  =>  complement = flip both bits.
// Base i of a read lives at bit (i % 32) of word (i / 32).  Any read character outside
// ACGTacgt sets its N bit and clears H and L (the reference treats every such character as
// one mismatch: ScanTemplate.hpp:157-166, MismatchTrie.hpp:452-453).  Bits past the end of a
// read are zero in all three planes.
// Word (plane p, word w) of the 32 reads of a tile is one contiguous 128-byte row, so a warp
// whose lane r owns read r loads every word fully coalesced:
//     index(tile, p, w, lane) = ((tile * 3 + p) * W + w) * 32 + lane
//
// KEYS (barcodes and variable regions) use the same H/L planes, KW = ceil(length / 32) words
// per plane; an extracted variable region also carries its N plane.
#pragma once

#if defined(__CUDACC_RTC__)
// run-time compilation: no standard headers, so the fixed-width types are spelled out
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <cstddef>
#include <cstdint>
#endif

#if defined(__CUDACC__)
#define SCG_HD __host__ __device__ __forceinline__
#else
#define SCG_HD inline
#endif

namespace scg {

constexpr int TILE = 32;
constexpr int PLANE_H = 0, PLANE_L = 1, PLANE_N = 2;
constexpr int MAX_TEMPLATE = 256;          // reference limit, src/count_single_barcodes.cpp:45-46
constexpr int MAX_TEMPLATE_WORDS = MAX_TEMPLATE / 32;
constexpr int MAX_REGIONS = 16;            // variable regions per template handled on the device
constexpr int MAX_KEY_WORDS = 16;          // keys up to 512 bases (two 256-base variable regions)
constexpr int MAX_SEEDS = 16;
constexpr int MAX_READ_LEN = 65535;        // lengths travel as uint16

SCG_HD size_t tile_words(int W) { return (size_t)3 * W * TILE; }

// 2-bit code of a base character, or -1 for anything that is not ACGTacgt.
SCG_HD int base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return -1;
}

// 64-bit finaliser (splitmix64); the one hash used for tables, seeds and the synthetic generator.
SCG_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Hash of a (masked) key given as kw words per plane: 32-bit multiply/xorshift rounds (cheap on the
// GPU's integer pipes; the same function builds the tables on the host).
SCG_HD uint32_t hash_key(const uint32_t* h, const uint32_t* l, int kw, uint32_t salt) {
    uint32_t acc = salt * 0x9E3779B1u + 0x7F4A7C15u;
    for (int i = 0; i < kw; ++i) {
        acc = (acc ^ h[i]) * 0x85EBCA6Bu;
        acc ^= acc >> 13;
        acc = (acc ^ l[i]) * 0xC2B2AE35u;
        acc ^= acc >> 16;
    }
    return acc;
}

// Second position of a key in the two-table cuckoo layout of the exact tables, derived from the first hash.
SCG_HD uint32_t hash_second(uint32_t acc) {
    const uint32_t x = acc * 0x9E3779B1u;
    return x ^ (x >> 15);
}

// Slack after every packed-read buffer: the scan may load up to two words past a read's last
// plane word (bits that are masked out afterwards), so the loads need no bounds checks.
constexpr size_t READ_GUARD_BYTES = 1024;

} // namespace scg
