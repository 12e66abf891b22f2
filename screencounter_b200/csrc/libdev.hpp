// Plain-old-data views of device-resident objects, readable by nvcc-built kernels and by the
// run-time specialised kernels (NVRTC) alike: no standard headers.
#pragma once

#include "layout.hpp"

namespace scg {

// Device view of a library (library.hpp builds it): plain pointers into device memory.
struct LibDev {
    int L;              // key length in bases
    int KW;             // words per plane
    int nentries;       // expanded (concrete) entries
    int dup_first;      // ties between different indices resolve to the lowest index
    // exact table, two-table cuckoo: T1 then T2, slot_mask + 1 slots of slot_words words each:
    // h[KW], l[KW], value (-1 = empty), padding; a key is at T1[hash & mask] or T2[hash_second(hash) & mask]
    const uint32_t* slots;
    uint32_t slot_mask;
    int slot_words;
    // expanded entries for verification: ent_keys[e*2KW ..] = h[KW], l[KW]; ent_idx[e] = pool index
    const uint32_t* ent_keys;
    const int32_t* ent_idx;
    // pigeonhole seeds
    int nseeds;
    const uint32_t* seed_masks;  // nseeds * KW words: base positions of each seed (same mask for H and L)
    const uint2* buckets;     // nseeds * (bucket_mask + 1) entries of (start, count) into cands
    uint32_t bucket_mask;
    const int32_t* cands;     // nseeds * nentries entry ids, grouped by bucket
    // KW == 1 only: the same candidates as rows (h, l, pool index, 0), so that a candidate costs one 16-byte load
    const uint4* cand_rows;
    // segmented search (dual paired-end): first segment = bases [0, seg1), second = [seg1, L)
    int seg1;
    // table of library rows with their last base dropped (SURVEY 8.1 T8 root rule)
    const uint32_t* prefix_slots;
    uint32_t prefix_mask;
    // the reference's flat trie (library.hpp), nullptr unless the first segment may take 2 or more mismatches
    const int32_t* trie;
};

// What the specialised single-barcode kernel (spec_single.cuh) needs of the two strands' libraries, passed by
// value as a kernel argument so that the table pointers sit in the constant bank.  [0] forward, [1] reverse.
struct SpecTables {
    const uint4* slots[2];      // exact cuckoo tables (16-byte slots)
    const uint2* buckets[2];    // pigeonhole seed buckets
    const uint4* cand_rows[2];  // candidates as rows
    uint32_t slot_mask[2];
    uint32_t bucket_mask[2];
    int nentries[2];
    const LibDev* libs;         // the full descriptors, for the generic search of the rare complicated reads
    // Exact table of BOTH strands for keys of up to 31 bases (uniform-length kernel): a two-table cuckoo hash like the
    // per-strand ones (16-byte slots: H word, L word, pool index, 0; empty = index -1) with the strand as bit 31 of the H
    // word, so that the table's address and size are the same for every lane, and with a hash of three multiply-adds.
    // Table 1 holds n = 1 << (32 - joint_shift) slots, table 2 the next n.
    const uint4* joint;
    uint32_t joint_shift;
    // Seed buckets with their first candidate inline: (H word, L word, pool index, start | count << 24) per bucket, same
    // indexing as `buckets`.  Nine in ten non-empty buckets hold one candidate: one load instead of bucket -> row.
    const uint4* ibuckets[2];
};

// the two homes of a key in the joint table (host builder and kernel alike): top bits of a multiplicative hash
SCG_HD uint32_t joint_hash(uint32_t kh_tagged, uint32_t kl) { return kh_tagged * 0x9E3779B1u + kl * 0x85EBCA6Bu; }
SCG_HD uint32_t joint_hash2(uint32_t x) { return x * 0xC2B2AE35u; }
constexpr int JOINT_MAX_KEYLEN = 31;

// Packed reads of one batch on the device.
struct ReadsDev {
    const uint32_t* data;   // tile-planar words
    const uint16_t* lens;   // nullptr when every read has length uniform_len
    int uniform_len;
    int W;
    long long n;
};

} // namespace scg
