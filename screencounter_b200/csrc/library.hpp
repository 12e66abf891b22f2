// Barcode library: host-side construction (validation, IUPAC expansion, duplicate policy) and
// the flat tables the kernels probe.  Replaces the reference's fill_library + MismatchTrie
// (inst/include/kaori/BarcodeSearch.hpp:23-60, MismatchTrie.hpp:93-205) with
//   * a two-table cuckoo hash of packed barcodes for exact hits (a key lives at one of exactly two
//     positions, so a probe is two independent 16-byte loads and no loop), and
//   * pigeonhole seed buckets for the mismatch-tolerant search: a barcode within `cap`
//     substitutions of the query agrees with it exactly on at least one of cap+1 disjoint
//     segments, so probing one bucket per segment enumerates every candidate; candidates are
//     verified by XOR/popcount on the packed planes and the best-unique / tie rules of
//     MismatchTrie.hpp:266-343 are applied to the verified distances.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "common.hpp"
#include "layout.hpp"
#include "libdev.hpp"

namespace scg {

enum class Duplicates { FIRST, ERROR };  // the two policies reachable from the R API

struct LibraryOptions {
    int max_mismatches = 0;          // any-mismatch search: cap never exceeds this
    bool segmented = false;          // two segments with their own caps
    int seg1 = 0;                    // length of the first segment
    int max_mismatches1 = 0, max_mismatches2 = 0;
    Duplicates duplicates = Duplicates::ERROR;
};

// Host image of the tables.
struct Library {
    int L = 0, KW = 0, nchoices = 0;
    LibraryOptions opt;
    std::vector<uint32_t> ent_keys;
    std::vector<int32_t> ent_idx;
    int slot_words = 0;
    std::vector<uint32_t> slots;        // tables T1 | T2, slot_mask + 1 slots each
    uint32_t slot_mask = 0;
    int nseeds = 0;
    std::vector<uint32_t> seed_masks;   // nseeds * KW
    uint32_t nbuckets = 0;
    std::vector<uint2> buckets;
    std::vector<int32_t> cands;
    std::vector<uint32_t> cand_rows;    // KW == 1: 4 words (h, l, pool index, 0) per (seed, candidate), in `cands` order
    std::vector<uint32_t> prefix_slots; // same layout, rows with the last base dropped
    uint32_t prefix_mask = 0;
    // kaori's flat 4-ary trie of the concrete rows (MismatchTrie.hpp:93-205: node = 4 child pointers, -1 = none, leaves hold pool
    // indices), built for segmented libraries whose first segment may take 2 or more mismatches: the search with caps [>= 2, 0]
    // can only be reproduced by walking it like the reference does
    std::vector<int32_t> trie;

    // `sequences` are the library rows as they must match the read (i.e. already
    // reverse-complemented by the caller when the reverse strand is searched).
    Library() {}
    Library(const std::vector<std::string>& sequences, int length, const LibraryOptions& options);

    size_t nentries() const { return ent_idx.size(); }
};

// Reverse complement of a pool sequence with IUPAC codes (utils.hpp:41-120, complement_base<true,true>).
std::string reverse_complement_iupac(const std::string& s);

// Pack a concrete ACGT string (any case) into planes; returns false if a non-ACGT char is present,
// in which case its bit is set in n.
bool pack_key(const char* s, int len, uint32_t* h, uint32_t* l, uint32_t* n);

} // namespace scg
