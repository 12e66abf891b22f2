// Plain-old-data parameter blocks shared by the host runtime and the kernels (handlers.cuh).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "layout.hpp"
#include "libdev.hpp"
#include "library.hpp"
#include "template_spec.hpp"

namespace scg {


struct SingleParams {
    ScanSpec spec;
    const LibDev* libs;   // device array: [0] forward library, [1] reverse-complemented library
    int kw;               // words per plane of the longest key (host-side dispatch)
    int max_mm;
    int use_first;
};

// CountSlot / CountTable64 / CountTable128 (the device count tables) live in libdev.hpp: the run-time compiled kernels
// use them too.

struct RandomParams {
    ScanSpec spec;
    int max_mm;
    int use_first;
    int key_len;   // length of the first variable region
};

struct ComboParams {
    ScanSpec spec;
    const LibDev* libs;          // device array indexed [2 * reverse + region]; reverse region r is built
                                 // from the reverse complement of pool[1 - r]
    int kw;
    int max_mm;
    int use_first;
    int n1, n2;                  // pool sizes
};

struct DualSEParams {
    ScanSpec spec;
    const LibDev* libs;   // device array: [0] concatenated rows, [1] their reverse complements
    int kw;
    int max_mm;
    int use_first;
};

struct ComboPEParams {
    SingleParams m1, m2;
    int randomized;
    int use_first;
};

struct DualPEParams {
    ScanSpec spec1, spec2;   // each searches exactly one strand
    const LibDev* lib;       // device: rows = [RC?]pool1[i] + [RC?]pool2[i], segments (len1, len2)
    int kw;
    int mm1, mm2;
    int randomized;
    int use_first;
    int len1, len2;
};

} // namespace scg
