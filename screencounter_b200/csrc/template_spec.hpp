// Template of constant flanks and variable regions: host-side parsing/validation and the POD the
// kernels read.  Follows the reference's ScanTemplate constructor
// (inst/include/kaori/ScanTemplate.hpp:53-95) for what is accepted, which strand needs what, and
// where the variable regions sit on each strand.
#pragma once

#include <string>
#include <vector>

#include "common.hpp"
#include "layout.hpp"

namespace scg {

// What the scan kernels need, passed by value as a kernel argument (constant bank).
// Bit s of word q describes template position 32*q + s.
struct ScanSpec {
    int T;                                  // template length
    int nwords;                             // ceil(T / 32)
    int fwd, rev;                           // strands searched
    int mm;                                 // mismatch budget for the constant part
    int cbits;                              // counter planes needed to count 0..mm+1 mismatches
    int nreg;                               // variable regions
    int fstart[MAX_REGIONS], rstart[MAX_REGIONS], rlen_f[MAX_REGIONS], rlen_r[MAX_REGIONS];
    uint32_t care_f[MAX_TEMPLATE_WORDS], hi_f[MAX_TEMPLATE_WORDS], lo_f[MAX_TEMPLATE_WORDS];
    uint32_t care_r[MAX_TEMPLATE_WORDS], hi_r[MAX_TEMPLATE_WORDS], lo_r[MAX_TEMPLATE_WORDS];
};

struct Region {
    int start, end;
};

struct TemplateSpec {
    std::string text;
    int length = 0;
    bool fwd = false, rev = false;
    std::string fwd_seq, rev_seq;             // '-' for variable positions
    std::vector<Region> fwd_regions, rev_regions;  // variable_regions<false/true>()
    int n_constant = 0;

    // strand: 0 original, 1 reverse, 2 both (src/utils.cpp:33-41).
    TemplateSpec() {}
    TemplateSpec(const std::string& constant, int strand);

    ScanSpec scan_spec(int mismatches) const;
};

} // namespace scg
