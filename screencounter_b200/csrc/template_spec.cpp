#include "template_spec.hpp"

#include <algorithm>
#include <cstring>

namespace scg {

namespace {

// utils.hpp:41-62 with allow_n = allow_iupac = false: only ACGTacgt have a complement.
char complement_strict(char b) {
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        case 'T': case 't': return 'A';
    }
    return 0;
}

void append_variable(std::vector<Region>& regions, int i) {  // ScanTemplate.hpp:289-298
    if (!regions.empty() && regions.back().end == i) {
        ++regions.back().end;
    } else {
        regions.push_back(Region{ i, i + 1 });
    }
}

} // namespace

TemplateSpec::TemplateSpec(const std::string& constant, int strand) : text(constant) {
    // src/count_single_barcodes.cpp:37-47: templates are dispatched to max_size 32/64/128/256.
    if (constant.size() > (size_t)MAX_TEMPLATE) {
        throw Error("lacking compile-time support for constant regions longer than 256 bp");
    }
    length = (int)constant.size();
    fwd = (strand == 0 || strand == 2);
    rev = (strand == 1 || strand == 2);
    fwd_seq.assign(length, '-');
    rev_seq.assign(length, '-');

    // Forward pass (ScanTemplate.hpp:59-81).  Constant bases are only validated when the forward
    // strand is searched; the forward variable regions are always recorded.
    for (int i = 0; i < length; ++i) {
        char b = constant[i];
        if (b == '-') {
            append_variable(fwd_regions, i);
        } else {
            if (fwd && base_code(b) < 0) {
                throw Error(std::string("unknown base '") + b + "'");  // utils.hpp:140-161
            }
            fwd_seq[i] = b;
            ++n_constant;
        }
    }
    // Reverse pass (ScanTemplate.hpp:82-94): reverse-complemented template, mirrored regions.
    if (rev) {
        for (int i = 0; i < length; ++i) {
            char b = constant[length - i - 1];
            if (b == '-') {
                append_variable(rev_regions, i);
            } else {
                char c = complement_strict(b);
                if (!c) {
                    throw Error(std::string("cannot complement unknown base '") + b + "'");  // utils.hpp:116-117
                }
                rev_seq[i] = c;
            }
        }
    }
}

ScanSpec TemplateSpec::scan_spec(int mismatches) const {
    ScanSpec s;
    std::memset(&s, 0, sizeof s);
    s.T = length;
    s.nwords = ceil_div(length, 32);
    s.fwd = fwd;
    s.rev = rev;
    // A budget at or above the number of constant positions accepts every window; clamp so the
    // bit-sliced counter stays small (the clamped value behaves identically).
    int mm = std::max(0, std::min(mismatches, n_constant));
    s.mm = mismatches < 0 ? -1 : mm;
    // The bit-sliced counter keeps `bits` planes plus a sticky overflow flag: it must count
    // 0..mm exactly (2^bits - 1 >= mm) and report anything above as overflow.
    int bits = 0;
    while ((1 << bits) < mm + 1) ++bits;
    s.cbits = bits;
    if ((int)fwd_regions.size() > MAX_REGIONS) {
        throw Error("templates with more than " + std::to_string(MAX_REGIONS) + " variable regions are not supported by this engine");
    }
    s.nreg = (int)fwd_regions.size();
    for (int r = 0; r < s.nreg; ++r) {
        s.fstart[r] = fwd_regions[r].start;
        s.rlen_f[r] = fwd_regions[r].end - fwd_regions[r].start;
        if (rev) {
            s.rstart[r] = rev_regions[r].start;
            s.rlen_r[r] = rev_regions[r].end - rev_regions[r].start;
        }
    }
    for (int i = 0; i < length; ++i) {
        int q = i / 32, b = i % 32;
        if (fwd && fwd_seq[i] != '-') {
            int code = base_code(fwd_seq[i]);
            s.care_f[q] |= 1u << b;
            if (code & 2) s.hi_f[q] |= 1u << b;
            if (code & 1) s.lo_f[q] |= 1u << b;
        }
        if (rev && rev_seq[i] != '-') {
            int code = base_code(rev_seq[i]);
            s.care_r[q] |= 1u << b;
            if (code & 2) s.hi_r[q] |= 1u << b;
            if (code & 1) s.lo_r[q] |= 1u << b;
        }
    }
    return s;
}

} // namespace scg
