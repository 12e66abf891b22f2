// The generic constant-flank scan (replaces ScanTemplate::next / strand_match,
// inst/include/kaori/ScanTemplate.hpp:183-252), bit-parallel ACROSS WINDOW POSITIONS: bit p of a
// 32-bit register is window position p, so one LOP3/SHF handles 32 windows.  The template arrives
// at run time as bit masks (ScanSpec); spec_single.cuh is the same scan with the template folded in
// at compile time.
#pragma once

#include "device_keys.cuh"
#include "template_spec.hpp"

namespace scg {

// Mismatch counts of the constant template positions for the 32 windows starting at
// 32*pb .. 32*pb+31, both strands in one pass over the read words.
template <int CB>
__device__ __forceinline__ void scan_block(const ReadView& rd, const ScanSpec& s, int pb, Counter<CB>& cf, Counter<CB>& cr) {
    cf.clear();
    cr.clear();
    uint32_t h0 = rd.word(PLANE_H, pb), l0 = rd.word(PLANE_L, pb), n0 = rd.word(PLANE_N, pb);
    for (int q = 0; q < s.nwords; ++q) {
        // template positions 32q .. 32q+31 look at read bits 32(pb+q) + sft + p
        const uint32_t h1 = rd.word(PLANE_H, pb + q + 1), l1 = rd.word(PLANE_L, pb + q + 1), n1 = rd.word(PLANE_N, pb + q + 1);
        const uint32_t care_f = s.care_f[q], care_r = s.care_r[q];
        const uint32_t hi_f = s.hi_f[q], lo_f = s.lo_f[q], hi_r = s.hi_r[q], lo_r = s.lo_r[q];
        const uint32_t any = care_f | care_r;
#pragma unroll
        for (int nib = 0; nib < 8; ++nib) {
            // four don't-care template positions in a row (inside a variable region, or past the
            // end of the template) are skipped with one uniform branch
            if (((any >> (4 * nib)) & 0xFu) == 0) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int sft = 4 * nib + k;
                if (!((any >> sft) & 1u)) continue;
                const uint32_t wh = __funnelshift_r(h0, h1, sft);
                const uint32_t wl = __funnelshift_r(l0, l1, sft);
                const uint32_t wn = __funnelshift_r(n0, n1, sft);
                if ((care_f >> sft) & 1u) {
                    const uint32_t th = ((hi_f >> sft) & 1u) ? 0xFFFFFFFFu : 0u;
                    const uint32_t tl = ((lo_f >> sft) & 1u) ? 0xFFFFFFFFu : 0u;
                    cf.add(wn | (wh ^ th) | (wl ^ tl));
                }
                if ((care_r >> sft) & 1u) {
                    const uint32_t th = ((hi_r >> sft) & 1u) ? 0xFFFFFFFFu : 0u;
                    const uint32_t tl = ((lo_r >> sft) & 1u) ? 0xFFFFFFFFu : 0u;
                    cr.add(wn | (wh ^ th) | (wl ^ tl));
                }
            }
        }
        h0 = h1;
        l0 = l1;
        n0 = n1;
    }
}

} // namespace scg
