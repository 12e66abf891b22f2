// Kernel launchers shared by the file-level entry points and the resident plans: each picks the run-time specialised
// kernel (spec_handlers.cuh, with its follow-up kernels) when the batch allows and the generic kernel otherwise, and
// records which in Context::kernel_note.  Everything is enqueued on `stream`; nothing synchronises.
#pragma once

#include "api_common.hpp"
#include "matchers.hpp"

namespace scg {

// countDualBarcodes over one batch of pairs (DualBarcodesPairedEnd::process, handlers/DualBarcodesPairedEnd.hpp:353-381)
void launch_dual_pe(Context& ctx, const ReadsDev& r1, const ReadsDev& r2, const DualPEMatcher& m, int32_t* d_counts, int32_t* d_index,
                    cudaStream_t stream);

// countComboBarcodes over one batch (CombinatorialBarcodesSingleEnd::process, handlers/CombinatorialBarcodesSingleEnd.hpp:200-258);
// skip_if_found: the diagnostics pass of the dual single-end design
void launch_combo(Context& ctx, const ReadsDev& reads, const ComboMatcher& m, const ComboSink& sink, const int32_t* skip_if_found,
                  int32_t* out_pairs, cudaStream_t stream);

// countRandomBarcodes over one batch (RandomBarcodeSingleEnd::process, handlers/RandomBarcodeSingleEnd.hpp:122-181);
// the caller has made room in `tab` (CountTable::ensure)
void launch_random(Context& ctx, const ReadsDev& reads, const RandomMatcher& m, CountTable& tab, const uint8_t* odd, OddOutcome* odd_out,
                   unsigned long long* odd_count, int32_t* out_index, cudaStream_t stream);

} // namespace scg
