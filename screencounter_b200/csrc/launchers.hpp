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

// countSingleBarcodes over one input on one device (runners_single.cu); counts stay on the device
long long count_single_core(Context& c, FastqReader* reader, const SingleMatcher& m, int nthreads, int32_t* d_counts, TraceSink& sink);

// countComboBarcodes / countRandomBarcodes over one input on one device (runners_combo.cu, runners_random.cu).  With want_sorted
// set, a result that can stay on the device as a sorted table is left there and *table stays null.
void count_combo_core(scg_ctx* ctx, const scg_source* src, const char* constant, int strand, const char* const* pool1, int npool1,
                      const char* const* pool2, int npool2, int mismatches, int use_first, int nthreads, int want_trace,
                      SortedTable* want_sorted, scg_result** table, int32_t* total);
void count_random_core(scg_ctx* ctx, const scg_source* src, const char* constant, int strand, int mismatches, int use_first,
                       int nthreads, SortedTable* want_sorted, scg_result** table, int32_t* total);

// countSingleBarcodes for one input (runners_single.cu): on the context's first device, or cut over all its devices
void count_single_file(scg_ctx* ctx, const scg_source* src, const char* constant, int strand, const char* const* pool, int npool,
                       int mismatches, int use_first, int nthreads, int32_t* counts, int32_t* total, scg_result** trace,
                       bool split_over_devices);

// The same over SEVERAL devices (runners_multi.cu): the text is cut at record boundaries into one contiguous part per device,
// every device counts its part on its own host thread, the count vectors are summed on the first device over peer memory.
// false = the input cannot be split (a gzip stream, too small, no safe cut, or a part did not parse cleanly): the caller runs
// the single-device path, which also produces the reference's error text where there is one.
bool count_single_multi(scg_ctx* ctx, FastqReader& reader, const char* constant, int strand, const char* const* pool, int npool,
                        int mismatches, bool use_first, int nthreads, int32_t* counts, int32_t* total, scg_result** trace);

} // namespace scg
