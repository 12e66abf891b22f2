// countDualBarcodes (reference src/count_dual_barcodes.cpp:12-116) and countPairedComboBarcodes
// (src/count_combo_barcodes_paired.cpp:12-95).
#include <algorithm>
#include <chrono>
#include <cstring>

#include "api_common.hpp"
#include "handlers.cuh"
#include "launchers.hpp"
#include "matchers.hpp"

namespace scg {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void launch_combo_pe(Context& ctx, const ReadsDev& r1, const ReadsDev& r2, const ComboPEParams& P, const ComboSink& sink,
                            int32_t* counters, const int32_t* skip_if_found, int32_t* out_pairs, int32_t* out_code) {
    if (r1.n <= 0) return;
    const long long ntiles = (r1.n + TILE - 1) / TILE;
    const int grid = ctx.grid_for(ntiles);
    const int cb = std::max(P.m1.spec.cbits, P.m2.spec.cbits);
    const int kw = std::max(P.m1.kw, P.m2.kw);
    dispatch_cb(cb, [&](auto CB) {
        dispatch_kw(kw, [&](auto KW) {
            combo_pe_kernel<decltype(CB)::value, decltype(KW)::value><<<grid, 128, 0, ctx.stream>>>(r1, r2, P, sink, counters, skip_if_found,
                                                                                                 out_pairs, out_code);
        });
    });
    SCG_CUDA_CHECK(cudaGetLastError());
    ++ctx.launches;
    ++ctx.timing.launches;
    ctx.kernel_note = "generic combo_pe_kernel (two single-barcode searches per pair)";
}

// SingleBarcodePairedEnd::process (reference handlers/SingleBarcodePairedEnd.hpp:93-124) from the two mates' single-barcode
// outcomes: first mode takes read 1's match, else read 2's; best mode takes the match with fewer mismatches, and on equal
// mismatches only when both reads name the same barcode.
__global__ void single_paired_combine(const int32_t* __restrict__ idx1, const uint32_t* __restrict__ info1, const int32_t* __restrict__ idx2,
                                      const uint32_t* __restrict__ info2, long long n, int use_first, int32_t* __restrict__ counts,
                                      int32_t* __restrict__ out_index) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = idx1[i], b = idx2[i];
    int chosen = -1;
    if (use_first) {
        chosen = a >= 0 ? a : b;
    } else if (a >= 0 && b < 0) {
        chosen = a;
    } else if (a < 0 && b >= 0) {
        chosen = b;
    } else if (a >= 0 && b >= 0) {
        const uint32_t m1 = (info1[i] >> 20) & 31u, m2 = (info2[i] >> 20) & 31u;
        if (m1 < m2) {
            chosen = a;
        } else if (m1 > m2) {
            chosen = b;
        } else if (a == b) {
            chosen = a;
        }
    }
    if (chosen >= 0) atomicAdd(counts + chosen, 1);
    if (out_index) out_index[i] = chosen;
}

struct PairedSources {
    Source s1, s2;
    PairedSources(const scg_source* a, const scg_source* b) : s1(a), s2(b) {}
};

} // namespace scg

using namespace scg;

extern "C" {

int scg_count_combo_paired(scg_ctx* ctx, const scg_source* src1, const char* constant1, int reverse1, int mismatches1,
                           const char* const* pool1, int npool1, const scg_source* src2, const char* constant2, int reverse2,
                           int mismatches2, const char* const* pool2, int npool2, int randomized, int use_first, int nthreads,
                           int want_trace, scg_result** table, int32_t* total, int32_t* barcode1_only, int32_t* barcode2_only) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        Source s1(src1);
        Pool p1(pool1, npool1);
        Source s2(src2);
        Pool p2(pool2, npool2);
        auto key_for = [&](int kind, int dup) {
            CacheKey k;
            const int header[8] = { kind, reverse1, mismatches1, reverse2, mismatches2, randomized, use_first, dup };
            k.feed(header, sizeof header);
            k.feed(std::string(constant1));
            k.feed(std::string(constant2));
            k.feed(p1);
            k.feed(p2);
            return k;
        };
        const std::shared_ptr<ComboPEMatcher> mp = cached_matcher<ComboPEMatcher>(c, key_for(/* combo paired */ 2, 0), [&] {
            auto built = std::make_shared<ComboPEMatcher>();
            built->prepare(constant1, reverse1 != 0, mismatches1, p1, constant2, reverse2 != 0, mismatches2, p2, randomized != 0,
                           use_first != 0, Duplicates::ERROR);
            c.ensure_ready();
            built->upload(c);
            return built;
        });
        ComboPEMatcher& m = *mp;
        c.ensure_ready();
        ComboTally tally;
        tally.init(c, npool1, npool2);
        DeviceBuffer d_counters, d_pairs, d_code;
        d_counters.alloc(2 * sizeof(int32_t), true);
        std::vector<int32_t> trace_pairs, trace_code;

        ReadPipeline pipe(c, s1.reader.get(), s2.reader.get(), nthreads, false);
        ReadPipeline::Batch b;
        long long npairs = 0;
        while (pipe.next(b)) {
            if (want_trace) {
                d_pairs.reserve((size_t)b.n * 2 * sizeof(int32_t));
                d_code.reserve((size_t)b.n * sizeof(int32_t));
            }
            launch_combo_pe(c, b.reads1, b.reads2, m.params, tally.sink(c, b.n), d_counters.as<int32_t>(), nullptr,
                            want_trace ? d_pairs.as<int32_t>() : nullptr, want_trace ? d_code.as<int32_t>() : nullptr);
            pipe.submitted(b);
            if (want_trace) {
                const size_t at = trace_pairs.size(), ac = trace_code.size();
                trace_pairs.resize(at + (size_t)b.n * 2);
                trace_code.resize(ac + (size_t)b.n);
                SCG_CUDA_CHECK(cudaMemcpyAsync(trace_pairs.data() + at, d_pairs.ptr, (size_t)b.n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaMemcpyAsync(trace_code.data() + ac, d_code.ptr, (size_t)b.n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            }
            npairs += b.n;
        }
        int32_t only[2] = { 0, 0 };
        SCG_CUDA_CHECK(cudaMemcpyAsync(only, d_counters.ptr, sizeof only, cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        auto* r = new scg_result;
        tally.harvest(c, *r);
        if (want_trace) {
            r->trace_width = 2;
            r->trace_index.swap(trace_pairs);
            r->trace_info.assign(trace_code.begin(), trace_code.end());
        }
        *table = r;
        *total = (int32_t)npairs;
        *barcode1_only = only[0];
        *barcode2_only = only[1];
        c.timing.parse_s = s1.reader->parse_seconds() + s2.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}

int scg_count_dual(scg_ctx* ctx, const scg_source* src1, const char* constant1, int reverse1, int mismatches1, const char* const* pool1,
                   int npool1, const scg_source* src2, const char* constant2, int reverse2, int mismatches2, const char* const* pool2,
                   int npool2, int randomized, int use_first, int diagnostics, int nthreads, int want_trace, int32_t* counts,
                   int32_t* total, scg_result** table, int32_t* barcode1_only, int32_t* barcode2_only) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        Source s1(src1);
        Pool p1(pool1, npool1);
        Source s2(src2);
        Pool p2(pool2, npool2);
        auto key_for = [&](int kind, int dup) {
            CacheKey k;
            const int header[8] = { kind, reverse1, mismatches1, reverse2, mismatches2, randomized, use_first, dup };
            k.feed(header, sizeof header);
            k.feed(std::string(constant1));
            k.feed(std::string(constant2));
            k.feed(p1);
            k.feed(p2);
            return k;
        };
        const std::shared_ptr<DualPEMatcher> mp = cached_matcher<DualPEMatcher>(c, key_for(/* dual paired */ 1, 0), [&] {
            auto built = std::make_shared<DualPEMatcher>();
            built->prepare(constant1, reverse1 != 0, mismatches1, p1, constant2, reverse2 != 0, mismatches2, p2, randomized != 0, use_first != 0);
            c.ensure_ready();
            built->upload(c);
            return built;
        });
        DualPEMatcher& m = *mp;
        std::shared_ptr<ComboPEMatcher> combop;
        if (diagnostics) {
            // DualBarcodesPairedEndWithDiagnostics (reference handlers/DualBarcodesPairedEndWithDiagnostics.hpp:53-72):
            // the combinatorial sub-handler allows duplicated single barcodes (DuplicateAction::FIRST)
            combop = cached_matcher<ComboPEMatcher>(c, key_for(/* combo paired */ 2, 1), [&] {
                auto built = std::make_shared<ComboPEMatcher>();
                built->prepare(constant1, reverse1 != 0, mismatches1, p1, constant2, reverse2 != 0, mismatches2, p2, randomized != 0,
                               use_first != 0, Duplicates::FIRST);
                c.ensure_ready();
                built->upload(c);
                return built;
            });
        }
        c.ensure_ready();
        ComboTally tally;
        DeviceBuffer d_counters;
        d_counters.alloc(2 * sizeof(int32_t), true);
        if (diagnostics) tally.init(c, npool1, npool2);
        DeviceBuffer d_counts, d_index;
        d_counts.alloc((size_t)std::max(npool1, 1) * sizeof(int32_t), true);
        const bool need_index = diagnostics || want_trace;
        std::vector<int32_t> trace_index;

        ReadPipeline pipe(c, s1.reader.get(), s2.reader.get(), nthreads, false);
        ReadPipeline::Batch b;
        long long npairs = 0;
        while (pipe.next(b)) {
            if (need_index) d_index.reserve((size_t)b.n * sizeof(int32_t));
            launch_dual_pe(c, b.reads1, b.reads2, m, d_counts.as<int32_t>(), need_index ? d_index.as<int32_t>() : nullptr, c.stream);
            const std::string dual_note = c.kernel_note;
            if (diagnostics) {
                // pairs without a valid combination go to the combinatorial handler (:115-120)
                launch_combo_pe(c, b.reads1, b.reads2, combop->params, tally.sink(c, b.n), d_counters.as<int32_t>(), d_index.as<int32_t>(),
                                nullptr, nullptr);
                c.kernel_note = dual_note + "; diagnostics: " + c.kernel_note;
            }
            pipe.submitted(b);
            if (want_trace) {
                const size_t at = trace_index.size();
                trace_index.resize(at + (size_t)b.n);
                SCG_CUDA_CHECK(cudaMemcpyAsync(trace_index.data() + at, d_index.ptr, (size_t)b.n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            }
            npairs += b.n;
        }
        SCG_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts.ptr, (size_t)npool1 * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        int32_t only[2] = { 0, 0 };
        SCG_CUDA_CHECK(cudaMemcpyAsync(only, d_counters.ptr, sizeof only, cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        *total = (int32_t)npairs;
        if (barcode1_only) *barcode1_only = only[0];
        if (barcode2_only) *barcode2_only = only[1];
        if (table) {
            auto* r = new scg_result;
            if (diagnostics) tally.harvest(c, *r);
            if (want_trace) {
                r->trace_width = 1;
                r->trace_index.swap(trace_index);
            }
            *table = r;
        }
        c.timing.parse_s = s1.reader->parse_seconds() + s2.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}


// Raw SegmentedBarcodeSearch<2>::search with per-query caps (reference BarcodeSearch.hpp:478-487, cache-free): the device
// search behind countDualBarcodes, exposed for the tests.  caps: two per query.
int scg_search_segmented(scg_ctx* ctx, const char* const* sequences, int nsequences, const int32_t* caps, const char* const* choices,
                         int nchoices, int len1, int len2, int max1, int max2, int32_t* index, int32_t* mismatches) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        Pool pool(choices, nchoices);
        if (pool.length != len1 + len2) throw Error("choices should be len1 + len2 bases long");
        if (max1 >= 2 && len1 + len2 > TRIE_MAX_LEN) throw Error("segments too long for the trie walk");
        LibraryOptions opt;
        opt.segmented = true;
        opt.seg1 = len1;
        opt.max_mismatches1 = max1;
        opt.max_mismatches2 = max2;
        opt.duplicates = Duplicates::ERROR;
        DeviceLibrary lib;
        lib.host = Library(pool.seqs, len1 + len2, opt);
        Pool queries(sequences, nsequences);
        if (nsequences == 0) return;
        if (queries.length < len1 + len2) throw Error("sequences are shorter than the choices");
        c.ensure_ready();
        lib.upload(c);
        const int KW = lib.host.KW;
        DeviceBuffer d_lib, d_q, d_caps, d_idx, d_mm;
        const LibDev* libp = upload_lib_array(c, std::vector<LibDev>{ lib.dev }, d_lib);
        d_caps.upload(caps, (size_t)nsequences * 2 * sizeof(int32_t), c.stream);
        d_idx.alloc((size_t)nsequences * sizeof(int32_t), false);
        d_mm.alloc((size_t)nsequences * sizeof(int32_t), false);
        const int grid = (nsequences + 127) / 128;
        dispatch_kw(KW, [&](auto KWC) {
            constexpr int K = decltype(KWC)::value;
            std::vector<uint32_t> qk((size_t)nsequences * 3 * K, 0);
            for (int i = 0; i < nsequences; ++i) {
                uint32_t* base = &qk[(size_t)i * 3 * K];
                pack_key(queries.seqs[i].data(), len1 + len2, base, base + K, base + 2 * K);
            }
            d_q.upload(qk.data(), qk.size() * sizeof(uint32_t), c.stream);
            SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            segmented_probe_kernel<K><<<grid, 128, 0, c.stream>>>(d_q.as<uint32_t>(), nsequences, libp, d_caps.as<int32_t>(), d_idx.as<int32_t>(),
                                                                  d_mm.as<int32_t>());
        });
        SCG_CUDA_CHECK(cudaGetLastError());
        ++c.launches;
        SCG_CUDA_CHECK(cudaMemcpyAsync(index, d_idx.ptr, (size_t)nsequences * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaMemcpyAsync(mismatches, d_mm.ptr, (size_t)nsequences * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}


// SingleBarcodePairedEnd (reference handlers/SingleBarcodePairedEnd.hpp:27-170): one barcode per PAIR, looked for on read 1
// and on read 2 with the same template and pool.  kaori has the handler; screenCounter exports no R function for it.
int scg_count_single_paired(scg_ctx* ctx, const scg_source* src1, const scg_source* src2, const char* constant, int strand,
                            const char* const* pool, int npool, int mismatches, int use_first, int nthreads, int32_t* counts,
                            int32_t* total, scg_result** trace) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        if (mismatches > 31) throw Error("SingleBarcodePairedEnd with more than 31 mismatches is not supported by this engine");
        Source s1(src1);
        Source s2(src2);
        const std::shared_ptr<SingleMatcher> matcher = cached_single_matcher(c, constant, strand, pool, npool, mismatches, use_first != 0);
        c.ensure_ready();
        DeviceBuffer d_counts, d_scratch, d_idx1, d_idx2, d_info1, d_info2, d_out;
        d_counts.alloc((size_t)std::max(npool, 1) * sizeof(int32_t), true);
        d_scratch.alloc((size_t)std::max(npool, 1) * sizeof(int32_t), true);   // the per-mate kernels count here; not reported
        std::vector<int32_t> trace_index;

        ReadPipeline pipe(c, s1.reader.get(), s2.reader.get(), nthreads, false);
        ReadPipeline::Batch b;
        long long npairs = 0;
        while (pipe.next(b)) {
            const size_t n = (size_t)b.n;
            d_idx1.reserve(n * sizeof(int32_t));
            d_idx2.reserve(n * sizeof(int32_t));
            d_info1.reserve(n * sizeof(uint32_t));
            d_info2.reserve(n * sizeof(uint32_t));
            if (trace) d_out.reserve(n * sizeof(int32_t));
            launch_single(c, b.reads1, *matcher, d_scratch.as<int32_t>(), d_idx1.as<int32_t>(), d_info1.as<uint32_t>(), c.stream);
            launch_single(c, b.reads2, *matcher, d_scratch.as<int32_t>(), d_idx2.as<int32_t>(), d_info2.as<uint32_t>(), c.stream);
            single_paired_combine<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(d_idx1.as<int32_t>(), d_info1.as<uint32_t>(), d_idx2.as<int32_t>(),
                                                                                     d_info2.as<uint32_t>(), b.n, use_first ? 1 : 0,
                                                                                     d_counts.as<int32_t>(), trace ? d_out.as<int32_t>() : nullptr);
            SCG_CUDA_CHECK(cudaGetLastError());
            ++c.launches;
            pipe.submitted(b);
            if (trace) {
                const size_t at = trace_index.size();
                trace_index.resize(at + n);
                SCG_CUDA_CHECK(cudaMemcpyAsync(trace_index.data() + at, d_out.ptr, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            }
            npairs += b.n;
        }
        SCG_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts.ptr, (size_t)npool * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        *total = (int32_t)npairs;
        if (trace) {
            auto* r = new scg_result;
            r->trace_width = 1;
            r->trace_index.swap(trace_index);
            r->trace_info.assign(r->trace_index.size(), 0u);
            *trace = r;
        }
        c.timing.parse_s = s1.reader->parse_seconds() + s2.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}

} // extern "C"
