// countComboBarcodes (reference src/count_combo_barcodes_single.cpp:12-70) and
// countDualBarcodesSingleEnd (src/count_dual_barcodes_single_end.cpp:12-88).
#include <algorithm>
#include <chrono>
#include <cstring>

#include "api_common.hpp"
#include "handlers.cuh"

namespace scg {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// CombinatorialBarcodesSingleEnd<_, 2> (reference handlers/CombinatorialBarcodesSingleEnd.hpp:66-119).
struct ComboMatcher {
    TemplateSpec tmpl;
    DeviceLibrary lib[4];   // [2 * reverse + region]
    DeviceBuffer libs_dev;
    ComboParams params;

    void prepare(const std::string& constant, int strand, const Pool& p1, const Pool& p2, int mismatches, bool use_first, Duplicates dup) {
        tmpl = TemplateSpec(constant, strand);
        const Pool* pools[2] = { &p1, &p2 };
        if (tmpl.fwd_regions.size() != 2) throw Error("expected 2 variable regions in the constant template");
        for (int i = 0; i < 2; ++i) {
            const int rlen = tmpl.fwd_regions[i].end - tmpl.fwd_regions[i].start;
            if (pools[i]->length != rlen) {
                throw Error("length of variable region " + std::to_string(i + 1) + " (" + std::to_string(rlen) +
                            ") should be the same as its sequences (" + std::to_string(pools[i]->length) + ")");
            }
        }
        LibraryOptions opt;
        opt.max_mismatches = mismatches;
        opt.duplicates = dup;
        if (tmpl.fwd) {
            for (int r = 0; r < 2; ++r) lib[r].host = Library(pools[r]->seqs, pools[r]->length, opt);
        }
        if (tmpl.rev) {  // reversed pool order on the reverse strand (:111-116)
            for (int r = 0; r < 2; ++r) lib[2 + r].host = Library(pools[1 - r]->reverse_complemented(), pools[1 - r]->length, opt);
        }
        std::memset(&params, 0, sizeof params);
        params.spec = tmpl.scan_spec(mismatches);
        params.max_mm = mismatches;
        params.use_first = use_first ? 1 : 0;
        params.n1 = (int)p1.seqs.size();
        params.n2 = (int)p2.seqs.size();
    }

    void upload(Context& ctx) {
        std::vector<LibDev> libs(4);
        std::memset(libs.data(), 0, 4 * sizeof(LibDev));
        params.kw = 1;
        for (int k = 0; k < 4; ++k) {
            const bool used = k < 2 ? tmpl.fwd : tmpl.rev;
            if (!used) continue;
            lib[k].upload(ctx);
            libs[k] = lib[k].dev;
            params.kw = std::max(params.kw, lib[k].dev.KW);
        }
        params.libs = upload_lib_array(ctx, libs, libs_dev);
    }
};

static void launch_combo(Context& ctx, const ReadsDev& reads, const ComboParams& P, const ComboSink& sink, const int32_t* skip_if_found,
                         int32_t* out_pairs) {
    if (reads.n <= 0) return;
    const long long ntiles = (reads.n + TILE - 1) / TILE;
    const int grid = ctx.grid_for(ntiles);
    dispatch_cb(P.spec.cbits, [&](auto CB) {
        dispatch_kw(P.kw, [&](auto KW) {
            combo_kernel<decltype(CB)::value, decltype(KW)::value><<<grid, 128, 0, ctx.stream>>>(reads, P, sink, skip_if_found, out_pairs);
        });
    });
    SCG_CUDA_CHECK(cudaGetLastError());
    ++ctx.launches;
    ++ctx.timing.launches;
}

// DualBarcodesSingleEnd (reference handlers/DualBarcodesSingleEnd.hpp:64-124).
struct DualSEMatcher {
    TemplateSpec tmpl;
    DeviceLibrary lib[2];
    DeviceBuffer libs_dev;
    DualSEParams params;
    int nchoices = 0;

    void prepare(const std::string& constant, const std::vector<Pool>& pools, int nchoices_, int strand, int mismatches, bool use_first) {
        tmpl = TemplateSpec(constant, strand);
        nchoices = nchoices_;
        if (pools.size() != tmpl.fwd_regions.size()) throw Error("length of 'barcode_pools' should equal the number of variable regions");
        int klen = 0;
        for (size_t i = 0; i < pools.size(); ++i) {
            const int rlen = tmpl.fwd_regions[i].end - tmpl.fwd_regions[i].start;
            if (pools[i].length != rlen) {
                throw Error("length of variable region " + std::to_string(i + 1) + " (" + std::to_string(rlen) +
                            ") should be the same as its sequences (" + std::to_string(pools[i].length) + ")");
            }
            klen += rlen;
        }
        std::vector<std::string> combined(nchoices);  // rows concatenated across the pools (:99-108)
        for (const auto& p : pools) {
            for (int c = 0; c < nchoices; ++c) combined[c] += p.seqs[c];
        }
        LibraryOptions opt;
        opt.max_mismatches = mismatches;
        opt.duplicates = Duplicates::ERROR;
        if (tmpl.fwd) lib[0].host = Library(combined, klen, opt);
        if (tmpl.rev) {  // reverse complement of the whole row (:117-120)
            std::vector<std::string> rc;
            rc.reserve(combined.size());
            for (const auto& s : combined) rc.push_back(reverse_complement_iupac(s));
            lib[1].host = Library(rc, klen, opt);
        }
        std::memset(&params, 0, sizeof params);
        params.spec = tmpl.scan_spec(mismatches);
        params.max_mm = mismatches;
        params.use_first = use_first ? 1 : 0;
    }

    void upload(Context& ctx) {
        std::vector<LibDev> libs(2);
        std::memset(libs.data(), 0, 2 * sizeof(LibDev));
        params.kw = 1;
        for (int k = 0; k < 2; ++k) {
            const bool used = k == 0 ? tmpl.fwd : tmpl.rev;
            if (!used) continue;
            lib[k].upload(ctx);
            libs[k] = lib[k].dev;
            params.kw = std::max(params.kw, lib[k].dev.KW);
        }
        params.libs = upload_lib_array(ctx, libs, libs_dev);
    }
};

} // namespace scg

using namespace scg;

extern "C" {

int scg_count_combo_single(scg_ctx* ctx, const scg_source* src, const char* constant, int strand, const char* const* pool1, int npool1,
                           const char* const* pool2, int npool2, int mismatches, int use_first, int nthreads, int want_trace,
                           scg_result** table, int32_t* total) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        Source source(src);
        Pool p1(pool1, npool1), p2(pool2, npool2);
        CacheKey key;
        {
            const int header[5] = { /* combo single */ 3, strand, mismatches, use_first, 0 };
            key.feed(header, sizeof header);
            key.feed(std::string(constant));
            key.feed(p1);
            key.feed(p2);
        }
        const std::shared_ptr<ComboMatcher> mp = cached_matcher<ComboMatcher>(c, key, [&] {
            auto built = std::make_shared<ComboMatcher>();
            built->prepare(constant, strand, p1, p2, mismatches, use_first != 0, Duplicates::ERROR);
            c.ensure_ready();
            built->upload(c);
            return built;
        });
        ComboMatcher& m = *mp;
        c.ensure_ready();
        ComboTally tally;
        tally.init(c, npool1, npool2);
        TraceSink trace;
        trace.enabled = want_trace != 0;
        trace.width = 2;

        ReadPipeline pipe(c, source.reader.get(), nullptr, nthreads, false);
        ReadPipeline::Batch b;
        long long nreads = 0;
        while (pipe.next(b)) {
            trace.prepare(b.n, false);
            launch_combo(c, b.reads1, m.params, tally.sink(c, b.n), nullptr, trace.enabled ? trace.d_index.as<int32_t>() : nullptr);
            pipe.submitted(b);
            trace.collect(c, b.n, false);
            nreads += b.n;
        }
        auto* r = new scg_result;
        tally.harvest(c, *r);
        if (trace.enabled) {
            r->trace_width = 2;
            r->trace_index.swap(trace.index);
        }
        *table = r;
        *total = (int32_t)nreads;
        c.timing.parse_s = source.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}

int scg_count_dual_single_end(scg_ctx* ctx, const scg_source* src, const char* constant, const char* const* pools_flat, int npools,
                              int nchoices, int strand, int mismatches, int use_first, int diagnostics, int nthreads, int want_trace,
                              int32_t* counts, int32_t* total, scg_result** table) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        Source source(src);
        std::vector<Pool> pools;
        for (int p = 0; p < npools; ++p) pools.emplace_back(pools_flat + (size_t)p * nchoices, nchoices);
        auto key_for = [&](int kind, int dup) {
            CacheKey k;
            const int header[6] = { kind, strand, mismatches, use_first, dup, npools };
            k.feed(header, sizeof header);
            k.feed(std::string(constant));
            for (const auto& p : pools) k.feed(p);
            return k;
        };
        const std::shared_ptr<DualSEMatcher> mp = cached_matcher<DualSEMatcher>(c, key_for(/* dual single-end */ 4, 0), [&] {
            auto built = std::make_shared<DualSEMatcher>();
            built->prepare(constant, pools, nchoices, strand, mismatches, use_first != 0);
            if (diagnostics && npools != 2) throw Error("expected 2 variable regions in the constant template");
            c.ensure_ready();
            built->upload(c);
            return built;
        });
        DualSEMatcher& m = *mp;
        std::shared_ptr<ComboMatcher> combop;
        if (diagnostics) {
            // DualBarcodesSingleEndWithDiagnostics<_, 2> (reference handlers/DualBarcodesSingleEndWithDiagnostics.hpp:44-58):
            // two variable regions, each pool searched on its own with DuplicateAction::FIRST
            if (npools != 2) throw Error("expected 2 variable regions in the constant template");
            combop = cached_matcher<ComboMatcher>(c, key_for(/* combo single */ 3, 1), [&] {
                auto built = std::make_shared<ComboMatcher>();
                built->prepare(constant, strand, pools[0], pools[1], mismatches, use_first != 0, Duplicates::FIRST);
                c.ensure_ready();
                built->upload(c);
                return built;
            });
        }
        c.ensure_ready();
        ComboTally tally;
        if (diagnostics) tally.init(c, nchoices, nchoices);
        DeviceBuffer d_counts, d_index;
        d_counts.alloc((size_t)std::max(nchoices, 1) * sizeof(int32_t), true);
        const bool need_index = diagnostics || want_trace;
        std::vector<int32_t> trace_index;

        ReadPipeline pipe(c, source.reader.get(), nullptr, nthreads, false);
        ReadPipeline::Batch b;
        long long nreads = 0;
        while (pipe.next(b)) {
            if (need_index) d_index.reserve((size_t)b.n * sizeof(int32_t));
            const long long ntiles = (b.n + TILE - 1) / TILE;
            const int grid = c.grid_for(ntiles);
            dispatch_cb(m.params.spec.cbits, [&](auto CB) {
                dispatch_kw(m.params.kw, [&](auto KW) {
                    dual_se_kernel<decltype(CB)::value, decltype(KW)::value><<<grid, 128, 0, c.stream>>>(
                        b.reads1, m.params, d_counts.as<int32_t>(), need_index ? d_index.as<int32_t>() : nullptr);
                });
            });
            SCG_CUDA_CHECK(cudaGetLastError());
            ++c.launches;
            ++c.timing.launches;
            if (diagnostics) {
                // only reads without a valid pair are tabulated (:99-104)
                launch_combo(c, b.reads1, combop->params, tally.sink(c, b.n), d_index.as<int32_t>(), nullptr);
            }
            pipe.submitted(b);
            if (want_trace) {
                const size_t at = trace_index.size();
                trace_index.resize(at + (size_t)b.n);
                SCG_CUDA_CHECK(cudaMemcpyAsync(trace_index.data() + at, d_index.ptr, (size_t)b.n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            }
            nreads += b.n;
        }
        SCG_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts.ptr, (size_t)nchoices * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        *total = (int32_t)nreads;
        if (table) {
            auto* r = new scg_result;
            if (diagnostics) tally.harvest(c, *r);
            if (want_trace) {
                r->trace_width = 1;
                r->trace_index.swap(trace_index);
            }
            *table = r;
        }
        c.timing.parse_s = source.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}

} // extern "C"
