// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Thin C-ABI harness around the UNMODIFIED reference (kaori v1.1.1 / byteme
// v1.0.1 as vendored in screenCounter 1.5.1).  The reference headers are
// included by -I path from /root/reference/inst/include at build time (see
// oracle/Makefile); nothing from the reference is copied into this repo.  The
// result, oracle/_ref/libkaori_ref.so, is the ground truth every parity test
// and the `cpu_baseline` leg of bench.py compare against.
//
// Each entry point constructs exactly the handler + options that the
// reference's Rcpp glue constructs (file:line cited per function) and drives it
// through kaori::process_{single,paired}_end_data, or -- for the per-read
// "trace" variants -- through handler.process() one read at a time.

// kaori relies on Rcpp.h having pulled these in first (src/count_single_barcodes.cpp:1-4).
#include <stdexcept>
#include <string>
#include <algorithm>
#include <numeric>
#include <vector>
#include <array>
#include <cstring>
#include <memory>
#include <unordered_map>

#include "kaori/kaori.hpp"
#include "byteme/byteme.hpp"

namespace {

thread_local std::string g_err;

struct Source {
    std::unique_ptr<byteme::Reader> reader;
    Source(const char* path, const char* data, size_t size) {
        if (path != nullptr) {
            reader.reset(new byteme::SomeFileReader(path)); // src/count_single_barcodes.cpp:30
        } else {
            reader.reset(new byteme::RawBufferReader(data, size));
        }
    }
    byteme::Reader* get() { return reader.get(); }
};

kaori::BarcodePool make_pool(const char* const* seqs, int n) {
    // src/utils.cpp:5-23 (format_pointers): all sequences must share one length.
    std::vector<const char*> ptrs(n);
    size_t size = 0;
    for (int i = 0; i < n; ++i) {
        size_t cur = std::strlen(seqs[i]);
        if (i == 0) {
            size = cur;
        } else if (cur != size) {
            throw std::runtime_error("variable regions should all have the same length (" + std::to_string(size) + ")");
        }
        ptrs[i] = seqs[i];
    }
    return kaori::BarcodePool(std::move(ptrs), size);
}

kaori::SearchStrand to_strand_int(int strand) { // src/utils.cpp:33-41
    if (strand == 0) return kaori::SearchStrand::FORWARD;
    if (strand == 1) return kaori::SearchStrand::REVERSE;
    return kaori::SearchStrand::BOTH;
}

kaori::SearchStrand to_strand_bool(int reverse) { // src/utils.cpp:25-31
    return reverse ? kaori::SearchStrand::REVERSE : kaori::SearchStrand::FORWARD;
}

// src/count_single_barcodes.cpp:37-47 -- run-time dispatch on template length.
template<class F>
void dispatch(size_t len, F&& f) {
    if (len <= 32) {
        f(std::integral_constant<size_t, 32>());
    } else if (len <= 64) {
        f(std::integral_constant<size_t, 64>());
    } else if (len <= 128) {
        f(std::integral_constant<size_t, 128>());
    } else if (len <= 256) {
        f(std::integral_constant<size_t, 256>());
    } else {
        throw std::runtime_error("lacking compile-time support for constant regions longer than 256 bp");
    }
}

template<class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

// Read the whole FASTQ into a ChunkOfReads-like flat store (for trace modes).
struct AllReads {
    std::vector<char> buf;
    std::vector<size_t> off{0};
    size_t size() const { return off.size() - 1; }
    std::pair<const char*, const char*> get(size_t i) const {
        return std::make_pair(buf.data() + off[i], buf.data() + off[i + 1]);
    }
};

void slurp(byteme::Reader* r, AllReads& out) {
    kaori::FastqReader fq(r);
    while (fq()) {
        const auto& s = fq.get_sequence();
        out.buf.insert(out.buf.end(), s.begin(), s.end());
        out.off.push_back(out.buf.size());
    }
}

} // namespace

struct KrefTable {
    // combos: keys = V ints per entry; random: strings of `width` chars per entry.
    int width = 0;
    std::vector<int> keys;
    std::vector<char> strings;
    std::vector<int> freq;
};

namespace {

// src/utils.h:14-45 (count_combinations): run-length encode the sorted combinations.
template<size_t V>
KrefTable* rle(const std::vector<std::array<int, V> >& sorted) {
    auto* t = new KrefTable;
    t->width = V;
    for (size_t i = 0; i < sorted.size(); ++i) {
        if (i == 0 || sorted[i] != sorted[i - 1]) {
            t->keys.insert(t->keys.end(), sorted[i].begin(), sorted[i].end());
            t->freq.push_back(1);
        } else {
            ++(t->freq.back());
        }
    }
    return t;
}

} // namespace

extern "C" {

const char* kref_last_error() { return g_err.c_str(); }

size_t kref_table_size(const KrefTable* t) { return t->freq.size(); }
int kref_table_width(const KrefTable* t) { return t->width; }
void kref_table_copy(const KrefTable* t, int* keys, char* strings, int* freq) {
    if (keys && !t->keys.empty()) std::memcpy(keys, t->keys.data(), t->keys.size() * sizeof(int));
    if (strings && !t->strings.empty()) std::memcpy(strings, t->strings.data(), t->strings.size());
    if (freq && !t->freq.empty()) std::memcpy(freq, t->freq.data(), t->freq.size() * sizeof(int));
}
void kref_table_free(KrefTable* t) { delete t; }

// Number of reads the reference's own parser sees (FastqReader.hpp:42-110).
int kref_count_reads(const char* path, const char* data, size_t size, long long* nreads, long long* nbases) {
    return guarded([&] {
        Source src(path, data, size);
        kaori::FastqReader fq(src.get());
        long long n = 0, b = 0;
        while (fq()) {
            ++n;
            b += fq.get_sequence().size();
        }
        *nreads = n;
        *nbases = b;
    });
}

// Parser output verbatim: concatenated sequences + offsets (caller sized via kref_count_reads).
int kref_parse(const char* path, const char* data, size_t size, char* bases, long long* offsets) {
    return guarded([&] {
        Source src(path, data, size);
        kaori::FastqReader fq(src.get());
        long long n = 0, b = 0;
        offsets[0] = 0;
        while (fq()) {
            const auto& s = fq.get_sequence();
            std::memcpy(bases + b, s.data(), s.size());
            b += s.size();
            offsets[++n] = b;
        }
    });
}

// ---------------------------------------------------------------------------
// countSingleBarcodes: src/count_single_barcodes.cpp:12-50
// ---------------------------------------------------------------------------
int kref_count_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                      const char* const* pool, int npool, int mismatches, int use_first, int nthreads,
                      int* counts, int* total) {
    return guarded([&] {
        Source src(path, data, size);
        auto bp = make_pool(pool, npool);
        std::string constant(tmpl);
        dispatch(constant.size(), [&](auto N) {
            typename kaori::SingleBarcodeSingleEnd<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            opt.use_first = use_first;
            kaori::SingleBarcodeSingleEnd<N.value> handler(constant.c_str(), constant.size(), bp, opt);
            kaori::process_single_end_data(src.get(), handler, nthreads);
            const auto& c = handler.get_counts();
            std::copy(c.begin(), c.end(), counts);
            *total = handler.get_total();
        });
    });
}

// Per-read outcome of SimpleSingleMatch::search_first / search_best
// (SimpleSingleMatch.hpp:200-306).  info = 4 ints per read:
// position, reverse, mismatches, variable_mismatches (valid only when index >= 0).
int kref_trace_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                      const char* const* pool, int npool, int mismatches, int use_first,
                      int* index, int* info, long long capacity, long long* nreads) {
    return guarded([&] {
        Source src(path, data, size);
        auto bp = make_pool(pool, npool);
        std::string constant(tmpl);
        AllReads reads;
        slurp(src.get(), reads);
        *nreads = reads.size();
        if ((long long)reads.size() > capacity) throw std::runtime_error("trace capacity too small");
        dispatch(constant.size(), [&](auto N) {
            typename kaori::SimpleSingleMatch<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            kaori::SimpleSingleMatch<N.value> matcher(constant.c_str(), constant.size(), bp, opt);
            auto state = matcher.initialize();
            for (size_t i = 0; i < reads.size(); ++i) {
                auto x = reads.get(i);
                bool found = use_first ? matcher.search_first(x.first, x.second - x.first, state)
                                       : matcher.search_best(x.first, x.second - x.first, state);
                index[i] = found ? state.index : -1;
                if (info) {
                    info[4 * i + 0] = found ? (int)state.position : -1;
                    info[4 * i + 1] = found ? (int)state.reverse : 0;
                    info[4 * i + 2] = found ? state.mismatches : -1;
                    info[4 * i + 3] = found ? state.variable_mismatches : -1;
                }
            }
        });
    });
}

// ---------------------------------------------------------------------------
// countRandomBarcodes: src/count_random_barcodes.cpp:12-60.  Output sorted by
// sequence (byte order) as R/countRandomBarcodes.R:73-74 does.
// ---------------------------------------------------------------------------
int kref_count_random(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                      int mismatches, int use_first, int nthreads, KrefTable** table, int* total) {
    return guarded([&] {
        Source src(path, data, size);
        std::string constant(tmpl);
        dispatch(constant.size(), [&](auto N) {
            typename kaori::RandomBarcodeSingleEnd<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            opt.use_first = use_first;
            kaori::RandomBarcodeSingleEnd<N.value> handler(constant.c_str(), constant.size(), opt);
            kaori::process_single_end_data(src.get(), handler, nthreads);
            const auto& c = handler.get_counts();
            std::vector<std::pair<std::string, int> > sorted(c.begin(), c.end());
            std::sort(sorted.begin(), sorted.end());
            auto* t = new KrefTable;
            t->width = sorted.empty() ? 0 : (int)sorted.front().first.size();
            for (const auto& p : sorted) {
                t->strings.insert(t->strings.end(), p.first.begin(), p.first.end());
                t->freq.push_back(p.second);
            }
            *table = t;
            *total = handler.get_total();
        });
    });
}

// ---------------------------------------------------------------------------
// countComboBarcodes: src/count_combo_barcodes_single.cpp:12-70 (V = 2 only)
// ---------------------------------------------------------------------------
int kref_count_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                            const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                            int mismatches, int use_first, int nthreads, KrefTable** table, int* total) {
    return guarded([&] {
        Source src(path, data, size);
        std::array<kaori::BarcodePool, 2> opts{ make_pool(pool1, npool1), make_pool(pool2, npool2) };
        std::string constant(tmpl);
        dispatch(constant.size(), [&](auto N) {
            typename kaori::CombinatorialBarcodesSingleEnd<N.value, 2>::Options opt;
            opt.max_mismatches = mismatches;
            opt.strand = to_strand_int(strand);
            opt.use_first = use_first;
            kaori::CombinatorialBarcodesSingleEnd<N.value, 2> handler(constant.c_str(), constant.size(), opts, opt);
            kaori::process_single_end_data(src.get(), handler, nthreads);
            handler.sort();
            *table = rle(handler.get_combinations());
            *total = handler.get_total();
        });
    });
}

// Per-read outcome: combo[2*i..] = (i1, i2) or (-1, -1).
int kref_trace_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                            const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                            int mismatches, int use_first, int* combo, long long capacity, long long* nreads) {
    return guarded([&] {
        Source src(path, data, size);
        std::array<kaori::BarcodePool, 2> opts{ make_pool(pool1, npool1), make_pool(pool2, npool2) };
        std::string constant(tmpl);
        AllReads reads;
        slurp(src.get(), reads);
        *nreads = reads.size();
        if ((long long)reads.size() > capacity) throw std::runtime_error("trace capacity too small");
        dispatch(constant.size(), [&](auto N) {
            typename kaori::CombinatorialBarcodesSingleEnd<N.value, 2>::Options opt;
            opt.max_mismatches = mismatches;
            opt.strand = to_strand_int(strand);
            opt.use_first = use_first;
            kaori::CombinatorialBarcodesSingleEnd<N.value, 2> handler(constant.c_str(), constant.size(), opts, opt);
            auto state = handler.initialize();
            for (size_t i = 0; i < reads.size(); ++i) {
                size_t before = state.collected.size();
                handler.process(state, reads.get(i));
                if (state.collected.size() > before) {
                    combo[2 * i] = state.collected.back()[0];
                    combo[2 * i + 1] = state.collected.back()[1];
                } else {
                    combo[2 * i] = combo[2 * i + 1] = -1;
                }
            }
        });
    });
}

// ---------------------------------------------------------------------------
// countDualBarcodesSingleEnd: src/count_dual_barcodes_single_end.cpp:12-88
// pools = npools arrays of nchoices strings, flattened pool-major.
// With diagnostics: table = invalid combinations (V = 2 only, as in the glue).
// ---------------------------------------------------------------------------
int kref_count_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                               const char* const* pools_flat, int npools, int nchoices, int strand,
                               int mismatches, int use_first, int diagnostics, int nthreads,
                               int* counts, int* total, KrefTable** table) {
    return guarded([&] {
        Source src(path, data, size);
        std::vector<kaori::BarcodePool> pools;
        for (int p = 0; p < npools; ++p) {
            pools.push_back(make_pool(pools_flat + (size_t)p * nchoices, nchoices));
        }
        std::string constant(tmpl);
        dispatch(constant.size(), [&](auto N) {
            typename kaori::DualBarcodesSingleEnd<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            opt.use_first = use_first;
            if (!diagnostics) {
                kaori::DualBarcodesSingleEnd<N.value> handler(constant.c_str(), constant.size(), pools, opt);
                kaori::process_single_end_data(src.get(), handler, nthreads);
                const auto& c = handler.get_counts();
                std::copy(c.begin(), c.end(), counts);
                *total = handler.get_total();
            } else {
                kaori::DualBarcodesSingleEndWithDiagnostics<N.value, 2> handler(constant.c_str(), constant.size(), pools, opt);
                kaori::process_single_end_data(src.get(), handler, nthreads);
                const auto& c = handler.get_counts();
                std::copy(c.begin(), c.end(), counts);
                handler.sort();
                *table = rle(handler.get_combinations());
                *total = handler.get_total();
            }
        });
    });
}

int kref_trace_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                               const char* const* pools_flat, int npools, int nchoices, int strand,
                               int mismatches, int use_first, int* index, long long capacity, long long* nreads) {
    return guarded([&] {
        Source src(path, data, size);
        std::vector<kaori::BarcodePool> pools;
        for (int p = 0; p < npools; ++p) {
            pools.push_back(make_pool(pools_flat + (size_t)p * nchoices, nchoices));
        }
        std::string constant(tmpl);
        AllReads reads;
        slurp(src.get(), reads);
        *nreads = reads.size();
        if ((long long)reads.size() > capacity) throw std::runtime_error("trace capacity too small");
        dispatch(constant.size(), [&](auto N) {
            typename kaori::DualBarcodesSingleEnd<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            opt.use_first = use_first;
            kaori::DualBarcodesSingleEnd<N.value> handler(constant.c_str(), constant.size(), pools, opt);
            auto state = handler.initialize();
            for (size_t i = 0; i < reads.size(); ++i) {
                bool found = handler.process(state, reads.get(i));
                index[i] = -1;
                if (found) {
                    for (size_t k = 0; k < state.counts.size(); ++k) {
                        if (state.counts[k]) {
                            index[i] = k;
                            state.counts[k] = 0;
                            break;
                        }
                    }
                }
            }
        });
    });
}

// ---------------------------------------------------------------------------
// countDualBarcodes (paired-end): src/count_dual_barcodes.cpp:12-116
// ---------------------------------------------------------------------------
int kref_count_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                    const char* const* pool1, int npool1,
                    const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                    const char* const* pool2, int npool2,
                    int randomized, int use_first, int diagnostics, int nthreads,
                    int* counts, int* total, KrefTable** table, int* b1only, int* b2only) {
    return guarded([&] {
        Source src1(path1, data1, size1);
        auto bp1 = make_pool(pool1, npool1);
        Source src2(path2, data2, size2);
        auto bp2 = make_pool(pool2, npool2);
        std::string constant1(tmpl1), constant2(tmpl2);
        dispatch(std::max(constant1.size(), constant2.size()), [&](auto N) {
            typename kaori::DualBarcodes<N.value>::Options opt;
            opt.strand1 = to_strand_bool(reverse1);
            opt.max_mismatches1 = mismatches1;
            opt.strand2 = to_strand_bool(reverse2);
            opt.max_mismatches2 = mismatches2;
            opt.random = randomized;
            opt.use_first = use_first;
            if (!diagnostics) {
                kaori::DualBarcodes<N.value> handler(constant1.c_str(), constant1.size(), bp1, constant2.c_str(), constant2.size(), bp2, opt);
                kaori::process_paired_end_data(src1.get(), src2.get(), handler, nthreads);
                const auto& c = handler.get_counts();
                std::copy(c.begin(), c.end(), counts);
                *total = handler.get_total();
            } else {
                kaori::DualBarcodesWithDiagnostics<N.value> handler(constant1.c_str(), constant1.size(), bp1, constant2.c_str(), constant2.size(), bp2, opt);
                kaori::process_paired_end_data(src1.get(), src2.get(), handler, nthreads);
                handler.sort();
                const auto& c = handler.get_counts();
                std::copy(c.begin(), c.end(), counts);
                *table = rle(handler.get_combinations());
                *total = handler.get_total();
                *b1only = handler.get_barcode1_only();
                *b2only = handler.get_barcode2_only();
            }
        });
    });
}

// Per-pair outcome of DualBarcodesPairedEnd::process (handlers/DualBarcodesPairedEnd.hpp:353-381).
// fresh_state != 0: the search cache is emptied before every pair, so a pair's outcome
// no longer depends on the pairs before it ("Quirk C"); the cache still
// acts WITHIN a pair (the same variable strings at two hit positions with different
// caps), which kaori_port.c reproduces in its CACHE_PER_PAIR mode.  fresh_state == 0:
// one state re-used for the whole file, which is what num.threads = 1 gives an R user.
int kref_trace_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                    const char* const* pool1, int npool1,
                    const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                    const char* const* pool2, int npool2,
                    int randomized, int use_first, int fresh_state,
                    int* index, long long capacity, long long* npairs) {
    return guarded([&] {
        Source src1(path1, data1, size1);
        auto bp1 = make_pool(pool1, npool1);
        Source src2(path2, data2, size2);
        auto bp2 = make_pool(pool2, npool2);
        std::string constant1(tmpl1), constant2(tmpl2);
        AllReads r1, r2;
        slurp(src1.get(), r1);
        slurp(src2.get(), r2);
        if (r1.size() != r2.size()) throw std::runtime_error("different number of reads in paired FASTQ files");
        *npairs = r1.size();
        if ((long long)r1.size() > capacity) throw std::runtime_error("trace capacity too small");
        dispatch(std::max(constant1.size(), constant2.size()), [&](auto N) {
            typename kaori::DualBarcodes<N.value>::Options opt;
            opt.strand1 = to_strand_bool(reverse1);
            opt.max_mismatches1 = mismatches1;
            opt.strand2 = to_strand_bool(reverse2);
            opt.max_mismatches2 = mismatches2;
            opt.random = randomized;
            opt.use_first = use_first;
            kaori::DualBarcodes<N.value> handler(constant1.c_str(), constant1.size(), bp1, constant2.c_str(), constant2.size(), bp2, opt);
            auto state = handler.initialize();
            for (size_t i = 0; i < r1.size(); ++i) {
                if (fresh_state) state.details.cache.clear();
                bool found = handler.process(state, r1.get(i), r2.get(i));
                index[i] = -1;
                if (found) {
                    for (size_t k = 0; k < state.counts.size(); ++k) {
                        if (state.counts[k]) {
                            index[i] = k;
                            state.counts[k] = 0;
                            break;
                        }
                    }
                }
            }
        });
    });
}

// ---------------------------------------------------------------------------
// countPairedComboBarcodes: src/count_combo_barcodes_paired.cpp:12-95
// ---------------------------------------------------------------------------
int kref_count_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                            const char* const* pool1, int npool1,
                            const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                            const char* const* pool2, int npool2,
                            int randomized, int use_first, int nthreads,
                            KrefTable** table, int* total, int* b1only, int* b2only) {
    return guarded([&] {
        Source src1(path1, data1, size1);
        auto bp1 = make_pool(pool1, npool1);
        Source src2(path2, data2, size2);
        auto bp2 = make_pool(pool2, npool2);
        std::string constant1(tmpl1), constant2(tmpl2);
        dispatch(std::max(constant1.size(), constant2.size()), [&](auto N) {
            typename kaori::CombinatorialBarcodesPairedEnd<N.value>::Options opt;
            opt.strand1 = to_strand_bool(reverse1);
            opt.max_mismatches1 = mismatches1;
            opt.strand2 = to_strand_bool(reverse2);
            opt.max_mismatches2 = mismatches2;
            opt.random = randomized;
            opt.use_first = use_first;
            kaori::CombinatorialBarcodesPairedEnd<N.value> handler(constant1.c_str(), constant1.size(), bp1, constant2.c_str(), constant2.size(), bp2, opt);
            kaori::process_paired_end_data(src1.get(), src2.get(), handler, nthreads);
            handler.sort();
            *table = rle(handler.get_combinations());
            *total = handler.get_total();
            *b1only = handler.get_barcode1_only();
            *b2only = handler.get_barcode2_only();
        });
    });
}

// Per-pair outcome: combo = (i1, i2) or (-1,-1); code: 0 none, 1 pair, 2 barcode1 only, 3 barcode2 only.
int kref_trace_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                            const char* const* pool1, int npool1,
                            const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                            const char* const* pool2, int npool2,
                            int randomized, int use_first,
                            int* combo, int* code, long long capacity, long long* npairs) {
    return guarded([&] {
        Source src1(path1, data1, size1);
        auto bp1 = make_pool(pool1, npool1);
        Source src2(path2, data2, size2);
        auto bp2 = make_pool(pool2, npool2);
        std::string constant1(tmpl1), constant2(tmpl2);
        AllReads r1, r2;
        slurp(src1.get(), r1);
        slurp(src2.get(), r2);
        if (r1.size() != r2.size()) throw std::runtime_error("different number of reads in paired FASTQ files");
        *npairs = r1.size();
        if ((long long)r1.size() > capacity) throw std::runtime_error("trace capacity too small");
        dispatch(std::max(constant1.size(), constant2.size()), [&](auto N) {
            typename kaori::CombinatorialBarcodesPairedEnd<N.value>::Options opt;
            opt.strand1 = to_strand_bool(reverse1);
            opt.max_mismatches1 = mismatches1;
            opt.strand2 = to_strand_bool(reverse2);
            opt.max_mismatches2 = mismatches2;
            opt.random = randomized;
            opt.use_first = use_first;
            kaori::CombinatorialBarcodesPairedEnd<N.value> handler(constant1.c_str(), constant1.size(), bp1, constant2.c_str(), constant2.size(), bp2, opt);
            auto state = handler.initialize();
            for (size_t i = 0; i < r1.size(); ++i) {
                size_t before = state.collected.size();
                int b1 = state.barcode1_only, b2 = state.barcode2_only;
                handler.process(state, r1.get(i), r2.get(i));
                combo[2 * i] = combo[2 * i + 1] = -1;
                code[i] = 0;
                if (state.collected.size() > before) {
                    combo[2 * i] = state.collected.back()[0];
                    combo[2 * i + 1] = state.collected.back()[1];
                    code[i] = 1;
                } else if (state.barcode1_only > b1) {
                    code[i] = 2;
                } else if (state.barcode2_only > b2) {
                    code[i] = 3;
                }
            }
        });
    });
}

// ---------------------------------------------------------------------------
// matchBarcodes: src/match_barcodes.cpp:7-37.  index is 0-based, -1 = NA.
// duplicates: 0 FIRST, 1 LAST, 2 NONE, 3 ERROR (utils.hpp DuplicateAction order);
// the Rcpp glue always uses ERROR (3).
// ---------------------------------------------------------------------------
int kref_match_barcodes(const char* const* seqs, int nseqs, const char* const* choices, int nchoices,
                        int substitutions, int reverse, int duplicates, int* index, int* mm) {
    return guarded([&] {
        typename kaori::SimpleBarcodeSearch::Options opt;
        opt.max_mismatches = substitutions;
        opt.reverse = reverse;
        opt.duplicates = static_cast<kaori::DuplicateAction>(duplicates);
        auto pool = make_pool(choices, nchoices);
        kaori::SimpleBarcodeSearch searcher(pool, opt);
        auto state = searcher.initialize();
        auto x = make_pool(seqs, nseqs);
        for (int i = 0; i < nseqs; ++i) {
            searcher.search(x.pool[i], state);
            if (state.index >= 0) {
                index[i] = state.index;
                mm[i] = state.mismatches;
            } else {
                index[i] = -1;
                mm[i] = -1;
            }
        }
    });
}

// Raw SimpleBarcodeSearch::search with a per-query cap (BarcodeSearch.hpp:243-251),
// fresh state per query.  index keeps kaori's -1 (missing) / -2 (ambiguous).
int kref_search_any(const char* const* seqs, int nseqs, const int* caps, const char* const* choices, int nchoices,
                    int max_mismatches, int reverse, int duplicates, int* index, int* mm) {
    return guarded([&] {
        typename kaori::SimpleBarcodeSearch::Options opt;
        opt.max_mismatches = max_mismatches;
        opt.reverse = reverse;
        opt.duplicates = static_cast<kaori::DuplicateAction>(duplicates);
        auto pool = make_pool(choices, nchoices);
        kaori::SimpleBarcodeSearch searcher(pool, opt);
        for (int i = 0; i < nseqs; ++i) {
            auto state = searcher.initialize();
            searcher.search(seqs[i], state, caps[i]);
            index[i] = state.index;
            mm[i] = state.mismatches;
        }
    });
}

// Raw SegmentedBarcodeSearch<2>::search with per-query caps (BarcodeSearch.hpp:478-487),
// fresh state per query (cache-free semantics).
int kref_search_segmented2(const char* const* seqs, int nseqs, const int* caps /* 2 per query */,
                           const char* const* choices, int nchoices, int len1, int len2,
                           int max1, int max2, int duplicates, int* index, int* mm) {
    return guarded([&] {
        typename kaori::SegmentedBarcodeSearch<2>::Options opt;
        opt.max_mismatches = { max1, max2 };
        opt.duplicates = static_cast<kaori::DuplicateAction>(duplicates);
        auto pool = make_pool(choices, nchoices);
        kaori::SegmentedBarcodeSearch<2> searcher(pool, std::array<int, 2>{ len1, len2 }, opt);
        for (int i = 0; i < nseqs; ++i) {
            auto state = searcher.initialize();
            searcher.search(seqs[i], state, std::array<int, 2>{ caps[2 * i], caps[2 * i + 1] });
            index[i] = state.index;
            mm[i] = state.mismatches;
        }
    });
}


// SingleBarcodePairedEnd (handlers/SingleBarcodePairedEnd.hpp:93-124): the single-barcode search on both mates.
int kref_count_single_paired(const char* path1, const char* data1, size_t size1, const char* path2, const char* data2, size_t size2,
                             const char* tmpl, int strand, const char* const* pool, int npool, int mismatches, int use_first,
                             int nthreads, int* counts, int* total) {
    return guarded([&] {
        Source src1(path1, data1, size1);
        Source src2(path2, data2, size2);
        auto bp = make_pool(pool, npool);
        std::string constant(tmpl);
        dispatch(constant.size(), [&](auto N) {
            typename kaori::SingleBarcodePairedEnd<N.value>::Options opt;
            opt.strand = to_strand_int(strand);
            opt.max_mismatches = mismatches;
            opt.use_first = use_first;
            kaori::SingleBarcodePairedEnd<N.value> handler(constant.c_str(), constant.size(), bp, opt);
            kaori::process_paired_end_data(src1.get(), src2.get(), handler, nthreads);
            const auto& c = handler.get_counts();
            std::copy(c.begin(), c.end(), counts);
            *total = handler.get_total();
        });
    });
}

} // extern "C"
