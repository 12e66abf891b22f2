// countRandomBarcodes (reference src/count_random_barcodes.cpp:12-60) and the device count tables
// (GPU open-addressing hash of packed keys) shared with the sparse combination tally.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>

#include <cub/device/device_radix_sort.cuh>

#include "api_common.hpp"
#include "handlers.cuh"
#include "hostpool.hpp"

namespace scg {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------
// CountTable
// ---------------------------------------------------------------------------------------
__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// re-insert the live entries of an old count table into a larger one
__global__ void rehash64_kernel(const unsigned long long* keys, const uint32_t* counts, size_t n, CountTable64 dst) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (keys[i] != ~0ull) count_insert64(dst, keys[i], counts[i]);
    }
}

__global__ void rehash128_kernel(const ulonglong2* keys, const uint32_t* counts, size_t n, CountTable128 dst) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!(keys[i].x == ~0ull && keys[i].y == ~0ull)) count_insert128(dst, keys[i], counts[i]);
    }
}

// compact the live entries of a count table: out_keys/out_counts sized by the live count
__global__ void compact64_kernel(const unsigned long long* keys, const uint32_t* counts, size_t n,
                                 unsigned long long* out_keys, uint32_t* out_counts, unsigned long long* cursor) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (keys[i] != ~0ull) {
            const unsigned long long at = atomicAdd(cursor, 1ull);
            out_keys[at] = keys[i];
            out_counts[at] = counts[i];
        }
    }
}

__global__ void compact128_kernel(const ulonglong2* keys, const uint32_t* counts, size_t n,
                                  ulonglong2* out_keys, uint32_t* out_counts, unsigned long long* cursor) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!(keys[i].x == ~0ull && keys[i].y == ~0ull)) {
            const unsigned long long at = atomicAdd(cursor, 1ull);
            out_keys[at] = keys[i];
            out_counts[at] = counts[i];
        }
    }
}

static void fill_empty(Context& ctx, void* keys, size_t n_u64) {
    fill_u64_kernel<<<(int)std::min<size_t>((n_u64 + 255) / 256, (size_t)ctx.sm_count * 32), 256, 0, ctx.stream>>>(
        static_cast<unsigned long long*>(keys), ~0ull, n_u64);
    SCG_CUDA_CHECK(cudaGetLastError());
    ++ctx.launches;
}

void CountTable::init(Context& ctx, bool wide128, size_t initial) {
    wide = wide128;
    capacity = std::max<size_t>(1024, next_pow2((uint32_t)std::min<size_t>(initial, 1u << 30)));
    keys.alloc(capacity * (wide ? 16 : 8), false);
    counts.alloc(capacity * sizeof(uint32_t), true);
    fill_empty(ctx, keys.ptr, capacity * (wide ? 2 : 1));
    upper_bound = 0;
}

CountTable64 CountTable::view64() const {
    CountTable64 v;
    v.keys = keys.as<unsigned long long>();
    v.counts = counts.as<uint32_t>();
    v.mask = capacity - 1;
    return v;
}

CountTable128 CountTable::view128() const {
    CountTable128 v;
    v.keys = keys.as<ulonglong2>();
    v.counts = counts.as<uint32_t>();
    v.mask = capacity - 1;
    return v;
}

void CountTable::ensure(Context& ctx, long long upcoming) {
    // every insert may be a new key: keep (keys so far + upcoming) <= capacity / 2
    const unsigned long long need = 2ull * (unsigned long long)(upper_bound + upcoming);
    if (need > capacity) {
        size_t ncap = capacity;
        while (ncap < need) ncap *= 2;
        DeviceBuffer nkeys, ncounts;
        nkeys.alloc(ncap * (wide ? 16 : 8), false);
        ncounts.alloc(ncap * sizeof(uint32_t), true);
        fill_empty(ctx, nkeys.ptr, ncap * (wide ? 2 : 1));
        const int grid = (int)std::min<size_t>((capacity + 255) / 256, (size_t)ctx.sm_count * 32);
        if (wide) {
            CountTable128 dst{ nkeys.as<ulonglong2>(), ncounts.as<uint32_t>(), ncap - 1 };
            rehash128_kernel<<<grid, 256, 0, ctx.stream>>>(keys.as<ulonglong2>(), counts.as<uint32_t>(), capacity, dst);
        } else {
            CountTable64 dst{ nkeys.as<unsigned long long>(), ncounts.as<uint32_t>(), ncap - 1 };
            rehash64_kernel<<<grid, 256, 0, ctx.stream>>>(keys.as<unsigned long long>(), counts.as<uint32_t>(), capacity, dst);
        }
        SCG_CUDA_CHECK(cudaGetLastError());
        ++ctx.launches;
        SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
        keys = std::move(nkeys);
        counts = std::move(ncounts);
        capacity = ncap;
    }
    upper_bound += upcoming;
}

void CountTable::download(Context& ctx, std::vector<unsigned long long>& keys_lo, std::vector<unsigned long long>& keys_hi,
                          std::vector<uint32_t>& out_counts) {
    DeviceBuffer d_keys, d_counts, d_cursor;
    const size_t maxlive = (size_t)std::min<unsigned long long>(capacity, (unsigned long long)std::max<long long>(upper_bound, 1));
    d_keys.alloc(maxlive * (wide ? 16 : 8), false);
    d_counts.alloc(maxlive * sizeof(uint32_t), false);
    d_cursor.alloc(sizeof(unsigned long long), true);
    const int grid = (int)std::min<size_t>((capacity + 255) / 256, (size_t)ctx.sm_count * 32);
    if (wide) {
        compact128_kernel<<<grid, 256, 0, ctx.stream>>>(keys.as<ulonglong2>(), counts.as<uint32_t>(), capacity, d_keys.as<ulonglong2>(),
                                                        d_counts.as<uint32_t>(), d_cursor.as<unsigned long long>());
    } else {
        compact64_kernel<<<grid, 256, 0, ctx.stream>>>(keys.as<unsigned long long>(), counts.as<uint32_t>(), capacity,
                                                       d_keys.as<unsigned long long>(), d_counts.as<uint32_t>(),
                                                       d_cursor.as<unsigned long long>());
    }
    SCG_CUDA_CHECK(cudaGetLastError());
    ++ctx.launches;
    unsigned long long live = 0;
    SCG_CUDA_CHECK(cudaMemcpyAsync(&live, d_cursor.ptr, sizeof live, cudaMemcpyDeviceToHost, ctx.stream));
    SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    out_counts.resize(live);
    keys_lo.resize(live);
    keys_hi.clear();
    if (live == 0) return;
    SCG_CUDA_CHECK(cudaMemcpyAsync(out_counts.data(), d_counts.ptr, live * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    if (wide) {
        std::vector<ulonglong2> tmp(live);
        SCG_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), d_keys.ptr, live * 16, cudaMemcpyDeviceToHost, ctx.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
        keys_hi.resize(live);
        for (size_t i = 0; i < live; ++i) {
            keys_lo[i] = tmp[i].x;
            keys_hi[i] = tmp[i].y;
        }
    } else {
        SCG_CUDA_CHECK(cudaMemcpyAsync(keys_lo.data(), d_keys.ptr, live * 8, cudaMemcpyDeviceToHost, ctx.stream));
        SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    }
}

// Rank of every base in the byte order of its letter (A < C < G < N < T, what R's order() sees,
// R/countRandomBarcodes.R:73), three bits per base, first base most significant: sorting these integers sorts the
// barcodes as text.  Keys of up to 21 bases (random_kernel's narrow layout: H | L << 21 | N << 42).
__global__ void text_order_kernel(const unsigned long long* __restrict__ keys, size_t n, int len, unsigned long long* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        const uint32_t H = (uint32_t)(k & 0x1FFFFFull), L = (uint32_t)((k >> 21) & 0x1FFFFFull), N = (uint32_t)((k >> 42) & 0x1FFFFFull);
        unsigned long long s = 0;
        for (int b = 0; b < len; ++b) {
            const uint32_t code = (((H >> b) & 1u) << 1) | ((L >> b) & 1u);
            const uint32_t rank = ((N >> b) & 1u) ? 3u : (code == 3u ? 4u : code);
            s = (s << 3) | rank;
        }
        out[i] = s;
    }
}

// Live entries of a narrow table as (text-order key, count), sorted by the key on the device.
void CountTable::download_sorted(Context& ctx, int key_len, std::vector<unsigned long long>& order_keys, std::vector<uint32_t>& out_counts) {
    DeviceBuffer d_keys, d_counts, d_cursor, d_order, d_order_sorted, d_counts_sorted, d_temp;
    const size_t maxlive = (size_t)std::min<unsigned long long>(capacity, (unsigned long long)std::max<long long>(upper_bound, 1));
    d_keys.alloc(maxlive * 8, false);
    d_counts.alloc(maxlive * sizeof(uint32_t), false);
    d_cursor.alloc(sizeof(unsigned long long), true);
    const int grid = (int)std::min<size_t>((capacity + 255) / 256, (size_t)ctx.sm_count * 32);
    compact64_kernel<<<grid, 256, 0, ctx.stream>>>(keys.as<unsigned long long>(), counts.as<uint32_t>(), capacity,
                                                   d_keys.as<unsigned long long>(), d_counts.as<uint32_t>(), d_cursor.as<unsigned long long>());
    SCG_CUDA_CHECK(cudaGetLastError());
    ++ctx.launches;
    unsigned long long live = 0;
    SCG_CUDA_CHECK(cudaMemcpyAsync(&live, d_cursor.ptr, sizeof live, cudaMemcpyDeviceToHost, ctx.stream));
    SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    order_keys.resize(live);
    out_counts.resize(live);
    if (live == 0) return;
    d_order.alloc(live * 8, false);
    d_order_sorted.alloc(live * 8, false);
    d_counts_sorted.alloc(live * sizeof(uint32_t), false);
    text_order_kernel<<<(int)std::min<size_t>((live + 255) / 256, (size_t)ctx.sm_count * 32), 256, 0, ctx.stream>>>(
        d_keys.as<unsigned long long>(), live, key_len, d_order.as<unsigned long long>());
    SCG_CUDA_CHECK(cudaGetLastError());
    size_t temp_bytes = 0;
    SCG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_order.as<unsigned long long>(), d_order_sorted.as<unsigned long long>(),
                                                   d_counts.as<uint32_t>(), d_counts_sorted.as<uint32_t>(), (long long)live, 0, 3 * key_len,
                                                   ctx.stream));
    d_temp.alloc(temp_bytes, false);
    SCG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_temp.ptr, temp_bytes, d_order.as<unsigned long long>(), d_order_sorted.as<unsigned long long>(),
                                                   d_counts.as<uint32_t>(), d_counts_sorted.as<uint32_t>(), (long long)live, 0, 3 * key_len,
                                                   ctx.stream));
    ctx.launches += 2;
    SCG_CUDA_CHECK(cudaMemcpyAsync(order_keys.data(), d_order_sorted.ptr, live * 8, cudaMemcpyDeviceToHost, ctx.stream));
    SCG_CUDA_CHECK(cudaMemcpyAsync(out_counts.data(), d_counts_sorted.ptr, live * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

// ---------------------------------------------------------------------------------------
// ComboTally
// ---------------------------------------------------------------------------------------
void ComboTally::init(Context& ctx, int a, int b) {
    n1 = a;
    n2 = b;
    dense = (long long)n1 * n2 <= (1ll << 24);
    if (dense) {
        matrix.alloc((size_t)std::max<long long>((long long)n1 * n2, 1) * sizeof(int32_t), true);
    } else {
        table.init(ctx, false, 1u << 20);
    }
}

ComboSink ComboTally::sink(Context& ctx, long long upcoming) {
    ComboSink s;
    std::memset(&s, 0, sizeof s);
    s.n2 = n2;
    if (dense) {
        s.dense = matrix.as<int32_t>();
    } else {
        table.ensure(ctx, upcoming);
        s.sparse = table.view64();
    }
    return s;
}

// sort_combinations + count_combinations (reference inst/include/kaori/utils.hpp:173-198, src/utils.h:14-45):
// rows sorted ascending by (first, second), one frequency per distinct combination
void ComboTally::harvest(Context& ctx, scg_result& out) {
    out.width = 2;
    if (dense) {
        std::vector<int32_t> host((size_t)n1 * n2);
        if (!host.empty()) {
            SCG_CUDA_CHECK(cudaMemcpyAsync(host.data(), matrix.ptr, host.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
            SCG_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
        }
        for (int i = 0; i < n1; ++i) {
            for (int j = 0; j < n2; ++j) {
                const int32_t f = host[(size_t)i * n2 + j];
                if (f) {
                    out.keys.push_back(i);
                    out.keys.push_back(j);
                    out.freq.push_back(f);
                }
            }
        }
    } else {
        std::vector<unsigned long long> lo, hi;
        std::vector<uint32_t> cnt;
        table.download(ctx, lo, hi, cnt);
        std::vector<size_t> order(lo.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return lo[a] < lo[b]; });
        for (size_t o : order) {
            out.keys.push_back((int32_t)(lo[o] >> 32));
            out.keys.push_back((int32_t)(lo[o] & 0xFFFFFFFFull));
            out.freq.push_back((int32_t)cnt[o]);
        }
    }
}

namespace {

// Decodes a packed random-barcode key (random_kernel's layout) to text.
std::string decode_key(unsigned long long lo, unsigned long long hi, int len, bool wide) {
    unsigned long long H, L, N;
    if (!wide) {
        H = lo & ((1ull << 21) - 1);
        L = (lo >> 21) & ((1ull << 21) - 1);
        N = (lo >> 42) & ((1ull << 21) - 1);
    } else {
        const unsigned long long m42 = (1ull << 42) - 1;
        H = lo & m42;
        L = ((lo >> 42) | (hi << 22)) & m42;
        N = (hi >> 20) & m42;
    }
    std::string s(len, 'A');
    for (int i = 0; i < len; ++i) {
        if ((N >> i) & 1ull) {
            s[i] = 'N';
        } else {
            s[i] = "ACGT"[(((H >> i) & 1ull) << 1) | ((L >> i) & 1ull)];
        }
    }
    return s;
}

} // namespace

} // namespace scg

using namespace scg;

extern "C" {

int scg_count_random(scg_ctx* ctx, const scg_source* src, const char* constant, int strand, int mismatches, int use_first,
                     int nthreads, scg_result** table, int32_t* total) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        const double t_start = now_s();
        c.timing = Timing();
        Source source(src);
        TemplateSpec tmpl(constant, strand);
        // the reference dereferences variable_regions()[0] unconditionally
        // (handlers/RandomBarcodeSingleEnd.hpp:93-96, :212-214)
        if (tmpl.fwd_regions.empty()) throw Error("expected at least one variable region in the constant template");
        const int key_len = tmpl.fwd_regions[0].end - tmpl.fwd_regions[0].start;
        if (key_len > 42) throw Error("random barcode regions longer than 42 bp are not supported by this engine");
        RandomParams P;
        std::memset(&P, 0, sizeof P);
        P.spec = tmpl.scan_spec(mismatches);
        P.max_mm = mismatches;
        P.use_first = use_first ? 1 : 0;
        P.key_len = key_len;
        const bool wide = key_len > 21;
        const int KWsel = key_len > 32 ? 2 : 1;

        c.ensure_ready();
        CountTable tab;
        tab.init(c, wide, 1u << 20);
        DeviceBuffer d_odd_out, d_odd_count;
        d_odd_count.alloc(sizeof(unsigned long long), true);
        std::map<std::string, int> extra;   // keys of reads that need their raw text

        ReadPipeline pipe(c, source.reader.get(), nullptr, nthreads, true);
        ReadPipeline::Batch b;
        long long nreads = 0;
        std::vector<OddOutcome> odd_host;
        while (pipe.next(b)) {
            {
                const double t0 = now_s();
                tab.ensure(c, b.n);
                c.timing.setup_s += now_s() - t0;
            }
            // reads flagged odd get their outcome reported back; the device-side reader only knows on the device how many
            // there are (count_odd < 0), so room is made for all
            const long long nodd_host = pipe.count_odd(b);
            const long long nodd = nodd_host < 0 ? b.n : nodd_host;
            d_odd_out.reserve((size_t)std::max<long long>(nodd, 1) * sizeof(OddOutcome));
            SCG_CUDA_CHECK(cudaMemsetAsync(d_odd_count.ptr, 0, sizeof(unsigned long long), c.stream));
            const long long ntiles = (b.n + TILE - 1) / TILE;
            const int grid = c.grid_for(ntiles);
            CountTable64 t64 = wide ? CountTable64{ nullptr, nullptr, 0 } : tab.view64();
            CountTable128 t128 = wide ? tab.view128() : CountTable128{ nullptr, nullptr, 0 };
            dispatch_cb(P.spec.cbits, [&](auto CB) {
                if (KWsel == 1) {
                    random_kernel<decltype(CB)::value, 1><<<grid, 128, 0, c.stream>>>(b.reads1, P, t64, t128, b.odd1, 0, d_odd_out.as<OddOutcome>(),
                                                                                    d_odd_count.as<unsigned long long>(), nullptr);
                } else {
                    random_kernel<decltype(CB)::value, 2><<<grid, 128, 0, c.stream>>>(b.reads1, P, t64, t128, b.odd1, 0, d_odd_out.as<OddOutcome>(),
                                                                                    d_odd_count.as<unsigned long long>(), nullptr);
                }
            });
            SCG_CUDA_CHECK(cudaGetLastError());
            ++c.launches;
            ++c.timing.launches;
            pipe.submitted(b);
            if (nodd > 0) {
                // reads holding lower-case letters or symbols other than N: the device decided where the
                // barcode sits, the host renders it from the raw read text exactly as the reference would
                unsigned long long got = 0;
                SCG_CUDA_CHECK(cudaMemcpyAsync(&got, d_odd_count.ptr, sizeof got, cudaMemcpyDeviceToHost, c.stream));
                SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
                odd_host.resize(got);
                if (got) {
                    SCG_CUDA_CHECK(cudaMemcpyAsync(odd_host.data(), d_odd_out.ptr, got * sizeof(OddOutcome), cudaMemcpyDeviceToHost, c.stream));
                    SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
                }
                std::sort(odd_host.begin(), odd_host.end(), [](const OddOutcome& x, const OddOutcome& y) { return x.read < y.read; });
                std::string key(key_len, ' ');
                std::string seq;
                for (const auto& o : odd_host) {
                    pipe.raw_read(b, (long long)o.read, seq);
                    const char* start = seq.data() + o.position + tmpl.fwd_regions[0].start;
                    if (!o.reverse) {  // forward_match: raw characters (handlers/RandomBarcodeSingleEnd.hpp:93-104)
                        key.assign(start, key_len);
                    } else {  // reverse_match: complement_base<true> (:106-120, utils.hpp:41-62)
                        for (int j = 0; j < key_len; ++j) {
                            const char ch = start[key_len - j - 1];
                            char out;
                            switch (ch) {
                                case 'A': case 'a': out = 'T'; break;
                                case 'C': case 'c': out = 'G'; break;
                                case 'G': case 'g': out = 'C'; break;
                                case 'T': case 't': out = 'A'; break;
                                case 'N': case 'n': out = 'N'; break;
                                default: throw Error(std::string("cannot complement unknown base '") + ch + "'");
                            }
                            key[j] = out;
                        }
                    }
                    ++extra[key];
                }
            }
            nreads += b.n;
        }
        const double t_harvest = now_s();
        auto* r = new scg_result;
        r->width = key_len;
        if (!wide) {
            // barcodes of up to 21 bases: sorted as text on the device (radix sort of rank-coded keys), rendered by the
            // host threads, merged with the few keys that came from raw read text
            std::vector<unsigned long long> order_keys;
            std::vector<uint32_t> cnt;
            tab.download_sorted(c, key_len, order_keys, cnt);
            const size_t n = order_keys.size();
            auto render = [&](unsigned long long k, char* out) {
                for (int b = key_len - 1; b >= 0; --b) {
                    out[b] = "ACGNT"[k & 7ull];
                    k >>= 3;
                }
            };
            if (extra.empty()) {
                r->strings.resize(n * (size_t)key_len);
                r->freq.resize(n);
                const int pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, nthreads), n >> 16));
                const size_t per = (n + pieces - 1) / pieces;
                HostPool::instance().parallel_for(pieces, pieces, [&](int k) {
                    const size_t b = (size_t)k * per, e = std::min(n, b + per);
                    for (size_t i = b; i < e; ++i) {
                        render(order_keys[i], &r->strings[i * (size_t)key_len]);
                        r->freq[i] = (int32_t)cnt[i];
                    }
                });
            } else {
                r->strings.reserve((n + extra.size()) * (size_t)key_len);
                r->freq.reserve(n + extra.size());
                std::string cur(key_len, ' ');
                auto it = extra.begin();
                auto emit = [&](const std::string& k, int f) {
                    r->strings.insert(r->strings.end(), k.begin(), k.end());
                    r->freq.push_back(f);
                };
                for (size_t i = 0; i < n; ++i) {
                    render(order_keys[i], &cur[0]);
                    while (it != extra.end() && it->first < cur) {
                        emit(it->first, it->second);
                        ++it;
                    }
                    int f = (int)cnt[i];
                    if (it != extra.end() && it->first == cur) {
                        f += it->second;
                        ++it;
                    }
                    emit(cur, f);
                }
                for (; it != extra.end(); ++it) emit(it->first, it->second);
            }
        } else {
            std::vector<unsigned long long> lo, hi;
            std::vector<uint32_t> cnt;
            tab.download(c, lo, hi, cnt);
            std::map<std::string, int> merged(std::move(extra));
            for (size_t i = 0; i < lo.size(); ++i) merged[decode_key(lo[i], hi[i], key_len, wide)] += (int)cnt[i];
            r->strings.reserve(merged.size() * (size_t)key_len);
            r->freq.reserve(merged.size());
            for (const auto& kv : merged) {  // std::map iterates in byte order = R's order(sequences) for ACGTN text
                r->strings.insert(r->strings.end(), kv.first.begin(), kv.first.end());
                r->freq.push_back(kv.second);
            }
        }
        c.timing.harvest_s = now_s() - t_harvest;
        *table = r;
        *total = (int32_t)nreads;
        c.timing.parse_s = source.reader->parse_seconds();
        c.timing.total_s = now_s() - t_start;
        c.finish_timing();
    });
}

} // extern "C"
