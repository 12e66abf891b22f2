// Synthetic workload generator (BASELINE.json configs): a counter-based,
// stateless function of (seed, global read index, position), so the device (packed reads in HBM)
// and the host (FASTQ text for the reference / the end-to-end path) produce the same reads and any
// shard of a multi-GPU run is reproducible on its own.
#include <algorithm>
#include <cstring>
#include <thread>

#include "api_common.hpp"

namespace scg {

struct SynthDev {
    unsigned long long seed;
    long long first_read, n_reads;
    int read_len, T;
    int n_pools, paired_rows, strand;
    int construct_permille, sub_per_10k, n_per_10k;
    long long random_space;
    int nreg;
    int reg_start[2], reg_len[2];
    int n_choices[2];
    const uint8_t* tmpl;      // T codes: 0..3 constant base, 4 variable position
    const uint8_t* pool[2];   // n_choices * reg_len codes
};

// uniform integer in [0, n) from 16 hash bits (n <= 65536) without a division
SCG_HD uint32_t pick16(uint32_t bits16, uint32_t n) { return ((bits16 & 0xFFFFu) * n) >> 16; }

struct SynthRead {
    bool has;
    int offset;
    bool reverse;
    long long choice[2];
    unsigned long long h;
};

SCG_HD SynthRead synth_read(const SynthDev& s, long long g) {
    SynthRead r;
    r.h = mix64(s.seed ^ mix64((unsigned long long)g));
    r.has = pick16((uint32_t)r.h, 1000) < (uint32_t)s.construct_permille && s.read_len >= s.T;
    const unsigned long long h1 = mix64(r.h + 1);
    r.offset = s.read_len >= s.T ? (int)(h1 % (unsigned long long)(s.read_len - s.T + 1)) : 0;
    const unsigned long long h2 = mix64(r.h + 2);
    r.reverse = s.strand == SCG_STRAND_BOTH ? ((h2 & 1ull) != 0) : (s.strand == SCG_STRAND_REVERSE);
    const unsigned long long h3 = mix64(r.h + 3), h4 = mix64(r.h + 4);
    if (s.n_pools == 0) {
        r.choice[0] = (long long)(h3 % (unsigned long long)(s.random_space > 1 ? s.random_space : 1));
        r.choice[1] = 0;
    } else {
        r.choice[0] = (long long)(h3 % (unsigned long long)(s.n_choices[0] > 1 ? s.n_choices[0] : 1));
        r.choice[1] = s.paired_rows ? r.choice[0] : (long long)(h4 % (unsigned long long)(s.n_choices[1] > 1 ? s.n_choices[1] : 1));
    }
    return r;
}

// code of the base at `pos`: 0..3, or 4 for N
SCG_HD int synth_base(const SynthDev& s, const SynthRead& r, int pos) {
    const unsigned long long hp = mix64(r.h ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(pos + 16)));
    int base = (int)(hp & 3ull);
    if (r.has && pos >= r.offset && pos < r.offset + s.T) {
        const int j = pos - r.offset;
        const int jf = r.reverse ? s.T - 1 - j : j;   // position on the forward construct
        int b = s.tmpl[jf];
        if (b == 4) {
            int v = (s.nreg > 1 && jf >= s.reg_start[1]) ? 1 : 0;
            const int k = jf - s.reg_start[v];
            if (s.n_pools == 0) {
                // "true" random barcode number choice[0]: 2 bits per base from its own hash stream
                const unsigned long long hb = mix64(0xB0A7C0DEull ^ mix64((unsigned long long)r.choice[0] * 2 + (unsigned long long)(k >> 5)));
                b = (int)((hb >> (2 * (k & 31))) & 3ull);
            } else {
                if (v >= s.n_pools) v = s.n_pools - 1;
                b = s.pool[v][(size_t)r.choice[v] * s.reg_len[v] + k];
            }
        }
        base = r.reverse ? 3 - b : b;
        if (pick16((uint32_t)(hp >> 8), 10000) < (uint32_t)s.sub_per_10k) {
            base = (base + 1 + (int)((hp >> 24) % 3ull)) & 3;   // substitution by one of the other three bases
        }
    }
    if (pick16((uint32_t)(hp >> 32), 10000) < (uint32_t)s.n_per_10k) base = 4;
    return base;
}

__global__ void __launch_bounds__(128) synth_kernel(SynthDev s, uint32_t* __restrict__ out, int W) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ntiles = (s.n_reads + TILE - 1) / TILE;
    for (long long tile = warp; tile < ntiles; tile += nwarps) {
        const long long i = tile * TILE + lane;
        uint32_t* base = out + (size_t)tile * tile_words(W) + lane;
        SynthRead r;
        if (i < s.n_reads) r = synth_read(s, s.first_read + i);
        for (int w = 0; w < W; ++w) {
            uint32_t h = 0, l = 0, n = 0;
            if (i < s.n_reads) {
                const int lim = min(32, s.read_len - 32 * w);
                for (int b = 0; b < lim; ++b) {
                    const int code = synth_base(s, r, 32 * w + b);
                    h |= (uint32_t)((code >> 1) & 1) << b;
                    l |= (uint32_t)(code & 1) << b;
                    n |= (uint32_t)(code >> 2) << b;
                }
                h &= ~n;
                l &= ~n;
            }
            base[(size_t)(PLANE_H * W + w) * TILE] = h;
            base[(size_t)(PLANE_L * W + w) * TILE] = l;
            base[(size_t)(PLANE_N * W + w) * TILE] = n;
        }
    }
}

namespace {

struct SynthHost {
    SynthDev dev;
    std::vector<uint8_t> tmpl;
    std::vector<uint8_t> pool[2];
};

void fill_host(const scg_synth_spec* spec, SynthHost& h) {
    if (!spec || !spec->constant) throw Error("null synthetic spec");
    std::memset(&h.dev, 0, sizeof h.dev);
    TemplateSpec t(spec->constant, SCG_STRAND_ORIGINAL);
    if (t.fwd_regions.size() > 2) throw Error("the synthetic generator handles at most two variable regions");
    if (spec->read_len <= 0 || spec->read_len > MAX_READ_LEN) throw Error("invalid synthetic read length");
    if (spec->n_pools < 0 || spec->n_pools > 2) throw Error("invalid number of synthetic pools");
    h.dev.seed = spec->seed;
    h.dev.first_read = spec->first_read;
    h.dev.n_reads = spec->n_reads;
    h.dev.read_len = spec->read_len;
    h.dev.T = t.length;
    h.dev.n_pools = spec->n_pools;
    h.dev.paired_rows = spec->paired_rows;
    h.dev.strand = spec->strand;
    h.dev.construct_permille = spec->construct_permille;
    h.dev.sub_per_10k = spec->sub_per_10k;
    h.dev.n_per_10k = spec->n_per_10k;
    h.dev.random_space = spec->random_space;
    h.dev.nreg = (int)t.fwd_regions.size();
    h.tmpl.resize(t.length);
    for (int i = 0; i < t.length; ++i) h.tmpl[i] = t.fwd_seq[i] == '-' ? 4 : (uint8_t)base_code(t.fwd_seq[i]);
    for (int v = 0; v < h.dev.nreg; ++v) {
        h.dev.reg_start[v] = t.fwd_regions[v].start;
        h.dev.reg_len[v] = t.fwd_regions[v].end - t.fwd_regions[v].start;
    }
    if (spec->n_pools == 0 && h.dev.nreg >= 1 && h.dev.reg_len[0] > 64) throw Error("synthetic random barcodes are limited to 64 bases");
    if (spec->n_pools > h.dev.nreg) throw Error("more synthetic pools than variable regions");
    for (int v = 0; v < spec->n_pools; ++v) {
        Pool p(spec->pools[v], spec->n_choices[v]);
        if (p.seqs.empty()) throw Error("empty synthetic pool");
        if (p.length != h.dev.reg_len[v]) throw Error("synthetic pool length does not match its variable region");
        h.dev.n_choices[v] = (int)p.seqs.size();
        h.pool[v].resize((size_t)p.seqs.size() * p.length);
        for (size_t c = 0; c < p.seqs.size(); ++c) {
            for (int k = 0; k < p.length; ++k) {
                const int code = base_code(p.seqs[c][k]);
                if (code < 0) throw Error("synthetic pools must be plain ACGT");
                h.pool[v][c * p.length + k] = (uint8_t)code;
            }
        }
    }
    if (spec->paired_rows && spec->n_pools == 2 && h.dev.n_choices[0] != h.dev.n_choices[1]) {
        throw Error("paired synthetic pools must have the same number of rows");
    }
    h.dev.tmpl = h.tmpl.data();
    h.dev.pool[0] = h.pool[0].data();
    h.dev.pool[1] = h.pool[1].data();
}

} // namespace

} // namespace scg

using namespace scg;

extern "C" {

int scg_reads_synthesize(scg_ctx* ctx, const scg_synth_spec* spec, scg_reads** out) {
    return guarded(ctx, [&] {
        Context& c = ctx->impl;
        SynthHost h;
        fill_host(spec, h);
        c.ensure_ready();
        DeviceBuffer d_tmpl, d_pool[2];
        d_tmpl.upload(h.tmpl.data(), h.tmpl.size(), c.stream);
        SynthDev dev = h.dev;
        dev.tmpl = d_tmpl.as<uint8_t>();
        for (int v = 0; v < 2; ++v) {
            d_pool[v].upload(h.pool[v].data(), h.pool[v].size(), c.stream);
            dev.pool[v] = d_pool[v].as<uint8_t>();
        }
        std::unique_ptr<scg_reads> reads(new scg_reads);
        reads->owner = ctx;
        const int W = ceil_div(spec->read_len, 32);
        // batches of at most 2^26 reads keep every allocation and grid well inside 32-bit tile counts
        const long long per_batch = 1ll << 26;
        for (long long at = 0; at < spec->n_reads; at += per_batch) {
            const long long n = std::min(per_batch, spec->n_reads - at);
            DeviceBatch b;
            const size_t words = (size_t)((n + TILE - 1) / TILE) * tile_words(W);
            b.data.alloc(words * sizeof(uint32_t) + READ_GUARD_BYTES, false);
            b.view.data = b.data.as<uint32_t>();
            b.view.lens = nullptr;
            b.view.uniform_len = spec->read_len;
            b.view.W = W;
            b.view.n = n;
            SynthDev part = dev;
            part.first_read = spec->first_read + at;
            part.n_reads = n;
            const long long ntiles = (n + TILE - 1) / TILE;
            synth_kernel<<<c.grid_for(ntiles), 128, 0, c.stream>>>(part, b.data.as<uint32_t>(), W);
            SCG_CUDA_CHECK(cudaGetLastError());
            ++c.launches;
            reads->n += n;
            reads->device_bytes += (long long)(words * sizeof(uint32_t));
            reads->batches.push_back(std::move(b));
        }
        SCG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        *out = reads.release();
    });
}

int scg_synth_fastq(const scg_synth_spec* spec, char* out, size_t capacity, size_t* used) {
    try {
        SynthHost h;
        fill_host(spec, h);
        const size_t R = (size_t)spec->read_len;
        const size_t per = 3 + R + 3 + R + 1;   // "@r\n" seq "\n+\n" qual "\n"
        const size_t need = per * (size_t)spec->n_reads;
        if (used) *used = need;
        if (!out) return 0;
        if (capacity < need) throw Error("buffer too small for the synthetic FASTQ");
        const int nt = (int)std::max<long long>(1, std::min<long long>(std::thread::hardware_concurrency(), spec->n_reads / 4096 + 1));
        auto work = [&](long long b, long long e) {
            for (long long i = b; i < e; ++i) {
                char* p = out + per * (size_t)i;
                p[0] = '@';
                p[1] = 'r';
                p[2] = '\n';
                const SynthRead r = synth_read(h.dev, spec->first_read + i);
                for (size_t k = 0; k < R; ++k) p[3 + k] = "ACGTN"[synth_base(h.dev, r, (int)k)];
                p[3 + R] = '\n';
                p[4 + R] = '+';
                p[5 + R] = '\n';
                std::memset(p + 6 + R, 'I', R);
                p[6 + 2 * R] = '\n';
            }
        };
        std::vector<std::thread> pool;
        const long long chunk = (spec->n_reads + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const long long b = t * chunk, e = std::min<long long>(spec->n_reads, b + chunk);
            if (b >= e) break;
            pool.emplace_back(work, b, e);
        }
        for (auto& th : pool) th.join();
        return 0;
    } catch (const std::exception& e) {
        creation_error() = e.what();
        return 1;
    }
}

} // extern "C"
