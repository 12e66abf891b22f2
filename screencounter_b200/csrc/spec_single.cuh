// countSingleBarcodes kernel with the TEMPLATE FOLDED IN AT COMPILE TIME.
//
// The generic kernel (handlers.cuh, single_kernel) receives the template as bit masks and spends
// most of its issue slots on uniform bookkeeping (which template position is constant, which base
// it holds).  Here every template position is a compile-time constant, so after unrolling the scan
// is exactly, per constant position and strand,
//     one funnel shift (the 32-window view of that base's mismatch plane) + the counter update
// -- the position-parallel formulation at its minimum instruction count.
//
// This file is compiled twice:
//   * by nvcc at build time for the default configuration below (BASELINE configs[1]'s template),
//     which proves it builds for sm_100a and gives that configuration a precompiled kernel;
//   * by NVRTC at run time (jit.cpp) with the SPEC_* macros of the template in use.
//
// Macros (all required when SPEC_CUSTOM is defined):
//   SPEC_T template length, SPEC_FBASES / SPEC_RBASES the forward / reverse-complemented template
//   as a string of A C G T and '-', SPEC_FWD / SPEC_REV strands searched, SPEC_W words per plane of
//   the batch, SPEC_NB window blocks (ceil((32*W - T + 1) / 32)), SPEC_CB counter planes, SPEC_MM
//   clamped scan budget, SPEC_MAXMM the caller's budget, SPEC_USE_FIRST, SPEC_FSTART / SPEC_RSTART /
//   SPEC_KEYLEN the variable region, SPEC_NAME the kernel's name.
#pragma once

#include "device_keys.cuh"

#ifndef SPEC_CUSTOM
#define SPEC_T 44
#define SPEC_FBASES "CAGCTACGTACG--------------------CCAGCTCGATCG"
#define SPEC_RBASES "CGATCGAGCTGG--------------------CGTACGTAGCTG"
#define SPEC_FWD 1
#define SPEC_REV 1
#define SPEC_W 3
#define SPEC_NB 2
#define SPEC_CB 1
#define SPEC_MM 1
#define SPEC_MAXMM 1
#define SPEC_USE_FIRST 1
#define SPEC_FSTART 12
#define SPEC_RSTART 12
#define SPEC_KEYLEN 20
#define SPEC_NAME spec_single_kernel_default
#endif

namespace scg {
namespace spec {

constexpr int T = SPEC_T;
constexpr int W = SPEC_W;
constexpr int NB = SPEC_NB;
constexpr int CB = SPEC_CB;
constexpr int KW = (SPEC_KEYLEN + 31) / 32;
constexpr char FB[] = SPEC_FBASES;
constexpr char RB[] = SPEC_RBASES;

// mismatch planes of one read: bit i of X?[w] is set when base 32*w + i is NOT that base
// (an N, or anything else that is not ACGT, mismatches all four)
struct Planes {
    uint32_t xa[W + 2], xc[W + 2], xg[W + 2], xt[W + 2];
};

template <char B>
__device__ __forceinline__ uint32_t window(const Planes& P, int word, int shift) {
    const uint32_t* x = B == 'A' || B == 'a' ? P.xa : (B == 'C' || B == 'c' ? P.xc : (B == 'G' || B == 'g' ? P.xg : P.xt));
    return __funnelshift_r(x[word], x[word + 1], shift);
}

template <int J>
struct ScanStep {
    static __device__ __forceinline__ void run(const Planes& P, int pb, Counter<CB>& cf, Counter<CB>& cr) {
        if constexpr (SPEC_FWD && FB[J] != '-') cf.add(window<FB[J]>(P, pb + J / 32, J % 32));
        if constexpr (SPEC_REV && RB[J] != '-') cr.add(window<RB[J]>(P, pb + J / 32, J % 32));
        ScanStep<J + 1>::run(P, pb, cf, cr);
    }
};

template <>
struct ScanStep<T> {
    static __device__ __forceinline__ void run(const Planes&, int, Counter<CB>&, Counter<CB>&) {}
};

} // namespace spec
} // namespace scg

extern "C" __global__ void __launch_bounds__(128) SPEC_NAME(scg::ReadsDev reads, const scg::LibDev* __restrict__ libs,
                                                            int32_t* __restrict__ counts, int32_t* __restrict__ out_index,
                                                            uint32_t* __restrict__ out_info) {
    using namespace scg;
    using namespace scg::spec;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ntiles = (reads.n + TILE - 1) / TILE;
    for (long long tile = warp; tile < ntiles; tile += nwarps) {
        const long long i = tile * TILE + lane;
        const ReadView rd = read_view(reads, tile, lane);

        Planes P;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t h = rd.word(PLANE_H, w), l = rd.word(PLANE_L, w), n = rd.word(PLANE_N, w);
            P.xa[w] = h | l | n;
            P.xc[w] = h | ~l | n;
            P.xg[w] = ~h | l | n;
            P.xt[w] = ~h | ~l | n;
        }
        P.xa[W] = P.xc[W] = P.xg[W] = P.xt[W] = 0;
        P.xa[W + 1] = P.xc[W + 1] = P.xg[W + 1] = P.xt[W + 1] = 0;

        SingleOut out{ false, -1, 0, false, 0, 0 };
        int best = SPEC_MAXMM + 1;
        bool done = false;
        const int nblocks = window_blocks(rd.len, T);
#pragma unroll
        for (int pb = 0; pb < NB; ++pb) {
            if (pb >= nblocks || done) continue;
            Counter<CB> cf, cr;
            cf.clear();
            cr.clear();
            ScanStep<0>::run(P, pb, cf, cr);
            const uint32_t valid = valid_windows(rd.len, T, pb);
            uint32_t okf = SPEC_FWD ? (cf.le(SPEC_MM) & valid) : 0u;
            uint32_t okr = SPEC_REV ? (cr.le(SPEC_MM) & valid) : 0u;
            // hits in the reference's order: positions ascending, forward before reverse
            // (SimpleSingleMatch.hpp:226-242); the strand is per-lane data
            while ((okf | okr) && !done) {
                const int p = __ffs(okf | okr) - 1;
                const bool rev = !((okf >> p) & 1u);
                if (rev) {
                    okr &= ~(1u << p);
                } else {
                    okf &= ~(1u << p);
                }
                const int c = rev ? cr.get(p) : cf.get(p);
                Key<KW> key;
                extract_region<KW>(rd, 32 * pb + p + (rev ? SPEC_RSTART : SPEC_FSTART), SPEC_KEYLEN, key);
                const Hit h = lookup_any<KW>(libs + (rev ? 1 : 0), key, SPEC_MAXMM - c);
                if (h.index < 0) continue;
                const int total = c + h.dist;
                if (SPEC_USE_FIRST) {
                    out.found = true;
                    out.index = h.index;
                    out.position = 32 * pb + p;
                    out.reverse = rev;
                    out.mismatches = total;
                    out.var_mismatches = h.dist;
                    done = true;
                } else if (total == best) {  // SimpleSingleMatch.hpp:270-275
                    if (out.index != h.index) {
                        out.found = false;
                        out.index = -1;
                    }
                } else if (total < best) {
                    best = total;
                    out.found = true;
                    out.index = h.index;
                    out.position = 32 * pb + p;
                    out.reverse = rev;
                    out.mismatches = total;
                    out.var_mismatches = h.dist;
                }
            }
        }
        if (i < reads.n) {
            if (out.found) atomicAdd(counts + out.index, 1);
            if (out_index) out_index[i] = out.found ? out.index : -1;
            if (out_info) out_info[i] = pack_info(out.found, out.reverse, out.mismatches, out.var_mismatches, out.position);
        }
    }
}
