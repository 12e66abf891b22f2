// The constant-flank scan with the TEMPLATE FOLDED IN AT COMPILE TIME, as a reusable component: the filter + verify
// scheme of the uniform-length single-barcode kernel (spec_single.cuh), parameterised by a traits type
// so that one kernel can scan with several templates (the two mates of a paired-end design) and so that every handler
// kernel of spec_handlers.cuh shares it.  Replaces ScanTemplate::next / strand_match
// (inst/include/kaori/ScanTemplate.hpp:183-252) for batches whose reads all have the same length.
//
// A traits type `Tr` provides
//     static constexpr int T;             template length
//     static constexpr char FB[], RB[];   forward / reverse-complemented template, '-' at variable positions
//     static constexpr bool FWD, REV;     strands searched
//     static constexpr int MM;            mismatch budget of the constant part (already clamped to the number of constant bases)
//     static constexpr int ULEN, W;       read length of the batch and its words per plane
//     static constexpr int SAMPLES;       sampled positions per pigeonhole group of the filter
//
// Scheme (all 32 windows of a block at once, bit p of a register = window p):
//   FILTER   a window with at most MM constant mismatches is mismatch-free in at least one of MM + 1 groups of constant
//            positions; per group up to SAMPLES positions are sampled, their mismatch planes (one funnel shift each) OR-ed,
//            and the windows where some group stayed clean are the candidates;
//   VERIFY   each lane cuts its candidate window out of its registers (funnel shifts by the lane's own position) and counts
//            the constant mismatches exactly with XOR / mask / POPC against the template's words, which are immediates.
// Candidates are visited in the reference's order: positions ascending, forward before reverse at a position
// (SimpleSingleMatch.hpp:226-242).
#pragma once

#include "device_keys.cuh"

namespace scg {
namespace sscan {

// one LOP3 with a chosen truth table (a = 0xF0, b = 0xCC, c = 0xAA)
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

__host__ __device__ constexpr int count_constant(const char* s, int T) {
    int n = 0;
    for (int j = 0; j < T; ++j) n += s[j] != '-';
    return n;
}
// template position of the k-th constant base
__host__ __device__ constexpr int kth_constant(const char* s, int T, int k) {
    int n = 0;
    for (int j = 0; j < T; ++j) {
        if (s[j] != '-') {
            if (n == k) return j;
            ++n;
        }
    }
    return 0;
}
// word k of a template: what = 0 constant-position mask, 1 high bits of the bases, 2 low bits
__host__ __device__ constexpr uint32_t template_word(const char* s, int T, int k, int what) {
    uint32_t w = 0;
    for (int j = 0; j < 32; ++j) {
        const int pos = 32 * k + j;
        if (pos >= T || s[pos] == '-') continue;
        const char c = s[pos];
        const uint32_t code = (c == 'A' || c == 'a') ? 0u : ((c == 'C' || c == 'c') ? 1u : ((c == 'G' || c == 'g') ? 2u : 3u));
        const uint32_t bit = what == 0 ? 1u : (what == 1 ? (code >> 1) : (code & 1u));
        w |= bit << j;
    }
    return w;
}

template <class Tr>
struct Dims {
    static constexpr int T = Tr::T, W = Tr::W;
    static constexpr int NCF = Tr::FWD ? count_constant(Tr::FB, Tr::T) : 0;
    static constexpr int NCR = Tr::REV ? count_constant(Tr::RB, Tr::T) : 0;
    static constexpr int NWIN = Tr::ULEN - Tr::T + 1;       // windows of a read
    static constexpr int NBLOCKS = (NWIN + 31) / 32;        // window blocks: block b holds windows [32 b, 32 b + 32)
    static constexpr int TW = (Tr::T + 31) / 32;            // words per window
    static constexpr int NGROUPS = Tr::MM + 1;              // pigeonhole groups of constant positions
    static constexpr int TILE_WORDS = 3 * Tr::W * TILE;     // one tile = 3 planes x W words x 32 lanes, contiguous
    static constexpr uint32_t TILE_BYTES = TILE_WORDS * 4u;
    static_assert(NWIN >= 1, "reads must be at least as long as the template");
    static_assert(Tr::W + 2 >= TW + 1, "window words plus their funnel partner must exist");
    static_assert(NBLOCKS - 1 + TW <= Tr::W, "the last block's window words and their funnel partners lie within the guarded words");
};

// The read's words in registers; two zero guard words so that every funnel shift has a partner.
template <int W>
struct Words {
    uint32_t h[W + 2], l[W + 2], n[W + 2];
};
// mismatch planes of one read: bit i of x?[w] is set when base 32*w + i is NOT that base (an N mismatches all four)
template <int W>
struct Planes {
    uint32_t xa[W + 2], xc[W + 2], xg[W + 2], xt[W + 2];
};

// a lane's words of its tile, from the tile's image in shared memory (conflict-free: layout.hpp)
template <int W>
__device__ __forceinline__ void load_words(const uint32_t* __restrict__ buf /* + lane */, Words<W>& R) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        R.h[w] = buf[(PLANE_H * W + w) * TILE];
        R.l[w] = buf[(PLANE_L * W + w) * TILE];
        R.n[w] = buf[(PLANE_N * W + w) * TILE];
    }
    R.h[W] = R.l[W] = R.n[W] = 0;
    R.h[W + 1] = R.l[W + 1] = R.n[W + 1] = 0;
}

template <int W>
__device__ __forceinline__ void make_planes(const Words<W>& R, Planes<W>& P) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t h = R.h[w], l = R.l[w], n = R.n[w];
        P.xa[w] = h | l | n;
        P.xc[w] = h | ~l | n;
        P.xg[w] = ~h | l | n;
        P.xt[w] = ~h | ~l | n;
    }
    P.xa[W] = P.xc[W] = P.xg[W] = P.xt[W] = 0;
    P.xa[W + 1] = P.xc[W + 1] = P.xg[W + 1] = P.xt[W + 1] = 0;
}

// 32-window mismatch plane of the K-th constant position of a strand, window block PB
template <class Tr, bool REV, int K, int PB>
__device__ __forceinline__ uint32_t cplane(const Planes<Tr::W>& P) {
    constexpr int j = REV ? kth_constant(Tr::RB, Tr::T, K) : kth_constant(Tr::FB, Tr::T, K);
    constexpr char b = REV ? Tr::RB[j] : Tr::FB[j];
    constexpr int word = PB + j / 32, shift = j % 32;
    const uint32_t* x = (b == 'A' || b == 'a') ? P.xa : ((b == 'C' || b == 'c') ? P.xc : ((b == 'G' || b == 'g') ? P.xg : P.xt));
    return __funnelshift_r(x[word], x[word + 1], shift);
}

// constant positions [lo, hi) of a strand's group G, and which of them are sampled
template <class Tr, bool REV, int G>
struct Group {
    static constexpr int NC = REV ? Dims<Tr>::NCR : Dims<Tr>::NCF;
    static constexpr int lo = NC * G / Dims<Tr>::NGROUPS, hi = NC * (G + 1) / Dims<Tr>::NGROUPS, size = hi - lo;
    static constexpr int S = size < Tr::SAMPLES ? size : Tr::SAMPLES;
    __host__ __device__ static constexpr int sample(int k) { return lo + (S > 0 ? k * size / S : 0); }
};

// OR of the mismatch planes of a group's sampled positions: bit p set = window p mismatches at a sampled position
template <class Tr, bool REV, int PB, int G, int K>
__device__ __forceinline__ uint32_t group_any(const Planes<Tr::W>& P) {
    using Gr = Group<Tr, REV, G>;
    if constexpr (K >= Gr::S) {
        return 0u;
    } else if constexpr (K + 3 <= Gr::S) {
        return cplane<Tr, REV, Gr::sample(K), PB>(P) | cplane<Tr, REV, Gr::sample(K + 1), PB>(P) | cplane<Tr, REV, Gr::sample(K + 2), PB>(P) |
               group_any<Tr, REV, PB, G, K + 3>(P);
    } else if constexpr (K + 2 <= Gr::S) {
        return cplane<Tr, REV, Gr::sample(K), PB>(P) | cplane<Tr, REV, Gr::sample(K + 1), PB>(P) | group_any<Tr, REV, PB, G, K + 2>(P);
    } else {
        return cplane<Tr, REV, Gr::sample(K), PB>(P) | group_any<Tr, REV, PB, G, K + 1>(P);
    }
}

// windows in which EVERY group shows a mismatch among its samples (those cannot be within the budget)
template <class Tr, bool REV, int PB, int G>
__device__ __forceinline__ uint32_t all_groups_dirty(const Planes<Tr::W>& P) {
    if constexpr (G >= Dims<Tr>::NGROUPS) {
        return 0xFFFFFFFFu;
    } else if constexpr (Group<Tr, REV, G>::S == 0) {
        return 0u;   // an empty group is trivially clean: nothing can be excluded
    } else {
        return group_any<Tr, REV, PB, G, 0>(P) & all_groups_dirty<Tr, REV, PB, G + 1>(P);
    }
}

// windows of block PB that exist in a read of ULEN bases
template <class Tr, int PB>
__host__ __device__ constexpr uint32_t block_mask() {
    return Dims<Tr>::NWIN - 32 * PB >= 32 ? 0xFFFFFFFFu : ((1u << (Dims<Tr>::NWIN - 32 * PB)) - 1u);
}

// Exact constant-mismatch count of window 32 PB + p on a strand; leaves the window's words in wh / wl / wn.
template <class Tr, int K, int PB>
__device__ __forceinline__ int verify_words(const Words<Tr::W>& R, int p, bool rev, uint32_t (&wh)[Dims<Tr>::TW + 1],
                                            uint32_t (&wl)[Dims<Tr>::TW + 1], uint32_t (&wn)[Dims<Tr>::TW + 1]) {
    if constexpr (K >= Dims<Tr>::TW) {
        return 0;
    } else {
        constexpr uint32_t fth = template_word(Tr::FB, Tr::T, K, 1), ftl = template_word(Tr::FB, Tr::T, K, 2), fcm = template_word(Tr::FB, Tr::T, K, 0);
        constexpr uint32_t rth = template_word(Tr::RB, Tr::T, K, 1), rtl = template_word(Tr::RB, Tr::T, K, 2), rcm = template_word(Tr::RB, Tr::T, K, 0);
        wh[K] = __funnelshift_r(R.h[PB + K], R.h[PB + K + 1], p);
        wl[K] = __funnelshift_r(R.l[PB + K], R.l[PB + K + 1], p);
        wn[K] = __funnelshift_r(R.n[PB + K], R.n[PB + K + 1], p);
        const uint32_t th = (Tr::FWD && Tr::REV) ? (rev ? rth : fth) : (Tr::REV ? rth : fth);
        const uint32_t tl = (Tr::FWD && Tr::REV) ? (rev ? rtl : ftl) : (Tr::REV ? rtl : ftl);
        const uint32_t cm = (Tr::FWD && Tr::REV) ? (rev ? rcm : fcm) : (Tr::REV ? rcm : fcm);
        return __popc(((wh[K] ^ th) | (wl[K] ^ tl) | wn[K]) & cm) + verify_words<Tr, K + 1, PB>(R, p, rev, wh, wl, wn);
    }
}

// bits [START, START + LEN) of a window given as words (LEN <= 32)
template <int START, int LEN, int NW>
__device__ __forceinline__ uint32_t window_bits(const uint32_t (&w)[NW]) {
    constexpr int a = START >> 5, sh = START & 31;
    constexpr uint32_t mask = LEN >= 32 ? 0xFFFFFFFFu : ((1u << LEN) - 1u);
    const uint32_t lo = w[a], hi = a + 1 < NW ? w[a + 1] : 0u;
    return (sh == 0 ? lo : __funnelshift_r(lo, hi, sh)) & mask;
}

// flags of the per-read outcome word (`meta`): low 16 bits = window position, bits 16..23 = constant mismatches
constexpr uint32_t SM_CAND = 1u << 31;   // a window passed the verify
constexpr uint32_t SM_REV = 1u << 28;    // the first one is on the reverse strand

// Filter + verify of window block PB and, recursively, the blocks after it.  `ncand` counts the verified windows; the
// FIRST verified window's strand, constant mismatches and position go to `meta`, and `on_first(wh, wl, wn, rev)` is
// called once with its words (variable regions are cut from them).  The first block's first round runs unconditionally
// (most reads carry a candidate); further rounds only while some lane still has one.  `live` = the lane holds a real read.
template <class Tr, int PB, class OnFirst>
__device__ __forceinline__ void scan_blocks(const Words<Tr::W>& R, const Planes<Tr::W>& P, bool live, int& ncand, uint32_t& meta,
                                            OnFirst&& on_first) {
    if constexpr (PB < Dims<Tr>::NBLOCKS) {
        constexpr int TW = Dims<Tr>::TW;
        const uint32_t windows = live ? block_mask<Tr, PB>() : 0u;
        uint32_t cf = 0u, cr = 0u;
        if constexpr (Tr::FWD) cf = ~all_groups_dirty<Tr, false, PB, 0>(P) & windows;
        if constexpr (Tr::REV) cr = ~all_groups_dirty<Tr, true, PB, 0>(P) & windows;
        bool more = PB == 0 ? true : __any_sync(0xFFFFFFFFu, (cf | cr) != 0u);
        while (more) {
            const uint32_t any = cf | cr;
            const uint32_t lowest = any & (0u - any);   // 0 when the lane has no candidate left
            const int p = 31 - __clz(lowest | 1u);
            const bool rev = Tr::FWD ? !(cf & lowest) : true;
            if (rev) {
                cr &= ~lowest;
            } else {
                cf &= ~lowest;
            }
            uint32_t wh[TW + 1], wl[TW + 1], wn[TW + 1];
            wh[TW] = wl[TW] = wn[TW] = 0;
            const int c = verify_words<Tr, 0, PB>(R, p, rev, wh, wl, wn);
            const bool ok = lowest != 0u && c <= Tr::MM;
            if (ok && ncand == 0) {
                meta = SM_CAND + (rev ? SM_REV : 0u) + ((uint32_t)c << 16) + (uint32_t)(32 * PB + p);
                on_first(wh, wl, wn, rev);
            }
            ncand += ok ? 1 : 0;
            more = __any_sync(0xFFFFFFFFu, (cf | cr) != 0u);
        }
        scan_blocks<Tr, PB + 1>(R, P, live, ncand, meta, on_first);
    }
}

// ---- TMA (1-D bulk copy global -> shared) signalled on an mbarrier ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// the packed reads pass through L2 once: their lines are marked evict-first so that the tables stay resident
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
// One elected lane arms the barrier for `bytes_total` and issues one bulk copy (called by the whole warp, converged).
__device__ __forceinline__ void tma_fetch(uint32_t bar, uint32_t dst, const void* src, uint32_t bytes, uint32_t bytes_total, uint64_t policy) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %5;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%2], [%3], %1, [%0], %4;\n\t}"
        ::"r"(bar), "r"(bytes), "r"(dst), "l"(src), "l"(policy), "r"(bytes_total)
        : "memory");
}
// a second copy signalled on a barrier that is already armed for it
__device__ __forceinline__ void tma_fetch_more(uint32_t bar, uint32_t dst, const void* src, uint32_t bytes, uint64_t policy) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%2], [%3], %1, [%0], %4;\n\t}"
        ::"r"(bar), "r"(bytes), "r"(dst), "l"(src), "l"(policy)
        : "memory");
}

} // namespace sscan
} // namespace scg
