// Device-side building blocks shared by every handler kernel (nvcc-built and NVRTC-specialised):
//   * the bit-sliced mismatch counter of the position-parallel scan;
//   * extraction of a variable region from the read's bit planes;
//   * the barcode lookups (replace SimpleBarcodeSearch::search / SegmentedBarcodeSearch::search,
//     BarcodeSearch.hpp:243-251, 478-487 and the trie searches of MismatchTrie.hpp:446-660).
#pragma once

#include "layout.hpp"
#include "libdev.hpp"

namespace scg {

// One lane's view of its read inside a tile (layout.hpp): word w of plane p is ptr[(p*W + w)*32].
struct ReadView {
    const uint32_t* __restrict__ ptr;
    int W;
    int len;
    // No bounds check: the scan and the extraction may read up to two words past the read's last
    // plane word; every buffer carries READ_GUARD_BYTES of slack and those bits are masked out.
    __device__ __forceinline__ uint32_t word(int plane, int w) const {
        return __ldg(ptr + (size_t)(plane * W + w) * TILE);
    }
};

// Bit-sliced saturating counter over 32 window positions: CB planes + a sticky overflow plane.
// add(m) adds 1 at every position whose bit is set in m.
template <int CB>
struct Counter {
    uint32_t c[CB > 0 ? CB : 1];
    uint32_t ovf;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < CB; ++i) c[i] = 0;
        ovf = 0;
    }
    __device__ __forceinline__ void add(uint32_t m) {
        uint32_t carry = m;
#pragma unroll
        for (int i = 0; i < CB; ++i) {
            uint32_t t = c[i] & carry;
            c[i] ^= carry;
            carry = t;
        }
        ovf |= carry;
    }
    // positions whose count is <= mm (mm < 2^CB; the caller clamps)
    __device__ __forceinline__ uint32_t le(int mm) const {
        if (mm < 0) return 0u;
        // count <= mm  <=>  not overflowed and not (count > mm); compare bit-sliced from the top
        uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
        for (int i = CB - 1; i >= 0; --i) {
            uint32_t mbit = ((mm >> i) & 1) ? 0xFFFFFFFFu : 0u;
            gt |= eq & c[i] & ~mbit;
            eq &= ~(c[i] ^ mbit);
        }
        return ~ovf & ~gt;
    }
    __device__ __forceinline__ int get(int p) const {
        int v = 0;
#pragma unroll
        for (int i = 0; i < CB; ++i) v |= ((c[i] >> p) & 1u) << i;
        return v;
    }
};

// Windows of block pb that exist in a read of length len: positions p with 32*pb + p + T <= len.
__device__ __forceinline__ uint32_t valid_windows(int len, int T, int pb) {
    int npos = len - T + 1 - 32 * pb;
    if (npos <= 0) return 0u;
    return npos >= 32 ? 0xFFFFFFFFu : ((1u << npos) - 1u);
}


__device__ __forceinline__ ReadView read_view(const ReadsDev& r, long long tile, int lane) {
    ReadView v;
    v.ptr = r.data + (size_t)tile * tile_words(r.W) + lane;
    v.W = r.W;
    const long long i = tile * TILE + lane;
    v.len = (i < r.n) ? (r.lens ? (int)r.lens[i] : r.uniform_len) : 0;
    return v;
}

__device__ __forceinline__ int window_blocks(int len, int T) {
    const int npos = len - T + 1;
    return npos <= 0 ? 0 : (npos + 31) >> 5;
}

// info word of the per-read trace (include/scg.h, scg_result_copy_trace)
__device__ __forceinline__ uint32_t pack_info(bool found, bool reverse, int mismatches, int var_mismatches, int position) {
    if (!found) return 0u;
    return 0x80000000u | (reverse ? 0x40000000u : 0u) | ((uint32_t)min(var_mismatches, 31) << 25) |
           ((uint32_t)min(mismatches, 31) << 20) | ((uint32_t)position & 0xFFFFFu);
}

// outcome of one single-barcode search (SimpleSingleMatch::State, SimpleSingleMatch.hpp:103-140)
struct SingleOut {
    bool found;
    int index;
    int position;
    bool reverse;
    int mismatches;
    int var_mismatches;
};

// ---- keys ---------------------------------------------------------------------------------

template <int KW>
struct Key {
    uint32_t h[KW], l[KW], n[KW];
};

// Extract `len` bases starting at read bit `start` as a key (KW static words; no dynamic register
// indexing, so the key stays in registers).
template <int KW>
__device__ __forceinline__ void extract_region(const ReadView& rd, int start, int len, Key<KW>& k) {
    const int a = start >> 5, sh = start & 31;
    uint32_t h0 = rd.word(PLANE_H, a), l0 = rd.word(PLANE_L, a), n0 = rd.word(PLANE_N, a);
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        const int rem = len - 32 * w;
        if (rem > 0) {
            const uint32_t h1 = rd.word(PLANE_H, a + w + 1), l1 = rd.word(PLANE_L, a + w + 1), n1 = rd.word(PLANE_N, a + w + 1);
            const uint32_t m = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
            k.h[w] = __funnelshift_r(h0, h1, sh) & m;
            k.l[w] = __funnelshift_r(l0, l1, sh) & m;
            k.n[w] = __funnelshift_r(n0, n1, sh) & m;
            h0 = h1;
            l0 = l1;
            n0 = n1;
        } else {
            k.h[w] = k.l[w] = k.n[w] = 0;
        }
    }
}

// OR `len` bases starting at read bit `start` into an existing key at bit offset dst_bit
// (regions concatenated for the dual designs).  Every destination word is visited statically.
template <int KW>
__device__ __forceinline__ void extract_into(const ReadView& rd, int start, int len, int dst_bit, Key<KW>& k) {
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        const int d0 = max(dst_bit, 32 * w), d1 = min(dst_bit + len, 32 * w + 32);
        if (d0 < d1) {
            const int sbit = start + (d0 - dst_bit);
            const int a = sbit >> 5, sh = sbit & 31;
            const int take = d1 - d0;
            const uint32_t m = take >= 32 ? 0xFFFFFFFFu : ((1u << take) - 1u);
            const int db = d0 - 32 * w;
            k.h[w] |= (__funnelshift_r(rd.word(PLANE_H, a), rd.word(PLANE_H, a + 1), sh) & m) << db;
            k.l[w] |= (__funnelshift_r(rd.word(PLANE_L, a), rd.word(PLANE_L, a + 1), sh) & m) << db;
            k.n[w] |= (__funnelshift_r(rd.word(PLANE_N, a), rd.word(PLANE_N, a + 1), sh) & m) << db;
        }
    }
}

template <int KW>
__device__ __forceinline__ void key_clear(Key<KW>& k) {
#pragma unroll
    for (int w = 0; w < KW; ++w) k.h[w] = k.l[w] = k.n[w] = 0;
}

// Reverse complement of a key of `len` bases (used by the random-barcode handler).
template <int KW>
__device__ __forceinline__ void key_revcomp(Key<KW>& k, int len) {
    // reverse all KW*32 bits, then shift right by (KW*32 - len); complement = flip both planes
    uint32_t rh[KW], rl[KW], rn[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        rh[w] = __brev(k.h[KW - 1 - w]);
        rl[w] = __brev(k.l[KW - 1 - w]);
        rn[w] = __brev(k.n[KW - 1 - w]);
    }
    const int shift = KW * 32 - len;
    const int ws = shift >> 5, bs = shift & 31;
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        uint32_t lo_h = (w + ws < KW) ? rh[w + ws] : 0u, hi_h = (w + ws + 1 < KW) ? rh[w + ws + 1] : 0u;
        uint32_t lo_l = (w + ws < KW) ? rl[w + ws] : 0u, hi_l = (w + ws + 1 < KW) ? rl[w + ws + 1] : 0u;
        uint32_t lo_n = (w + ws < KW) ? rn[w + ws] : 0u, hi_n = (w + ws + 1 < KW) ? rn[w + ws + 1] : 0u;
        k.h[w] = __funnelshift_r(lo_h, hi_h, bs);
        k.l[w] = __funnelshift_r(lo_l, hi_l, bs);
        k.n[w] = __funnelshift_r(lo_n, hi_n, bs);
    }
    // complement the called bases only (an N stays an N with H = L = 0)
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        const int rem = len - 32 * w;
        const uint32_t m = rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
        k.h[w] = (~k.h[w]) & m & ~k.n[w];
        k.l[w] = (~k.l[w]) & m & ~k.n[w];
    }
}

// ---- lookups --------------------------------------------------------------------------------

struct Hit {
    int index;  // pool index, or -1 (missing or ambiguous; the handlers treat both alike)
    int dist;
};

template <int KW>
__device__ __forceinline__ bool key_has_n(const Key<KW>& k) {
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < KW; ++w) any |= k.n[w];
    return any != 0;
}

// Probe a two-table cuckoo hash of packed keys (library.cpp CuckooTable): the key, if present, sits
// at one of exactly two slots, so both are fetched at once -- two independent loads, no loop, no
// divergence between lanes.  Returns the slot's value or -1.  hm/lm are the (masked) key planes to
// look for; kw (<= KW) is the table's own number of words per plane.
template <int KW>
__device__ __forceinline__ int probe_table(const uint32_t* __restrict__ slots, uint32_t mask, int slot_words, int kw,
                                           const uint32_t* hm, const uint32_t* lm) {
    const uint32_t acc = hash_key(hm, lm, KW == 1 ? 1 : kw, 0);
    const uint32_t* s1 = slots + (size_t)(acc & mask) * slot_words;
    const uint32_t* s2 = slots + ((size_t)(mask + 1) + (hash_second(acc) & mask)) * slot_words;
    if (KW == 1) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(s1));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(s2));
        // values are >= 0; an empty slot holds -1 (and zero key planes, which a poly-A query also has)
        const int ra = (a.x == hm[0] && a.y == lm[0]) ? (int)a.z : -1;
        const int rb = (b.x == hm[0] && b.y == lm[0]) ? (int)b.z : -1;
        return max(ra, rb);
    } else {
        int out = -1;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint32_t* s = t ? s2 : s1;
            const int val = (int)__ldg(s + 2 * kw);
            bool same = val != -1;
#pragma unroll
            for (int w = 0; w < KW; ++w) {
                if (w < kw) same &= (__ldg(s + w) == hm[w]) & (__ldg(s + kw + w) == lm[w]);
            }
            if (same) out = val;
        }
        return out;
    }
}

// Candidate enumeration through the pigeonhole seeds + verification (the mismatch-tolerant part of
// both searches).  seg1 < 0: one cap (c1) on the total distance; seg1 >= 0: caps (c1, c2) on the
// two segments [0, seg1) and [seg1, L).  Applies the best-unique / tie rules of
// MismatchTrie.hpp:266-343 to the verified distances.
template <int KW>
__device__ __forceinline__ Hit lookup_seeded_body(const LibDev* __restrict__ libp, const Key<KW>& q, int c1, int c2, int seg1) {
    const LibDev lib = *libp;
    Hit out{ -1, 0 };
    const int kw = KW == 1 ? 1 : lib.KW;
    const bool segmented = seg1 >= 0;
    uint32_t s1m[KW];  // base positions of the first segment
#pragma unroll
    for (int w = 0; w < KW; ++w) {
        const int rem = (segmented ? seg1 : 0) - 32 * w;
        s1m[w] = rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
    }
    const int cap = segmented ? c1 + c2 : c1;
    int best = cap + 1, bidx = -1;
    bool ambiguous = false;
    for (int sd = 0; sd < lib.nseeds; ++sd) {
        const uint32_t* sm = lib.seed_masks + sd * kw;
        uint32_t mh[KW], ml[KW], bad = 0;
#pragma unroll
        for (int w = 0; w < KW; ++w) {
            const uint32_t m = (w < kw) ? __ldg(sm + w) : 0u;
            mh[w] = q.h[w] & m;
            ml[w] = q.l[w] & m;
            bad |= q.n[w] & m;
        }
        if (bad) continue;  // an N inside the seed: no barcode agrees with the query there
        const uint32_t b = hash_key(mh, ml, kw, 0x5EED0000u + sd) & lib.bucket_mask;
        const uint2 bk = __ldg(lib.buckets + (size_t)sd * (lib.bucket_mask + 1) + b);
        const int32_t* cd = lib.cands + (size_t)sd * lib.nentries + bk.x;
        for (uint32_t c = 0; c < bk.y; ++c) {
            const int e = __ldg(cd + c);
            const uint32_t* ek = lib.ent_keys + (size_t)e * 2 * kw;
            int d1 = 0, d2 = 0;
#pragma unroll
            for (int w = 0; w < KW; ++w) {
                if (w < kw) {
                    const uint32_t diff = (q.h[w] ^ __ldg(ek + w)) | (q.l[w] ^ __ldg(ek + kw + w)) | q.n[w];
                    d1 += __popc(diff & s1m[w]);
                    d2 += __popc(diff & ~s1m[w]);
                }
            }
            if (segmented ? (d1 > c1 || d2 > c2) : (d2 > c1)) continue;
            const int d = d1 + d2;
            if (d > best) continue;
            const int idx = __ldg(lib.ent_idx + e);
            if (d < best) {
                best = d;
                bidx = idx;
                ambiguous = false;
            } else if (idx != bidx) {  // d == best
                if (lib.dup_first) {
                    bidx = min(bidx, idx);
                } else {
                    ambiguous = true;
                }
            }
        }
    }
    if (bidx >= 0 && !ambiguous) {
        out.index = bidx;
        out.dist = best;
    }
    return out;
}

// Out of line by default (the handler kernels call it from several places); INLINE = true folds it into the caller.
template <int KW>
__device__ __noinline__ Hit lookup_seeded(const LibDev* __restrict__ libp, const Key<KW>& q, int c1, int c2, int seg1) {
    return lookup_seeded_body<KW>(libp, q, c1, c2, seg1);
}

// Best-unique search with a single cap (AnyMismatches::search semantics, MismatchTrie.hpp:446-501):
// the minimum distance over the library if it is <= cap and attained by one pool index, else a miss.
// A query position holding N mismatches every barcode.
template <int KW, bool INLINE = false>
__device__ __forceinline__ Hit lookup_any(const LibDev* __restrict__ lib, const Key<KW>& q, int cap) {
    Hit out{ -1, 0 };
    const int kw = KW == 1 ? 1 : lib->KW;
    if (!key_has_n(q)) {
        const int v = probe_table<KW>(lib->slots, lib->slot_mask, KW == 1 ? 4 : lib->slot_words, kw, q.h, q.l);
        if (v >= 0) {
            out.index = v;
            return out;
        }
    }
    if (cap <= 0 || lib->nseeds == 0) return out;
    int nbad = 0;
#pragma unroll
    for (int w = 0; w < KW; ++w) nbad += __popc(q.n[w]);
    if (nbad > cap) return out;
    if (INLINE) return lookup_seeded_body<KW>(lib, q, min(cap, lib->L), 0, -1);
    return lookup_seeded<KW>(lib, q, min(cap, lib->L), 0, -1);
}

// SegmentedMismatches<2>::search (MismatchTrie.hpp:577-660) walked over the reference's own trie, frame by frame: the exact
// child first, whose result becomes `best` by plain assignment (:624-627 -- this is where the phantom "missing, total + 1"
// of a dead end enters), then the alternates in A, C, G, T order while the shared mismatch cap (lowered by every hit, :602)
// still allows them, merged with replace_best_with_chosen (:266-299).  Only the caps the table search cannot reproduce come
// here ([>= 2, 0]); it is correct for any caps, which is how the tests check it.  One lane, iterative, frames in local memory.
constexpr int TRIE_MAX_LEN = 64;
template <int KW>
__device__ __noinline__ Hit trie_search_segmented(const LibDev* __restrict__ libp, const Key<KW>& q, int c1, int c2) {
    const int32_t* __restrict__ ptr = libp->trie;
    const int L = libp->L, seg1 = libp->seg1;
    constexpr int MISSING = -1, AMBIGUOUS = -2;
    int cap = c1 + c2;
    int node[TRIE_MAX_LEN], tot[TRIE_MAX_LEN], bidx[TRIE_MAX_LEN], btot[TRIE_MAX_LEN];
    signed char m0[TRIE_MAX_LEN], m1[TRIE_MAX_LEN], sh[TRIE_MAX_LEN], snext[TRIE_MAX_LEN], phase[TRIE_MAX_LEN], alts_ok[TRIE_MAX_LEN];
    int d = 0, ret_idx = MISSING, ret_tot = 0;
    bool entering = true;
    node[0] = 0;
    tot[0] = 0;
    m0[0] = m1[0] = 0;
    auto base_at = [&](int pos) -> int {
        const int w = pos >> 5, bit = pos & 31;
        uint32_t h = 0, l = 0, n = 0;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            if (k == w) {
                h = q.h[k];
                l = q.l[k];
                n = q.n[k];
            }
        }
        if ((n >> bit) & 1u) return -1;
        return (int)((((h >> bit) & 1u) << 1) | ((l >> bit) & 1u));
    };
    while (d >= 0) {
        const int seg = d < seg1 ? 0 : 1;
        const int segcap = seg == 0 ? c1 : c2;
        if (entering) {
            const int shift = base_at(d);
            const int current = shift >= 0 ? ptr[node[d] + shift] : MISSING;
            if (d + 1 == L) {
                // the last position (:595-613)
                if (current >= 0 || current == AMBIGUOUS) {
                    cap = tot[d];
                    ret_idx = current;
                    ret_tot = tot[d];
                } else {
                    int idx = MISSING;
                    const int t = tot[d] + 1;
                    const int segmm = (seg == 0 ? m0[d] : m1[d]) + 1;
                    if (t <= cap && segmm <= segcap) {
                        // scan_final_position_with_mismatch (:302-343)
                        bool found = false;
                        for (int s = 0; s < 4; ++s) {
                            if (s == shift) continue;
                            const int cand = ptr[node[d] + s];
                            if (cand >= 0) {
                                if (found) {
                                    if (cand != idx) {
                                        if (libp->dup_first) {
                                            if (idx > cand) idx = cand;
                                        } else {
                                            idx = AMBIGUOUS;
                                            break;
                                        }
                                    }
                                } else {
                                    idx = cand;
                                    cap = t;
                                    found = true;
                                }
                            } else if (cand == AMBIGUOUS) {
                                idx = AMBIGUOUS;
                                cap = t;
                                break;
                            }
                        }
                    }
                    ret_idx = idx;
                    ret_tot = t;
                }
                entering = false;
                --d;
                continue;
            }
            sh[d] = (signed char)shift;
            bidx[d] = MISSING;
            btot[d] = cap + 1;
            if (current >= 0) {
                phase[d] = 1;
                node[d + 1] = current;
                tot[d + 1] = tot[d];
                m0[d + 1] = m0[d];
                m1[d + 1] = m1[d];
                ++d;
                continue;   // entering the exact child
            }
            phase[d] = 2;
            const int t = tot[d] + 1, segmm = (seg == 0 ? m0[d] : m1[d]) + 1;
            alts_ok[d] = (t <= cap && segmm <= segcap) ? 1 : 0;
            snext[d] = 0;
        } else if (phase[d] == 1) {
            // back from the exact child: its result IS best (:624-627)
            bidx[d] = ret_idx;
            btot[d] = ret_tot;
            phase[d] = 2;
            const int t = tot[d] + 1, segmm = (seg == 0 ? m0[d] : m1[d]) + 1;
            alts_ok[d] = (t <= cap && segmm <= segcap) ? 1 : 0;
            snext[d] = 0;
        } else {
            // back from an alternate: replace_best_with_chosen (:266-299)
            if (ret_idx >= 0) {
                if (ret_tot < btot[d]) {
                    bidx[d] = ret_idx;
                    btot[d] = ret_tot;
                } else if (ret_tot == btot[d] && ret_idx != bidx[d]) {
                    if (libp->dup_first) {
                        if (ret_idx < bidx[d]) bidx[d] = ret_idx;
                    } else {
                        bidx[d] = AMBIGUOUS;
                    }
                }
            } else if (ret_idx == AMBIGUOUS) {
                if (ret_tot < btot[d]) {
                    bidx[d] = ret_idx;
                    btot[d] = ret_tot;
                } else if (ret_tot == btot[d]) {
                    bidx[d] = AMBIGUOUS;
                }
            }
        }
        // the alternates of frame d (:633-650)
        entering = false;
        if (alts_ok[d]) {
            const int t = tot[d] + 1;
            while (snext[d] < 4) {
                const int s = snext[d]++;
                if (s == sh[d]) continue;
                const int alt = ptr[node[d] + s];
                if (alt < 0) continue;
                if (t <= cap) {
                    node[d + 1] = alt;
                    tot[d + 1] = t;
                    m0[d + 1] = (signed char)(m0[d] + (seg == 0 ? 1 : 0));
                    m1[d + 1] = (signed char)(m1[d] + (seg == 1 ? 1 : 0));
                    ++d;
                    entering = true;
                    break;
                }
            }
        }
        if (entering) continue;
        ret_idx = bidx[d];
        ret_tot = btot[d];
        --d;
    }
    Hit out{ -1, 0 };
    if (ret_idx >= 0) {
        out.index = ret_idx;
        out.dist = ret_tot;
    }
    return out;
}

// Best-unique search with one cap per segment (SegmentedMismatches<2>::search,
// MismatchTrie.hpp:577-660), cache-free semantics, including the phantom result of :608-617
// when the second segment's cap is 0 ("Quirk A"):
//   c2 == 0, c1 == 1 : if the query minus its last base is free of N and is a prefix of a library
//                      row, the search reports no match ("root rule");
//   c2 == 0, c1 >= 2 : the reference's own walk over its trie (trie_search_segmented above).
// the part of the search that follows a missed (or impossible) exact probe
template <int KW>
__device__ __forceinline__ Hit lookup_segmented_inexact(const LibDev* __restrict__ libp, const Key<KW>& q, int c1, int c2) {
    Hit out{ -1, 0 };
    const LibDev lib = *libp;
    const int kw = KW == 1 ? 1 : lib.KW;
    if ((c1 <= 0 && c2 <= 0) || lib.nseeds == 0) return out;
    // caps [>= 2, 0]: only the reference's own walk over its trie gives the reference's answer
    if (c2 == 0 && c1 >= 2 && lib.trie != nullptr) return trie_search_segmented<KW>(libp, q, c1, 0);
    if (c2 == 0 && c1 >= 1) {
        // root rule: the exact chain down to the last base exists -> the phantom (MISSING, 1) ties or beats every hit
        uint32_t ph[KW], pl[KW], pn = 0;
        const int lw = (lib.L - 1) >> 5;
        const uint32_t lastbit = 1u << ((lib.L - 1) & 31);
#pragma unroll
        for (int w = 0; w < KW; ++w) {
            ph[w] = q.h[w];
            pl[w] = q.l[w];
            uint32_t nn = q.n[w];
            if (w == lw) {
                ph[w] &= ~lastbit;
                pl[w] &= ~lastbit;
                nn &= ~lastbit;
            }
            pn |= nn;
        }
        if (!pn && probe_table<KW>(lib.prefix_slots, lib.prefix_mask, lib.slot_words, kw, ph, pl) >= 0) return out;
    }
    return lookup_seeded<KW>(libp, q, min(c1, lib.seg1), min(c2, lib.L - lib.seg1), lib.seg1);
}

template <int KW>
__device__ __forceinline__ Hit lookup_segmented(const LibDev* __restrict__ libp, const Key<KW>& q, int c1, int c2) {
    if (!key_has_n(q)) {
        const int kw = KW == 1 ? 1 : libp->KW;
        const int v = probe_table<KW>(libp->slots, libp->slot_mask, libp->slot_words, kw, q.h, q.l);
        if (v >= 0) return Hit{ v, 0 };
    }
    return lookup_segmented_inexact<KW>(libp, q, c1, c2);
}

} // namespace scg
