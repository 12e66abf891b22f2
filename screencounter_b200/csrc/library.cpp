#include "library.hpp"

#include <algorithm>
#include <cstring>
#include <unordered_map>

namespace scg {

namespace {

// Expansion of an IUPAC code in the order the reference inserts it (MismatchTrie.hpp:152-188).
const char* iupac_expansion(char b) {
    switch (b) {
        case 'R': case 'r': return "AG";
        case 'Y': case 'y': return "CT";
        case 'S': case 's': return "CG";
        case 'W': case 'w': return "AT";
        case 'K': case 'k': return "GT";
        case 'M': case 'm': return "AC";
        case 'B': case 'b': return "CGT";
        case 'D': case 'd': return "AGT";
        case 'H': case 'h': return "ACT";
        case 'V': case 'v': return "ACG";
        case 'N': case 'n': return "ACGT";
    }
    return nullptr;
}

struct Builder {
    int L, KW;
    Duplicates dup;
    std::unordered_map<std::string, int> seen;  // concrete sequence -> pool index
    std::vector<std::string> concrete;          // insertion order
    std::vector<int32_t> index;
    int counter = 0;
    std::string work;

    // MismatchTrie::end (MismatchTrie.hpp:77-105) for one concrete leaf.
    void leaf() {
        auto it = seen.find(work);
        if (it != seen.end()) {
            if (it->second == counter) return;  // cannot happen: expansions of one row are distinct
            if (dup == Duplicates::ERROR) {
                throw Error("duplicate sequences detected (" + std::to_string(it->second + 1) + ", " +
                            std::to_string(counter + 1) + ") when constructing the trie");
            }
            return;  // FIRST: the earlier row keeps the leaf
        }
        seen.emplace(work, counter);
        concrete.push_back(work);
        index.push_back(counter);
    }

    // MismatchTrie::recursive_add (MismatchTrie.hpp:107-190).
    void expand(const std::string& row, int i) {
        while (i < L && base_code(row[i]) >= 0) {
            work[i] = "ACGT"[base_code(row[i])];
            ++i;
        }
        if (i == L) {
            leaf();
            return;
        }
        const char* alts = iupac_expansion(row[i]);
        if (!alts) {
            throw Error(std::string("unknown base '") + row[i] + "' detected when constructing the trie");
        }
        for (; *alts; ++alts) {
            work[i] = *alts;
            expand(row, i + 1);
        }
    }

    void add(const std::string& row) {
        if (L > 0) {
            work.assign(L, 'A');
            expand(row, 0);
        }
        ++counter;
    }
};

// Two-table cuckoo hash (device_keys.cuh probe_table reads it): slot = h[KW], l[KW], value (-1 = empty),
// padding.  A key lives in T1 at hash & mask or in T2 at hash_second(hash) & mask.
struct CuckooTable {
    int KW, slot_words;
    uint32_t n = 0;   // slots per table
    std::vector<uint32_t> slots;

    void reset(uint32_t per_table) {
        n = per_table;
        slots.assign((size_t)2 * n * slot_words, 0);
        for (size_t s = 0; s < (size_t)2 * n; ++s) slots[s * slot_words + 2 * KW] = 0xFFFFFFFFu;
    }
    uint32_t* at(int table, uint32_t pos) { return &slots[((size_t)table * n + pos) * slot_words]; }
    uint32_t position(int table, const uint32_t* key) const {
        const uint32_t acc = hash_key(key, key + KW, KW, 0);
        return (table == 0 ? acc : hash_second(acc)) & (n - 1);
    }
    bool same(const uint32_t* slot, const uint32_t* key) const { return std::memcmp(slot, key, 2 * KW * sizeof(uint32_t)) == 0; }
    static bool empty(const uint32_t* slot, int KW) { return (int32_t)slot[2 * KW] == -1; }

    // false = gave up after too many evictions (the caller rebuilds with larger tables)
    bool insert(const uint32_t* h, const uint32_t* l, int32_t value, bool keep_first) {
        std::vector<uint32_t> cur(2 * KW + 1), tmp(2 * KW + 1);
        std::memcpy(cur.data(), h, KW * sizeof(uint32_t));
        std::memcpy(cur.data() + KW, l, KW * sizeof(uint32_t));
        cur[2 * KW] = (uint32_t)value;
        for (int t = 0; t < 2; ++t) {
            uint32_t* s = at(t, position(t, cur.data()));
            if (!empty(s, KW) && same(s, cur.data())) {
                if (!keep_first) s[2 * KW] = (uint32_t)value;
                return true;
            }
        }
        int table = 0;
        for (int kicks = 0; kicks < 2000; ++kicks) {
            uint32_t* s = at(table, position(table, cur.data()));
            if (empty(s, KW)) {
                std::memcpy(s, cur.data(), (2 * KW + 1) * sizeof(uint32_t));
                return true;
            }
            if (kicks == 0) {  // try the other home before evicting anyone
                uint32_t* o = at(1, position(1, cur.data()));
                if (empty(o, KW)) {
                    std::memcpy(o, cur.data(), (2 * KW + 1) * sizeof(uint32_t));
                    return true;
                }
            }
            std::memcpy(tmp.data(), s, (2 * KW + 1) * sizeof(uint32_t));
            std::memcpy(s, cur.data(), (2 * KW + 1) * sizeof(uint32_t));
            cur.swap(tmp);
            table ^= 1;
        }
        return false;
    }
};

// Builds the table for `count` keys produced by key(e, h, l) / value(e).
template <class KeyFn, class ValueFn>
void build_cuckoo(CuckooTable& t, size_t count, KeyFn key, ValueFn value) {
    uint32_t per_table = std::max<uint32_t>(16, next_pow2((uint32_t)std::min<size_t>(count + 1, 1u << 29)));
    std::vector<uint32_t> h(t.KW), l(t.KW);
    for (;;) {
        t.reset(per_table);
        bool ok = true;
        for (size_t e = 0; e < count && ok; ++e) {
            key(e, h.data(), l.data());
            ok = t.insert(h.data(), l.data(), value(e), true);
        }
        if (ok) return;
        if (per_table >= (1u << 30)) throw Error("could not build the barcode hash table");
        per_table *= 2;
    }
}

} // namespace

std::string reverse_complement_iupac(const std::string& s) {
    std::string out(s.size(), 'N');
    for (size_t j = 0; j < s.size(); ++j) {
        char b = s[s.size() - j - 1];
        char c = 0;
        switch (b) {
            case 'A': case 'a': c = 'T'; break;
            case 'C': case 'c': c = 'G'; break;
            case 'G': case 'g': c = 'C'; break;
            case 'T': case 't': c = 'A'; break;
            case 'N': case 'n': c = 'N'; break;
            case 'R': case 'r': c = 'Y'; break;
            case 'Y': case 'y': c = 'R'; break;
            case 'S': case 's': c = 'S'; break;
            case 'W': case 'w': c = 'W'; break;
            case 'K': case 'k': c = 'M'; break;
            case 'M': case 'm': c = 'K'; break;
            case 'B': case 'b': c = 'V'; break;
            case 'D': case 'd': c = 'H'; break;
            case 'H': case 'h': c = 'D'; break;
            case 'V': case 'v': c = 'B'; break;
        }
        if (!c) throw Error(std::string("cannot complement unknown base '") + b + "'");  // utils.hpp:116-117
        out[j] = c;
    }
    return out;
}

bool pack_key(const char* s, int len, uint32_t* h, uint32_t* l, uint32_t* n) {
    int kw = ceil_div(len, 32);
    for (int w = 0; w < kw; ++w) {
        h[w] = l[w] = 0;
        if (n) n[w] = 0;
    }
    bool clean = true;
    for (int i = 0; i < len; ++i) {
        int code = base_code(s[i]);
        uint32_t bit = 1u << (i & 31);
        if (code < 0) {
            clean = false;
            if (n) n[i >> 5] |= bit;
        } else {
            if (code & 2) h[i >> 5] |= bit;
            if (code & 1) l[i >> 5] |= bit;
        }
    }
    return clean;
}

Library::Library(const std::vector<std::string>& sequences, int length, const LibraryOptions& options)
    : L(length), KW(std::max(1, ceil_div(length, 32))), nchoices((int)sequences.size()), opt(options) {
    if (KW > MAX_KEY_WORDS) {
        throw Error("barcodes longer than " + std::to_string(MAX_KEY_WORDS * 32) + " bp are not supported by this engine");
    }

    Builder b;
    b.L = L;
    b.KW = KW;
    b.dup = options.duplicates;
    for (const auto& row : sequences) b.add(row);

    const size_t E = b.concrete.size();
    ent_keys.assign(E * 2 * KW, 0);
    ent_idx = b.index;
    for (size_t e = 0; e < E; ++e) {
        pack_key(b.concrete[e].data(), L, &ent_keys[e * 2 * KW], &ent_keys[e * 2 * KW + KW], nullptr);
    }

    // Exact table.  Slot = h[KW], l[KW], value, padded to a multiple of 4 words (16-byte loads).
    slot_words = ((2 * KW + 1) + 3) / 4 * 4;
    {
        CuckooTable t;
        t.KW = KW;
        t.slot_words = slot_words;
        build_cuckoo(
            t, E,
            [&](size_t e, uint32_t* h, uint32_t* l) {
                std::memcpy(h, &ent_keys[e * 2 * KW], KW * sizeof(uint32_t));
                std::memcpy(l, &ent_keys[e * 2 * KW + KW], KW * sizeof(uint32_t));
            },
            [&](size_t e) { return ent_idx[e]; });
        slots.swap(t.slots);
        slot_mask = t.n - 1;
    }

    // Seeds.
    std::vector<std::vector<int> > seed_positions;  // base positions of each seed
    auto split = [](int from, int to, int parts) {
        std::vector<std::pair<int, int> > out;
        int len = to - from;
        for (int p = 0; p < parts; ++p) {
            out.emplace_back(from + (int)((long long)len * p / parts), from + (int)((long long)len * (p + 1) / parts));
        }
        return out;
    };
    auto range_positions = [](std::pair<int, int> r) {
        std::vector<int> out;
        for (int i = r.first; i < r.second; ++i) out.push_back(i);
        return out;
    };
    if (!options.segmented) {
        int cap = std::min(std::max(options.max_mismatches, 0), L);
        if (cap > 0) {
            if (cap + 1 > MAX_SEEDS) {
                seed_positions.push_back({});  // one empty seed = every entry is a candidate
            } else {
                for (auto r : split(0, L, cap + 1)) seed_positions.push_back(range_positions(r));
            }
        }
    } else {
        if (options.seg1 < 0 || options.seg1 > L) throw Error("invalid segment length");
        int cap1 = std::min(std::max(options.max_mismatches1, 0), options.seg1);
        int cap2 = std::min(std::max(options.max_mismatches2, 0), L - options.seg1);
        if (cap1 > 0 || cap2 > 0) {
            if ((cap1 + 1) * (cap2 + 1) > MAX_SEEDS) {
                seed_positions.push_back({});
            } else {
                // a candidate within (cap1, cap2) agrees exactly with one part of EACH segment
                for (auto r1 : split(0, options.seg1, cap1 + 1)) {
                    for (auto r2 : split(options.seg1, L, cap2 + 1)) {
                        auto pos = range_positions(r1);
                        auto pos2 = range_positions(r2);
                        pos.insert(pos.end(), pos2.begin(), pos2.end());
                        seed_positions.push_back(pos);
                    }
                }
            }
        }
    }
    nseeds = (int)seed_positions.size();
    seed_masks.assign((size_t)std::max(nseeds, 1) * KW, 0);
    for (int s = 0; s < nseeds; ++s) {
        for (int p : seed_positions[s]) seed_masks[(size_t)s * KW + (p >> 5)] |= 1u << (p & 31);
    }
    if (nseeds > 0) {
        // four buckets per entry: a bucket holding an entry holds a third one 1 % of the time, so the device searches fetch
        // two candidate rows per seed without looping (more buckets would push the tables out of L2)
        nbuckets = std::max<uint32_t>(16, next_pow2((uint32_t)std::min<size_t>(E * 4 + 1, 1u << 22)));
        buckets.assign((size_t)nseeds * nbuckets, make_uint2(0, 0));
        cands.assign((size_t)nseeds * E, 0);
        std::vector<uint32_t> mh(KW), ml(KW), which(E);
        for (int s = 0; s < nseeds; ++s) {
            uint2* bk = &buckets[(size_t)s * nbuckets];
            const uint32_t* mask = &seed_masks[(size_t)s * KW];
            for (size_t e = 0; e < E; ++e) {
                for (int w = 0; w < KW; ++w) {
                    mh[w] = ent_keys[e * 2 * KW + w] & mask[w];
                    ml[w] = ent_keys[e * 2 * KW + KW + w] & mask[w];
                }
                which[e] = hash_key(mh.data(), ml.data(), KW, 0x5EED0000u + s) & (nbuckets - 1);
                ++bk[which[e]].y;
            }
            uint32_t run = 0;
            for (uint32_t k = 0; k < nbuckets; ++k) {
                bk[k].x = run;
                run += bk[k].y;
                bk[k].y = 0;
            }
            int32_t* cd = &cands[(size_t)s * E];
            for (size_t e = 0; e < E; ++e) {
                uint2& slot = bk[which[e]];
                cd[slot.x + slot.y] = (int32_t)e;
                ++slot.y;
            }
        }
        if (KW == 1) {
            cand_rows.assign((size_t)nseeds * E * 4, 0);
            for (size_t k = 0; k < (size_t)nseeds * E; ++k) {
                const size_t e = (size_t)cands[k];
                cand_rows[4 * k + 0] = ent_keys[e * 2];
                cand_rows[4 * k + 1] = ent_keys[e * 2 + 1];
                cand_rows[4 * k + 2] = (uint32_t)ent_idx[e];
            }
        }
    }

    // Rows with the last base dropped (only the segmented search consults it).
    if (options.segmented && L > 0) {
        CuckooTable t;
        t.KW = KW;
        t.slot_words = slot_words;
        const uint32_t lastbit = 1u << ((L - 1) & 31);
        build_cuckoo(
            t, E,
            [&](size_t e, uint32_t* h, uint32_t* l) {
                std::memcpy(h, &ent_keys[e * 2 * KW], KW * sizeof(uint32_t));
                std::memcpy(l, &ent_keys[e * 2 * KW + KW], KW * sizeof(uint32_t));
                h[(L - 1) >> 5] &= ~lastbit;
                l[(L - 1) >> 5] &= ~lastbit;
            },
            [&](size_t) { return 0; });
        prefix_slots.swap(t.slots);
        prefix_mask = t.n - 1;
    }

    // The reference's trie, node for node (MismatchTrie::next / end, MismatchTrie.hpp:66-105), from the concrete rows in
    // insertion order; duplicates were settled by the builder above, so a leaf is a pool index or missing.
    if (options.segmented && options.max_mismatches1 >= 2 && L > 0) {
        trie.assign(4, -1);
        for (size_t e = 0; e < E; ++e) {
            const std::string& row = b.concrete[e];
            int position = 0;
            for (int i = 0; i < L; ++i) {
                const int shift = base_code(row[i]);
                if (i + 1 == L) {
                    if (trie[position + shift] < 0) trie[position + shift] = ent_idx[e];
                } else if (trie[position + shift] < 0) {
                    const int fresh = (int)trie.size();
                    trie[position + shift] = fresh;
                    trie.resize(trie.size() + 4, -1);
                    position = fresh;
                } else {
                    position = trie[position + shift];
                }
            }
        }
    }
}

} // namespace scg
