/* TEST INFRASTRUCTURE ONLY -- see kaori_port.h.  Plain-C restatement of the
 * reference algorithm; citations are to /root/reference/inst/include/kaori/...
 * (kaori v1.1.1) unless a longer path is given.
 *
 * Deliberately simple: char-level loops, one flat trie, no caches.  The
 * reference's search caches (BarcodeSearch.hpp:62-93) are transparent for the
 * any-mismatch search and are NOT reproduced for the segmented search, where
 * they make the reference order-dependent: this port states
 * the cache-free semantics.
 */
#define _GNU_SOURCE
#include "kaori_port.h"

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define MISSING (-1)   /* MismatchTrie.hpp:27 */
#define AMBIGUOUS (-2) /* MismatchTrie.hpp:28 */

enum { DUP_FIRST = 0, DUP_LAST = 1, DUP_NONE = 2, DUP_ERROR = 3 }; /* utils.hpp:18 */

static __thread char g_err[1024];

const char* kport_last_error(void) { return g_err; }

static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* bases: utils.hpp:41-133                                                    */
/* ------------------------------------------------------------------------- */

static int base_code(char b) { /* MismatchTrie.hpp:232-254 base_shift; -1 = non-standard */
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return -1;
}

/* complement_base<allow_n, allow_iupac>, utils.hpp:41-120.  Returns 0 for "throw". */
static char complement(char b, int allow_n, int allow_iupac) {
    switch (toupper((unsigned char)b)) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        case 'N': return (allow_n || allow_iupac) ? 'N' : 0;
    }
    if (!allow_iupac) return 0;
    switch (toupper((unsigned char)b)) {
        case 'R': return 'Y';
        case 'Y': return 'R';
        case 'S': return 'S';
        case 'W': return 'W';
        case 'K': return 'M';
        case 'M': return 'K';
        case 'B': return 'V';
        case 'D': return 'H';
        case 'H': return 'D';
        case 'V': return 'B';
    }
    return 0;
}

/* IUPAC expansion in the order MismatchTrie.hpp:152-188 visits it. */
static const char* iupac_expansion(char b) {
    switch (toupper((unsigned char)b)) {
        case 'R': return "AG";
        case 'Y': return "CT";
        case 'S': return "CG";
        case 'W': return "AT";
        case 'K': return "GT";
        case 'M': return "AC";
        case 'B': return "CGT";
        case 'D': return "AGT";
        case 'H': return "ACT";
        case 'V': return "ACG";
        case 'N': return "ACGT";
    }
    return NULL;
}

/* ------------------------------------------------------------------------- */
/* FASTQ: FastqReader.hpp:42-110 over byteme::PerByte (byteme/PerByte.hpp:48-110) */
/* ------------------------------------------------------------------------- */

typedef struct {
    char* buf;
    size_t* off;
    size_t n, cap_n, nbuf, cap_buf;
} reads_t;

static void reads_free(reads_t* r) {
    free(r->buf);
    free(r->off);
    memset(r, 0, sizeof *r);
}

static void reads_push(reads_t* r, const char* s, size_t len) {
    if (r->off == NULL) {
        r->cap_n = 1024;
        r->off = malloc((r->cap_n + 1) * sizeof(size_t));
        r->off[0] = 0;
    }
    if (r->n + 1 > r->cap_n) {
        r->cap_n *= 2;
        r->off = realloc(r->off, (r->cap_n + 1) * sizeof(size_t));
    }
    if (r->nbuf + len > r->cap_buf) {
        r->cap_buf = (r->nbuf + len) * 2 + 1024;
        r->buf = realloc(r->buf, r->cap_buf);
    }
    memcpy(r->buf + r->nbuf, s, len);
    r->nbuf += len;
    r->off[++r->n] = r->nbuf;
}

/* byteme/SomeFileReader.hpp:25-66: gzip is sniffed from the magic bytes; zlib's
 * gzread does the same sniffing and passes plain files through. */
static int slurp_file(const char* path, char** out, size_t* n) {
    gzFile f = gzopen(path, "rb");
    if (!f) return fail("failed to open file at '%s'", path);
    size_t cap = 1 << 20, len = 0;
    char* buf = malloc(cap);
    for (;;) {
        if (len == cap) {
            cap *= 2;
            buf = realloc(buf, cap);
        }
        int got = gzread(f, buf + len, (unsigned)((cap - len) > (1u << 30) ? (1u << 30) : (cap - len)));
        if (got < 0) {
            gzclose(f);
            free(buf);
            return fail("failed to read file at '%s'", path);
        }
        if (got == 0) break;
        len += (size_t)got;
    }
    gzclose(f);
    *out = buf;
    *n = len;
    return 0;
}

static int parse_fastq(const char* path, const char* data, size_t size, reads_t* out) {
    char* owned = NULL;
    if (path) {
        if (slurp_file(path, &owned, &size)) return 1;
        data = owned;
    }
    memset(out, 0, sizeof *out);
    reads_push(out, "", 0); /* allocate, then rewind */
    out->n = 0;
    out->nbuf = 0;

    size_t pos = 0;
    int okay = size > 0; /* PerByte::valid() at construction, FastqReader.hpp:27-31 */
    int line_count = 0;
    size_t seq_cap = 256, seq_len;
    char* seq = malloc(seq_cap);
    int status = 0;

#define ADVANCE_OR_FAIL()                                                                   \
    do {                                                                                    \
        if (++pos >= size) {                                                                \
            status = fail("premature end of the file at line %d", line_count + 1);          \
            goto done;                                                                      \
        }                                                                                   \
    } while (0)

    while (okay) {
        int init_line = line_count;
        if (data[pos] != '@') { /* FastqReader.hpp:54-57 */
            status = fail("read name should start with '@' (starting line %d)", init_line + 1);
            goto done;
        }
        ADVANCE_OR_FAIL();
        while (!isspace((unsigned char)data[pos]) || (unsigned char)data[pos] >= 0x80) ADVANCE_OR_FAIL(); /* name, :59-63 */
        while (data[pos] != '\n') ADVANCE_OR_FAIL();                                                      /* rest of line 1, :65-67 */
        ++line_count;

        seq_len = 0; /* sequence up to '+', newlines dropped, everything else (incl. '\r') kept, :70-78 */
        ADVANCE_OR_FAIL();
        while (data[pos] != '+') {
            if (data[pos] != '\n') {
                if (seq_len == seq_cap) {
                    seq_cap *= 2;
                    seq = realloc(seq, seq_cap);
                }
                seq[seq_len++] = data[pos];
            }
            ADVANCE_OR_FAIL();
        }
        ++line_count;

        ADVANCE_OR_FAIL(); /* '+' line, :81-85 */
        while (data[pos] != '\n') ADVANCE_OR_FAIL();
        ++line_count;

        size_t qual = 0; /* qualities, :91-105 */
        okay = 0;
        while (++pos < size) {
            if (data[pos] != '\n') {
                ++qual;
            } else if (qual >= seq_len) {
                okay = (++pos < size);
                break;
            }
        }
        if (qual != seq_len) {
            status = fail("non-equal lengths for quality and sequence strings (starting line %d)", init_line + 1);
            goto done;
        }
        ++line_count;
        reads_push(out, seq, seq_len);
    }

done:
#undef ADVANCE_OR_FAIL
    free(seq);
    free(owned);
    if (status) reads_free(out);
    return status;
}

/* ------------------------------------------------------------------------- */
/* Template: ScanTemplate.hpp:53-95 (constructor), :233-252 (strand_match)    */
/* ------------------------------------------------------------------------- */

#define MAX_TEMPLATE 256
#define MAX_REGIONS 128

typedef struct {
    int T;
    int do_fwd, do_rev;
    char fwd[MAX_TEMPLATE + 1], rev[MAX_TEMPLATE + 1];
    int nreg;
    int fstart[MAX_REGIONS], fend[MAX_REGIONS]; /* variable_regions<false>() */
    int rstart[MAX_REGIONS], rend[MAX_REGIONS]; /* variable_regions<true>()  */
} tmpl_t;

static int tmpl_init(tmpl_t* t, const char* s, int strand /* 0 fwd, 1 rev, 2 both: utils.hpp:23-37 */) {
    size_t len = strlen(s);
    if (len > MAX_TEMPLATE) { /* src/count_single_barcodes.cpp:45-46 */
        return fail("lacking compile-time support for constant regions longer than 256 bp");
    }
    memset(t, 0, sizeof *t);
    t->T = (int)len;
    t->do_fwd = (strand == 0 || strand == 2);
    t->do_rev = (strand == 1 || strand == 2);
    /* forward pass, ScanTemplate.hpp:59-81: bases are only validated when the forward strand is searched */
    for (int i = 0; i < t->T; ++i) {
        char b = s[i];
        if (b == '-') {
            t->fwd[i] = '-';
            if (t->nreg && t->fend[t->nreg - 1] == i) {
                ++t->fend[t->nreg - 1];
            } else {
                if (t->nreg == MAX_REGIONS) return fail("too many variable regions");
                t->fstart[t->nreg] = i;
                t->fend[t->nreg] = i + 1;
                ++t->nreg;
            }
        } else {
            if (t->do_fwd && base_code(b) < 0) return fail("unknown base '%c'", b); /* utils.hpp:140-161 */
            t->fwd[i] = b;
        }
    }
    /* reverse pass, ScanTemplate.hpp:82-94 */
    if (t->do_rev) {
        int nr = 0;
        for (int i = 0; i < t->T; ++i) {
            char b = s[t->T - i - 1];
            if (b == '-') {
                t->rev[i] = '-';
                if (nr && t->rend[nr - 1] == i) {
                    ++t->rend[nr - 1];
                } else {
                    t->rstart[nr] = i;
                    t->rend[nr] = i + 1;
                    ++nr;
                }
            } else {
                char c = complement(b, 0, 0);
                if (!c) return fail("cannot complement unknown base '%c'", b); /* utils.hpp:116-117 */
                t->rev[i] = c;
            }
        }
    }
    return 0;
}

/* Net effect of ScanTemplate::next + strand_match (ScanTemplate.hpp:183-252):
 * Hamming distance over the constant positions; any read char outside ACGTacgt
 * is one mismatch. */
static int const_mm(const char* read, const char* pattern, int T) {
    int mm = 0;
    for (int i = 0; i < T; ++i) {
        if (pattern[i] == '-') continue;
        int a = base_code(read[i]);
        if (a < 0 || a != base_code(pattern[i])) ++mm;
    }
    return mm;
}

/* ------------------------------------------------------------------------- */
/* Trie: MismatchTrie.hpp                                                     */
/* ------------------------------------------------------------------------- */

typedef struct {
    int* p;
    int n, cap;
    int length;
    int dup;
    int counter;
    /* segmented search only */
    int nseg;
    int boundaries[8];
} trie_t;

static void trie_init(trie_t* t, int length, int dup) { /* MismatchTrie.hpp:45-50 */
    memset(t, 0, sizeof *t);
    t->length = length;
    t->dup = dup;
    t->cap = 1024;
    t->p = malloc(t->cap * sizeof(int));
    for (int i = 0; i < 4; ++i) t->p[i] = MISSING;
    t->n = 4;
}

static void trie_free(trie_t* t) {
    free(t->p);
    t->p = NULL;
}

static int trie_descend(trie_t* t, int node, int shift) { /* MismatchTrie.hpp:66-75 next() */
    int cur = t->p[node + shift];
    if (cur >= 0) return cur;
    if (t->n + 4 > t->cap) {
        t->cap *= 2;
        t->p = realloc(t->p, t->cap * sizeof(int));
    }
    int fresh = t->n;
    for (int i = 0; i < 4; ++i) t->p[fresh + i] = MISSING;
    t->n += 4;
    t->p[node + shift] = fresh;
    return fresh;
}

static int trie_end(trie_t* t, int node, int shift) { /* MismatchTrie.hpp:77-105 end() */
    int* cur = &t->p[node + shift];
    if (*cur >= 0) {
        switch (t->dup) {
            case DUP_FIRST: break;
            case DUP_LAST: *cur = t->counter; break;
            case DUP_NONE: *cur = AMBIGUOUS; break;
            default:
                return fail("duplicate sequences detected (%d, %d) when constructing the trie", *cur + 1, t->counter + 1);
        }
    } else if (*cur == MISSING) {
        *cur = t->counter;
    }
    return 0;
}

static int trie_add_from(trie_t* t, const char* seq, int i, int node) { /* MismatchTrie.hpp:107-190 recursive_add */
    for (;;) {
        int shift = base_code(seq[i]);
        if (shift < 0) break;
        if (++i == t->length) return trie_end(t, node, shift);
        node = trie_descend(t, node, shift);
    }
    const char* alts = iupac_expansion(seq[i]);
    if (!alts) return fail("unknown base '%c' detected when constructing the trie", seq[i]);
    for (; *alts; ++alts) {
        int shift = base_code(*alts);
        if (i + 1 == t->length) {
            if (trie_end(t, node, shift)) return 1;
        } else {
            int child = trie_descend(t, node, shift);
            if (trie_add_from(t, seq, i + 1, child)) return 1;
        }
    }
    return 0;
}

static int trie_add(trie_t* t, const char* seq) { /* MismatchTrie.hpp:200-205 */
    if (t->length == 0) {
        ++t->counter;
        return 0;
    }
    if (trie_add_from(t, seq, 0, 0)) return 1;
    ++t->counter;
    return 0;
}

typedef struct {
    int index;
    int total;
    int per_segment[8];
} hit_t;

/* MismatchTrie.hpp:266-297 replace_best_with_chosen */
static void merge_chosen(const trie_t* t, hit_t* best, const hit_t* chosen) {
    if (chosen->index >= 0) {
        if (chosen->total < best->total) {
            *best = *chosen;
        } else if (chosen->total == best->total && chosen->index != best->index) {
            if (t->dup == DUP_FIRST) {
                if (chosen->index < best->index) best->index = chosen->index;
            } else if (t->dup == DUP_LAST) {
                if (chosen->index > best->index) best->index = chosen->index;
            } else {
                best->index = AMBIGUOUS;
            }
        }
    } else if (chosen->index == AMBIGUOUS) {
        if (chosen->total < best->total) {
            *best = *chosen;
        } else if (chosen->total == best->total) {
            best->index = AMBIGUOUS;
        }
    }
}

/* MismatchTrie.hpp:300-343 scan_final_position_with_mismatch */
static void scan_last(const trie_t* t, int node, int refshift, int* index, int mismatches, int* cap) {
    int found = 0;
    for (int s = 0; s < 4; ++s) {
        if (s == refshift) continue;
        int cand = t->p[node + s];
        if (cand >= 0) {
            if (found) {
                if (cand != *index) {
                    if (t->dup == DUP_FIRST) {
                        if (*index > cand) *index = cand;
                    } else if (t->dup == DUP_LAST) {
                        if (*index < cand) *index = cand;
                    } else {
                        *index = AMBIGUOUS;
                        break;
                    }
                }
            } else {
                *index = cand;
                *cap = mismatches;
                found = 1;
            }
        } else if (cand == AMBIGUOUS) {
            *index = AMBIGUOUS;
            *cap = mismatches;
            break;
        }
    }
}

/* AnyMismatches::search, MismatchTrie.hpp:446-501 */
static hit_t any_dfs(const trie_t* t, const char* q, int pos, int node, int mm, int* cap) {
    int shift = base_code(q[pos]);
    int cur = shift >= 0 ? t->p[node + shift] : MISSING;
    hit_t out;
    memset(&out, 0, sizeof out);
    if (pos + 1 == t->length) {
        if (cur >= 0 || cur == AMBIGUOUS) {
            *cap = mm;
            out.index = cur;
            out.total = mm;
            return out;
        }
        out.index = MISSING;
        out.total = mm + 1;
        if (mm + 1 <= *cap) scan_last(t, node, shift, &out.index, mm + 1, cap);
        return out;
    }
    out.index = MISSING;
    out.total = *cap + 1;
    if (cur >= 0) out = any_dfs(t, q, pos + 1, cur, mm, cap);
    ++mm;
    if (mm <= *cap) {
        for (int s = 0; s < 4; ++s) {
            if (s == shift) continue;
            int alt = t->p[node + s];
            if (alt < 0) continue;
            if (mm <= *cap) {
                hit_t chosen = any_dfs(t, q, pos + 1, alt, mm, cap);
                merge_chosen(t, &out, &chosen);
            }
        }
    }
    return out;
}

static hit_t trie_search_any(const trie_t* t, const char* q, int cap) {
    hit_t out;
    memset(&out, 0, sizeof out);
    if (t->length == 0) { /* degenerate; not reachable from the R API */
        out.index = MISSING;
        return out;
    }
    return any_dfs(t, q, 0, 0, 0, &cap);
}

/* SegmentedMismatches::search, MismatchTrie.hpp:577-660 (incl. the phantom result
 * of :608-617 when a per-segment cap fails at the last position -- "Quirk A"). */
static hit_t seg_dfs(const trie_t* t, const char* q, int pos, int seg, int node, hit_t state, const int* segcap, int* cap) {
    int shift = base_code(q[pos]);
    int cur = shift >= 0 ? t->p[node + shift] : MISSING;
    if (pos + 1 == t->length) {
        if (cur >= 0 || cur == AMBIGUOUS) {
            *cap = state.total;
            state.index = cur;
            return state;
        }
        state.index = MISSING;
        ++state.total;
        ++state.per_segment[seg];
        if (state.total <= *cap && state.per_segment[seg] <= segcap[seg]) {
            scan_last(t, node, shift, &state.index, state.total, cap);
        }
        return state;
    }
    int next_seg = seg;
    if (pos + 1 == t->boundaries[seg]) ++next_seg;
    hit_t best;
    memset(&best, 0, sizeof best);
    best.index = MISSING;
    best.total = *cap + 1;
    if (cur >= 0) best = seg_dfs(t, q, pos + 1, next_seg, cur, state, segcap, cap);
    ++state.total;
    ++state.per_segment[seg];
    if (state.total <= *cap && state.per_segment[seg] <= segcap[seg]) {
        for (int s = 0; s < 4; ++s) {
            if (s == shift) continue;
            int alt = t->p[node + s];
            if (alt < 0) continue;
            if (state.total <= *cap) {
                hit_t chosen = seg_dfs(t, q, pos + 1, next_seg, alt, state, segcap, cap);
                merge_chosen(t, &best, &chosen);
            }
        }
    }
    return best;
}

static hit_t trie_search_segmented(const trie_t* t, const char* q, const int* segcap) {
    int cap = 0;
    for (int s = 0; s < t->nseg; ++s) cap += segcap[s];
    hit_t start;
    memset(&start, 0, sizeof start);
    return seg_dfs(t, q, 0, 0, 0, start, segcap, &cap);
}

/* fill_library, BarcodeSearch.hpp:23-60.  The `exact` map is an accelerator: a
 * hit there equals the trie's distance-0 answer, so only the trie is kept. */
static int trie_fill(trie_t* t, const char* const* pool, int npool, int reverse) {
    int len = t->length;
    char* tmp = malloc((size_t)len + 1);
    for (int i = 0; i < npool; ++i) {
        const char* s = pool[i];
        if (reverse) {
            for (int j = 0; j < len; ++j) {
                char c = complement(s[len - j - 1], 1, 1);
                if (!c) {
                    free(tmp);
                    return fail("cannot complement unknown base '%c'", s[len - j - 1]);
                }
                tmp[j] = c;
            }
            tmp[len] = 0;
            s = tmp;
        }
        if (trie_add(t, s)) {
            free(tmp);
            return 1;
        }
    }
    free(tmp);
    return 0;
}

/* format_pointers, screenCounter src/utils.cpp:5-23 */
static int pool_length(const char* const* pool, int npool, int* len) {
    size_t size = 0;
    for (int i = 0; i < npool; ++i) {
        size_t cur = strlen(pool[i]);
        if (i == 0) {
            size = cur;
        } else if (cur != size) {
            return fail("variable regions should all have the same length (%zu)", size);
        }
    }
    *len = (int)size;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* SimpleSingleMatch: SimpleSingleMatch.hpp                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
    tmpl_t tmpl;
    trie_t flib, rlib;
    int have_f, have_r;
    int mm;
    int npool;
} single_t;

typedef struct {
    int found, index, position, reverse, mismatches, variable_mismatches;
} single_hit_t;

static void single_free(single_t* m) {
    if (m->have_f) trie_free(&m->flib);
    if (m->have_r) trie_free(&m->rlib);
    m->have_f = m->have_r = 0;
}

static int single_init(single_t* m, const char* tmpl, int strand, const char* const* pool, int npool, int mm, int dup) {
    memset(m, 0, sizeof *m);
    int plen;
    if (pool_length(pool, npool, &plen)) return 1; /* format_pointers runs before the handler is built */
    if (tmpl_init(&m->tmpl, tmpl, strand)) return 1;
    if (m->tmpl.nreg != 1) return fail("expected one variable region in the constant template"); /* SimpleSingleMatch.hpp:75-77 */
    int vlen = m->tmpl.fend[0] - m->tmpl.fstart[0];
    if (vlen != plen) { /* :80-83 */
        return fail("length of barcode_pool sequences (%d) should be the same as the barcode_pool region (%d)", plen, vlen);
    }
    m->mm = mm;
    m->npool = npool;
    if (m->tmpl.do_fwd) {
        trie_init(&m->flib, plen, dup);
        m->have_f = 1;
        if (trie_fill(&m->flib, pool, npool, 0)) {
            single_free(m);
            return 1;
        }
    }
    if (m->tmpl.do_rev) {
        trie_init(&m->rlib, plen, dup);
        m->have_r = 1;
        if (trie_fill(&m->rlib, pool, npool, 1)) {
            single_free(m);
            return 1;
        }
    }
    return 0;
}

/* search_first (:200-245) when first != 0, search_best (:259-306) otherwise. */
static single_hit_t single_search(const single_t* m, const char* read, size_t len, int first) {
    single_hit_t out = { 0, -1, 0, 0, 0, 0 };
    const tmpl_t* t = &m->tmpl;
    int best = m->mm + 1;
    if ((size_t)t->T > len) return out; /* ScanTemplate.hpp:153,168-170 */
    for (size_t p = 0; p + t->T <= len; ++p) {
        for (int rev = 0; rev < 2; ++rev) {
            if (rev ? !t->do_rev : !t->do_fwd) continue;
            int c = const_mm(read + p, rev ? t->rev : t->fwd, t->T);
            if (c > m->mm) continue;
            int start = rev ? t->rstart[0] : t->fstart[0];
            hit_t h = trie_search_any(rev ? &m->rlib : &m->flib, read + p + start, m->mm - c);
            if (h.index < 0) continue;
            int total = c + h.total;
            if (first) {
                if (total > m->mm) continue;
                out.found = 1;
                out.index = h.index;
                out.position = (int)p;
                out.reverse = rev;
                out.mismatches = total;
                out.variable_mismatches = h.total;
                return out;
            }
            if (total == best) {
                if (out.index != h.index) {
                    out.found = 0;
                    out.index = -1;
                }
            } else if (total < best) {
                best = total;
                out.found = 1;
                out.index = h.index;
                out.position = (int)p;
                out.reverse = rev;
                out.mismatches = total;
                out.variable_mismatches = h.total;
            }
        }
    }
    return out;
}

/* ------------------------------------------------------------------------- */
/* result tables                                                              */
/* ------------------------------------------------------------------------- */

struct kport_table {
    int width;
    size_t n;
    int* keys;
    char* strings;
    int* freq;
};

size_t kport_table_size(const kport_table* t) { return t->n; }
int kport_table_width(const kport_table* t) { return t->width; }
void kport_table_copy(const kport_table* t, int* keys, char* strings, int* freq) {
    if (keys && t->keys) memcpy(keys, t->keys, t->n * t->width * sizeof(int));
    if (strings && t->strings) memcpy(strings, t->strings, t->n * t->width);
    if (freq && t->freq) memcpy(freq, t->freq, t->n * sizeof(int));
}
void kport_table_free(kport_table* t) {
    if (!t) return;
    free(t->keys);
    free(t->strings);
    free(t->freq);
    free(t);
}

typedef struct {
    int* v; /* pairs */
    size_t n, cap;
} pairs_t;

static void pairs_push(pairs_t* c, int a, int b) {
    if (c->n == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 1024;
        c->v = realloc(c->v, c->cap * 2 * sizeof(int));
    }
    c->v[2 * c->n] = a;
    c->v[2 * c->n + 1] = b;
    ++c->n;
}

static int cmp_pair(const void* a, const void* b) {
    const int* x = a;
    const int* y = b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}

/* sort_combinations (utils.hpp:173-198) followed by count_combinations
 * (screenCounter src/utils.h:14-45): ascending (first, second), run-length encoded. */
static kport_table* pairs_to_table(pairs_t* c) {
    qsort(c->v, c->n, 2 * sizeof(int), cmp_pair);
    kport_table* t = calloc(1, sizeof *t);
    t->width = 2;
    t->keys = malloc((c->n ? c->n : 1) * 2 * sizeof(int));
    t->freq = malloc((c->n ? c->n : 1) * sizeof(int));
    for (size_t i = 0; i < c->n; ++i) {
        if (i && cmp_pair(c->v + 2 * i, c->v + 2 * (i - 1)) == 0) {
            ++t->freq[t->n - 1];
        } else {
            t->keys[2 * t->n] = c->v[2 * i];
            t->keys[2 * t->n + 1] = c->v[2 * i + 1];
            t->freq[t->n] = 1;
            ++t->n;
        }
    }
    return t;
}

/* ------------------------------------------------------------------------- */
/* entry points                                                               */
/* ------------------------------------------------------------------------- */

int kport_count_reads(const char* path, const char* data, size_t size, long long* nreads, long long* nbases) {
    reads_t r;
    if (parse_fastq(path, data, size, &r)) return 1;
    *nreads = (long long)r.n;
    *nbases = (long long)r.nbuf;
    reads_free(&r);
    return 0;
}

int kport_parse(const char* path, const char* data, size_t size, char* bases, long long* offsets) {
    reads_t r;
    if (parse_fastq(path, data, size, &r)) return 1;
    memcpy(bases, r.buf, r.nbuf);
    for (size_t i = 0; i <= r.n; ++i) offsets[i] = (long long)r.off[i];
    reads_free(&r);
    return 0;
}

/* SingleBarcodeSingleEnd::process, handlers/SingleBarcodeSingleEnd.hpp:93-104 */
static int run_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                      const char* const* pool, int npool, int mismatches, int use_first,
                      int* counts, int* total, int* index, int* info, long long capacity, long long* nreads) {
    /* the handler is constructed before the first record is read (src/count_single_barcodes.cpp:18-19) */
    single_t m;
    if (single_init(&m, tmpl, strand, pool, npool, mismatches, DUP_ERROR)) return 1;
    reads_t r;
    if (parse_fastq(path, data, size, &r)) {
        single_free(&m);
        return 1;
    }
    if (nreads) *nreads = (long long)r.n;
    if (index && (long long)r.n > capacity) {
        single_free(&m);
        reads_free(&r);
        return fail("trace capacity too small");
    }
    if (counts) memset(counts, 0, (size_t)npool * sizeof(int));
    for (size_t i = 0; i < r.n; ++i) {
        single_hit_t h = single_search(&m, r.buf + r.off[i], r.off[i + 1] - r.off[i], use_first);
        if (h.found && counts) ++counts[h.index];
        if (index) index[i] = h.found ? h.index : -1;
        if (info) {
            info[4 * i + 0] = h.found ? h.position : -1;
            info[4 * i + 1] = h.found ? h.reverse : 0;
            info[4 * i + 2] = h.found ? h.mismatches : -1;
            info[4 * i + 3] = h.found ? h.variable_mismatches : -1;
        }
    }
    if (total) *total = (int)r.n;
    single_free(&m);
    reads_free(&r);
    return 0;
}

int kport_count_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       const char* const* pool, int npool, int mismatches, int use_first, int nthreads,
                       int* counts, int* total) {
    (void)nthreads;
    return run_single(path, data, size, tmpl, strand, pool, npool, mismatches, use_first, counts, total, NULL, NULL, 0, NULL);
}

int kport_trace_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       const char* const* pool, int npool, int mismatches, int use_first,
                       int* index, int* info, long long capacity, long long* nreads) {
    return run_single(path, data, size, tmpl, strand, pool, npool, mismatches, use_first, NULL, NULL, index, info, capacity, nreads);
}

/* --- random barcodes: handlers/RandomBarcodeSingleEnd.hpp:93-181 ----------- */

typedef struct {
    char* keys; /* width bytes each */
    int* counts;
    size_t n, cap;
    int width;
} strmap_t;

static int g_sort_width;
static int cmp_fixed(const void* a, const void* b) { return memcmp(a, b, (size_t)g_sort_width); }

static void strmap_add(strmap_t* m, const char* key) {
    /* append; duplicates are merged after sorting (the reference's unordered_map
     * has no observable order: R sorts by sequence, R/countRandomBarcodes.R:73-74) */
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 1024;
        m->keys = realloc(m->keys, m->cap * (size_t)(m->width ? m->width : 1));
    }
    memcpy(m->keys + m->n * (size_t)m->width, key, (size_t)m->width);
    ++m->n;
}

int kport_count_random(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       int mismatches, int use_first, int nthreads, kport_table** table, int* total) {
    (void)nthreads;
    tmpl_t t;
    if (tmpl_init(&t, tmpl, strand)) return 1;
    /* the handler dereferences variable_regions()[0] unconditionally (:93-96, :212-214) */
    if (t.nreg < 1) return fail("expected at least one variable region in the constant template");
    reads_t r;
    if (parse_fastq(path, data, size, &r)) return 1;
    int vstart = t.fstart[0], vlen = t.fend[0] - t.fstart[0];
    strmap_t map;
    memset(&map, 0, sizeof map);
    map.width = vlen;
    char* key = malloc((size_t)vlen + 1);
    int status = 0;

    for (size_t i = 0; i < r.n && !status; ++i) {
        const char* read = r.buf + r.off[i];
        size_t len = r.off[i + 1] - r.off[i];
        int best = mismatches + 1, best_fwd = 1, tied = 0, have = 0;
        size_t best_pos = 0;
        if ((size_t)t.T <= len) {
            for (size_t p = 0; p + t.T <= len && !(use_first && have); ++p) {
                for (int rev = 0; rev < 2; ++rev) {
                    if (rev ? !t.do_rev : !t.do_fwd) continue;
                    int c = const_mm(read + p, rev ? t.rev : t.fwd, t.T);
                    if (c > mismatches) continue;
                    if (use_first) { /* :126-137 */
                        have = 1;
                        best = c;
                        best_pos = p;
                        best_fwd = !rev;
                        break;
                    }
                    if (c < best) { /* :139-170 */
                        best = c;
                        best_pos = p;
                        best_fwd = !rev;
                        tied = 0;
                    } else if (c == best) {
                        tied = 1;
                    }
                }
            }
        }
        int counted = use_first ? have : (!tied && best <= mismatches);
        if (counted) {
            const char* start = read + best_pos + vstart; /* forward coordinates on BOTH strands (:106-108, "Quirk B") */
            if (best_fwd) {
                memcpy(key, start, (size_t)vlen);
            } else {
                for (int j = 0; j < vlen; ++j) {
                    char c = complement(start[vlen - j - 1], 1, 0);
                    if (!c) {
                        status = fail("cannot complement unknown base '%c'", start[vlen - j - 1]);
                        break;
                    }
                    key[j] = c;
                }
            }
            if (!status) strmap_add(&map, key);
        }
    }
    free(key);
    if (status) {
        free(map.keys);
        reads_free(&r);
        return 1;
    }

    g_sort_width = vlen;
    if (vlen > 0) qsort(map.keys, map.n, (size_t)vlen, cmp_fixed);
    kport_table* out = calloc(1, sizeof *out);
    out->width = vlen;
    out->strings = malloc((map.n ? map.n : 1) * (size_t)(vlen ? vlen : 1));
    out->freq = malloc((map.n ? map.n : 1) * sizeof(int));
    for (size_t i = 0; i < map.n; ++i) {
        const char* k = map.keys + i * (size_t)vlen;
        if (out->n && memcmp(out->strings + (out->n - 1) * (size_t)vlen, k, (size_t)vlen) == 0) {
            ++out->freq[out->n - 1];
        } else {
            memcpy(out->strings + out->n * (size_t)vlen, k, (size_t)vlen);
            out->freq[out->n] = 1;
            ++out->n;
        }
    }
    free(map.keys);
    *table = out;
    *total = (int)r.n;
    reads_free(&r);
    return 0;
}

/* --- combinatorial single-end: handlers/CombinatorialBarcodesSingleEnd.hpp ---- */

#define MAX_V 8

typedef struct {
    tmpl_t tmpl;
    int V;
    trie_t flib[MAX_V], rlib[MAX_V];
    int have_f, have_r;
    int mm;
} combo_t;

static void combo_free(combo_t* c) {
    for (int i = 0; i < c->V; ++i) {
        if (c->have_f) trie_free(&c->flib[i]);
        if (c->have_r) trie_free(&c->rlib[i]);
    }
    c->have_f = c->have_r = 0;
}

static int combo_init(combo_t* c, const char* tmpl, int strand, const char* const* const* pools, const int* npools, int V, int mm, int dup) {
    memset(c, 0, sizeof *c);
    int plen[MAX_V];
    for (int i = 0; i < V; ++i) {
        if (pool_length(pools[i], npools[i], &plen[i])) return 1;
    }
    if (tmpl_init(&c->tmpl, tmpl, strand)) return 1;
    if (c->tmpl.nreg != V) return fail("expected %d variable regions in the constant template", V); /* :79-81 */
    for (int i = 0; i < V; ++i) { /* :85-92 */
        int rlen = c->tmpl.fend[i] - c->tmpl.fstart[i];
        if (rlen != plen[i]) {
            return fail("length of variable region %d (%d) should be the same as its sequences (%d)", i + 1, rlen, plen[i]);
        }
    }
    c->V = 0;
    c->mm = mm;
    c->have_f = c->tmpl.do_fwd;
    c->have_r = c->tmpl.do_rev;
    for (int i = 0; i < V; ++i) {
        if (c->have_f) trie_init(&c->flib[i], plen[i], dup);
        if (c->have_r) trie_init(&c->rlib[i], plen[V - i - 1], dup); /* reversed pool order, :111-116 */
        c->V = i + 1;
    }
    for (int i = 0; i < V; ++i) {
        if (c->have_f && trie_fill(&c->flib[i], pools[i], npools[i], 0)) goto bad;
    }
    for (int i = 0; i < V; ++i) {
        if (c->have_r && trie_fill(&c->rlib[i], pools[V - i - 1], npools[V - i - 1], 1)) goto bad;
    }
    return 0;
bad:
    combo_free(c);
    return 1;
}

/* find_match<reverse>, :149-186.  Returns total mismatches or -1. */
static int combo_match(const combo_t* c, const char* window, int rev, int obs, int* ids) {
    const tmpl_t* t = &c->tmpl;
    for (int r = 0; r < c->V; ++r) {
        int start = rev ? t->rstart[r] : t->fstart[r];
        hit_t h = trie_search_any(rev ? &c->rlib[r] : &c->flib[r], window + start, c->mm - obs);
        if (h.index < 0) return -1;
        obs += h.total;
        if (obs > c->mm) return -1;
        ids[rev ? c->V - r - 1 : r] = h.index;
    }
    return obs;
}

/* process_first :197-217 / process_best :219-258 */
static int combo_search(const combo_t* c, const char* read, size_t len, int first, int* out_ids) {
    const tmpl_t* t = &c->tmpl;
    int found = 0, best = c->mm + 1;
    int ids[MAX_V], best_ids[MAX_V];
    if ((size_t)t->T > len) return 0;
    for (size_t p = 0; p + t->T <= len; ++p) {
        for (int rev = 0; rev < 2; ++rev) {
            if (rev ? !t->do_rev : !t->do_fwd) continue;
            int cm = const_mm(read + p, rev ? t->rev : t->fwd, t->T);
            if (cm > c->mm) continue;
            int tot = combo_match(c, read + p, rev, cm, ids);
            if (tot < 0) continue;
            if (first) {
                memcpy(out_ids, ids, sizeof(int) * c->V);
                return 1;
            }
            if (tot == best) {
                if (memcmp(best_ids, ids, sizeof(int) * c->V) != 0) found = 0;
            } else if (tot < best) {
                found = 1;
                best = tot;
                memcpy(best_ids, ids, sizeof(int) * c->V);
            }
        }
    }
    if (found) memcpy(out_ids, best_ids, sizeof(int) * c->V);
    return found;
}

static int run_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                            const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                            int mismatches, int use_first, kport_table** table, int* total,
                            int* combo, long long capacity, long long* nreads) {
    const char* const* pools[2] = { pool1, pool2 };
    int npools[2] = { npool1, npool2 };
    combo_t c;
    if (combo_init(&c, tmpl, strand, pools, npools, 2, mismatches, DUP_ERROR)) return 1;
    reads_t r;
    if (parse_fastq(path, data, size, &r)) {
        combo_free(&c);
        return 1;
    }
    if (nreads) *nreads = (long long)r.n;
    if (combo && (long long)r.n > capacity) {
        combo_free(&c);
        reads_free(&r);
        return fail("trace capacity too small");
    }
    pairs_t collected = { 0, 0, 0 };
    for (size_t i = 0; i < r.n; ++i) {
        int ids[MAX_V];
        int found = combo_search(&c, r.buf + r.off[i], r.off[i + 1] - r.off[i], use_first, ids);
        if (found) pairs_push(&collected, ids[0], ids[1]);
        if (combo) {
            combo[2 * i] = found ? ids[0] : -1;
            combo[2 * i + 1] = found ? ids[1] : -1;
        }
    }
    if (table) *table = pairs_to_table(&collected);
    if (total) *total = (int)r.n;
    free(collected.v);
    combo_free(&c);
    reads_free(&r);
    return 0;
}

int kport_count_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                             const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                             int mismatches, int use_first, int nthreads, kport_table** table, int* total) {
    (void)nthreads;
    return run_combo_single(path, data, size, tmpl, strand, pool1, npool1, pool2, npool2, mismatches, use_first, table, total, NULL, 0, NULL);
}

int kport_trace_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                             const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                             int mismatches, int use_first, int* combo, long long capacity, long long* nreads) {
    return run_combo_single(path, data, size, tmpl, strand, pool1, npool1, pool2, npool2, mismatches, use_first, NULL, NULL, combo, capacity, nreads);
}

/* --- dual barcodes, single-end: handlers/DualBarcodesSingleEnd.hpp ------------ */

typedef struct {
    tmpl_t tmpl;
    trie_t flib, rlib;
    int have_f, have_r;
    int mm;
    int klen;
    int nchoices;
} dualse_t;

static void dualse_free(dualse_t* d) {
    if (d->have_f) trie_free(&d->flib);
    if (d->have_r) trie_free(&d->rlib);
    d->have_f = d->have_r = 0;
}

static int dualse_init(dualse_t* d, const char* tmpl, const char* const* pools_flat, int npools, int nchoices, int strand, int mm) {
    memset(d, 0, sizeof *d);
    int plen[MAX_REGIONS];
    if (npools > MAX_REGIONS) return fail("too many pools");
    for (int p = 0; p < npools; ++p) {
        if (pool_length(pools_flat + (size_t)p * nchoices, nchoices, &plen[p])) return 1;
    }
    if (tmpl_init(&d->tmpl, tmpl, strand)) return 1;
    if (npools != d->tmpl.nreg) return fail("length of 'barcode_pools' should equal the number of variable regions"); /* :75-77 */
    int klen = 0;
    for (int p = 0; p < npools; ++p) { /* :79-86 */
        int rlen = d->tmpl.fend[p] - d->tmpl.fstart[p];
        if (rlen != plen[p]) {
            return fail("length of variable region %d (%d) should be the same as its sequences (%d)", p + 1, rlen, plen[p]);
        }
        klen += plen[p];
    }
    /* concatenate the rows, :99-108 */
    char** rows = malloc((size_t)(nchoices ? nchoices : 1) * sizeof(char*));
    for (int c = 0; c < nchoices; ++c) {
        rows[c] = malloc((size_t)klen + 1);
        int o = 0;
        for (int p = 0; p < npools; ++p) {
            memcpy(rows[c] + o, pools_flat[(size_t)p * nchoices + c], (size_t)plen[p]);
            o += plen[p];
        }
        rows[c][klen] = 0;
    }
    d->mm = mm;
    d->klen = klen;
    d->nchoices = nchoices;
    int status = 0;
    if (d->tmpl.do_fwd) {
        trie_init(&d->flib, klen, DUP_ERROR);
        d->have_f = 1;
        status = trie_fill(&d->flib, (const char* const*)rows, nchoices, 0);
    }
    if (!status && d->tmpl.do_rev) {
        trie_init(&d->rlib, klen, DUP_ERROR);
        d->have_r = 1;
        status = trie_fill(&d->rlib, (const char* const*)rows, nchoices, 1); /* RC of the whole row, :117-120 */
    }
    for (int c = 0; c < nchoices; ++c) free(rows[c]);
    free(rows);
    if (status) dualse_free(d);
    return status;
}

/* process_first :170-190 / process_best :192-231; returns index or -1 */
static int dualse_search(const dualse_t* d, const char* read, size_t len, int first, char* buffer) {
    const tmpl_t* t = &d->tmpl;
    int found = 0, best = d->mm + 1, best_id = -1;
    if ((size_t)t->T > len) return -1;
    for (size_t p = 0; p + t->T <= len; ++p) {
        for (int rev = 0; rev < 2; ++rev) {
            if (rev ? !t->do_rev : !t->do_fwd) continue;
            int cm = const_mm(read + p, rev ? t->rev : t->fwd, t->T);
            if (cm > d->mm) continue;
            int o = 0;
            for (int r = 0; r < t->nreg; ++r) { /* find_match, :144-160 */
                int s = rev ? t->rstart[r] : t->fstart[r], e = rev ? t->rend[r] : t->fend[r];
                memcpy(buffer + o, read + p + s, (size_t)(e - s));
                o += e - s;
            }
            hit_t h = trie_search_any(rev ? &d->rlib : &d->flib, buffer, d->mm - cm);
            if (h.index < 0) continue;
            if (first) return h.index;
            int tot = cm + h.total;
            if (tot == best) {
                if (best_id != h.index) found = 0;
            } else if (tot < best) {
                found = 1;
                best = tot;
                best_id = h.index;
            }
        }
    }
    return found ? best_id : -1;
}

static int run_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                               const char* const* pools_flat, int npools, int nchoices, int strand,
                               int mismatches, int use_first, int diagnostics,
                               int* counts, int* total, kport_table** table,
                               int* index, long long capacity, long long* nreads) {
    dualse_t d;
    if (dualse_init(&d, tmpl, pools_flat, npools, nchoices, strand, mismatches)) return 1;
    combo_t c;
    int have_combo = 0;
    if (diagnostics) { /* handlers/DualBarcodesSingleEndWithDiagnostics.hpp:44-58: V = 2, DuplicateAction::FIRST */
        const char* const* pools[2] = { pools_flat, pools_flat + nchoices };
        int np[2] = { nchoices, nchoices };
        if (npools != 2 || combo_init(&c, tmpl, strand, pools, np, 2, mismatches, DUP_FIRST)) {
            if (npools != 2) fail("expected 2 variable regions in the constant template");
            dualse_free(&d);
            return 1;
        }
        have_combo = 1;
    }
    reads_t r;
    if (parse_fastq(path, data, size, &r)) {
        if (have_combo) combo_free(&c);
        dualse_free(&d);
        return 1;
    }
    if (nreads) *nreads = (long long)r.n;
    if (index && (long long)r.n > capacity) {
        dualse_free(&d);
        reads_free(&r);
        return fail("trace capacity too small");
    }
    if (counts) memset(counts, 0, (size_t)nchoices * sizeof(int));
    char* buffer = malloc((size_t)d.klen + 1);
    pairs_t collected = { 0, 0, 0 };
    for (size_t i = 0; i < r.n; ++i) {
        const char* read = r.buf + r.off[i];
        size_t len = r.off[i + 1] - r.off[i];
        int id = dualse_search(&d, read, len, use_first, buffer);
        if (id >= 0 && counts) ++counts[id];
        if (index) index[i] = id;
        if (id < 0 && have_combo) { /* :99-104 */
            int ids[MAX_V];
            if (combo_search(&c, read, len, use_first, ids)) pairs_push(&collected, ids[0], ids[1]);
        }
    }
    free(buffer);
    if (table && diagnostics) *table = pairs_to_table(&collected);
    free(collected.v);
    if (total) *total = (int)r.n;
    if (have_combo) combo_free(&c);
    dualse_free(&d);
    reads_free(&r);
    return 0;
}

int kport_count_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                                const char* const* pools_flat, int npools, int nchoices, int strand,
                                int mismatches, int use_first, int diagnostics, int nthreads,
                                int* counts, int* total, kport_table** table) {
    (void)nthreads;
    return run_dual_single_end(path, data, size, tmpl, pools_flat, npools, nchoices, strand, mismatches, use_first, diagnostics,
                               counts, total, table, NULL, 0, NULL);
}

int kport_trace_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                                const char* const* pools_flat, int npools, int nchoices, int strand,
                                int mismatches, int use_first, int* index, long long capacity, long long* nreads) {
    return run_dual_single_end(path, data, size, tmpl, pools_flat, npools, nchoices, strand, mismatches, use_first, 0,
                               NULL, NULL, NULL, index, capacity, nreads);
}

/* --- dual barcodes, paired-end: handlers/DualBarcodesPairedEnd.hpp ------------ */

typedef struct {
    tmpl_t t1, t2;
    int rev1, rev2;
    int mm1, mm2;
    int len1, len2;
    trie_t lib;
    int npairs;
} dualpe_t;

static int dualpe_init(dualpe_t* d, const char* tmpl1, int reverse1, int mm1, const char* const* pool1, int npool1,
                       const char* tmpl2, int reverse2, int mm2, const char* const* pool2, int npool2) {
    memset(d, 0, sizeof *d);
    int plen1, plen2;
    if (pool_length(pool1, npool1, &plen1)) return 1;
    if (pool_length(pool2, npool2, &plen2)) return 1;
    {
        size_t l1 = strlen(tmpl1), l2 = strlen(tmpl2);
        if ((l1 > l2 ? l1 : l2) > MAX_TEMPLATE) return fail("lacking compile-time support for constant regions longer than 256 bp");
    }
    if (tmpl_init(&d->t1, tmpl1, reverse1 ? 1 : 0)) return 1;
    if (tmpl_init(&d->t2, tmpl2, reverse2 ? 1 : 0)) return 1;
    if (npool1 != npool2) return fail("both barcode pools should be of the same length"); /* :107-109 */
    if (d->t1.nreg != 1) return fail("expected one variable region in the first constant template");
    d->len1 = d->t1.fend[0] - d->t1.fstart[0];
    if (d->len1 != plen1) return fail("length of variable sequences (%d) should be the same as the variable region (%d)", plen1, d->len1);
    if (d->t2.nreg != 1) return fail("expected one variable region in the second constant template");
    d->len2 = d->t2.fend[0] - d->t2.fstart[0];
    if (d->len2 != plen2) return fail("length of variable sequences (%d) should be the same as the variable region (%d)", plen2, d->len2);
    d->rev1 = reverse1 != 0;
    d->rev2 = reverse2 != 0;
    d->mm1 = mm1;
    d->mm2 = mm2;
    d->npairs = npool1;

    /* combined strings, :139-164: each half reverse-complemented on its own when its strand is reverse */
    int klen = d->len1 + d->len2;
    trie_init(&d->lib, klen, DUP_ERROR);
    d->lib.nseg = 2;
    d->lib.boundaries[0] = d->len1;
    d->lib.boundaries[1] = klen;
    char* row = malloc((size_t)klen + 1);
    int status = 0;
    for (int i = 0; i < npool1 && !status; ++i) {
        for (int j = 0; j < d->len1; ++j) {
            char b = d->rev1 ? complement(pool1[i][d->len1 - j - 1], 1, 1) : pool1[i][j];
            if (!b) status = fail("cannot complement unknown base '%c'", pool1[i][d->len1 - j - 1]);
            row[j] = b;
        }
        for (int j = 0; j < d->len2 && !status; ++j) {
            char b = d->rev2 ? complement(pool2[i][d->len2 - j - 1], 1, 1) : pool2[i][j];
            if (!b) status = fail("cannot complement unknown base '%c'", pool2[i][d->len2 - j - 1]);
            row[d->len1 + j] = b;
        }
        row[klen] = 0;
        if (!status) status = trie_add(&d->lib, row);
    }
    free(row);
    if (status) trie_free(&d->lib);
    return status;
}

typedef struct {
    int pos;
    int mm;
} chit_t;

/* all constant-region hits on one read for one configured strand: inner_process, :229-256 */
static size_t dualpe_hits(const tmpl_t* t, int rev, int mm, const char* read, size_t len, chit_t** out, size_t* cap) {
    size_t n = 0;
    if ((size_t)t->T > len) return 0;
    for (size_t p = 0; p + t->T <= len; ++p) {
        int c = const_mm(read + p, rev ? t->rev : t->fwd, t->T);
        if (c <= mm) {
            if (n == *cap) {
                *cap = *cap ? *cap * 2 : 64;
                *out = realloc(*out, *cap * sizeof(chit_t));
            }
            (*out)[n].pos = (int)p + (rev ? t->rstart[0] : t->fstart[0]);
            (*out)[n].mm = c;
            ++n;
        }
    }
    return n;
}

/* Optional emulation of the reference's result cache for the segmented search
 * (matcher_in_the_rye, BarcodeSearch.hpp:62-93; Methods::update, :435-448).  The cache is
 * NOT semantically transparent there ("Quirk C"), so the port can state
 * three semantics:
 *   CACHE_NONE     every search is fresh -- the cache-free definition the CUDA path implements;
 *   CACHE_PER_PAIR cache emptied before each read pair -- what ref_harness.cpp's
 *                  kref_trace_dual(fresh_state=1) measures on the real reference;
 *   CACHE_FILE     one cache for the whole file -- what an R user gets with num.threads = 1. */
enum { CACHE_FILE = 0, CACHE_PER_PAIR = 1, CACHE_NONE = 2 };

typedef struct {
    char* key;
    hit_t res;
} centry_t;

typedef struct {
    centry_t* slots;
    size_t cap, n;
    int klen;
} cache_t;

static size_t cache_hash(const char* k, int len) {
    size_t h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) h = (h ^ (unsigned char)k[i]) * 1099511628211ull;
    return h;
}

static void cache_clear(cache_t* c) {
    for (size_t i = 0; i < c->cap; ++i) free(c->slots[i].key);
    free(c->slots);
    c->slots = NULL;
    c->cap = c->n = 0;
}

static centry_t* cache_find(cache_t* c, const char* k) {
    if (!c->cap) return NULL;
    size_t i = cache_hash(k, c->klen) & (c->cap - 1);
    while (c->slots[i].key) {
        if (memcmp(c->slots[i].key, k, (size_t)c->klen) == 0) return &c->slots[i];
        i = (i + 1) & (c->cap - 1);
    }
    return NULL;
}

static void cache_put(cache_t* c, const char* k, const hit_t* res) {
    if ((c->n + 1) * 2 > c->cap) {
        size_t ncap = c->cap ? c->cap * 2 : 64;
        centry_t* old = c->slots;
        size_t ocap = c->cap;
        c->slots = calloc(ncap, sizeof(centry_t));
        c->cap = ncap;
        for (size_t i = 0; i < ocap; ++i) {
            if (!old[i].key) continue;
            size_t j = cache_hash(old[i].key, c->klen) & (ncap - 1);
            while (c->slots[j].key) j = (j + 1) & (ncap - 1);
            c->slots[j] = old[i];
        }
        free(old);
    }
    size_t i = cache_hash(k, c->klen) & (c->cap - 1);
    while (c->slots[i].key) i = (i + 1) & (c->cap - 1);
    c->slots[i].key = malloc((size_t)c->klen);
    memcpy(c->slots[i].key, k, (size_t)c->klen);
    c->slots[i].res = *res;
    ++c->n;
}

typedef struct {
    chit_t *h1, *h2;
    size_t c1, c2;
    char* key;
    int cache_mode;
    cache_t cache;
} dualpe_scratch_t;

/* SegmentedBarcodeSearch<2>::search, BarcodeSearch.hpp:478-487, with the cache when asked for. */
static hit_t dualpe_lookup(const dualpe_t* d, dualpe_scratch_t* s, const char* key, const int* caps) {
    if (s->cache_mode == CACHE_NONE) return trie_search_segmented(&d->lib, key, caps);
    centry_t* e = cache_find(&s->cache, key);
    if (e) { /* Methods::update(state, cached, caps), :443-448 */
        hit_t out = e->res;
        if (out.per_segment[0] > caps[0] || out.per_segment[1] > caps[1]) out.index = -1;
        return out;
    }
    hit_t h = trie_search_segmented(&d->lib, key, caps);
    if (h.index >= 0 || (caps[0] == d->mm1 && caps[1] == d->mm2)) cache_put(&s->cache, key, &h); /* :81-83 */
    return h;
}

/* process_first :258-308 (first != 0) -> (index or -1, 0);  process_best :310-347 -> (chosen, best_mismatches) */
static void dualpe_search(const dualpe_t* d, dualpe_scratch_t* s, const char* ra, size_t la, const char* rb, size_t lb,
                          int first, int* out_index, int* out_score) {
    size_t n1 = dualpe_hits(&d->t1, d->rev1, d->mm1, ra, la, &s->h1, &s->c1);
    size_t n2 = dualpe_hits(&d->t2, d->rev2, d->mm2, rb, lb, &s->h2, &s->c2);
    int chosen = -1, best = d->mm1 + d->mm2 + 1;
    /* process_first enumerates (hit1, hit2) in lexicographic order and stops at the first valid
     * pair (the lazy discovery of hit2 in :283-305 only changes when work is done, not the order);
     * process_best visits every pair when read 2 has at least one hit. */
    if (n2 > 0) {
        for (size_t i = 0; i < n1; ++i) {
            for (size_t j = 0; j < n2; ++j) {
                memcpy(s->key, ra + s->h1[i].pos, (size_t)d->len1);
                memcpy(s->key + d->len1, rb + s->h2[j].pos, (size_t)d->len2);
                int caps[2] = { d->mm1 - s->h1[i].mm, d->mm2 - s->h2[j].mm };
                hit_t h = dualpe_lookup(d, s, s->key, caps);
                if (h.index < 0) continue;
                if (first) {
                    *out_index = h.index;
                    *out_score = 0;
                    return;
                }
                int cur = h.total + s->h1[i].mm + s->h2[j].mm;
                if (cur < best) {
                    chosen = h.index;
                    best = cur;
                } else if (cur == best && chosen != h.index) {
                    chosen = -1;
                }
            }
        }
    }
    *out_index = chosen;
    *out_score = best;
}

/* process, :353-381 */
static int dualpe_process(const dualpe_t* d, dualpe_scratch_t* s, const char* r1, size_t l1, const char* r2, size_t l2,
                          int randomized, int use_first) {
    int idx, score;
    if (use_first) {
        dualpe_search(d, s, r1, l1, r2, l2, 1, &idx, &score);
        if (idx < 0 && randomized) dualpe_search(d, s, r2, l2, r1, l1, 1, &idx, &score);
        return idx;
    }
    dualpe_search(d, s, r1, l1, r2, l2, 0, &idx, &score);
    if (randomized) {
        int idx2, score2;
        dualpe_search(d, s, r2, l2, r1, l1, 0, &idx2, &score2);
        if (idx < 0 || score > score2) {
            idx = idx2;
            score = score2;
        } else if (score == score2 && idx != idx2) {
            idx = -1;
        }
    }
    return idx;
}

/* --- combinatorial paired-end: handlers/CombinatorialBarcodesPairedEnd.hpp:167-242 ---- */

/* returns code: 0 none, 1 pair collected (ids set), 2 barcode1 only, 3 barcode2 only */
static int combope_process(const single_t* m1, const single_t* m2, const char* r1, size_t l1, const char* r2, size_t l2,
                           int randomized, int use_first, int* ids) {
    single_hit_t a = single_search(m1, r1, l1, use_first);
    single_hit_t b = single_search(m2, r2, l2, use_first);
    if (use_first) {
        if (a.found && b.found) {
            ids[0] = a.index;
            ids[1] = b.index;
            return 1;
        }
        if (randomized) {
            single_hit_t n1 = single_search(m1, r2, l2, 1);
            single_hit_t n2 = single_search(m2, r1, l1, 1);
            if (n1.found && n2.found) {
                ids[0] = n1.index;
                ids[1] = n2.index;
                return 1;
            }
            if (a.found || n1.found) return 2;
            if (b.found || n2.found) return 3;
            return 0;
        }
        if (a.found) return 2;
        if (b.found) return 3;
        return 0;
    }
    if (!randomized) {
        if (a.found && b.found) {
            ids[0] = a.index;
            ids[1] = b.index;
            return 1;
        }
        if (a.found) return 2;
        if (b.found) return 3;
        return 0;
    }
    single_hit_t n1 = single_search(m1, r2, l2, 0);
    single_hit_t n2 = single_search(m2, r1, l1, 0);
    if (a.found && b.found) {
        int mm = a.mismatches + b.mismatches;
        if (n1.found && n2.found) {
            int rmm = n1.mismatches + n2.mismatches;
            if (mm > rmm) {
                ids[0] = n1.index;
                ids[1] = n2.index;
                return 1;
            } else if (mm < rmm) {
                ids[0] = a.index;
                ids[1] = b.index;
                return 1;
            } else if (a.index == n1.index && b.index == n2.index) {
                ids[0] = a.index;
                ids[1] = b.index;
                return 1;
            }
            return 0;
        }
        ids[0] = a.index;
        ids[1] = b.index;
        return 1;
    }
    if (n1.found && n2.found) {
        ids[0] = n1.index;
        ids[1] = n2.index;
        return 1;
    }
    if (a.found || n1.found) return 2;
    if (b.found || n2.found) return 3;
    return 0;
}

static int load_pair(const char* path1, const char* data1, size_t size1, const char* path2, const char* data2, size_t size2,
                     reads_t* r1, reads_t* r2) {
    if (parse_fastq(path1, data1, size1, r1)) return 1;
    if (parse_fastq(path2, data2, size2, r2)) {
        reads_free(r1);
        return 1;
    }
    return 0;
}

static int run_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                    const char* const* pool1, int npool1,
                    const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                    const char* const* pool2, int npool2,
                    int randomized, int use_first, int diagnostics, int cache_mode,
                    int* counts, int* total, kport_table** table, int* b1only, int* b2only,
                    int* index, long long capacity, long long* npairs) {
    reads_t r1, r2;
    if (load_pair(path1, data1, size1, path2, data2, size2, &r1, &r2)) return 1;
    dualpe_t d;
    if (dualpe_init(&d, tmpl1, reverse1, mismatches1, pool1, npool1, tmpl2, reverse2, mismatches2, pool2, npool2)) {
        reads_free(&r1);
        reads_free(&r2);
        return 1;
    }
    single_t m1, m2;
    int have_combo = 0;
    int status = 0;
    if (diagnostics) { /* handlers/DualBarcodesPairedEndWithDiagnostics.hpp:53-72: DuplicateAction::FIRST */
        if (single_init(&m1, tmpl1, reverse1 ? 1 : 0, pool1, npool1, mismatches1, DUP_FIRST)) {
            status = 1;
        } else if (single_init(&m2, tmpl2, reverse2 ? 1 : 0, pool2, npool2, mismatches2, DUP_FIRST)) {
            single_free(&m1);
            status = 1;
        } else {
            have_combo = 1;
        }
    }
    if (!status && r1.n != r2.n) status = fail("different number of reads in paired FASTQ files"); /* process_data.hpp:284-285 */
    if (!status && index && (long long)r1.n > capacity) status = fail("trace capacity too small");
    if (!status) {
        if (npairs) *npairs = (long long)r1.n;
        if (counts) memset(counts, 0, (size_t)npool1 * sizeof(int));
        dualpe_scratch_t s;
        memset(&s, 0, sizeof s);
        s.key = malloc((size_t)(d.len1 + d.len2) + 1);
        s.cache_mode = cache_mode;
        s.cache.klen = d.len1 + d.len2;
        pairs_t collected = { 0, 0, 0 };
        int only1 = 0, only2 = 0;
        for (size_t i = 0; i < r1.n; ++i) {
            const char* a = r1.buf + r1.off[i];
            size_t la = r1.off[i + 1] - r1.off[i];
            const char* b = r2.buf + r2.off[i];
            size_t lb = r2.off[i + 1] - r2.off[i];
            if (cache_mode == CACHE_PER_PAIR) cache_clear(&s.cache);
            int id = dualpe_process(&d, &s, a, la, b, lb, randomized, use_first);
            if (id >= 0 && counts) ++counts[id];
            if (index) index[i] = id;
            if (id < 0 && have_combo) {
                int ids[2];
                int code = combope_process(&m1, &m2, a, la, b, lb, randomized, use_first, ids);
                if (code == 1) pairs_push(&collected, ids[0], ids[1]);
                if (code == 2) ++only1;
                if (code == 3) ++only2;
            }
        }
        if (diagnostics && table) *table = pairs_to_table(&collected);
        if (diagnostics && b1only) *b1only = only1;
        if (diagnostics && b2only) *b2only = only2;
        if (total) *total = (int)r1.n;
        free(collected.v);
        free(s.h1);
        free(s.h2);
        free(s.key);
        cache_clear(&s.cache);
    }
    if (have_combo) {
        single_free(&m1);
        single_free(&m2);
    }
    trie_free(&d.lib);
    reads_free(&r1);
    reads_free(&r2);
    return status;
}

int kport_count_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                     const char* const* pool1, int npool1,
                     const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                     const char* const* pool2, int npool2,
                     int randomized, int use_first, int diagnostics, int nthreads,
                     int* counts, int* total, kport_table** table, int* b1only, int* b2only) {
    (void)nthreads;
    return run_dual(path1, data1, size1, tmpl1, reverse1, mismatches1, pool1, npool1,
                    path2, data2, size2, tmpl2, reverse2, mismatches2, pool2, npool2,
                    randomized, use_first, diagnostics, CACHE_NONE, counts, total, table, b1only, b2only, NULL, 0, NULL);
}

int kport_trace_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                     const char* const* pool1, int npool1,
                     const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                     const char* const* pool2, int npool2,
                     int randomized, int use_first, int fresh_state,
                     int* index, long long capacity, long long* npairs) {
    /* fresh_state: 0 = one cache for the file, 1 = cache emptied per pair, 2 = no cache at all */
    return run_dual(path1, data1, size1, tmpl1, reverse1, mismatches1, pool1, npool1,
                    path2, data2, size2, tmpl2, reverse2, mismatches2, pool2, npool2,
                    randomized, use_first, 0, fresh_state, NULL, NULL, NULL, NULL, NULL, index, capacity, npairs);
}

static int run_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                            const char* const* pool1, int npool1,
                            const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                            const char* const* pool2, int npool2,
                            int randomized, int use_first,
                            kport_table** table, int* total, int* b1only, int* b2only,
                            int* combo, int* code_out, long long capacity, long long* npairs) {
    reads_t r1, r2;
    if (load_pair(path1, data1, size1, path2, data2, size2, &r1, &r2)) return 1;
    {
        size_t l1 = strlen(tmpl1), l2 = strlen(tmpl2);
        if ((l1 > l2 ? l1 : l2) > MAX_TEMPLATE) {
            reads_free(&r1);
            reads_free(&r2);
            return fail("lacking compile-time support for constant regions longer than 256 bp");
        }
    }
    single_t m1, m2;
    if (single_init(&m1, tmpl1, reverse1 ? 1 : 0, pool1, npool1, mismatches1, DUP_ERROR)) {
        reads_free(&r1);
        reads_free(&r2);
        return 1;
    }
    if (single_init(&m2, tmpl2, reverse2 ? 1 : 0, pool2, npool2, mismatches2, DUP_ERROR)) {
        single_free(&m1);
        reads_free(&r1);
        reads_free(&r2);
        return 1;
    }
    int status = 0;
    if (r1.n != r2.n) status = fail("different number of reads in paired FASTQ files");
    if (!status && combo && (long long)r1.n > capacity) status = fail("trace capacity too small");
    if (!status) {
        if (npairs) *npairs = (long long)r1.n;
        pairs_t collected = { 0, 0, 0 };
        int only1 = 0, only2 = 0;
        for (size_t i = 0; i < r1.n; ++i) {
            int ids[2] = { -1, -1 };
            int code = combope_process(&m1, &m2, r1.buf + r1.off[i], r1.off[i + 1] - r1.off[i],
                                       r2.buf + r2.off[i], r2.off[i + 1] - r2.off[i], randomized, use_first, ids);
            if (code == 1) pairs_push(&collected, ids[0], ids[1]);
            if (code == 2) ++only1;
            if (code == 3) ++only2;
            if (combo) {
                combo[2 * i] = code == 1 ? ids[0] : -1;
                combo[2 * i + 1] = code == 1 ? ids[1] : -1;
                code_out[i] = code;
            }
        }
        if (table) *table = pairs_to_table(&collected);
        if (total) *total = (int)r1.n;
        if (b1only) *b1only = only1;
        if (b2only) *b2only = only2;
        free(collected.v);
    }
    single_free(&m1);
    single_free(&m2);
    reads_free(&r1);
    reads_free(&r2);
    return status;
}

int kport_count_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                             const char* const* pool1, int npool1,
                             const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                             const char* const* pool2, int npool2,
                             int randomized, int use_first, int nthreads,
                             kport_table** table, int* total, int* b1only, int* b2only) {
    (void)nthreads;
    return run_combo_paired(path1, data1, size1, tmpl1, reverse1, mismatches1, pool1, npool1,
                            path2, data2, size2, tmpl2, reverse2, mismatches2, pool2, npool2,
                            randomized, use_first, table, total, b1only, b2only, NULL, NULL, 0, NULL);
}

int kport_trace_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                             const char* const* pool1, int npool1,
                             const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                             const char* const* pool2, int npool2,
                             int randomized, int use_first,
                             int* combo, int* code, long long capacity, long long* npairs) {
    return run_combo_paired(path1, data1, size1, tmpl1, reverse1, mismatches1, pool1, npool1,
                            path2, data2, size2, tmpl2, reverse2, mismatches2, pool2, npool2,
                            randomized, use_first, NULL, NULL, NULL, NULL, combo, code, capacity, npairs);
}

/* --- raw searches: screenCounter src/match_barcodes.cpp:7-37 -------------------- */

static int any_search_batch(const char* const* seqs, int nseqs, const int* caps, int cap_all,
                            const char* const* choices, int nchoices, int reverse, int duplicates,
                            int* index, int* mm, int na_style) {
    int clen, slen;
    if (pool_length(choices, nchoices, &clen)) return 1;
    trie_t t;
    trie_init(&t, clen, duplicates);
    if (trie_fill(&t, choices, nchoices, reverse)) {
        trie_free(&t);
        return 1;
    }
    if (pool_length(seqs, nseqs, &slen)) {
        trie_free(&t);
        return 1;
    }
    for (int i = 0; i < nseqs; ++i) {
        hit_t h = trie_search_any(&t, seqs[i], caps ? caps[i] : cap_all);
        if (na_style) { /* src/match_barcodes.cpp:24-30 */
            index[i] = h.index >= 0 ? h.index : -1;
            mm[i] = h.index >= 0 ? h.total : -1;
        } else {
            index[i] = h.index;
            mm[i] = h.total;
        }
    }
    trie_free(&t);
    return 0;
}

int kport_match_barcodes(const char* const* seqs, int nseqs, const char* const* choices, int nchoices,
                         int substitutions, int reverse, int duplicates, int* index, int* mm) {
    return any_search_batch(seqs, nseqs, NULL, substitutions, choices, nchoices, reverse, duplicates, index, mm, 1);
}

int kport_search_any(const char* const* seqs, int nseqs, const int* caps, const char* const* choices, int nchoices,
                     int max_mismatches, int reverse, int duplicates, int* index, int* mm) {
    (void)max_mismatches;
    return any_search_batch(seqs, nseqs, caps, 0, choices, nchoices, reverse, duplicates, index, mm, 0);
}

int kport_search_segmented2(const char* const* seqs, int nseqs, const int* caps,
                            const char* const* choices, int nchoices, int len1, int len2,
                            int max1, int max2, int duplicates, int* index, int* mm) {
    (void)max1;
    (void)max2;
    int clen;
    if (pool_length(choices, nchoices, &clen)) return 1;
    if (clen != len1 + len2) return fail("variable sequences should have the same length as the sum of segment lengths");
    trie_t t;
    trie_init(&t, clen, duplicates);
    t.nseg = 2;
    t.boundaries[0] = len1;
    t.boundaries[1] = len1 + len2;
    if (trie_fill(&t, choices, nchoices, 0)) {
        trie_free(&t);
        return 1;
    }
    for (int i = 0; i < nseqs; ++i) {
        hit_t h = trie_search_segmented(&t, seqs[i], caps + 2 * i);
        index[i] = h.index;
        mm[i] = h.total;
    }
    trie_free(&t);
    return 0;
}
