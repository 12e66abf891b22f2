/* TEST INFRASTRUCTURE ONLY -- the product path never links or calls this.
 *
 * Plain-C restatement of the reference's barcode-counting algorithm (kaori
 * v1.1.1 as vendored in crisprVerse/screenCounter 1.5.1).  Every function in
 * kaori_port.c cites the reference file:line it follows (paths relative to
 * /root/reference/inst/include/kaori unless written out).
 *
 * PINNED: tests/test_oracle_golden.py checks this port against (a) every
 * known-answer vector of the reference's own testthat suite and (b) the
 * unmodified reference compiled in oracle/_ref (fuzzed, all handlers).
 *
 * The exported signatures mirror oracle/ref_harness.cpp one for one (prefix
 * kport_ instead of kref_) so that the same ctypes binding drives either.
 */
#ifndef KAORI_PORT_H
#define KAORI_PORT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kport_table kport_table;

const char* kport_last_error(void);

size_t kport_table_size(const kport_table* t);
int kport_table_width(const kport_table* t);
void kport_table_copy(const kport_table* t, int* keys, char* strings, int* freq);
void kport_table_free(kport_table* t);

int kport_count_reads(const char* path, const char* data, size_t size, long long* nreads, long long* nbases);
int kport_parse(const char* path, const char* data, size_t size, char* bases, long long* offsets);

int kport_count_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       const char* const* pool, int npool, int mismatches, int use_first, int nthreads,
                       int* counts, int* total);
int kport_trace_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       const char* const* pool, int npool, int mismatches, int use_first,
                       int* index, int* info, long long capacity, long long* nreads);

int kport_count_random(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                       int mismatches, int use_first, int nthreads, kport_table** table, int* total);

int kport_count_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                             const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                             int mismatches, int use_first, int nthreads, kport_table** table, int* total);
int kport_trace_combo_single(const char* path, const char* data, size_t size, const char* tmpl, int strand,
                             const char* const* pool1, int npool1, const char* const* pool2, int npool2,
                             int mismatches, int use_first, int* combo, long long capacity, long long* nreads);

int kport_count_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                                const char* const* pools_flat, int npools, int nchoices, int strand,
                                int mismatches, int use_first, int diagnostics, int nthreads,
                                int* counts, int* total, kport_table** table);
int kport_trace_dual_single_end(const char* path, const char* data, size_t size, const char* tmpl,
                                const char* const* pools_flat, int npools, int nchoices, int strand,
                                int mismatches, int use_first, int* index, long long capacity, long long* nreads);

int kport_count_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                     const char* const* pool1, int npool1,
                     const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                     const char* const* pool2, int npool2,
                     int randomized, int use_first, int diagnostics, int nthreads,
                     int* counts, int* total, kport_table** table, int* b1only, int* b2only);
int kport_trace_dual(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                     const char* const* pool1, int npool1,
                     const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                     const char* const* pool2, int npool2,
                     int randomized, int use_first, int fresh_state,
                     int* index, long long capacity, long long* npairs);

int kport_count_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                             const char* const* pool1, int npool1,
                             const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                             const char* const* pool2, int npool2,
                             int randomized, int use_first, int nthreads,
                             kport_table** table, int* total, int* b1only, int* b2only);
int kport_trace_combo_paired(const char* path1, const char* data1, size_t size1, const char* tmpl1, int reverse1, int mismatches1,
                             const char* const* pool1, int npool1,
                             const char* path2, const char* data2, size_t size2, const char* tmpl2, int reverse2, int mismatches2,
                             const char* const* pool2, int npool2,
                             int randomized, int use_first,
                             int* combo, int* code, long long capacity, long long* npairs);

int kport_match_barcodes(const char* const* seqs, int nseqs, const char* const* choices, int nchoices,
                         int substitutions, int reverse, int duplicates, int* index, int* mm);
int kport_search_any(const char* const* seqs, int nseqs, const int* caps, const char* const* choices, int nchoices,
                     int max_mismatches, int reverse, int duplicates, int* index, int* mm);
int kport_search_segmented2(const char* const* seqs, int nseqs, const int* caps,
                            const char* const* choices, int nchoices, int len1, int len2,
                            int max1, int max2, int duplicates, int* index, int* mm);

#ifdef __cplusplus
}
#endif

#endif
