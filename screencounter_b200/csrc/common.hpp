// Shared host-side plumbing: error type, CUDA status checks, small helpers.
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace scg {

// Every failure inside the library is an scg::Error; the C ABI (api.cpp) turns it into a
// status code + message, mirroring how BEGIN_RCPP/END_RCPP turns kaori's exceptions into
// R errors (reference src/RcppExports.cpp:16,33).
struct Error : public std::runtime_error {
    explicit Error(const std::string& what) : std::runtime_error(what) {}
};

#define SCG_CUDA_CHECK(expr)                                                                       \
    do {                                                                                           \
        cudaError_t scg_status__ = (expr);                                                         \
        if (scg_status__ != cudaSuccess) {                                                         \
            throw ::scg::Error(std::string("CUDA error: ") + cudaGetErrorString(scg_status__) +    \
                               " (" #expr ") at " __FILE__ ":" + std::to_string(__LINE__));        \
        }                                                                                          \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

inline uint32_t next_pow2(uint32_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

} // namespace scg
