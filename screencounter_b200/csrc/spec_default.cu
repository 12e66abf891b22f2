// Build-time instantiation of the specialised single-barcode kernel for its default configuration
// (the 12 + 20 + 12 template of BASELINE configs[1], one mismatch, both strands, 75-bp reads): the
// same source NVRTC compiles at run time for any other template (jit.cpp).
#include "spec_single.cuh"

namespace scg {
const void* spec_single_default_kernel() { return reinterpret_cast<const void*>(&spec_single_kernel_default); }
} // namespace scg
