// Build-time instantiation of the specialised countDualBarcodes kernel (spec_handlers.cuh, SPH_KIND 1) for BASELINE
// configs[2]'s shape (two 12 + 20 + 12 templates, forward strands, one mismatch per read, 75-base reads): proves the source
// NVRTC compiles at run time builds for sm_100a, and shows its registers in the build log (-Xptxas -v).
#define SPH_KIND 1
#define SPH_MIN_BLOCKS 8
#define SPH_STAGES 2
#define SPH_GROUP 1
#define SPH_SAMPLES 8
#define SPH_USE_FIRST 1
#define SPH_HAS_INDEX 1
#define SPH_A_T 44
#define SPH_A_FB "CAGCTACGTACG--------------------CCAGCTCGATCG"
#define SPH_A_RB "--------------------------------------------"
#define SPH_A_FWD 1
#define SPH_A_REV 0
#define SPH_A_MM 1
#define SPH_A_MAXMM 1
#define SPH_A_ULEN 75
#define SPH_A_W 3
#define SPH_A_FSTART0 12
#define SPH_A_FLEN0 20
#define SPH_A_RSTART0 12
#define SPH_A_RLEN0 20
#define SPH_B_T 44
#define SPH_B_FB "GATTACAGGCTA--------------------TTGACCGTAGCA"
#define SPH_B_RB "--------------------------------------------"
#define SPH_B_FWD 1
#define SPH_B_REV 0
#define SPH_B_MM 1
#define SPH_B_MAXMM 1
#define SPH_B_ULEN 75
#define SPH_B_W 3
#define SPH_B_FSTART0 12
#define SPH_B_FLEN0 20
#define SPH_B_RSTART0 12
#define SPH_B_RLEN0 20
#include "spec_handlers.cuh"

namespace scg {
const void* spec_dual_default_kernel() { return reinterpret_cast<const void*>(&spec_dual_pe_kernel); }
} // namespace scg
