// Build-time instantiation of the specialised countRandomBarcodes kernel (spec_handlers.cuh, SPH_KIND 3) for BASELINE
// configs[4]'s shape (12 + 16 + 12 template, both strands, one mismatch, 75-base reads).
#define SPH_KIND 3
#define SPH_MIN_BLOCKS 8
#define SPH_STAGES 2
#define SPH_GROUP 2
#define SPH_SAMPLES 8
#define SPH_USE_FIRST 1
#define SPH_HAS_INDEX 1
#define SPH_A_T 40
#define SPH_A_FB "CAGCTACGTACG----------------CCAGCTCGATCG"
#define SPH_A_RB "CGATCGAGCTGG----------------CGTACGTAGCTG"
#define SPH_A_FWD 1
#define SPH_A_REV 1
#define SPH_A_MM 1
#define SPH_A_MAXMM 1
#define SPH_A_ULEN 75
#define SPH_A_W 3
#define SPH_A_FSTART0 12
#define SPH_A_FLEN0 16
#define SPH_A_RSTART0 12
#define SPH_A_RLEN0 16
#include "spec_handlers.cuh"

namespace scg {
const void* spec_random_default_kernel() { return reinterpret_cast<const void*>(&spec_random_kernel); }
} // namespace scg
