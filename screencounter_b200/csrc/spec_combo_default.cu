// Build-time instantiation of the specialised countComboBarcodes kernel (spec_handlers.cuh, SPH_KIND 2) for BASELINE
// configs[3]'s shape (8 + 20 + 8 + 20 + 8 template, both strands, one mismatch, 75-base reads).
#define SPH_KIND 2
#define SPH_MIN_BLOCKS 8
#define SPH_STAGES 2
#define SPH_GROUP 2
#define SPH_SAMPLES 8
#define SPH_USE_FIRST 1
#define SPH_HAS_INDEX 1
#define SPH_A_T 64
#define SPH_A_FB "CAGCTACG--------------------GGTACCTT--------------------CGATCGAG"
#define SPH_A_RB "CTCGATCG--------------------AAGGTACC--------------------CGTAGCTG"
#define SPH_A_FWD 1
#define SPH_A_REV 1
#define SPH_A_MM 1
#define SPH_A_MAXMM 1
#define SPH_A_ULEN 75
#define SPH_A_W 3
#define SPH_A_FSTART0 8
#define SPH_A_FLEN0 20
#define SPH_A_RSTART0 8
#define SPH_A_RLEN0 20
#define SPH_A_FSTART1 36
#define SPH_A_FLEN1 20
#define SPH_A_RSTART1 36
#define SPH_A_RLEN1 20
#include "spec_handlers.cuh"

namespace scg {
const void* spec_combo_default_kernel() { return reinterpret_cast<const void*>(&spec_combo_kernel); }
} // namespace scg
