// K4 — tcgen05/TMEM matcher (placeholder until the kernel lands; the dispatcher
// routes every batch through the streaming scan while gemm_supported() is false).
#include "frp_internal.h"

namespace frp {

bool gemm_supported(const frp_gallery* /*g*/) { return false; }
int gemm_num_lists(const frp_gallery* g) { return g->sm_count; }
int launch_match_gemm(frp_gallery* /*g*/, const __half* /*q16*/, const QParams* /*qp*/, int /*B*/,
                      int /*KP*/, uint64_t* /*cand*/, int /*nlists*/, cudaStream_t /*st*/) {
    set_error("tcgen05 matcher not built");
    return FRP_E_UNSUPPORTED;
}

}  // namespace frp
