// tcgen05 engine (placeholder until the TMA/UMMA kernels land): reports "unsupported" so that AUTO falls
// back to the CUDA-core engine and ENGINE_TC fails loudly.
#include "common.cuh"
namespace vp {
bool tc_available() { return false; }
int launch_tapgemm_tc(const TapGemm&, cudaStream_t) { set_error("tcgen05 engine not built"); return VP_EUNSUPPORTED; }
int launch_tapwgrad_tc(const TapWgrad&, cudaStream_t) { set_error("tcgen05 engine not built"); return VP_EUNSUPPORTED; }
}  // namespace vp
