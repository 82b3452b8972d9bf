// gemm_tc.cu -- TMA-staged tcgen05 (3xTF32) GEMM engine for the mel projection and its
// pseudo-inverse.  (placeholder until the tensor-core path lands; the fused FFMA kernels in
// spectral.cu are the product path meanwhile)
#include "spev_internal.cuh"

namespace spev {
int gemm_tc_init(spev_ctx*) { return SPEV_OK; }
void gemm_tc_destroy(spev_ctx*) {}
int launch_mel_project_tc(spev_ctx*, const float*, int64_t, float*, int, float, float, float, cudaStream_t) {
    set_error("spev_mel_project: tensor-core path not built yet");
    return SPEV_E_UNSUPPORTED;
}
}  // namespace spev
