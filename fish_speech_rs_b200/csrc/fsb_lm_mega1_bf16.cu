// bf16-weight instantiation of the single-row decode megakernel (TMA weight ring).
#include "fsb_lm_mega1.cuh"
namespace fsb {
cudaError_t mega1_launch_bf16(const MegaParams &mp, int grid, size_t smem, cudaStream_t st) {
    return mega1_launch_impl<__nv_bfloat16>(mp, grid, smem, st);
}
}  // namespace fsb
