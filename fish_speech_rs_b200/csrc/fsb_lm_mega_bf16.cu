// bf16-weight instantiations of the persistent decode megakernel.
#include "fsb_lm_mega.cuh"
namespace fsb {
cudaError_t mega_launch_bf16(int NB, const MegaParams &mp, int grid, size_t smem, cudaStream_t st) {
    return mega_launch_impl<__nv_bfloat16>(NB, mp, grid, smem, st);
}
}  // namespace fsb
