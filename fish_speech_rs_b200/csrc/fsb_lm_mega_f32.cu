// fp32-weight (parity mode) instantiations of the persistent decode megakernel.
#include "fsb_lm_mega.cuh"
namespace fsb {
cudaError_t mega_launch_f32(int NB, const MegaParams &mp, int grid, size_t smem, cudaStream_t st) {
    return mega_launch_impl<float>(NB, mp, grid, smem, st);
}
}  // namespace fsb
