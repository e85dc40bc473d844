// Instantiation + launcher of the wide-batch decode megakernel (tcgen05 + TMA weight ring, 9..32 rows).
#include "fsb_lm_megab.cuh"
namespace fsb {
int megab_max_stages(int npad) { return npad <= 16 ? 8 : 7; }
size_t megab_smem_bytes(int npad, int nstages) {
    // ring | operand / staging region | barriers, tables, sampler state (< 6 KB) | slack for the 1024-byte alignment
    return (size_t)nstages * kMBStage + (npad <= 16 ? kMBXsBytes16 : kMBXsBytes32) + 6144 + 1024;
}
cudaError_t megab_launch(const MegaParams &mp, const MegaBExtra &ex, int npad, int grid, size_t smem, cudaStream_t st) {
    const void *kern = npad <= 16 ? (const void *)megab_decode_kernel<16> : (const void *)megab_decode_kernel<32>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&mp, (void *)&ex};
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kMBThreads), args, smem, st);
}
}  // namespace fsb
