// Probe: which f32 2-D tensor-map loads does the B200 TMA accept?  (box wider than the tensor, negative coordinates, ...)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int box_cols, int box_rows, float *out) {
    extern __shared__ float raw[];
    float *sm = raw + (((128u - (s32(raw) & 127u)) & 127u) >> 2);
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(box_cols * box_rows * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sm)),
                     "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1)
                     : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < box_cols * box_rows; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN_enc enc = (PFN_enc)fp;
    struct Case { int rows, cols, bc, br, c0, c1; const char *what; };
    const Case cases[] = {
        {256, 32, 132, 16, -2, 0, "box wider than tensor, negative start (the failing vocoder case)"},
        {256, 32, 132, 16, 0, 0, "box wider than tensor, start 0"},
        {256, 32, 32, 16, -2, 0, "box == tensor width, negative start"},
        {256, 512, 132, 16, -2, 0, "box narrower than tensor, negative start"},
        {256, 512, 132, 16, 4, 0, "box narrower than tensor, positive start"},
        {256, 512, 128, 16, -4, 0, "box 128 cols (512 B rows), negative start"},
        {256, 512, 180, 16, -52, 16, "box 180 cols, start -52, row 16"},
        {256, 512, 64, 16, -4, 0, "box 64 cols (256 B rows), negative start"},
        {256, 512, 256, 16, -52, 0, "box 256 cols, start -52"},
        {256, 512, 132, 16, 0, 0, "box 132 cols, start 0"},
        {256, 512, 128, 16, 0, 0, "box 128 cols, start 0"},
    };
    for (const Case &cs : cases) {
        if (++idx != only && only >= 0) continue;
        std::vector<float> h((size_t)cs.rows * cs.cols);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) + 1.f;
        float *d, *o;
        cudaMalloc(&d, h.size() * 4);
        cudaMalloc(&o, 256 * 256 * 4);
        cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)cs.cols, (cuuint64_t)cs.rows}, strides[1] = {(cuuint64_t)cs.cols * 4};
        cuuint32_t box[2] = {(cuuint32_t)cs.bc, (cuuint32_t)cs.br}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("%-70s encode=%d ", cs.what, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        probe<<<1, 128, cs.bc * cs.br * 4 + 256>>>(tm, cs.c0, cs.c1, cs.bc, cs.br, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch -> %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> got((size_t)cs.bc * cs.br);
        cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < cs.br; ++rr)
            for (int cc = 0; cc < cs.bc; ++cc) {
                const int gc = cs.c0 + cc, gr = cs.c1 + rr;
                const float exp = (gc >= 0 && gc < cs.cols && gr < cs.rows) ? h[(size_t)gr * cs.cols + gc] : 0.f;
                if (got[(size_t)rr * cs.bc + cc] != exp) ++bad;
            }
        printf("ok, mismatches=%d\n", bad);
        cudaFree(d);
        cudaFree(o);
    }
    return 0;
}
