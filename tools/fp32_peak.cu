// Measured peaks the vocoder roofline needs (BASELINE.md section 2 asked for the first one):
//   * FP32 FMA (CUDA cores): 8 independent FFMA chains per thread, 148 x 4 CTAs x 256 threads
//   * legacy warp-level tensor-core path: mma.sync.m16n8k16 bf16 -> f32 (what a register-fragment conv kernel could use)
// Prints one JSON line; bench.py reads profiles/r02_fp32_peak.json written from it.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters) {
    float a[8], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mma_kernel(float *out, int iters) {
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 4, threads = 256;
    float *out;
    cudaMalloc(&out, (size_t)grid * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best_f = 0, best_m = 0;
    for (int rep = 0; rep < 6; ++rep) {
        const int iters = 20000;
        cudaEventRecord(e0);
        ffma_kernel<<<grid, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 64 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best_f) best_f = tf;
        cudaEventRecord(e0);
        mma_kernel<<<grid, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double tm = 2.0 * 16 * 8 * 16 * 8 * iters * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12;
        if (rep > 0 && tm > best_m) best_m = tm;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_tflops\": %.2f, \"mma_sync_bf16_tflops\": %.2f, "
           "\"how\": \"tools/fp32_peak.cu: 8 independent FFMA chains x 64 FFMA per loop, 4 CTAs x 256 threads per SM, best of 5; "
           "mma.sync.m16n8k16 bf16->f32, 8 independent accumulators per warp, 32 warps per SM, best of 5\"}\n",
           prop.name, prop.multiProcessorCount, best_f, best_m);
    return 0;
}
