// Latency probe for the megakernel design: L2 round trip (ld.cg pointer chase), DRAM round trip,
// and grid-barrier cost for 148 co-resident CTAs with different arrive/wait flavours.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

__global__ void chase(const int *buf, int steps, long long *out, int *sink) {
    int idx = threadIdx.x == 0 ? blockIdx.x * 64 : 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; ++i) idx = __ldcg(buf + idx);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; sink[blockIdx.x] = idx; }
}

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void barrier_loop(unsigned *bar, int iters, long long *out, float *data) {
    unsigned target = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // a little "work": every thread writes one float other CTAs will read next round
        data[(blockIdx.x * blockDim.x + threadIdx.x)] = (float)it;
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            if (MODE == 0) { __threadfence(); atomicAdd(bar, 1u); while ((int)(ld_relaxed(bar) - target) < 0) {} __threadfence(); }
            if (MODE == 1) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory"); while ((int)(ld_acquire(bar) - target) < 0) {} }
            if (MODE == 2) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory"); while ((int)(ld_relaxed(bar) - target) < 0) {} asm volatile("fence.acquire.gpu;" ::: "memory"); }
        }
        __syncthreads();
        if (MODE == 3) cg::this_grid().sync();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// barrier + one dependent L2 read of data written by another CTA in the previous round
template <int MODE>
__global__ void barrier_read_loop(unsigned *bar, int iters, long long *out, float *data) {
    unsigned target = 0;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        data[(blockIdx.x * blockDim.x + threadIdx.x)] = (float)it + acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while ((int)(ld_relaxed(bar) - target) < 0) {}
            asm volatile("fence.acquire.gpu;" ::: "memory");
        }
        __syncthreads();
        const int other = (blockIdx.x + 37) % gridDim.x;
        acc += __ldcg(data + other * blockDim.x + threadIdx.x) * 1e-9f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0 + (long long)(acc * 0.f);
}

// One synthetic "phase" of the decode megakernel: grid barrier -> reload a 4 KB activation written by
// another CTA -> every warp streams `tasks` x 2 KB of never-reused weights -> write 1 result per warp.
// MODE 0: weight loads issued after the barrier; 1: first task preloaded between arrive and wait;
// 2: like 1 plus prefetch.global.L2 of the NEXT iteration's weights; 3: no weights at all.
template <int MODE>
__global__ void phase_loop(unsigned *bar, int iters, long long *out, float *act, const uint4 *w, size_t w_elems,
                           int tasks, long long *split) {
    __shared__ float xs[1024];
    unsigned target = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float acc = 0.f;
    long long t_x = 0, t_w = 0, t_b = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // this CTA's slice of iteration `it`: contiguous, never reused
        const size_t base = (((size_t)it * gridDim.x + blockIdx.x) * nw * tasks * 128) % (w_elems - (size_t)nw * tasks * 128 * 2);
        const uint4 *wp = w + base + (size_t)warp * tasks * 128 + lane;
        uint4 pre[4];
        long long a = clock64();
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[i] = __ldcs(wp + i * 32);
        }
        if (MODE == 2) {
            const size_t nb = (((size_t)(it + 1) * gridDim.x + blockIdx.x) * nw * tasks * 128) % (w_elems - (size_t)nw * tasks * 128 * 2);
            const char *np = (const char *)(w + nb);
            for (int ln = threadIdx.x; ln < nw * tasks * 128 * 16 / 128; ln += blockDim.x)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)ln * 128));
        }
        if (threadIdx.x == 0) {
            while ((int)(ld_relaxed(bar) - target) < 0) {}
            asm volatile("fence.acquire.gpu;" ::: "memory");
        }
        __syncthreads();
        long long b = clock64();
        const int other = (blockIdx.x + 37) % gridDim.x;
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = __ldcg(act + other * 1024 + i);
        __syncthreads();
        long long c = clock64();
        if (MODE != 3) {
            for (int t = 0; t < tasks; ++t) {
                if (!((MODE == 1 || MODE == 2) && t == 0)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) pre[i] = __ldcs(wp + (size_t)t * 128 + i * 32);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    acc += __uint_as_float(pre[i].x) * xs[(lane * 4 + i) & 1023] + __uint_as_float(pre[i].w) * 1e-30f;
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) act[blockIdx.x * 1024 + warp + (it & 1) * 512] = acc * 1e-30f + (float)it;
        long long d = clock64();
        t_b += b - a; t_x += c - b; t_w += d - c;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) {
        out[blockIdx.x] = t1 - t0;
        if (blockIdx.x == 0) { split[0] = t_b; split[1] = t_x; split[2] = t_w; }
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int G = prop.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", prop.name, G, prop.clockRate);
    // pointer chase: 8 MB (L2 resident) and 2 GB (DRAM), stride-chained
    for (size_t bytes : {(size_t)8 << 20, (size_t)2 << 30}) {
        const size_t n = bytes / 4;
        std::vector<int> h(n);
        const size_t stride = 4099 * 32;  // ints; co-prime-ish walk
        for (size_t i = 0; i < n; ++i) h[i] = (int)((i + stride) % n);
        int *d; long long *out; int *sink;
        cudaMalloc(&d, bytes); cudaMalloc(&out, 8 * G); cudaMalloc(&sink, 4 * G);
        cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; ++rep) chase<<<1, 32>>>(d, 2000, out, sink);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("ld.cg chase over %zu MB: %.1f cycles/load (1 warp)\n", bytes >> 20, c / 2000.0);
        chase<<<G, 32>>>(d, 2000, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("  with %d CTAs chasing: %.1f cycles/load\n", G, c / 2000.0);
        cudaFree(d); cudaFree(out); cudaFree(sink);
    }
    unsigned *bar; long long *out; float *data;
    cudaMalloc(&bar, 4); cudaMalloc(&out, 8 * G); cudaMalloc(&data, 4 * G * 512);
    const int iters = 2000;
    auto run = [&](const char *name, const void *k, int threads) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(bar, 0, 4);
            int it = iters;
            void *args[] = {&bar, &it, &out, &data};
            cudaLaunchCooperativeKernel(k, dim3(G), dim3(threads), args, 0, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        }
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("%-46s %d thr: %.0f cycles/iter = %.2f us\n", name, threads, c / (double)iters, c / (double)iters / (prop.clockRate * 1e-3));
    };
    for (int threads : {128, 512}) {
        run("fence+atomicAdd / relaxed spin / fence", (const void *)barrier_loop<0>, threads);
        run("red.release / ld.acquire spin", (const void *)barrier_loop<1>, threads);
        run("red.release / relaxed spin / fence.acquire", (const void *)barrier_loop<2>, threads);
        run("cooperative_groups grid.sync()", (const void *)barrier_loop<3>, threads);
        run("barrier(mode 2) + dependent ld.cg of remote data", (const void *)barrier_read_loop<2>, threads);
    }
    {
        const size_t w_bytes = (size_t)3 << 30;
        uint4 *w; float *act; long long *split;
        cudaMalloc(&w, w_bytes); cudaMemset(w, 0, w_bytes);
        cudaMalloc(&act, 4 * G * 1024); cudaMemset(act, 0, 4 * G * 1024);
        cudaMalloc(&split, 64);
        size_t w_elems = w_bytes / 16;
        const int it2 = 1500;
        auto runp = [&](const char *name, const void *k, int tasks) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(bar, 0, 4);
                int it = it2;
                void *args[] = {&bar, &it, &out, &act, &w, &w_elems, &tasks, &split};
                cudaLaunchCooperativeKernel(k, dim3(G), dim3(512), args, 0, 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            }
            long long c, sp[3];
            cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(sp, split, 24, cudaMemcpyDeviceToHost);
            const double f = 1.0 / it2 / (prop.clockRate * 1e-3);
            const double bytes = (double)G * 16 * tasks * 2048;
            printf("%-34s tasks/warp %d: %.2f us/phase (barrier %.2f, x reload %.2f, weights %.2f) -> %.0f GB/s\n", name, tasks,
                   c * f, sp[0] * f, sp[1] * f, sp[2] * f, bytes / (c * f * 1e-6) / 1e9);
        };
        for (int tasks : {1, 4}) {
            runp("weights after barrier", (const void *)phase_loop<0>, tasks);
            runp("first task preloaded", (const void *)phase_loop<1>, tasks);
            runp("preload + L2 prefetch of next", (const void *)phase_loop<2>, tasks);
        }
        runp("no weights", (const void *)phase_loop<3>, 1);
    }
    return 0;
}
