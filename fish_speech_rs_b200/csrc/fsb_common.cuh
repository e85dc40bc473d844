// Shared host/device helpers for libfsb (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/fsb.h"

namespace fsb {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
const char *get_error();

#define FSB_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            fsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return FSB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define FSB_TRY(expr)                \
    do {                             \
        int _s = (expr);             \
        if (_s != FSB_OK) return _s; \
    } while (0)

#define FSB_REQUIRE(cond, code, ...)    \
    do {                                \
        if (!(cond)) {                  \
            fsb::set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

// Picks the device and refuses anything that is not sm_100 (no fallback path exists).
int select_device(int device);

// ---------------------------------------------------------------- weights
struct DevTensor {
    void *ptr = nullptr;
    int dtype = FSB_F32;
    int ndim = 0;
    int64_t shape[4] = {0, 0, 0, 0};
    size_t numel() const {
        size_t n = 1;
        for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
        return n;
    }
};

// Looks up `name` in the caller's table, checks the shape, uploads it as `want_dtype`.
int upload_tensor(const fsb_tensor *table, size_t n, const std::string &name, std::vector<int64_t> shape,
                  int want_dtype, cudaStream_t stream, DevTensor *out, std::vector<void *> *owned);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte streaming loads of weights: 4 floats, or 8 bf16 widened to floats.
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_as(float *p, float v) { *p = v; }
__device__ __forceinline__ void store_as(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// silu as the reference writes it: x / (1 + exp(-x))  (candle_nn::ops::silu)
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
#endif

}  // namespace fsb
