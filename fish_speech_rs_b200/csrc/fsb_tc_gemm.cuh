// tcgen05 / TMA / TMEM GEMM used by prefill (fsb_tc_gemm.cu).
#pragma once
#include "fsb_common.cuh"

namespace fsb {

// opaque, 64-byte aligned storage for a CUtensorMap (128 B)
struct alignas(64) TcMap {
    unsigned char bytes[128];
};

// tensor map of a row-major (rows, K) bf16 matrix, box = (64 k, box_rows), 128-byte swizzle
int tc_make_map_bf16(TcMap *out, const void *base, int rows, int K, int box_rows);
// tensor map of a row-major (rows, cols) f32 matrix, box = (box_cols, box_rows), no swizzle, out-of-bounds elements
// read as zero (the vocoder's causal left padding)
int tc_make_map_f32_2d(TcMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
// N tile (prompt positions per CTA) for a prompt chunk of P rows
int tc_pick_bn(int P);
// x (n elements, f32) -> hi | mid | lo bf16 copies at element offsets 0, seg_elems, 2 * seg_elems
int tc_split3(const float *x, __nv_bfloat16 *out, size_t n, size_t seg_elems, cudaStream_t st);
// fused producers of the split operand
int tc_rmsnorm_split3(const float *x, const float *g, float eps, int M, int D, __nv_bfloat16 *out, size_t seg_elems,
                      cudaStream_t st);
int tc_swiglu_split3(const float *g1, const float *g3, size_t n, __nv_bfloat16 *out, size_t seg_elems, cudaStream_t st);
// one-time kernel attribute setup (dynamic shared memory opt-in)
int tc_init();
// C[p, n] = sum_k W[n, k] * (hi + mid + lo)[p, k] (+ resid[p, n]); mx built with box_rows == bn
// ws (optional, ws_floats capacity): split-K workspace; used when the output has too few tiles to fill the GPU
int tc_gemm(const TcMap &mw, const TcMap &mx, int bn, float *C, const float *resid, int P, int N, int K, int x_seg_rows,
            int ldc, cudaStream_t st, float *ws = nullptr, size_t ws_floats = 0);

}  // namespace fsb
