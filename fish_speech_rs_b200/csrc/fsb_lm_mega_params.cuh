// Parameters of the persistent decode megakernel (fsb_lm_mega.cuh) shared with the host code.
#pragma once
#include <cuda.h>

#include "fsb_common.cuh"
#include "fsb_sample.cuh"

namespace fsb {

constexpr int kMegaThreads = 512;
constexpr int kMegaWarps = kMegaThreads / 32;
constexpr int kMegaPre = 4;        // 16-byte weight loads per lane kept in flight across a grid barrier (one task)
constexpr int kMegaChunk = 128;    // cached positions one attention work item covers (staged in shared memory)
constexpr int kMegaKvStride = 80;  // floats per staged K/V row (64 + pad: conflict-free LDS.128 for 8 rows at once)

// single-row kernel (fsb_lm_mega1.cuh)
constexpr int kM1Warps = 16;                      // compute warps
constexpr int kM1Threads = kM1Warps * 32;         // compute threads (named barrier 1)
constexpr int kM1AllThreads = kM1Threads + 32;    // + the TMA producer warp
constexpr int kM1Slice = 1024;                    // K elements of one task
constexpr int kM1ChunkBytes = 32768;              // one ring slot
constexpr int kM1MaxDepth = 6;
constexpr int kM1AttChunk = 64;                   // cached positions per attention item
constexpr int kM1KvStride = 80;                   // floats per staged K/V row (conflict-free LDS.128)
constexpr int kM1ValFloats = 128;
constexpr int kM1Rep = 8;  // replicas of the activation vectors every CTA reads after a barrier

// word offsets of the LL arena (MegaParams::ll); kM1FlagStride flag slots per replica
constexpr int kM1FlagStride = 160;
struct LLLayout {
    size_t xt, ht, qt, nkv, fkv, pt, lt, ct, fl, total;
    __host__ __device__ LLLayout(int D, int I, int Hhd, int KV, int hd, int NFL, int fast_len, int H, int slots, int ldl) {
        size_t o = 0;
        xt = o; o += (size_t)2 * kM1Rep * D;
        ht = o; o += (size_t)kM1Rep * I;
        qt = o; o += Hhd;
        nkv = o; o += (size_t)2 * KV * hd;
        fkv = o; o += (size_t)NFL * 2 * KV * fast_len * hd;
        pt = o; o += (size_t)H * slots * (hd + 4);
        lt = o; o += ldl;
        ct = o; o += 16;
        fl = o; o += (size_t)kM1Rep * kM1FlagStride;
        total = o;
    }
};

struct MegaLayer {
    const void *wqkv, *wo, *w1, *w3, *w2;
    const float *attn_norm, *ffn_norm;
};

struct MegaParams {
    const MegaLayer *slow, *fast;
    int NL, NFL, D, I, H, KV, hd, QKV, C, CS;
    int n_slow_logits, slow_row0, slow_rest_base;
    const void *emb, *cb_emb, *out_w, *fast_emb, *fast_out;
    const float *norm, *fast_norm;
    float eps;
    float *kc, *vc, *fkc, *fvc;
    int max_len, fast_len, max_batch;
    const float *cosT, *sinT;
    float *x;        // (B, D) slow residual stream == pre-norm hidden
    float *fx;       // (B, D) fast residual stream
    float *q;        // (B, H*hd) roped q
    float *partial;  // (B, H, 2*n_chunks_max, hd+4): o[hd], m, l, pad (16-byte aligned slots)
    int n_chunks_max;  // ceil(max_len / kMegaChunk)
    float *h;        // (B, I)
    float *logits;   // (B, ldl)
    int ldl;
    GenState st;  // by value: every field is launch-constant (the arrays it points to are not)
    unsigned int *bar;
    int nb, nframes, first_is_tail;
    int row0;  // global index of local row 0 (groups of <= 8 rows of a larger batch): Philox row = row0 + b
    uint32_t sem_start, sem_end;
    int has_end;
    int val_floats;  // capacity of the per-task result array (floats)
    int xs_floats;   // capacity of the activation staging area (floats)
    int ring_depth;  // single-row kernel (fsb_lm_mega1.cuh): 32 KB slots of the TMA weight ring
    int kvs_floats;  // single-row kernel: K/V staging area == sampler scratch (floats)
    float *rep;      // single-row kernel: replicas 1..kM1Rep-1 of x | fx | h  ((kM1Rep-1) * (2 D + I) floats)
    unsigned long long *ll;  // single-row kernel: arena of tagged words (flag-in-data synchronisation), null = grid barriers
    unsigned long long *ll_xt, *ll_ht, *ll_qt, *ll_nkv, *ll_fkv, *ll_pt, *ll_lt, *ll_ct, *ll_fl;  // its regions (LLLayout)
    size_t slow_kv_stride, fast_kv_stride;  // floats per layer of the slow / fast K (or V) cache
    int sampler_cta; // single-row kernel: CTA that only samples (-1: CTA 0 samples and streams)
    unsigned long long *dbg;  // optional (FSB_MEGA_TIMERS=1): per phase kind {work ns, barrier ns, count} of CTA 0 and the last CTA
};

// ---- wide-batch kernel (fsb_lm_megab.cuh): 9..32 rows, bf16 weights, tcgen05 + TMA weight ring
constexpr int kMBWorkers = 256;               // worker threads (warps 0-7, named barrier 1)
constexpr int kMBThreads = kMBWorkers + 64;   // + TMA producer warp + MMA warp
constexpr int kMBStage = 16384;               // one ring stage: 128 rows x 64 k bf16
constexpr int kMBXsBytes16 = 81920;           // NPAD 16: two activation slice buffers | two K/V chunk buffers | sampler scratch
constexpr int kMBXsBytes32 = 98304;           // NPAD 32 (two 48 KB slice buffers)
constexpr int kMBMaxStages = 8;
constexpr int kMBChunk = 64;                  // cached positions staged at once
constexpr int kMBKvStride = 80;               // floats per staged K/V row
constexpr int kMBMaxSplit = 16;               // position ranges per (row, kv head)
constexpr int kMBSsq = 32;                    // sum-of-squares slots per row
constexpr int kMBCntStride = 80;              // arrival counters per phase kind
constexpr long long kMBSpinLimit = 6000000000ll;  // ~3 s: a broken protocol traps instead of hanging the GPU

struct MegaBExtra {
    const CUtensorMap *maps;  // 5 per layer (wqkv, wo, w1, w3, w2), slow then fast, then out_w, fast_out
    float *ws;                // partial outputs of the fused FFN [64 blocks][NPAD][dim]
    unsigned *cnt;            // scratch counters: [B * KV] attention arrival counters, "go" flags
    unsigned *att_cnt;
    unsigned *go;             // [4] "some row continues into frame f" flags
    float *att;               // (B, H * hd) attention output
    float *apart;             // (B, H, kMBMaxSplit, hd + 4) partial attention (o, m, l)
    float *ssq_x, *ssq_fx;    // (B, kMBSsq) partial sums of squares of the slow / fast stream
    // operand images (K = 1024: 16 k-blocks of [3 NPAD rows][64] bf16, 128-byte swizzle) of the three activations the
    // projections consume, written by their PRODUCERS with the consumer's norm weight folded in (x * g split into
    // hi + mid + lo); a consumer stages a 256-column slice with one cp.async.bulk
    unsigned char *xop_x, *xop_fx, *xop_att;
    int nstages;
    int head_tiles;           // 128-row tiles of the constrained slow head (without the extra <|im_end|> tile)
    int head_extra;           // 1: <|im_end|> is not adjacent to the semantic range -> one more tile for logit 0
};


// Launchers, one translation unit per weight dtype (fsb_lm_mega_{bf16,f32}.cu).  NB in {1,2,4,8}.
cudaError_t mega_launch_bf16(int NB, const MegaParams &mp, int grid, size_t smem, cudaStream_t st);
cudaError_t mega_launch_f32(int NB, const MegaParams &mp, int grid, size_t smem, cudaStream_t st);
// single-row kernel with the TMA weight ring (fsb_lm_mega1.cuh)
cudaError_t mega1_launch_bf16(const MegaParams &mp, int grid, size_t smem, cudaStream_t st);
cudaError_t mega1_launch_f32(const MegaParams &mp, int grid, size_t smem, cudaStream_t st);

// wide-batch kernel; npad in {16, 32}
cudaError_t megab_launch(const MegaParams &mp, const MegaBExtra &ex, int npad, int grid, size_t smem, cudaStream_t st);
size_t megab_smem_bytes(int npad, int nstages);
int megab_max_stages(int npad);

}  // namespace fsb
