// Shared declarations of the persistent TMEM-A 3xTF32 GEMM (tca_gemm.cu) and its users.
#pragma once
#include <cuda.h>

#include "ppd_common.cuh"

namespace ppd {
namespace tca {

constexpr int BM = 128;            // UMMA M
constexpr int BK = 32;             // floats per k-block (one 128-byte row)
constexpr int kSA = 8;             // shared-memory A stages (128 KB in flight per SM covers the HBM latency)
constexpr int kMaxSB = 8;          // shared-memory B stages ([hi | lo] each): as many as fit, at most 8
constexpr size_t kSmemBudget = 225 * 1024;
constexpr int kTA = 4;             // tensor-memory A stages (64 columns each: 32 hi + 32 lo)
constexpr int kXformWarps = 8;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (4 + kXformWarps + kEpiWarps);   // + A producer, MMA issuer, B producer, spare
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = 0, kAccStride = 128;   // two accumulators of up to 128 columns
constexpr uint32_t kTaCol0 = 256;

struct Args {
    float* C; int64_t ldc;
    int64_t I, J, KK;
    const float* bias; const float* mask; int64_t ldm;
    int relu, accumulate, transpose_out;
    int bn, a_mn, b_mn, b_presplit;
    int num_m, num_n, splits;
    int64_t kk_per_split;
    float* partial;
    int total_items;
    int sb_stages;
};

struct Plan { int bn, num_m, num_n, splits; int64_t kk_per_split; size_t ws; };

Plan make_plan(int64_t I, int64_t J, int64_t KK, size_t ws_avail, bool limit);
int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                CUtensorMapSwizzle swz);
// Enqueues the GEMM kernel only; with plan.splits > 1 the raw partial products are in `workspace` and the caller
// reduces them (tc_splitk_reduce_kernel).  b_lo != NULL: B is pre-split (g->B holds the hi parts, b_lo the residuals).
int launch(const ppd_gemm_args* g, int transpose_out, const float* b_lo, void* workspace, size_t workspace_bytes,
           cudaStream_t s, Plan* plan_out);

// hi = TF32(x), lo = TF32(x - hi) with the kernel's rounding; n a multiple of 4, pointers 16-byte aligned.
int split_operand(const float* x, float* hi, float* lo, int64_t n, cudaStream_t s);

}  // namespace tca
}  // namespace ppd
