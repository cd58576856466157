// Shared declarations of the persistent TMEM-A 3xTF32 GEMM (tca_gemm.cu) and its users.
#pragma once
#include <cuda.h>

#include "ppd_common.cuh"

namespace ppd {
namespace tca {

constexpr int BM = 128;            // UMMA M
constexpr int BK = 32;             // floats per k-block (one 128-byte row)
constexpr int kSA = 8;             // shared-memory A stages (128 KB in flight per SM covers the HBM latency)
constexpr int kMaxTaps = 64;       // k-blocks of a tile-resident convolution (modes 6 / 7): per-k-block line offsets live in shared memory
constexpr int kMaxSB = 18;         // shared-memory B stages ([hi | lo] each): a ring of as many as fit (at most 8), or all k-blocks resident
constexpr size_t kSmemBudget = 225 * 1024;
constexpr int kMaxChainKb = 176;   // weight gradients: k-blocks accumulated into one TMEM accumulator before the partial tile is flushed
constexpr int kTA = 4;             // tensor-memory A stages (64 columns each: 32 hi + 32 lo) ...
constexpr int kTP = kTA / 2;       // ... handed over in pairs: slots of two stages (two k-blocks), filled by one transform group
#ifndef PPD_GROUPS
#define PPD_GROUPS 2
#endif
constexpr int kGroups = PPD_GROUPS;   // transform groups of four warps taking k-blocks in turn
constexpr int kXformWarps = 4 * kGroups;
constexpr int kEpiWarps = 4;
#ifndef PPD_APROD
#define PPD_APROD 2
#endif
constexpr int kAProd = PPD_APROD;      // A-producer warps of a convolution: warps 0, 3 (and, with 4, two extra warps: measured no faster)
constexpr int kFirstExtra = 4 + kXformWarps + kEpiWarps;
constexpr int kThreads = 32 * (kFirstExtra + (kAProd == 4 ? 2 : 0));   // A producer, MMA issuer, B producer, A producer 2, transform, epilogue (, A producers 3-4)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = 0, kAccStride = 128;   // two accumulators of up to 128 columns
constexpr uint32_t kTaCol0 = 256;

// Implicit-GEMM convolution: the A tile is not a slice of a matrix but `nseg` SEGMENTS of `segw` pixels, each loaded by
// one 4-D TMA box {32 floats, segw pixels, 1, 1} straight from the NHWC activation tensor (OOB coordinates read as zero).
//   mode 1, forward:  segment = output row (b, oy); k-block kb = 32 floats of the (ky, kx, c) patch: filter row
//                     ky = kb / kpk, offset (kb % kpk) * 32 inside the KW*C contiguous floats of that row.
//                     A map: dims {KW*C, OW (stride s*C), Hin, B} -- overlapping strides, an im2col VIEW of x.
//   mode 2, dgrad:    gather form, no atomics.  dx pixels are grouped by their parity class (py, px) = (y % s, x % s);
//                     segment = (b, i) with dx row y = s*i + py and the Wq = Win/s pixels x = s*j + px; k-block kb =
//                     tap (dky, dkx) of the T x T taps that reach this class (ky = py + s*dky) x 32-channel chunk of dY.
//                     A map: dY dims {Cout, OWy, OHy, B}; coordinates (chunk*32, -dkx, i - dky, b).
//   mode 3, wgrad:    dW^T[(ky,kx,c), cout] = sum over pixels of patch[pixel, (ky,kx,c)] * dY[pixel, cout].  The 128 tile rows are
//                     two CHUNKS of 64 consecutive patch floats; a k-block is 32 pixel slots = nseg segments of segw pixels
//                     (pad slots are zeroed by the transform warps); boxes {64 floats, segw pixels} land as [pixel][64] so
//                     that consecutive lanes (= consecutive patch floats) read consecutive words.  NCHW inputs (the
//                     observations, C = 3, one chunk = one channel's 8x8 patch) use a 5-D view {kx, ky, ox, oy, b*C+c}.
//                     The contraction over all pixels is split over CTAs; partial products are reduced afterwards.
//   mode 4, forward over NCHW observations (conv1: C = 3, 8x8, stride 4): the contiguous runs of a patch are only kw floats,
//                     too short for swizzled TMA rows, so the k-block (channel c, 32/kw filter rows) stages the RAW image rows
//                     y = s*oy + ky as boxes {W floats, 32/kw rows}; the window expansion happens on the way to registers:
//                     the thread of pixel ox reads floats [s*ox, s*ox + kw) of each row -- consecutive lanes read consecutive
//                     16-byte words (s = 4), conflict-free, and every input element crosses L2 -> SM once, not kw/s times.
//   mode 6, forward NHWC with a TILE-RESIDENT raw input: the input rows a tile of output rows needs are staged ONCE per tile
//                     (plain NHWC boxes {32 channels, Win pixels} per image row and channel plane, 128B-swizzled lines) and every
//                     k-block (filter tap x channel chunk) reads its shifted window of the same stage: pixel ox of tap (ky, kx)
//                     is line (row0 + ky) * Win + s*ox + kx.  TMA moves each input element once per tile instead of once per tap
//                     (4x fewer bytes for the 4x4 stride-2 layer): the implicit convolutions are bound by TMA ingest of the
//                     expanded tile stream (profiles/r1c_tca_ablation.md).
struct ConvA {
    int mode;                // 0 = plain GEMM
    int segw, nseg;          // pixels per segment, segments per tile (nseg * segw <= 128)
    int nseg_class;          // segments per class
    int ntile_class;         // tiles per class
    int rows_per_img;        // forward: OH; dgrad: Hq = Hin / s
    int s;                   // convolution stride
    int kpk;                 // forward: k-blocks per filter row; dgrad: k-blocks (32-channel chunks) per tap
    int T;                   // dgrad: taps per dimension per class (k / s)
    int KW, Cin;             // dgrad: filter width, input channels
    int Hin, Win;            // dgrad: dx height / width
    int nkb;                 // k-blocks per tile
    int spr;                 // wgrad: segments per output row (OW / segw)
    int cpr;                 // wgrad: chunks per filter row (NHWC) -- unused for NCHW
    int nchunks;             // wgrad: K / 64
    int nchw;                // wgrad: 5-D NCHW view (conv1) instead of the 4-D NHWC view
    int raw;                 // wgrad NCHW: stage raw image rows with 1-D bulk copies, windows expanded on the register read
    int KH, planes, nrows_max;   // mode 6: filter height, C / 32 channel planes, input rows a stage can hold per plane
    uint32_t tile_stage_bytes;   // mode 6: bytes of one tile-resident A stage (two of them)
    int C;                   // wgrad NCHW: channels
    int total_kb, kbps;      // wgrad: k-blocks in total / per split
    int total_seg;           // wgrad: B * OH * spr
};

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
    const float* a_ptr;      // A tensor base (1-D bulk copies of the raw-row wgrad)
    uint32_t a_region_bytes; // bytes of the A region of shared memory (0 = kSA stages of 16 KB)
    int b_resident;          // convolutions: the B (weight) tiles of ALL k-blocks stay in shared memory; reloaded only when the tile class changes
    int tma_store;           // epilogue: stage the finished tile in shared memory (128B-swizzled rows) and write it with ONE TMA store per
                             // 32-column block instead of row-per-thread global stores (forward convolutions, plain GEMMs without split-K)
    uint32_t stage_off;      // byte offset of that staging tile in dynamic shared memory (1024-aligned)
    ConvA conv;
};

extern int g_conv_resident, g_b_resident, g_max_ctas;

struct Plan { int bn, num_m, num_n, splits; int64_t kk_per_split; size_t ws; };

Plan make_plan(int64_t I, int64_t J, int64_t KK, size_t ws_avail, bool limit);
int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                CUtensorMapSwizzle swz);
// Enqueues the GEMM kernel only; with plan.splits > 1 the raw partial products are in `workspace` and the caller
// reduces them (tc_splitk_reduce_kernel).  b_lo != NULL: B is pre-split (g->B holds the hi parts, b_lo the residuals).
int launch(const ppd_gemm_args* g, int transpose_out, const float* b_lo, void* workspace, size_t workspace_bytes,
           cudaStream_t s, Plan* plan_out);

// hi = TF32(x), lo = TF32(x - hi) with the kernel's rounding; n a multiple of 4, pointers 16-byte aligned.
// Implicit-GEMM NHWC convolution forward / dgrad on the same kernel (see ConvA).
int conv_forward(const float* x, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* bias, int relu,
                 float* out, cudaStream_t s);
int conv_dgrad(const float* dy, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* act_mask,
               float* dx, cudaStream_t s);
int conv_forward_nchw(const float* x, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* bias,
                      int relu, float* out, cudaStream_t s);
// dW[Cout, K] (+)= dY^T patches(x): x NHWC (nchw = 0, patch order (ky,kx,c)) or NCHW (nchw = 1, patch order (c,ky,kx)).
size_t conv_wgrad_workspace(const ppd_conv_geom* g, int Cout);
int conv_wgrad(const float* x, const ppd_conv_geom* g, int nchw, const float* dy, int Cout, float* dW, int accumulate,
               void* workspace, size_t workspace_bytes, cudaStream_t s, int* splits_out);
int split_operand(const float* x, float* hi, float* lo, int64_t n, cudaStream_t s);

}  // namespace tca
}  // namespace ppd
