// fp32 SIMT GEMM family used by the parity ("fp32") mode of the policy network
// (conv-as-GEMM, FC, GRU input projection, heads; forward, dgrad and wgrad).
//
//   C[i, j] (+)= sum_kk OpA(i, kk) * OpB(j, kk)        i < I, j < J, kk < KK
//   OpA(i,kk) = a_kmajor ? A[i*lda + kk] : A[kk*lda + i]
//   OpB(j,kk) = b_kmajor ? B[j*ldb + kk] : B[kk*ldb + j]
// so with row-major matrices:
//   forward  Y[M,N] = X[M,K] W[N,K]^T   -> A=X (k-major), B=W (k-major)
//   dgrad    dX[M,K] = dY[M,N] W[N,K]   -> A=dY (k-major), B=W (j-major), contraction over N
//   wgrad    dW[N,K] = dY[M,N]^T X[M,K] -> A=dY (i-major), B=X (j-major), contraction over M
// Epilogue: + bias[j], ReLU, multiply by (mask[i,j] > 0) (ReLU backward), accumulate into C.
// A long contraction (wgrad: KK = rows of the minibatch) is split over gridDim.z; the partial
// tiles go to the workspace and are summed in a fixed order by a second kernel, so results are
// deterministic (no float atomics).
//
// Tiling: BI x BJ output tile per CTA of 256 threads, BK = 16, register micro-tile TI x TJ,
// shared-memory tiles stored kk-major ([BK][BI+4]) so the inner product reads float4 rows;
// global loads are 16-byte vectors whenever the operand's leading dimension allows it, with the
// next tile prefetched into registers while the current one is multiplied.
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int BK = 16;

struct Args {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    int64_t I, J, KK;
    const float* bias;
    const float* mask; int64_t ldm;
    int relu, accumulate;
    int vecA, vecB;          // 16-byte global loads allowed
    int64_t kk_per_split;    // multiple of BK
    float* partial;          // != nullptr -> write raw partial tiles [z][I][J]
};

// Load one BROWS x BK operand tile into registers.  KMAJ: contraction index contiguous in memory.
template <int BROWS, bool KMAJ>
struct TileLoader {
    static constexpr int kTileVecs = BROWS * BK / 4;                       // float4 per tile
    static constexpr int kVecs = (kTileVecs + kThreads - 1) / kThreads;    // float4 per thread (1 or 2)
    float v[4 * kVecs];

    __device__ __forceinline__ void load(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t nrows,
                                         int64_t kk0, int64_t kk_end, bool vec) {
        const int tid = threadIdx.x;
#pragma unroll
        for (int q = 0; q < kVecs; ++q) {
            const int e = tid + q * kThreads;          // vector index within the tile
            if (kTileVecs % kThreads != 0 && e >= kTileVecs) break;
            if (KMAJ) {
                const int r = e / (BK / 4), c = (e % (BK / 4)) * 4;
                const int64_t row = row0 + r, kk = kk0 + c;
                const float* p = P + row * ld + kk;
                if (vec && row < nrows && kk + 3 < kk_end) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[4 * q + u] = (row < nrows && kk + u < kk_end) ? __ldg(p + u) : 0.f;
                }
            } else {
                const int c = e / (BROWS / 4), r = (e % (BROWS / 4)) * 4;
                const int64_t row = row0 + r, kk = kk0 + c;
                const float* p = P + kk * ld + row;
                if (vec && kk < kk_end && row + 3 < nrows) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[4 * q + u] = (kk < kk_end && row + u < nrows) ? __ldg(p + u) : 0.f;
                }
            }
        }
    }

    // smem tile layout: S[kk][row], row stride BROWS + 4
    __device__ __forceinline__ void store(float* __restrict__ S) const {
        constexpr int LD = BROWS + 4;
        const int tid = threadIdx.x;
#pragma unroll
        for (int q = 0; q < kVecs; ++q) {
            const int e = tid + q * kThreads;
            if (kTileVecs % kThreads != 0 && e >= kTileVecs) break;
            if (KMAJ) {
                const int r = e / (BK / 4), c = (e % (BK / 4)) * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u) S[(c + u) * LD + r] = v[4 * q + u];
            } else {
                const int c = e / (BROWS / 4), r = (e % (BROWS / 4)) * 4;
                *reinterpret_cast<float4*>(S + c * LD + r) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    }
};

template <int BI, int BJ, int TI, int TJ, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(kThreads) sgemm_kernel(const Args a) {
    static_assert((BI / TI) * (BJ / TJ) == kThreads, "thread grid must cover the tile");
    static_assert(TI % 4 == 0 && (TJ % 4 == 0 || TJ == 2), "micro-tile");
    constexpr int LDA = BI + 4, LDB = BJ + 4;
    __shared__ __align__(16) float As[2][BK * LDA];
    __shared__ __align__(16) float Bs[2][BK * LDB];

    const int tj = threadIdx.x % (BJ / TJ);
    const int ti = threadIdx.x / (BJ / TJ);
    const int64_t i0 = (int64_t)blockIdx.y * BI;
    const int64_t j0 = (int64_t)blockIdx.x * BJ;
    const int64_t kk_begin = (int64_t)blockIdx.z * a.kk_per_split;
    const int64_t kk_end = min(a.KK, kk_begin + a.kk_per_split);

    float acc[TI][TJ];
#pragma unroll
    for (int x = 0; x < TI; ++x)
#pragma unroll
        for (int y = 0; y < TJ; ++y) acc[x][y] = 0.f;

    TileLoader<BI, A_KMAJ> la;
    TileLoader<BJ, B_KMAJ> lb;
    const int64_t ntiles = (kk_end - kk_begin + BK - 1) / BK;
    if (ntiles > 0) {
        la.load(a.A, a.lda, i0, a.I, kk_begin, kk_end, a.vecA);
        lb.load(a.B, a.ldb, j0, a.J, kk_begin, kk_end, a.vecB);
        la.store(As[0]);
        lb.store(Bs[0]);
    }
    __syncthreads();
    for (int64_t t = 0; t < ntiles; ++t) {
        const int cur = t & 1;
        if (t + 1 < ntiles) {
            la.load(a.A, a.lda, i0, a.I, kk_begin + (t + 1) * BK, kk_end, a.vecA);
            lb.load(a.B, a.ldb, j0, a.J, kk_begin + (t + 1) * BK, kk_end, a.vecB);
        }
        const float* __restrict__ sa = As[cur];
        const float* __restrict__ sb = Bs[cur];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float ra[TI], rb[TJ];
#pragma unroll
            for (int x = 0; x < TI; x += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(sa + k * LDA + ti * TI + x);
                ra[x] = t4.x; ra[x + 1] = t4.y; ra[x + 2] = t4.z; ra[x + 3] = t4.w;
            }
            if (TJ == 2) {
                const float2 t2 = *reinterpret_cast<const float2*>(sb + k * LDB + tj * TJ);
                rb[0] = t2.x; rb[1] = t2.y;
            } else {
#pragma unroll
                for (int y = 0; y < TJ; y += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(sb + k * LDB + tj * TJ + y);
                    rb[y] = t4.x; rb[y + 1] = t4.y; rb[y + 2] = t4.z; rb[y + 3] = t4.w;
                }
            }
#pragma unroll
            for (int x = 0; x < TI; ++x)
#pragma unroll
                for (int y = 0; y < TJ; ++y) acc[x][y] = fmaf(ra[x], rb[y], acc[x][y]);
        }
        if (t + 1 < ntiles) {
            la.store(As[cur ^ 1]);
            lb.store(Bs[cur ^ 1]);
        }
        __syncthreads();
    }

    // ---- epilogue
    if (a.partial) {
        float* P = a.partial + (int64_t)blockIdx.z * a.I * a.J;
#pragma unroll
        for (int x = 0; x < TI; ++x) {
            const int64_t i = i0 + ti * TI + x;
            if (i >= a.I) continue;
#pragma unroll
            for (int y = 0; y < TJ; ++y) {
                const int64_t j = j0 + tj * TJ + y;
                if (j < a.J) P[i * a.J + j] = acc[x][y];
            }
        }
        return;
    }
#pragma unroll
    for (int x = 0; x < TI; ++x) {
        const int64_t i = i0 + ti * TI + x;
        if (i >= a.I) continue;
#pragma unroll
        for (int y = 0; y < TJ; ++y) {
            const int64_t j = j0 + tj * TJ + y;
            if (j >= a.J) continue;
            float v = acc[x][y];
            if (a.bias) v += __ldg(a.bias + j);
            if (a.relu) v = fmaxf(v, 0.f);
            if (a.mask) v = (__ldg(a.mask + i * a.ldm + j) > 0.f) ? v : 0.f;
            float* c = a.C + i * a.ldc + j;
            *c = a.accumulate ? (*c + v) : v;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t I, int64_t J, float* __restrict__ C,
                     int64_t ldc, const float* __restrict__ bias, const float* __restrict__ mask, int64_t ldm,
                     int relu, int accumulate) {
    const int64_t n = I * J;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < n; e += (int64_t)gridDim.x * kThreads) {
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += partial[(int64_t)z * n + e];
        const int64_t i = e / J, j = e - i * J;
        if (bias) v += __ldg(bias + j);
        if (relu) v = fmaxf(v, 0.f);
        if (mask) v = (__ldg(mask + i * ldm + j) > 0.f) ? v : 0.f;
        float* c = C + i * ldc + j;
        *c = accumulate ? (*c + v) : v;
    }
}

// Column sums: out[j] (+)= sum_i X[i*ld + j]   (bias gradients).  Two deterministic stages.
//   narrow matrices (256 % J == 0 and dense rows, the conv bias gradients: J = 32 / 64 with up to 819 200
//   rows): the matrix is one contiguous stream; a thread striding by 256 always lands on column
//   tid % J, so loads are fully coalesced and the per-column fold happens once per block in shared memory;
//   wide matrices: one thread per column over a slab of rows.
constexpr int kColRows = 64;     // rows per partial block (wide path)
constexpr int kNarrowElems = 256 * 256;  // elements per partial block (narrow path): few partials -- the final stage is a latency chain over them
__global__ void __launch_bounds__(kThreads)
colsum_partial_kernel(const float* __restrict__ X, int64_t ld, int64_t I, int64_t J, float* __restrict__ partial) {
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= J) return;
    const int64_t r0 = (int64_t)blockIdx.y * kColRows, r1 = min(I, r0 + kColRows);
    float s = 0.f;
    for (int64_t i = r0; i < r1; ++i) s += __ldg(X + i * ld + j);
    partial[(int64_t)blockIdx.y * J + j] = s;
}
__global__ void __launch_bounds__(kThreads)
colsum_narrow_partial_kernel(const float* __restrict__ X, int64_t n, int J, float* __restrict__ partial) {
    __shared__ float sm[kThreads];
    const int64_t e0 = (int64_t)blockIdx.x * kNarrowElems, e1 = min(n, e0 + kNarrowElems);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t e = e0 + threadIdx.x;
    for (; e + 3 * kThreads < e1; e += 4 * kThreads) {
        s0 += __ldg(X + e); s1 += __ldg(X + e + kThreads); s2 += __ldg(X + e + 2 * kThreads); s3 += __ldg(X + e + 3 * kThreads);
    }
    for (; e < e1; e += kThreads) s0 += __ldg(X + e);
    sm[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.x < J) {            // e0 is a multiple of 256, hence of J: thread t holds column t % J
        float s = 0.f;
        for (int t = threadIdx.x; t < kThreads; t += J) s += sm[t];
        partial[(int64_t)blockIdx.x * J + threadIdx.x] = s;
    }
}
__global__ void __launch_bounds__(kThreads)
colsum_final_kernel(const float* __restrict__ partial, int64_t nparts, int64_t J, float* __restrict__ out, int accumulate) {
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= J) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t p = 0;
    for (; p + 3 < nparts; p += 4) {
        s0 += partial[p * J + j]; s1 += partial[(p + 1) * J + j]; s2 += partial[(p + 2) * J + j]; s3 += partial[(p + 3) * J + j];
    }
    for (; p < nparts; ++p) s0 += partial[p * J + j];
    const float s = (s0 + s1) + (s2 + s3);
    out[j] = accumulate ? (out[j] + s) : s;
}

bool colsum_narrow(int64_t ld, int64_t J) { return ld == J && J <= kThreads && (kThreads % J) == 0; }
int64_t colsum_parts(int64_t ld, int64_t I, int64_t J) {
    return colsum_narrow(ld, J) ? (I * J + kNarrowElems - 1) / kNarrowElems : (I + kColRows - 1) / kColRows;
}

// ---- several column sums in two launches (all bias gradients of a minibatch at once)
constexpr int kMaxSegs = 8;
struct MultiSegs {
    const float* X[kMaxSegs]; float* out[kMaxSegs];
    int64_t ld[kMaxSegs], I[kMaxSegs], J[kMaxSegs], parts[kMaxSegs], part_off[kMaxSegs];   // part_off: floats into the workspace
    int block_off[kMaxSegs + 1], final_off[kMaxSegs + 1];
    int narrow[kMaxSegs], accumulate[kMaxSegs], gx[kMaxSegs];
    int n;
};

__global__ void __launch_bounds__(kThreads) colsum_multi_partial_kernel(const MultiSegs m, float* __restrict__ ws) {
    __shared__ float sm[kThreads];
    int s = 0;
    while (s + 1 < m.n && (int)blockIdx.x >= m.block_off[s + 1]) ++s;
    const int lb = blockIdx.x - m.block_off[s];
    const float* __restrict__ X = m.X[s];
    const int64_t J = m.J[s];
    float* partial = ws + m.part_off[s];
    if (m.narrow[s]) {
        const int64_t n = m.I[s] * J;
        const int64_t e0 = (int64_t)lb * kNarrowElems, e1 = min(n, e0 + kNarrowElems);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int64_t e = e0 + threadIdx.x;
        for (; e + 3 * kThreads < e1; e += 4 * kThreads) {
            s0 += __ldg(X + e); s1 += __ldg(X + e + kThreads); s2 += __ldg(X + e + 2 * kThreads); s3 += __ldg(X + e + 3 * kThreads);
        }
        for (; e < e1; e += kThreads) s0 += __ldg(X + e);
        sm[threadIdx.x] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (threadIdx.x < J) {
            float t = 0.f;
            for (int q = threadIdx.x; q < kThreads; q += (int)J) t += sm[q];
            partial[(int64_t)lb * J + threadIdx.x] = t;
        }
    } else {
        const int bx = lb % m.gx[s], by = lb / m.gx[s];
        const int64_t j = (int64_t)bx * kThreads + threadIdx.x;
        if (j >= J) return;
        const int64_t r0 = (int64_t)by * kColRows, r1 = min(m.I[s], r0 + kColRows);
        float t = 0.f;
        for (int64_t i = r0; i < r1; ++i) t += __ldg(X + i * m.ld[s] + j);
        partial[(int64_t)by * J + j] = t;
    }
}

__global__ void __launch_bounds__(kThreads) colsum_multi_final_kernel(const MultiSegs m, const float* __restrict__ ws) {
    // one block per 256 columns of a segment; when the segment is narrower the spare threads split the partials between them
    // (a 32-column bias over ~800k rows has thousands of partials: one thread per column walking them all took 180 us)
    __shared__ float sm[kThreads];
    int s = 0;
    while (s + 1 < m.n && (int)blockIdx.x >= m.final_off[s + 1]) ++s;
    const int64_t J = m.J[s];
    const int64_t jb = (int64_t)(blockIdx.x - m.final_off[s]) * kThreads;
    const int cols = (int)min((int64_t)kThreads, J - jb);
    const int G = kThreads / cols;
    const int g = threadIdx.x / cols, c = threadIdx.x - g * cols;
    const float* partial = ws + m.part_off[s] + jb + c;
    // four accumulators, eight loads in flight: the walk over the partials is a latency chain (L2 round trips), not bandwidth
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (g < G) {
        int64_t p = g;
        const int64_t P = m.parts[s];
#pragma unroll 2
        for (; p + 3 * G < P; p += 4 * G) {
            s0 += partial[p * J]; s1 += partial[(p + G) * J]; s2 += partial[(p + 2 * G) * J]; s3 += partial[(p + 3 * G) * J];
        }
        for (; p < P; p += G) s0 += partial[p * J];
    }
    sm[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0) {
        float t = 0.f;
        for (int q = 0; q < G; ++q) t += sm[q * cols + c];
        float* o = m.out[s] + jb + c;
        *o = m.accumulate[s] ? (*o + t) : t;
    }
}

// Very short contractions (the 9-wide policy / value heads' input gradient: dhs[B, 512] = dz[B, 9] . W[9, 512]): a tiled GEMM spends its
// time filling 128 x 128 tiles for nine multiply-adds per output (13.5 us for the 2048-row minibatch).  Here a thread owns four
// adjacent outputs: KK broadcast loads of its A row, KK coalesced 16-byte loads of B (L1-resident), 4 KK FMAs, one 16-byte store.
// A [I, KK] contraction-contiguous, B [KK, J] row-major; same epilogue as the tiled kernel (bias, ReLU mask, relu, accumulate).
constexpr int kSkinnyK = 16;
__global__ void __launch_bounds__(kThreads) sgemm_skinny_k_kernel(const Args a) {
    const int64_t J4 = a.J >> 2, total = a.I * J4;
    const int KK = (int)a.KK;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
        const int64_t i = e / J4, j = (e - i * J4) << 2;
        const float* arow = a.A + i * a.lda;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < KK; ++k) {
            const float av = __ldg(arow + k);
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.B + (int64_t)k * a.ldb + j));
            acc.x = fmaf(av, b.x, acc.x); acc.y = fmaf(av, b.y, acc.y); acc.z = fmaf(av, b.z, acc.z); acc.w = fmaf(av, b.w, acc.w);
        }
        if (a.bias) { acc.x += __ldg(a.bias + j); acc.y += __ldg(a.bias + j + 1); acc.z += __ldg(a.bias + j + 2); acc.w += __ldg(a.bias + j + 3); }
        if (a.mask) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(a.mask + i * a.ldm + j));
            acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f; acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
        }
        if (a.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        float4* c = reinterpret_cast<float4*>(a.C + i * a.ldc + j);
        if (a.accumulate) { const float4 o = *c; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
        *c = acc;
    }
}

struct Plan { int bi, bj; int64_t gx, gy; int splits; int64_t kk_per_split; size_t ws; };

Plan make_plan(int64_t I, int64_t J, int64_t KK, size_t ws_bytes_avail, bool have_ws_limit) {
    Plan p;
    if (I >= 1024 && J >= 128) { p.bi = 128; p.bj = 128; }
    else if (J <= 32) { p.bi = 64; p.bj = 32; }
    else { p.bi = 64; p.bj = 64; }
    p.gx = (J + p.bj - 1) / p.bj;
    p.gy = (I + p.bi - 1) / p.bi;
    const int64_t tiles = p.gx * p.gy;
    int64_t splits = 1;
    const int64_t target = 2 * ppd::kNumSMs;
    if (tiles < target && KK >= 8 * BK) {
        splits = (target + tiles - 1) / tiles;
        const int64_t max_splits = KK / (4 * BK);
        if (splits > max_splits) splits = max_splits;
        if (splits > 512) splits = 512;
        if (splits < 1) splits = 1;
    }
    if (have_ws_limit) {
        while (splits > 1 && (size_t)splits * I * J * sizeof(float) > ws_bytes_avail) --splits;
    }
    int64_t per = (KK + splits - 1) / splits;
    per = (per + BK - 1) / BK * BK;
    splits = (KK + per - 1) / per;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;
    p.kk_per_split = per;
    p.ws = splits > 1 ? (size_t)splits * I * J * sizeof(float) : 0;
    return p;
}

template <int BI, int BJ, int TI, int TJ>
void launch_tile(const Args& a, dim3 grid, cudaStream_t s, bool ak, bool bk) {
    if (ak && bk)        sgemm_kernel<BI, BJ, TI, TJ, true, true><<<grid, kThreads, 0, s>>>(a);
    else if (ak && !bk)  sgemm_kernel<BI, BJ, TI, TJ, true, false><<<grid, kThreads, 0, s>>>(a);
    else if (!ak && bk)  sgemm_kernel<BI, BJ, TI, TJ, false, true><<<grid, kThreads, 0, s>>>(a);
    else                 sgemm_kernel<BI, BJ, TI, TJ, false, false><<<grid, kThreads, 0, s>>>(a);
}

}  // namespace

extern "C" size_t ppd_sgemm_workspace(int64_t I, int64_t J, int64_t KK) {
    if (I <= 0 || J <= 0 || KK <= 0) return 0;
    return make_plan(I, J, KK, 0, false).ws;
}

extern "C" int ppd_sgemm(const ppd_gemm_args* g, void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(g && g->A && g->B && g->C, "null pointer");
    PPD_REQUIRE(g->I > 0 && g->J > 0 && g->KK > 0, "sizes must be positive");
    PPD_REQUIRE(g->lda > 0 && g->ldb > 0 && g->ldc >= g->J, "bad leading dimension");
    PPD_REQUIRE(!g->mask || g->ldm >= g->J, "bad mask leading dimension");
    const Plan p = make_plan(g->I, g->J, g->KK, workspace ? workspace_bytes : 0, true);
    PPD_REQUIRE(p.gy <= 65535 && p.gx <= 0x7fffffffLL, "grid too large");
    Args a;
    a.A = g->A; a.lda = g->lda; a.B = g->B; a.ldb = g->ldb; a.C = g->C; a.ldc = g->ldc;
    a.I = g->I; a.J = g->J; a.KK = g->KK;
    a.bias = g->bias; a.mask = g->mask; a.ldm = g->ldm; a.relu = g->relu; a.accumulate = g->accumulate;
    a.vecA = (g->lda % 4 == 0) && ((uintptr_t)g->A % 16 == 0);
    a.vecB = (g->ldb % 4 == 0) && ((uintptr_t)g->B % 16 == 0);
    a.kk_per_split = p.kk_per_split;
    a.partial = p.splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
    cudaStream_t s = ppd::as_stream(stream);
    dim3 grid((unsigned)p.gx, (unsigned)p.gy, (unsigned)p.splits);
    const bool ak = g->a_kmajor != 0, bk = g->b_kmajor != 0;
    if (ak && !bk && g->KK <= kSkinnyK && g->J % 4 == 0 && g->ldb % 4 == 0 && g->ldc % 4 == 0 &&
        !(((uintptr_t)g->B | (uintptr_t)g->C) & 15) && (!g->mask || (g->ldm % 4 == 0 && !((uintptr_t)g->mask & 15)))) {
        int64_t nb = (g->I * (g->J / 4) + kThreads - 1) / kThreads;
        if (nb > 16 * ppd::kNumSMs) nb = 16 * ppd::kNumSMs;
        sgemm_skinny_k_kernel<<<(unsigned)nb, kThreads, 0, s>>>(a);
        return ppd::launch_status("sgemm_skinny_k_kernel");
    }
    if (p.bi == 128)      launch_tile<128, 128, 8, 8>(a, grid, s, ak, bk);
    else if (p.bj == 32)  launch_tile<64, 32, 4, 2>(a, grid, s, ak, bk);
    else                  launch_tile<64, 64, 4, 4>(a, grid, s, ak, bk);
    int rc = ppd::launch_status("sgemm_kernel");
    if (rc || p.splits == 1) return rc;
    int64_t nb = (g->I * g->J + kThreads - 1) / kThreads;
    if (nb > 4 * ppd::kNumSMs) nb = 4 * ppd::kNumSMs;
    splitk_reduce_kernel<<<(unsigned)nb, kThreads, 0, s>>>(a.partial, p.splits, g->I, g->J, g->C, g->ldc, g->bias,
                                                           g->mask, g->ldm, g->relu, g->accumulate);
    return ppd::launch_status("splitk_reduce_kernel");
}

extern "C" size_t ppd_colsum_workspace(int64_t I, int64_t J) {
    if (I <= 0 || J <= 0) return 0;
    const int64_t a = (I * J + kNarrowElems - 1) / kNarrowElems, b = (I + kColRows - 1) / kColRows;
    return (size_t)(a > b ? a : b) * J * sizeof(float);
}

extern "C" int ppd_colsum(const float* X, int64_t ld, int64_t I, int64_t J, float* out, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(X && out && workspace, "null pointer");
    PPD_REQUIRE(I > 0 && J > 0 && ld >= J, "bad sizes");
    const int64_t parts = colsum_parts(ld, I, J);
    if (workspace_bytes < (size_t)parts * J * sizeof(float)) {
        ppd::set_error("ppd_colsum: workspace too small");
        return PPD_EWORKSPACE;
    }
    cudaStream_t s = ppd::as_stream(stream);
    const unsigned gx = (unsigned)((J + kThreads - 1) / kThreads);
    if (colsum_narrow(ld, J)) {
        PPD_REQUIRE(parts <= 0x7fffffffLL, "too many rows");
        colsum_narrow_partial_kernel<<<(unsigned)parts, kThreads, 0, s>>>(X, I * J, (int)J, (float*)workspace);
    } else {
        PPD_REQUIRE(parts <= 65535, "too many rows");
        colsum_partial_kernel<<<dim3(gx, (unsigned)parts), kThreads, 0, s>>>(X, ld, I, J, (float*)workspace);
    }
    int rc = ppd::launch_status("colsum_partial_kernel");
    if (rc) return rc;
    colsum_final_kernel<<<gx, kThreads, 0, s>>>((const float*)workspace, parts, J, out, accumulate);
    return ppd::launch_status("colsum_final_kernel");
}

extern "C" size_t ppd_colsum_multi_workspace(const ppd_colsum_seg* segs, int n) {
    size_t tot = 0;
    for (int s = 0; segs && s < n; ++s) tot += ppd_colsum_workspace(segs[s].I, segs[s].J);
    return tot;
}

extern "C" int ppd_colsum_multi(const ppd_colsum_seg* segs, int n, void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(segs && workspace, "null pointer");
    PPD_REQUIRE(n >= 1 && n <= kMaxSegs, "between 1 and 8 segments");
    MultiSegs m;
    m.n = n;
    int64_t off = 0;
    m.block_off[0] = 0;
    m.final_off[0] = 0;
    for (int s = 0; s < n; ++s) {
        const ppd_colsum_seg& g = segs[s];
        PPD_REQUIRE(g.X && g.out && g.I > 0 && g.J > 0 && g.ld >= g.J, "bad segment");
        m.X[s] = g.X; m.out[s] = g.out; m.ld[s] = g.ld; m.I[s] = g.I; m.J[s] = g.J; m.accumulate[s] = g.accumulate;
        m.narrow[s] = colsum_narrow(g.ld, g.J) ? 1 : 0;
        m.parts[s] = colsum_parts(g.ld, g.I, g.J);
        m.gx[s] = (int)((g.J + kThreads - 1) / kThreads);
        m.part_off[s] = off;
        off += m.parts[s] * g.J;
        const int64_t blocks = m.narrow[s] ? m.parts[s] : m.parts[s] * m.gx[s];
        PPD_REQUIRE(m.block_off[s] + blocks <= 0x7fffffffLL, "too many blocks");
        m.block_off[s + 1] = m.block_off[s] + (int)blocks;
        m.final_off[s + 1] = m.final_off[s] + m.gx[s];
    }
    if (workspace_bytes < (size_t)off * sizeof(float)) {
        ppd::set_error("ppd_colsum_multi: workspace too small");
        return PPD_EWORKSPACE;
    }
    cudaStream_t st = ppd::as_stream(stream);
    colsum_multi_partial_kernel<<<m.block_off[n], kThreads, 0, st>>>(m, (float*)workspace);
    int rc = ppd::launch_status("colsum_multi_partial_kernel");
    if (rc) return rc;
    colsum_multi_final_kernel<<<m.final_off[n], kThreads, 0, st>>>(m, (const float*)workspace);
    return ppd::launch_status("colsum_multi_final_kernel");
}
