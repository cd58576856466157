// GRU with in-kernel mask reset: h_t = GRU(x_t, h_{t-1} * m_t)   (NNBase._forward_gru,
// PKG/model.py:111-166; nn.GRU gate order r, z, n).  The input projection gi = x W_ih^T + b_ih for
// all T steps is one big GEMM done by the caller; these kernels are the strictly sequential part.
//
// Persistent cooperative kernels: one launch runs all T steps, with one grid-wide barrier per
// step, instead of the reference's host-side segmentation (a .cpu() sync per minibatch,
// model.py:129-133) and per-segment cuDNN calls.
//   CTA (bx, by) owns hidden units [bx*HU, bx*HU+HU) for envs [by*ET, by*ET+ET).
//   forward : the 3*HU rows of W_hh it needs stay in shared memory for the whole sequence; per step
//             it reads the masked previous state of its envs (E*H floats, L2-resident), forms the
//             3*HU*ET dot products warp-wise, applies the gates and writes h_t for its units.
//   backward: the HU columns of W_hh it needs stay in shared memory; per step (a) gate backward for
//             its units -> dgi[t], dgh_n[t]; grid barrier; (b) carry dh_{t-1} for its units =
//             (dgh[t] . W_hh[:, units] + dh*z) * m_t.  Weight gradients are GEMMs over all T
//             afterwards (caller).
#include <cooperative_groups.h>

#include "ppd_common.cuh"

namespace cg = cooperative_groups;

namespace ppd {
// gru_cluster.cu: cluster/DSMEM path for small E; returns -1 when it does not apply
int gru_forward_cluster(const float* gi, const float* h0, const float* masks, const float* w_hh, const float* b_hh,
                        int T, int E, int H, float* hs, float* h_last, float* sr, float* sz, float* sn, float* sghn,
                        cudaStream_t s);
int gru_backward_cluster(const float* dhs, const float* masks, const float* w_hh, const float* h0, const float* hs,
                         const float* sr, const float* sz, const float* sn, const float* sghn, int T, int E, int H,
                         float* dgi, float* dghn, float* dh0, cudaStream_t s);
extern int g_reg_kernels;
extern int g_multi_min_e;
extern int g_multi_clusters;
static int g_gru_mode = 0;   // 0 = auto (cluster path when it applies), 1 = always the grid-cooperative kernels
}  // namespace ppd

namespace {

constexpr int kThreads = 384;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxItems = 8;     // (row|unit, env-quad) work items per warp, backward

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct FwdArgs {
    const float* gi; const float* h0; const float* masks; const float* w_hh; const float* b_hh;
    float* hs; float* h_last; float* sr; float* sz; float* sn; float* sghn;
    int T, E, H, HU, ET;
};

__global__ void __launch_bounds__(kThreads) gru_fwd_kernel(const FwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HU = a.HU, ET = a.ET, E = a.E;
    float* Wr = smem;                    // [3*HU][H]
    float* hb = Wr + 3 * HU * H;         // [ET][H]   masked previous state (rows >= ne stay zero)
    float* gh = hb + ET * H;             // [3*HU][ET]
    const int j0 = blockIdx.x * HU, e0 = blockIdx.y * ET;
    const int nu = min(HU, H - j0), ne = min(ET, E - e0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < 3 * HU * H; idx += kThreads) {
        const int r = idx / H, k = idx - r * H;
        const int g = r / HU, u = r - g * HU;
        Wr[idx] = (u < nu) ? __ldg(a.w_hh + (size_t)(g * H + j0 + u) * H + k) : 0.f;
    }
    for (int idx = tid; idx < ET * H; idx += kThreads) hb[idx] = 0.f;
    __syncthreads();

    const int nq = (ne + 3) >> 2;
    const int items = 3 * HU * nq;
    for (int t = 0; t < a.T; ++t) {
        // ---- masked previous hidden state of this CTA's envs (written by other CTAs last step)
        for (int idx = tid; idx < ne * H; idx += kThreads) {
            const int e = idx / H, k = idx - e * H;
            const float m = __ldg(a.masks + (size_t)t * E + e0 + e);
            const float hp = (t == 0) ? __ldg(a.h0 + (size_t)(e0 + e) * H + k)
                                      : __ldcg(a.hs + ((size_t)(t - 1) * E + e0 + e) * H + k);
            hb[idx] = hp * m;
        }
        __syncthreads();
        // ---- gh[row][e] = W_hh[row,:] . hm[e,:]
        for (int it = warp; it < items; it += kWarps) {
            const int r = it / nq, q = it - r * nq;
            const float* wrow = Wr + r * H;
            const float* h4 = hb + (q * 4) * H;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int k = lane; k < H; k += 32) {
                const float w = wrow[k];
                a0 = fmaf(w, h4[k], a0);
                a1 = fmaf(w, h4[H + k], a1);
                a2 = fmaf(w, h4[2 * H + k], a2);
                a3 = fmaf(w, h4[3 * H + k], a3);
            }
            a0 = ppd::warp_sum(a0); a1 = ppd::warp_sum(a1); a2 = ppd::warp_sum(a2); a3 = ppd::warp_sum(a3);
            if (lane == 0) {
                float* o = gh + r * ET + q * 4;
                o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
            }
        }
        __syncthreads();
        // ---- gates for (unit u, env e)
        for (int idx = tid; idx < nu * ne; idx += kThreads) {
            const int e = idx / nu, u = idx - e * nu;
            const int j = j0 + u;
            const size_t row = (size_t)t * E + e0 + e;
            const float* gir = a.gi + row * 3 * H;
            const float ghr = gh[(0 * HU + u) * ET + e] + __ldg(a.b_hh + j);
            const float ghz = gh[(1 * HU + u) * ET + e] + __ldg(a.b_hh + H + j);
            const float ghn = gh[(2 * HU + u) * ET + e] + __ldg(a.b_hh + 2 * H + j);
            const float r = sigmoidf_(__ldg(gir + j) + ghr);
            const float z = sigmoidf_(__ldg(gir + H + j) + ghz);
            const float n = tanhf(__ldg(gir + 2 * H + j) + r * ghn);
            const float hm = hb[e * H + j];
            const float hn = n + z * (hm - n);
            a.hs[row * H + j] = hn;
            if (a.sr) { a.sr[row * H + j] = r; a.sz[row * H + j] = z; a.sn[row * H + j] = n; a.sghn[row * H + j] = ghn; }
            if (a.h_last && t == a.T - 1) a.h_last[(size_t)(e0 + e) * H + j] = hn;
        }
        if (t + 1 < a.T) grid.sync();
    }
}

struct BwdArgs {
    const float* dhs; const float* masks; const float* w_hh; const float* h0; const float* hs;
    const float* sr; const float* sz; const float* sn; const float* sghn;
    float* dgi; float* dghn; float* dh0;
    int T, E, H, HU, ET, JC;
};

__global__ void __launch_bounds__(kThreads) gru_bwd_kernel(const BwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HU = a.HU, ET = a.ET, E = a.E, JC = a.JC, H3 = 3 * a.H;
    float* WT = smem;                    // [HU][3H]   WT[u][j] = W_hh[j][j0+u]
    float* chunk = WT + HU * H3;         // [ET][JC]   slice of dgh[t] for this CTA's envs (rows >= ne zero)
    float* carry = chunk + ET * JC;      // [ET][HU]   dh flowing into step t from step t+1
    float* dhz = carry + ET * HU;        // [ET][HU]
    const int j0 = blockIdx.x * HU, e0 = blockIdx.y * ET;
    const int nu = min(HU, H - j0), ne = min(ET, E - e0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < HU * H3; idx += kThreads) {
        const int u = idx / H3, j = idx - u * H3;
        WT[idx] = (u < nu) ? __ldg(a.w_hh + (size_t)j * H + j0 + u) : 0.f;
    }
    for (int idx = tid; idx < ET * JC; idx += kThreads) chunk[idx] = 0.f;
    for (int idx = tid; idx < ET * HU; idx += kThreads) { carry[idx] = 0.f; dhz[idx] = 0.f; }
    __syncthreads();

    const int nq = (ne + 3) >> 2;
    const int items = HU * nq;           // launcher guarantees items <= kMaxItems * kWarps
    for (int t = a.T - 1; t >= 0; --t) {
        // ---- (a) gate backward for this CTA's (unit, env) pairs
        for (int idx = tid; idx < nu * ne; idx += kThreads) {
            const int e = idx / nu, u = idx - e * nu;
            const int j = j0 + u;
            const size_t row = (size_t)t * E + e0 + e;
            const float dh = __ldg(a.dhs + row * H + j) + carry[e * HU + u];
            const float r = __ldg(a.sr + row * H + j), z = __ldg(a.sz + row * H + j);
            const float n = __ldg(a.sn + row * H + j), ghn = __ldg(a.sghn + row * H + j);
            const float m = __ldg(a.masks + row);
            const float hp = (t == 0) ? __ldg(a.h0 + (size_t)(e0 + e) * H + j)
                                      : __ldg(a.hs + ((size_t)(t - 1) * E + e0 + e) * H + j);
            const float hm = hp * m;
            const float dz = dh * (hm - n);
            const float dn = dh * (1.f - z);
            const float dpn = dn * (1.f - n * n);
            const float dpz = dz * z * (1.f - z);
            const float dpr = (dpn * ghn) * r * (1.f - r);
            float* g = a.dgi + row * H3;
            g[j] = dpr; g[H + j] = dpz; g[2 * H + j] = dpn;
            a.dghn[row * H + j] = dpn * r;
            dhz[e * HU + u] = dh * z;
        }
        grid.sync();      // every unit's dgh[t] is visible
        // ---- (b) carry[e][u] = (sum_j dgh[t][e][j] * W_hh[j][j0+u] + dh*z) * m_t[e]
        float acc[kMaxItems][4];
#pragma unroll
        for (int i = 0; i < kMaxItems; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        for (int c0 = 0; c0 < H3; c0 += JC) {
            const int jc = min(JC, H3 - c0);
            __syncthreads();
            for (int idx = tid; idx < ne * jc; idx += kThreads) {
                const int e = idx / jc, jj = idx - e * jc;
                const int j = c0 + jj;
                const size_t row = (size_t)t * E + e0 + e;
                chunk[e * JC + jj] = (j < 2 * H) ? __ldcg(a.dgi + row * H3 + j) : __ldcg(a.dghn + row * H + (j - 2 * H));
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kMaxItems; ++i) {
                const int it = warp + i * kWarps;
                if (it < items) {
                    const int u = it / nq, q = it - u * nq;
                    const float* wcol = WT + u * H3 + c0;
                    const float* d4 = chunk + (q * 4) * JC;
                    for (int jj = lane; jj < jc; jj += 32) {
                        const float w = wcol[jj];
                        acc[i][0] = fmaf(w, d4[jj], acc[i][0]);
                        acc[i][1] = fmaf(w, d4[JC + jj], acc[i][1]);
                        acc[i][2] = fmaf(w, d4[2 * JC + jj], acc[i][2]);
                        acc[i][3] = fmaf(w, d4[3 * JC + jj], acc[i][3]);
                    }
                }
            }
        }
        __syncthreads();      // all reads of carry/dhz from phase (a) and chunk are done
#pragma unroll
        for (int i = 0; i < kMaxItems; ++i) {
            const int it = warp + i * kWarps;
            if (it < items) {
                const int u = it / nq, q = it - u * nq;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float s = ppd::warp_sum(acc[i][x]);
                    const int e = q * 4 + x;
                    if (lane == 0 && e < ne && u < nu) {
                        const float m = __ldg(a.masks + (size_t)t * E + e0 + e);
                        const float v = (s + dhz[e * HU + u]) * m;
                        carry[e * HU + u] = v;
                        if (t == 0 && a.dh0) a.dh0[(size_t)(e0 + e) * H + j0 + u] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
masked_prev_kernel(const float* __restrict__ hs, const float* __restrict__ h0, const float* __restrict__ masks,
                   int T, int E, int H, float* __restrict__ hm) {
    const int64_t total = (int64_t)T * E * H;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int64_t row = idx / H;
        const int k = (int)(idx - row * H);
        const int64_t t = row / E, e = row - t * E;
        const float hp = (t == 0) ? __ldg(h0 + e * H + k) : __ldg(hs + (row - E) * H + k);
        hm[idx] = hp * __ldg(masks + row);
    }
}

struct Cfg { int HU, ET, JC; size_t smem; dim3 grid; };

size_t fwd_smem(int H, int HU, int ET) { return (size_t)(3 * HU * H + ET * H + 3 * HU * ET) * sizeof(float); }
size_t bwd_smem(int H, int HU, int ET, int JC) { return (size_t)(HU * 3 * H + ET * JC + 2 * ET * HU) * sizeof(float); }

// Pick the smallest tiles (most CTAs) that still fit co-resident on the device: a cooperative
// launch needs every CTA resident at once.
template <typename K>
int pick(K kernel, bool fwd, int E, int H, Cfg* out) {
    static const int cand[][2] = {{4, 4}, {4, 8}, {8, 8}, {8, 16}, {8, 32}, {16, 32}, {16, 64}, {32, 64}, {32, 128}};
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = ppd::kNumSMs;
    for (const auto& c : cand) {
        int HU = c[0], ET = c[1];
        if (HU > H) HU = H;
        if (ET >= 4 * ((E + 3) / 4) && ET > 4) {      // do not pad envs needlessly
            const int need = 4 * ((E + 3) / 4);
            if (ET / 2 >= need) continue;
        }
        const int JC = 256;
        const size_t smem = fwd ? fwd_smem(H, HU, ET) : bwd_smem(H, HU, ET, JC);
        if (smem > 227 * 1024) continue;
        if (!fwd && HU * (ET / 4) > kMaxItems * kWarps) continue;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const int gx = (H + HU - 1) / HU, gy = (E + ET - 1) / ET;
        if ((int64_t)gx * gy <= (int64_t)per_sm * sms) {
            out->HU = HU; out->ET = ET; out->JC = JC; out->smem = smem; out->grid = dim3(gx, gy);
            return 0;
        }
    }
    return -1;
}

}  // namespace

extern "C" int ppd_gru_forward(const float* gi, const float* h0, const float* masks, const float* w_hh,
                               const float* b_hh, int T, int E, int H, float* hs, float* h_last,
                               float* save_r, float* save_z, float* save_n, float* save_ghn, void* stream) {
    PPD_REQUIRE(gi && h0 && masks && w_hh && b_hh && hs, "null pointer");
    PPD_REQUIRE(T > 0 && E > 0 && H > 0, "sizes must be positive");
    PPD_REQUIRE((save_r != nullptr) == (save_z != nullptr) && (save_r != nullptr) == (save_n != nullptr) &&
                (save_r != nullptr) == (save_ghn != nullptr), "save buffers must be all set or all NULL");
    if (ppd::g_gru_mode == 0) {
        const int rc = ppd::gru_forward_cluster(gi, h0, masks, w_hh, b_hh, T, E, H, hs, h_last, save_r, save_z, save_n,
                                                save_ghn, ppd::as_stream(stream));
        if (rc >= 0) return rc;
    }
    Cfg c;
    if (pick(gru_fwd_kernel, true, E, H, &c)) {
        ppd::set_error("ppd_gru_forward: no co-resident tiling for E=%d H=%d", E, H);
        return PPD_EINVAL;
    }
    FwdArgs a{gi, h0, masks, w_hh, b_hh, hs, h_last, save_r, save_z, save_n, save_ghn, T, E, H, c.HU, c.ET};
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)gru_fwd_kernel, c.grid, dim3(kThreads), params, c.smem,
                                                ppd::as_stream(stream));
    if (e != cudaSuccess) {
        ppd::set_error("ppd_gru_forward: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return ppd::launch_status("gru_fwd_kernel");
}

extern "C" int ppd_gru_backward(const float* dhs, const float* masks, const float* w_hh, const float* h0,
                                const float* hs, const float* save_r, const float* save_z, const float* save_n,
                                const float* save_ghn, int T, int E, int H, float* dgi, float* dghn, float* dh0,
                                void* stream) {
    PPD_REQUIRE(dhs && masks && w_hh && h0 && hs && save_r && save_z && save_n && save_ghn && dgi && dghn,
                "null pointer");
    PPD_REQUIRE(T > 0 && E > 0 && H > 0, "sizes must be positive");
    if (ppd::g_gru_mode == 0) {
        const int rc = ppd::gru_backward_cluster(dhs, masks, w_hh, h0, hs, save_r, save_z, save_n, save_ghn, T, E, H, dgi,
                                                 dghn, dh0, ppd::as_stream(stream));
        if (rc >= 0) return rc;
    }
    Cfg c;
    if (pick(gru_bwd_kernel, false, E, H, &c)) {
        ppd::set_error("ppd_gru_backward: no co-resident tiling for E=%d H=%d", E, H);
        return PPD_EINVAL;
    }
    BwdArgs a{dhs, masks, w_hh, h0, hs, save_r, save_z, save_n, save_ghn, dgi, dghn, dh0, T, E, H, c.HU, c.ET, c.JC};
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)gru_bwd_kernel, c.grid, dim3(kThreads), params, c.smem,
                                                ppd::as_stream(stream));
    if (e != cudaSuccess) {
        ppd::set_error("ppd_gru_backward: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return ppd::launch_status("gru_bwd_kernel");
}

extern "C" void ppd_gru_set_mode(int mode) {
    if (mode >= 100) { ppd::g_multi_clusters = mode - 100; return; }
    ppd::g_gru_mode = (mode == 1) ? 1 : 0;
    ppd::g_reg_kernels = (mode == 2) ? 0 : 1;
    ppd::g_multi_min_e = (mode == 3) ? 1 : (mode == 4) ? (1 << 30) : 9;
}

extern "C" int ppd_gru_masked_prev(const float* hs, const float* h0, const float* masks, int T, int E, int H,
                                   float* hm, void* stream) {
    PPD_REQUIRE(hs && h0 && masks && hm, "null pointer");
    PPD_REQUIRE(T > 0 && E > 0 && H > 0, "sizes must be positive");
    int64_t nb = ((int64_t)T * E * H + 255) / 256;
    if (nb > 16 * ppd::kNumSMs) nb = 16 * ppd::kNumSMs;
    masked_prev_kernel<<<(unsigned)nb, 256, 0, ppd::as_stream(stream)>>>(hs, h0, masks, T, E, H, hm);
    return ppd::launch_status("masked_prev_kernel");
}
