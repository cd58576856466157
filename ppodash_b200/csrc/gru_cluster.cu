// GRU recurrence for small env counts (E <= 8 per minibatch: the PPO-Dash configuration has E = 4)
// on thread-block clusters with distributed shared memory.
//
// At E = 4 the recurrence is a chain of T = 512 tiny mat-vecs: latency, not FLOPs.  The
// grid-cooperative kernels in gru.cu pay one grid-wide barrier through L2 per timestep (~5 us).
// Here one cluster of 16 CTAs (16 SMs of one GPC) runs one env's whole sequence:
//   * CTA r keeps the 3*H/16 rows of W_hh that produce its H/16 hidden units resident in shared
//     memory (96 rows x 512 x 4 B = 192 KB at H = 512) for all T steps -- W_hh is read from HBM once;
//   * forward : per step each CTA forms its 3*HU dot products with the (masked) previous state,
//     applies the gates and writes its HU new state values straight into the shared memory of all
//     16 CTAs with st.async (data + mbarrier complete_tx in one DSMEM transaction), so each CTA only
//     waits on a local mbarrier for "all 512 state values of step t are here" -- no cluster barrier;
//   * backward: per step each CTA turns dh of its units into d(gates), multiplies by its W_hh rows
//     to get its partial sum of dh_{t-1} for ALL units and scatters the partials to their owner CTAs
//     through DSMEM (a reduce-scatter) with the same st.async + local mbarrier handshake.
// Envs are independent sequences, so E envs run as E clusters side by side (E*16 SMs busy).
// Global loads of the next step's operands are issued one step ahead (software pipelining).
#include <cooperative_groups.h>

#include "ppd_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int CS = 16;             // CTAs per cluster (non-portable size, needs opt-in)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- DSMEM push with completion: st.async writes a value into a peer CTA's shared memory and
// performs complete_tx on that CTA's mbarrier, so data and "it has arrived" travel together and the
// receiver only waits on a local mbarrier (no cluster-wide barrier per step).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arm(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

struct FwdArgs {
    const float* gi; const float* h0; const float* masks; const float* w_hh; const float* b_hh;
    float* hs; float* h_last; float* sr; float* sz; float* sn; float* sghn;
    int T, E, H;
};

// One cluster = one env (env index = cluster id).  KI = H/32 when known at compile time (the state
// is then held in registers during the dot products), 0 = generic H.
template <int KI>
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_cluster_kernel(const FwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HU = H / CS, R = 3 * HU, E = a.E;
    const int rank = (int)cluster.block_rank();
    const int env = blockIdx.x / CS;
    const int j0 = rank * HU;
    float* W = smem;                       // [R][H]  row (g*HU + u) = W_hh[g*H + j0 + u, :]
    float* hb = W + (size_t)R * H;         // [2][H]  masked previous state, double-buffered by step parity
    float* gh = hb + 2 * H;                // [R]
    float* stage = gh + R;                 // [HU]    new state of this CTA's units, masked for the next step
    __shared__ __align__(8) uint64_t hbar[2];   // hbar[b]: all H values of state buffer b have arrived
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t step_bytes = (uint32_t)H * 4;
    if (tid == 0) {
        mbar_init(&hbar[0], 1);
        mbar_init(&hbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(&hbar[1], step_bytes);        // first use: step 1
        mbar_arm(&hbar[0], step_bytes);        // first use: step 2
    }

    for (int idx = tid; idx < R * H; idx += kThreads) {
        const int r = idx / H, k = idx - r * H;
        const int g = r / HU, u = r - g * HU;
        W[idx] = __ldg(a.w_hh + (size_t)(g * H + j0 + u) * H + k);
    }
    {
        const float m0 = __ldg(a.masks + env);
        for (int k = tid; k < H; k += kThreads) hb[k] = __ldg(a.h0 + (size_t)env * H + k) * m0;
    }
    // gate-phase threads: one per own unit
    const bool gate_thread = tid < HU;
    const int ju = j0 + tid;
    float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f, gi_r = 0.f, gi_z = 0.f, gi_n = 0.f, m_next = 0.f;
    if (gate_thread) {
        bh_r = __ldg(a.b_hh + ju); bh_z = __ldg(a.b_hh + H + ju); bh_n = __ldg(a.b_hh + 2 * H + ju);
        const float* g = a.gi + (size_t)env * 3 * H;
        gi_r = __ldg(g + ju); gi_z = __ldg(g + H + ju); gi_n = __ldg(g + 2 * H + ju);
        m_next = (a.T > 1) ? __ldg(a.masks + (size_t)E + env) : 0.f;
    }
    __syncthreads();
    cluster.sync();          // every CTA's shared memory is initialised before remote stores start

    const int rows_per_warp = (R + kWarps - 1) / kWarps;
    for (int t = 0; t < a.T; ++t) {
        const float* hcur = hb + (t & 1) * H;
        if (t > 0) {
            // the previous step's state: pushed by all 16 CTAs into buffer t&1; use u of that buffer -> parity u&1
            mbar_wait(&hbar[t & 1], (uint32_t)((t - 1) >> 1) & 1u);
        }
        // ---- dot products: warp w owns rows [w*rows_per_warp, ...)
        float hreg[KI > 0 ? KI : 1];
        if (KI > 0) {
#pragma unroll
            for (int i = 0; i < KI; ++i) hreg[i] = hcur[lane + 32 * i];
        }
        for (int rr = 0; rr < rows_per_warp; ++rr) {
            const int r = warp * rows_per_warp + rr;
            if (r >= R) break;
            const float* wrow = W + (size_t)r * H;
            float acc = 0.f;
            if (KI > 0) {
#pragma unroll
                for (int i = 0; i < KI; ++i) acc = fmaf(wrow[lane + 32 * i], hreg[i], acc);
            } else {
                for (int k = lane; k < H; k += 32) acc = fmaf(wrow[k], hcur[k], acc);
            }
            acc = ppd::warp_sum(acc);
            if (lane == 0) gh[r] = acc;
        }
        __syncthreads();
        // ---- gates for this CTA's units; prefetch next step's gi / mask
        const size_t row = (size_t)t * E + env;
        if (gate_thread) {
            const float ghr = gh[tid] + bh_r, ghz = gh[HU + tid] + bh_z, ghn = gh[2 * HU + tid] + bh_n;
            const float r = sigmoidf_(gi_r + ghr);
            const float z = sigmoidf_(gi_z + ghz);
            const float n = tanhf(gi_n + r * ghn);
            const float hm = hcur[ju];
            const float hn = n + z * (hm - n);
            stage[tid] = hn * m_next;
            a.hs[row * H + ju] = hn;
            if (a.sr) { a.sr[row * H + ju] = r; a.sz[row * H + ju] = z; a.sn[row * H + ju] = n; a.sghn[row * H + ju] = ghn; }
            if (a.h_last && t == a.T - 1) a.h_last[(size_t)env * H + ju] = hn;
            if (t + 1 < a.T) {
                const float* g = a.gi + (row + E) * 3 * H;
                gi_r = __ldg(g + ju); gi_z = __ldg(g + H + ju); gi_n = __ldg(g + 2 * H + ju);
                m_next = (t + 2 < a.T) ? __ldg(a.masks + row + 2 * (size_t)E) : 0.f;
            }
        }
        if (t + 1 == a.T) break;
        __syncthreads();           // stage[] written; every warp is done reading buffer t&1
        if (tid == 0 && t >= 1 && t + 2 < a.T) mbar_arm(&hbar[t & 1], step_bytes);   // re-arm for step t+2
        // ---- push the masked new state of own units into every CTA's next-step buffer
        float* hnext = hb + ((t + 1) & 1) * H;
        const uint32_t local_dst = smem_u32(hnext + j0);
        const uint32_t local_bar = smem_u32(&hbar[(t + 1) & 1]);
        for (int idx = tid; idx < CS * HU; idx += kThreads) {
            const int dst = idx / HU, u = idx - dst * HU;
            st_async_f32(map_to_rank(local_dst + 4u * u, (uint32_t)dst), stage[u], map_to_rank(local_bar, (uint32_t)dst));
        }
    }
    cluster.sync();          // do not exit while a peer may still be writing into this CTA's smem
}

struct BwdArgs {
    const float* dhs; const float* masks; const float* w_hh; const float* h0; const float* hs;
    const float* sr; const float* sz; const float* sn; const float* sghn;
    float* dgi; float* dghn; float* dh0;
    int T, E, H;
};

__global__ void __launch_bounds__(kThreads, 1) gru_bwd_cluster_kernel(const BwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HU = H / CS, R = 3 * HU, E = a.E;
    const int rank = (int)cluster.block_rank();
    const int env = blockIdx.x / CS;
    const int j0 = rank * HU;
    float* W = smem;                       // [R][H]   same slice as forward: rows of this CTA's units
    float* dgh = W + (size_t)R * H;        // [R]      d(hidden-side pre-activations) of own units at step t
    float* recv = dgh + R;                 // [2][CS][HU]  partial dh_{t-1} for own units from every CTA
    float* part = recv + 2 * CS * HU;      // [H]      this CTA's partial dh_{t-1} for all units
    __shared__ __align__(8) uint64_t rbar[2];   // rbar[b]: all 16 partial slices of recv buffer b have arrived
    const int tid = threadIdx.x;
    const uint32_t step_bytes = (uint32_t)(CS * HU) * 4;
    if (tid == 0) {
        mbar_init(&rbar[0], 1);
        mbar_init(&rbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(&rbar[0], step_bytes);
        mbar_arm(&rbar[1], step_bytes);
    }

    for (int idx = tid; idx < R * H; idx += kThreads) {
        const int r = idx / H, k = idx - r * H;
        const int g = r / HU, u = r - g * HU;
        W[idx] = __ldg(a.w_hh + (size_t)(g * H + j0 + u) * H + k);
    }
    const bool gate_thread = tid < HU;
    const int ju = j0 + tid;
    float carry = 0.f;                      // dh flowing into step t from step t+1, own unit
    // operands of the current step, prefetched one step ahead
    float p_dh = 0.f, p_r = 0.f, p_z = 0.f, p_n = 0.f, p_ghn = 0.f, p_hp = 0.f, p_m = 0.f;
    auto prefetch = [&](int t) {
        const size_t row = (size_t)t * E + env;
        p_dh = __ldg(a.dhs + row * H + ju);
        p_r = __ldg(a.sr + row * H + ju); p_z = __ldg(a.sz + row * H + ju);
        p_n = __ldg(a.sn + row * H + ju); p_ghn = __ldg(a.sghn + row * H + ju);
        p_m = __ldg(a.masks + row);
        p_hp = (t == 0) ? __ldg(a.h0 + (size_t)env * H + ju) : __ldg(a.hs + (row - E) * H + ju);
    };
    if (gate_thread) prefetch(a.T - 1);
    __syncthreads();
    cluster.sync();

    for (int t = a.T - 1; t >= 0; --t) {
        const size_t row = (size_t)t * E + env;
        float dhz = 0.f, m_t = 0.f;
        // ---- (a) gate backward for own units
        if (gate_thread) {
            const float dh = p_dh + carry;
            const float r = p_r, z = p_z, n = p_n, ghn = p_ghn;
            m_t = p_m;
            const float hm = p_hp * m_t;
            const float dz = dh * (hm - n);
            const float dn = dh * (1.f - z);
            const float dpn = dn * (1.f - n * n);
            const float dpz = dz * z * (1.f - z);
            const float dpr = (dpn * ghn) * r * (1.f - r);
            float* g = a.dgi + row * 3 * H;
            g[ju] = dpr; g[H + ju] = dpz; g[2 * H + ju] = dpn;
            const float dgn = dpn * r;
            a.dghn[row * H + ju] = dgn;
            dgh[tid] = dpr; dgh[HU + tid] = dpz; dgh[2 * HU + tid] = dgn;
            dhz = dh * z;
            if (t > 0) prefetch(t - 1);
        }
        __syncthreads();
        // ---- (b) partial[k] = sum over own rows of dgh[row] * W_hh[row][k], for every unit k
        for (int k = tid; k < H; k += kThreads) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < R; ++r) s = fmaf(dgh[r], W[(size_t)r * H + k], s);
            part[k] = s;
        }
        __syncthreads();
        // ---- (c) reduce-scatter through DSMEM: unit k's partial goes to its owner CTA
        float* rbuf = recv + (t & 1) * CS * HU;
        {
            const uint32_t local_dst = smem_u32(rbuf + rank * HU);
            const uint32_t local_bar = smem_u32(&rbar[t & 1]);
            for (int k = tid; k < H; k += kThreads) {
                const int dst = k / HU, u = k - dst * HU;
                st_async_f32(map_to_rank(local_dst + 4u * u, (uint32_t)dst), part[k], map_to_rank(local_bar, (uint32_t)dst));
            }
        }
        // use u of buffer t&1 (uses go t = T-1, T-3, ... / T-2, T-4, ...) -> parity u&1
        mbar_wait(&rbar[t & 1], (uint32_t)((a.T - 1 - t) >> 1) & 1u);
        // ---- (d) carry for own units
        if (gate_thread) {
            float s = 0.f;
#pragma unroll
            for (int src = 0; src < CS; ++src) s += rbuf[src * HU + tid];
            carry = (s + dhz) * m_t;
            if (t == 0 && a.dh0) a.dh0[(size_t)env * H + ju] = carry;
        }
        __syncthreads();           // every reader of recv buffer t&1 is done before it is re-armed
        if (tid == 0 && t >= 2) mbar_arm(&rbar[t & 1], step_bytes);   // for step t-2
    }
    cluster.sync();
}

size_t fwd_smem(int H) { const int HU = H / CS; return (size_t)(3 * HU * H + 2 * H + 3 * HU + HU) * sizeof(float); }
size_t bwd_smem(int H) { const int HU = H / CS; return (size_t)(3 * HU * H + 3 * HU + 2 * CS * HU + H) * sizeof(float); }

template <typename K>
int launch_cluster(K kernel, const void* args_struct, size_t args_size, int E, size_t smem, cudaStream_t s, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * E);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg);
    if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return -1; }
    (void)args_size;
    void* params[] = {const_cast<void*>(args_struct)};
    e = cudaLaunchKernelExC(&cfg, (const void*)kernel, params);
    if (e != cudaSuccess) {
        ppd::set_error("%s: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return ppd::launch_status(what);
}

}  // namespace

namespace ppd {

// Return 0 on success, -1 if the cluster path does not apply (caller falls back to gru.cu), >0 on CUDA error.
int gru_forward_cluster(const float* gi, const float* h0, const float* masks, const float* w_hh, const float* b_hh,
                        int T, int E, int H, float* hs, float* h_last, float* sr, float* sz, float* sn, float* sghn,
                        cudaStream_t s) {
    if (E > 8 || H % CS != 0 || H / CS > kThreads || fwd_smem(H) > 227 * 1024) return -1;
    FwdArgs a{gi, h0, masks, w_hh, b_hh, hs, h_last, sr, sz, sn, sghn, T, E, H};
    if (H == 512) return launch_cluster(gru_fwd_cluster_kernel<16>, &a, sizeof(a), E, fwd_smem(H), s, "gru_fwd_cluster_kernel");
    return launch_cluster(gru_fwd_cluster_kernel<0>, &a, sizeof(a), E, fwd_smem(H), s, "gru_fwd_cluster_kernel");
}

int gru_backward_cluster(const float* dhs, const float* masks, const float* w_hh, const float* h0, const float* hs,
                         const float* sr, const float* sz, const float* sn, const float* sghn, int T, int E, int H,
                         float* dgi, float* dghn, float* dh0, cudaStream_t s) {
    if (E > 8 || H % CS != 0 || H / CS > kThreads || bwd_smem(H) > 227 * 1024) return -1;
    BwdArgs a{dhs, masks, w_hh, h0, hs, sr, sz, sn, sghn, dgi, dghn, dh0, T, E, H};
    return launch_cluster(gru_bwd_cluster_kernel, &a, sizeof(a), E, bwd_smem(H), s, "gru_bwd_cluster_kernel");
}

}  // namespace ppd
