// GRU recurrence, one thread-block cluster per env (the PPO-Dash configuration has E = 4 envs per minibatch; larger E runs as waves of clusters)
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
#include <stdlib.h>

#include "ppd_common.cuh"

namespace cg = cooperative_groups;

namespace ppd { extern int g_multi_clusters; }

namespace {

constexpr int CS = 16;             // CTAs per cluster (non-portable size, needs opt-in)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// Gate non-linearities of the register kernels: the gate phase is one warp's dependent chain inside every recurrence step (all
// other warps wait at the barrier), so its length is step latency.  ex2.approx (2^-22 relative) and rcp.approx (1 ulp) keep the
// absolute error of both functions below 2e-7 -- under the fp32 rounding noise of the 512-term dot products feeding them.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - 2.f * rcp_approx(1.f + ex2_approx(2.8853900817779268f * x)); }

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

// =====================================================================================================
// H = 512 specialisations with the W_hh slice held in REGISTERS (the generic kernels above stream the
// 192 KB slice from shared memory every step: 1536 cycles at 128 B/clk, the dominant per-step cost).
//   forward : 384 threads; thread (row group rg = tid/16, column chunk cc = tid%16) keeps the 4 x 32 block
//             W_hh[rows 4rg..4rg+3][32cc .. +32) in 128 registers; per step 8 LDS.128 of the state (each value feeds
//             4 rows: 384 shared-memory wavefronts per step instead of the 1536 of a one-row-per-thread mapping, which
//             was the dominant per-step cost), 128 FMAs and a 16-lane shuffle reduction of the 4 partial sums.
//   backward: 512 threads; thread (column group cg = tid/4, row quarter rq = tid%4) keeps the 24 x 4 block
//             W_hh[own rows 24rq..24rq+23][4cg .. +4) in 96 registers; per step 6 LDS.128 of d(gates) (each value feeds
//             4 columns: 384 shared-memory wavefronts per step instead of 1536), 96 FMAs and a 3-shuffle reduction over the
//             4 row quarters that leaves thread tid with the partial sum of unit tid, pushed straight to its owner.
// =====================================================================================================
constexpr int kH = 512, kHU = kH / CS, kR = 3 * kHU;      // 32 units, 96 rows per CTA
constexpr int kFwdThreads = kR * 4;                        // 384
constexpr int kCW = 32;                                    // state values per column chunk
constexpr int kNC = kH / kCW;                              // 16 chunks
constexpr int kCPad = kCW + 4;                             // chunk stride in smem: the 8 lanes of an LDS.128 phase hit 32 banks

__device__ __forceinline__ int hpad(int j) { return (j / kCW) * kCPad + (j % kCW); }

// Packed fp32 FMA (sm_100a FFMA2): two independent fused multiply-adds per lane and instruction, c = a * b + c elementwise.
// The mat-vec of a step is 49 152 FMAs per CTA; as scalar FFMA that alone is several hundred issue cycles of every step.
__device__ __forceinline__ void ffma2(float2& c, const float2 a, const float2 b) {
    unsigned long long cc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    c = *reinterpret_cast<float2*>(&cc);
}

#ifdef PPD_GRU_TRACE
// clock stamps into shared memory (CS2R + STS; a traced step is ~14 % longer), dumped at exit: [kernel][step - 256][slot], CTA 0,
// thread 0 (gate warp) / first thread of the last warp.  tools/probes/gru_trace.py
__device__ unsigned g_gru_trace[2][32][8];
#define GRU_TR(tt, slot) do { if (blockIdx.x == 0 && (tt) >= 256 && (tt) < 288) trs[((tt) - 256) * 8 + (slot)] = (unsigned)clock64(); } while (0)
#else
#define GRU_TR(tt, slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kFwdThreads, 1) gru_fwd_cluster512_kernel(const FwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
#ifdef PPD_GRU_TRACE
    __shared__ unsigned trs[32 * 8];
#endif
    __shared__ __align__(16) float hb[2][kNC * kCPad];    // masked previous state, padded chunks, double-buffered
    __shared__ float gh[kR];
    __shared__ float stage[kHU];
    __shared__ __align__(8) uint64_t hbar[2];
    const int E = a.E;
    const int rank = (int)cluster.block_rank();
    const int env = blockIdx.x / CS;
    const int j0 = rank * kHU;
    const int tid = threadIdx.x;
    const int rg = tid >> 4, cc = tid & 15;
    const uint32_t step_bytes = kH * 4;
    if (tid == 0) {
        mbar_init(&hbar[0], 1);
        mbar_init(&hbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(&hbar[1], step_bytes);
        mbar_arm(&hbar[0], step_bytes);
    }
    float2 w[4][kCW / 2];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = 4 * rg + rr;
        const int g = r / kHU, u = r - g * kHU;
        const float4* src = reinterpret_cast<const float4*>(a.w_hh + (size_t)(g * kH + j0 + u) * kH + cc * kCW);
#pragma unroll
        for (int i = 0; i < kCW / 4; ++i) {
            const float4 v = __ldg(src + i);
            w[rr][2 * i] = make_float2(v.x, v.y); w[rr][2 * i + 1] = make_float2(v.z, v.w);
        }
    }
    {
        const float m0 = __ldg(a.masks + env);
        for (int k = tid; k < kH; k += kFwdThreads) hb[0][hpad(k)] = __ldg(a.h0 + (size_t)env * kH + k) * m0;
    }
    const bool gate_thread = tid < kHU;
    const int ju = j0 + tid;
    float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f, gi_r = 0.f, gi_z = 0.f, gi_n = 0.f, m_next = 0.f;
    if (gate_thread) {
        bh_r = __ldg(a.b_hh + ju); bh_z = __ldg(a.b_hh + kH + ju); bh_n = __ldg(a.b_hh + 2 * kH + ju);
        const float* g = a.gi + (size_t)env * 3 * kH;
        gi_r = __ldg(g + ju); gi_z = __ldg(g + kH + ju); gi_n = __ldg(g + 2 * kH + ju);
        m_next = (a.T > 1) ? __ldg(a.masks + (size_t)E + env) : 0.f;
    }
    __syncthreads();
    cluster.sync();

    for (int t = 0; t < a.T; ++t) {
        const float* hcur = hb[t & 1];
        if (tid == 0) GRU_TR(t, 0);
        if (t > 0) mbar_wait(&hbar[t & 1], (uint32_t)((t - 1) >> 1) & 1u);
        if (tid == 0) GRU_TR(t, 1);
        if (tid == kFwdThreads - 32) GRU_TR(t, 6);
        // ---- 4 x 32 block of the mat-vec: every state value read from shared memory feeds four rows
        {
            const float4* h4 = reinterpret_cast<const float4*>(hcur + cc * kCPad);
            float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;       // (even-k, odd-k) partial sums of the 4 rows
#pragma unroll
            for (int i = 0; i < kCW / 4; ++i) {
                const float4 hv = h4[i];
                const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
                ffma2(q0, w[0][2 * i], h01); ffma2(q1, w[1][2 * i], h01); ffma2(q2, w[2][2 * i], h01); ffma2(q3, w[3][2 * i], h01);
                ffma2(q0, w[0][2 * i + 1], h23); ffma2(q1, w[1][2 * i + 1], h23); ffma2(q2, w[2][2 * i + 1], h23); ffma2(q3, w[3][2 * i + 1], h23);
            }
            const float s0 = q0.x + q0.y, s1 = q1.x + q1.y, s2 = q2.x + q2.y, s3 = q3.x + q3.y;
            // sum over the 16 column chunks (lanes of one half-warp); after each exchange a lane keeps the sums it still owns:
            // xor 8: rows {0,1} <-> {2,3}; xor 4: row pairs; then plain butterflies -- 4 + 2 + 1 + 1 = 8 shuffles instead of 16
            {
                const bool up = (cc & 8) != 0;
                const float k0 = up ? s2 : s0, k1 = up ? s3 : s1, g0 = up ? s0 : s2, g1 = up ? s1 : s3;
                const float r0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 8), r1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 8);
                const bool up4 = (cc & 4) != 0;
                const float kk = up4 ? r1 : r0, gg = up4 ? r0 : r1;
                float acc = kk + __shfl_xor_sync(0xffffffffu, gg, 4);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                // lanes with cc % 4 == 0 hold row 4rg + 2*(cc>>3 & 1) + (cc>>2 & 1)
                if ((cc & 3) == 0) gh[4 * rg + ((cc >> 3) & 1) * 2 + ((cc >> 2) & 1)] = acc;
            }
        }
        if (tid == 0) GRU_TR(t, 2);
        if (tid == kFwdThreads - 32) GRU_TR(t, 7);
        __syncthreads();
        if (tid == 0) GRU_TR(t, 3);
        const size_t row = (size_t)t * E + env;
        if (gate_thread) {
            const float ghr = gh[tid] + bh_r, ghz = gh[kHU + tid] + bh_z, ghn = gh[2 * kHU + tid] + bh_n;
#ifdef PPD_GRU_EXACT_GATES
            const float rg = sigmoidf_(gi_r + ghr);
            const float z = sigmoidf_(gi_z + ghz);
            const float n = tanhf(gi_n + rg * ghn);
#else
            const float rg = fast_sigmoid(gi_r + ghr);
            const float z = fast_sigmoid(gi_z + ghz);
            const float n = fast_tanh(gi_n + rg * ghn);
#endif
            const float hm = hcur[hpad(ju)];
            const float hn = n + z * (hm - n);
            stage[tid] = hn * m_next;
            a.hs[row * kH + ju] = hn;
            if (a.sr) { a.sr[row * kH + ju] = rg; a.sz[row * kH + ju] = z; a.sn[row * kH + ju] = n; a.sghn[row * kH + ju] = ghn; }
            if (a.h_last && t == a.T - 1) a.h_last[(size_t)env * kH + ju] = hn;
            if (t + 1 < a.T) {
                const float* g = a.gi + (row + E) * 3 * kH;
                gi_r = __ldg(g + ju); gi_z = __ldg(g + kH + ju); gi_n = __ldg(g + 2 * kH + ju);
                m_next = (t + 2 < a.T) ? __ldg(a.masks + row + 2 * (size_t)E) : 0.f;
            }
        }
        if (t + 1 == a.T) break;
        if (tid == 0) GRU_TR(t, 4);
        __syncthreads();
        if (tid == 0) GRU_TR(t, 5);
        if (tid == 0 && t >= 1 && t + 2 < a.T) mbar_arm(&hbar[t & 1], step_bytes);
        {
            const uint32_t local_base = smem_u32(&hb[(t + 1) & 1][0]);
            const uint32_t local_bar = smem_u32(&hbar[(t + 1) & 1]);
            for (int idx = tid; idx < CS * kHU; idx += kFwdThreads) {
                const int dst = idx / kHU, u = idx - dst * kHU;
                st_async_f32(map_to_rank(local_base + 4u * (uint32_t)hpad(j0 + u), (uint32_t)dst), stage[u],
                             map_to_rank(local_bar, (uint32_t)dst));
            }
        }
    }
    cluster.sync();
#ifdef PPD_GRU_TRACE
    if (blockIdx.x == 0) for (int i = tid; i < 32 * 8; i += kFwdThreads) (&g_gru_trace[0][0][0])[i] = trs[i];
#endif
}

__global__ void __launch_bounds__(kH, 1) gru_bwd_cluster512_kernel(const BwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
#ifdef PPD_GRU_TRACE
    __shared__ unsigned trs[32 * 8];
#endif
    __shared__ __align__(16) float dgh[kR];
    __shared__ float recv[2][CS * kHU];
    __shared__ __align__(8) uint64_t rbar[2];
    const int E = a.E;
    const int rank = (int)cluster.block_rank();
    const int env = blockIdx.x / CS;
    const int j0 = rank * kHU;
    const int tid = threadIdx.x;                 // = unit index k this thread produces partial sums for
    const uint32_t step_bytes = CS * kHU * 4;
    if (tid == 0) {
        mbar_init(&rbar[0], 1);
        mbar_init(&rbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(&rbar[0], step_bytes);
        mbar_arm(&rbar[1], step_bytes);
    }
    const int cg4 = tid >> 2, rq = tid & 3;
    float2 wc[kR / 4][2];                         // W_hh[own row 24rq + i][4cg4 + (0,1) | (2,3)]
#pragma unroll
    for (int i = 0; i < kR / 4; ++i) {
        const int r = (kR / 4) * rq + i;
        const int g = r / kHU, u = r - g * kHU;
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.w_hh + (size_t)(g * kH + j0 + u) * kH + 4 * cg4));
        wc[i][0] = make_float2(v.x, v.y); wc[i][1] = make_float2(v.z, v.w);
    }
    const bool gate_thread = tid < kHU;
    const int ju = j0 + tid;
    float carry = 0.f;
    float p_dh = 0.f, p_r = 0.f, p_z = 0.f, p_n = 0.f, p_ghn = 0.f, p_hp = 0.f, p_m = 0.f;
    auto prefetch = [&](int t) {
        const size_t row = (size_t)t * E + env;
        p_dh = __ldg(a.dhs + row * kH + ju);
        p_r = __ldg(a.sr + row * kH + ju); p_z = __ldg(a.sz + row * kH + ju);
        p_n = __ldg(a.sn + row * kH + ju); p_ghn = __ldg(a.sghn + row * kH + ju);
        p_m = __ldg(a.masks + row);
        p_hp = (t == 0) ? __ldg(a.h0 + (size_t)env * kH + ju) : __ldg(a.hs + (row - E) * kH + ju);
    };
    if (gate_thread) prefetch(a.T - 1);
    // owner of this thread's unit and the slot of this CTA's contribution in the owner's receive buffers
    const uint32_t dst_rank = (uint32_t)(tid / kHU);
    const uint32_t slot = (uint32_t)(rank * kHU + (tid % kHU));
    __syncthreads();
    cluster.sync();

    for (int t = a.T - 1; t >= 0; --t) {
        const size_t row = (size_t)t * E + env;
        float dhz = 0.f, m_t = 0.f;
        if (tid == 0) GRU_TR(a.T - 1 - t, 0);
        if (gate_thread) {
            const float dh = p_dh + carry;
            const float rg = p_r, z = p_z, n = p_n, ghn = p_ghn;
            m_t = p_m;
            const float hm = p_hp * m_t;
            const float dz = dh * (hm - n);
            const float dn = dh * (1.f - z);
            const float dpn = dn * (1.f - n * n);
            const float dpz = dz * z * (1.f - z);
            const float dpr = (dpn * ghn) * rg * (1.f - rg);
            float* g = a.dgi + row * 3 * kH;
            g[ju] = dpr; g[kH + ju] = dpz; g[2 * kH + ju] = dpn;
            const float dgn = dpn * rg;
            a.dghn[row * kH + ju] = dgn;
            dgh[tid] = dpr; dgh[kHU + tid] = dpz; dgh[2 * kHU + tid] = dgn;
            dhz = dh * z;
            if (t > 0) prefetch(t - 1);
        }
        if (tid == 0) GRU_TR(a.T - 1 - t, 1);
        __syncthreads();
        if (tid == 0) GRU_TR(a.T - 1 - t, 2);
        // ---- partial dh_{t-1}[k] over this CTA's 96 rows
        float s;
        {
            const float4* d4 = reinterpret_cast<const float4*>(dgh + (kR / 4) * rq);
            float2 c01 = make_float2(0.f, 0.f), c23 = c01;      // columns 4cg4 .. +3 over this thread's 24 rows
#pragma unroll
            for (int i = 0; i < kR / 16; ++i) {
                const float4 dv = d4[i];
                const float2 dx = make_float2(dv.x, dv.x), dy = make_float2(dv.y, dv.y), dz = make_float2(dv.z, dv.z), dw = make_float2(dv.w, dv.w);
                ffma2(c01, dx, wc[4 * i][0]); ffma2(c23, dx, wc[4 * i][1]);
                ffma2(c01, dy, wc[4 * i + 1][0]); ffma2(c23, dy, wc[4 * i + 1][1]);
                ffma2(c01, dz, wc[4 * i + 2][0]); ffma2(c23, dz, wc[4 * i + 2][1]);
                ffma2(c01, dw, wc[4 * i + 3][0]); ffma2(c23, dw, wc[4 * i + 3][1]);
            }
            const float c0 = c01.x, c1 = c01.y, c2 = c23.x, c3 = c23.y;
            // sum over the 4 row quarters (adjacent lanes); each exchange halves what a lane keeps: lane rq ends with column rq,
            // i.e. thread tid with unit 4cg4 + rq = tid
            const bool up2 = (rq & 2) != 0;
            const float k0 = up2 ? c2 : c0, k1 = up2 ? c3 : c1, g0 = up2 ? c0 : c2, g1 = up2 ? c1 : c3;
            const float r0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 2), r1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 2);
            const bool up1 = (rq & 1) != 0;
            s = (up1 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, up1 ? r0 : r1, 1);
        }
        // ---- reduce-scatter: push to the owner of unit k
        if (tid == 0) GRU_TR(a.T - 1 - t, 3);
        st_async_f32(map_to_rank(smem_u32(&recv[t & 1][slot]), dst_rank), s, map_to_rank(smem_u32(&rbar[t & 1]), dst_rank));
        if (tid == 0) GRU_TR(a.T - 1 - t, 4);
        mbar_wait(&rbar[t & 1], (uint32_t)((a.T - 1 - t) >> 1) & 1u);
        if (tid == 0) GRU_TR(a.T - 1 - t, 5);
        if (gate_thread) {
            float acc = 0.f;
#pragma unroll
            for (int src = 0; src < CS; ++src) acc += recv[t & 1][src * kHU + tid];
            carry = (acc + dhz) * m_t;
            if (t == 0 && a.dh0) a.dh0[(size_t)env * kH + ju] = carry;
        }
        if (tid == 0) GRU_TR(a.T - 1 - t, 6);
        __syncthreads();           // dgh and recv[t&1] are free again
        if (tid == 0) GRU_TR(a.T - 1 - t, 7);
        if (tid == 0 && t >= 2) mbar_arm(&rbar[t & 1], step_bytes);
    }
    cluster.sync();
#ifdef PPD_GRU_TRACE
    if (blockIdx.x == 0) for (int i = tid; i < 32 * 8; i += kH) (&g_gru_trace[1][0][0])[i] = trs[i];
#endif
}

// =====================================================================================================
// Backward over many envs (E > 8, the 128-env minibatches of the 8 x 1024 configuration): PERSISTENT clusters, INTERLEAVED envs.
// One cluster per env leaves the recurrence latency-bound -- of the ~1700 clocks of a backward step only ~400 are FMAs, the rest
// is the gate chain of one warp, two block barriers and the DSMEM exchange -- and E envs cost ceil(E / 7 resident clusters)
// sequential waves of that.  Envs are independent sequences and W_hh is the same for all of them, so here every resident cluster
// keeps its register copy of W_hh for the whole launch and walks NE env slots in lock step as a two-stage software pipeline:
//   G(i)  ONE designated warp per slot (warps 12..15: the highest warp id of each scheduler, which the arbiter favours; lane =
//         hidden unit): carry sum of the 16 received partials, gate backward, the step's results to HBM, the next operands;
//   M(i)  every warp: the 24 x 4 register-block mat-vec and the reduce-scatter push of the one-env kernel.
// G runs one slot-step AHEAD of M, so the exchange latency and the gate chain hide behind the other slots' mat-vecs; hand-offs
// inside the CTA are mbarriers (full / empty per slot), across CTAs st.async + complete_tx as in the one-env kernel.
// Measured at E = 128, T = 512: 6.2 ms against 8.3 ms for 19 waves of one-env clusters.  The same treatment of the FORWARD
// recurrence did not pay and is not kept: its mat-vec needs a 12-shuffle butterfly per lane (3 gate rows per hidden unit) and
// an extra staged hand-off for the coalesced push; three variants (gate phase replicated over the half-warps, on designated
// warps, as a 3-stage pipeline) ran at 1700-3000 clocks per slot-step, no better than the 1600 of the one-env kernel -- every
// mbarrier wait costs a warp ~250 clocks even when its phase is already complete, and replicated gate / address arithmetic
// makes the kernel issue-bound (~350 instructions per warp and slot-step against the 57 of the mat-vec itself; clock64 stamps).
// (A 17th warp for the gate phase would cap every thread at 96 registers: five warps on one scheduler's 16 K registers.)
// Buffers are double-buffered by step parity; a CTA can only receive step t-1 data for a buffer after every CTA has consumed
// step t+1 from it (the same argument as for the one-env kernels).  Rounds (groups of NE envs) are separated by cluster.sync().
// =====================================================================================================
constexpr int kMaxNE = 4;

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NE>
__global__ void __launch_bounds__(kH, 1) gru_bwd_multi_kernel(const BwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float dgh[NE][kR];              // d(hidden-side pre-activations) of own units
    __shared__ float recv[NE][2][CS * kHU];                  // partial dh_{t-1} of own units from every CTA, by step parity
    __shared__ __align__(8) uint64_t rbar[NE][2], full[NE], empty[NE];
    const int E = a.E, T = a.T;
    const int rank = (int)cluster.block_rank();
    const int NC = gridDim.x / CS, cid = blockIdx.x / CS;
    const int j0 = rank * kHU;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t step_bytes = CS * kHU * 4;
    if (tid == 0) {
        for (int k = 0; k < NE; ++k) {
            mbar_init(&rbar[k][0], 1); mbar_init(&rbar[k][1], 1);
            mbar_init(&full[k], 32); mbar_init(&empty[k], kH / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int cg4 = tid >> 2, rq = tid & 3;
    float2 wc[kR / 4][2];                         // W_hh[own row 24rq + i][4cg4 + (0,1) | (2,3)]
#pragma unroll
    for (int i = 0; i < kR / 4; ++i) {
        const int r = (kR / 4) * rq + i;
        const int g = r / kHU, uu = r - g * kHU;
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.w_hh + (size_t)(g * kH + j0 + uu) * kH + 4 * cg4));
        wc[i][0] = make_float2(v.x, v.y); wc[i][1] = make_float2(v.z, v.w);
    }
    int myslot = -1;
#pragma unroll
    for (int k = 0; k < NE; ++k) if (warp == (12 + k)) myslot = k;
    const int ju = j0 + lane;
    // thread tid ends up with the partial sum of unit tid: its owner CTA and this CTA's slot in the owner's receive buffers
    const uint32_t dst_rank = (uint32_t)(tid / kHU);
    const uint32_t slot = (uint32_t)(rank * kHU + (tid % kHU));
    uint32_t ph_r = 0, ph_full = 0, ph_empty = 1u;           // rbar bit b and empty of the designated slot; full bit k (all warps)
    __syncthreads();

    const int n_mine = cid < E ? (E - cid + NC - 1) / NC : 0;
    const int n_max = (E + NC - 1) / NC;
    // operands of the designated slot's next gate phase
    float o_dh = 0.f, o_r = 0.f, o_z = 0.f, o_n = 0.f, o_ghn = 0.f, o_hp = 0.f, o_m = 0.f, dhz = 0.f, m_prev = 0.f;
    for (int base = 0; base < n_max; base += NE) {
        const int nact = min(NE, n_mine - base);
        auto load_ops = [&](int t, int env) {
            const size_t row = (size_t)t * E + env;
            o_dh = __ldg(a.dhs + row * kH + ju);
            o_r = __ldg(a.sr + row * kH + ju); o_z = __ldg(a.sz + row * kH + ju);
            o_n = __ldg(a.sn + row * kH + ju); o_ghn = __ldg(a.sghn + row * kH + ju);
            o_m = __ldg(a.masks + row);
            o_hp = (t == 0) ? __ldg(a.h0 + (size_t)env * kH + ju) : __ldg(a.hs + (row - E) * kH + ju);
        };
        for (int k = 0; k < nact; ++k) {
            if (myslot == k) { load_ops(T - 1, cid + NC * (base + k)); dhz = 0.f; m_prev = 0.f; }
            if (tid == 0) {
                mbar_arm(&rbar[k][(T - 1) & 1], step_bytes);
                if (T > 1) mbar_arm(&rbar[k][(T - 2) & 1], step_bytes);
            }
        }
        __syncthreads();
        cluster.sync();
        // carry into step tpush-1 of slot k: the partial sums pushed at step tpush (buffer tpush & 1), summed over the 16 CTAs
        auto take_carry = [&](const int k, const int tpush) -> float {
            const int b = tpush & 1;
            mbar_wait(&rbar[k][b], (ph_r >> b) & 1u);
            ph_r ^= 1u << b;
            if (lane == 0 && tpush >= 2) mbar_arm(&rbar[k][b], step_bytes);     // phase complete: arm it for the pushes of step tpush-2
            const float* rb = &recv[k][b][lane];
            float v[CS];
#pragma unroll
            for (int src = 0; src < CS; ++src) v[src] = rb[src * kHU];
#pragma unroll
            for (int w2 = CS / 2; w2 > 0; w2 >>= 1)
#pragma unroll
                for (int i = 0; i < w2; ++i) v[i] += v[i + w2];
            return (v[0] + dhz) * m_prev;
        };
        auto gate = [&](const int k, const int t) {              // designated warp of slot k: carry, gate backward, results, next operands
            const int env = cid + NC * (base + k);
            const float carry = (t == T - 1) ? 0.f : take_carry(k, t + 1);
            const float dh = o_dh + carry;
            const float rg = o_r, z = o_z, n = o_n, ghn = o_ghn;
            const float m_t = o_m;
            const float hm = o_hp * m_t;
            const float dz = dh * (hm - n);
            const float dn = dh * (1.f - z);
            const float dpn = dn * (1.f - n * n);
            const float dpz = dz * z * (1.f - z);
            const float dpr = (dpn * ghn) * rg * (1.f - rg);
            const float dgn = dpn * rg;
            dhz = dh * z; m_prev = m_t;
            mbar_wait(&empty[k], ph_empty & 1u);                 // the mat-vecs are done with the previous step's values
            ph_empty ^= 1u;
            dgh[k][lane] = dpr; dgh[k][kHU + lane] = dpz; dgh[k][2 * kHU + lane] = dgn;
            mbar_arrive_local(&full[k]);
            const size_t row = (size_t)t * E + env;
            float* g = a.dgi + row * 3 * kH;
            g[ju] = dpr; g[kH + ju] = dpz; g[2 * kH + ju] = dpn;
            a.dghn[row * kH + ju] = dgn;
            if (t > 0) load_ops(t - 1, env);
        };
        auto matvec = [&](const int k, const int t) {            // partial dh_{t-1}[unit tid] over this CTA's 96 rows, pushed to the unit's owner
            const int b = t & 1;
            mbar_wait(&full[k], (ph_full >> k) & 1u);
            ph_full ^= 1u << k;
            const float4* d4 = reinterpret_cast<const float4*>(&dgh[k][(kR / 4) * rq]);
            float2 c01 = make_float2(0.f, 0.f), c23 = c01;
#pragma unroll
            for (int i = 0; i < kR / 16; ++i) {
                const float4 dv = d4[i];
                const float2 dx = make_float2(dv.x, dv.x), dy = make_float2(dv.y, dv.y), dz = make_float2(dv.z, dv.z), dw = make_float2(dv.w, dv.w);
                ffma2(c01, dx, wc[4 * i][0]); ffma2(c23, dx, wc[4 * i][1]);
                ffma2(c01, dy, wc[4 * i + 1][0]); ffma2(c23, dy, wc[4 * i + 1][1]);
                ffma2(c01, dz, wc[4 * i + 2][0]); ffma2(c23, dz, wc[4 * i + 2][1]);
                ffma2(c01, dw, wc[4 * i + 3][0]); ffma2(c23, dw, wc[4 * i + 3][1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&empty[k]);         // dgh[k] has been consumed
            const float c0 = c01.x, c1 = c01.y, c2 = c23.x, c3 = c23.y;
            const bool up2 = (rq & 2) != 0;
            const float k0 = up2 ? c2 : c0, k1 = up2 ? c3 : c1, g0 = up2 ? c0 : c2, g1 = up2 ? c1 : c3;
            const float r0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 2), r1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 2);
            const bool up1 = (rq & 1) != 0;
            const float sum = (up1 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, up1 ? r0 : r1, 1);
            st_async_f32(map_to_rank(smem_u32(&recv[k][b][slot]), dst_rank), sum, map_to_rank(smem_u32(&rbar[k][b]), dst_rank));
        };
        // Software pipeline over the slot-steps i = (T - 1 - t) * nact + k: the gate phase of i + 1 is issued BEFORE the mat-vec of i
        // (its carry comes from the mat-vec of i + 1 - nact, issued earlier whenever nact > 1)
        if (nact > 0) {
            const int n_it = T * nact;
            int km = 0, tm = T - 1, kg = 0, tg = T - 1;
            if (nact > 1) {
                if (warp == 12) gate(0, T - 1);
                if (++kg == nact) { kg = 0; --tg; }
            }
            for (int i = 0; i < n_it; ++i) {
                if (nact == 1) {
                    if (warp == 12) gate(0, tm);
                    matvec(0, tm);
                    --tm;
                } else {
                    if (tg >= 0 && warp == (12 + kg)) gate(kg, tg);
                    matvec(km, tm);
                    if (++km == nact) { km = 0; --tm; }
                    if (++kg == nact) { kg = 0; --tg; }
                }
            }
            // dh0: the carry out of step 0
            if (myslot >= 0 && myslot < nact) {
                const float carry = take_carry(myslot, 0);
                if (a.dh0) a.dh0[(size_t)(cid + NC * (base + myslot)) * kH + ju] = carry;
            }
        }
        cluster.sync();
    }
}

// Persistent launch: as many clusters as the device holds (at most one per env), NE envs interleaved per cluster and round.
template <typename K>
int launch_multi(K kernel, const void* args_struct, int nclusters, cudaStream_t s, const char* what) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * nclusters);
    cfg.blockDim = dim3(kH);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* params[] = {const_cast<void*>(args_struct)};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kernel, params);
    if (e != cudaSuccess) {
        ppd::set_error("%s: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return ppd::launch_status(what);
}

// resident 16-CTA clusters of the multi-env kernels on this device (0: the cluster launch is not possible)
template <typename K>
int multi_clusters(K kernel) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 16);
    cfg.blockDim = dim3(kH);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// envs per cluster and round: the fewest rounds, then the smallest NE that still needs that many
inline int pick_ne(int E, int nclusters) {
    const int per = (E + nclusters - 1) / nclusters;
    const int rounds = (per + kMaxNE - 1) / kMaxNE;
    return (per + rounds - 1) / rounds;
}

int multi_backward(const BwdArgs& a, cudaStream_t s) {
    static int ncl = -1;
    if (ncl < 0) {
        ncl = multi_clusters(gru_bwd_multi_kernel<4>);
        int n;
        n = multi_clusters(gru_bwd_multi_kernel<3>); if (n < ncl) ncl = n;
        n = multi_clusters(gru_bwd_multi_kernel<2>); if (n < ncl) ncl = n;
        n = multi_clusters(gru_bwd_multi_kernel<1>); if (n < ncl) ncl = n;
    }
    if (ncl < 1) return -1;
    const int want = ppd::g_multi_clusters > 0 ? ppd::g_multi_clusters : ncl;
    const int nc = a.E < want ? a.E : want;
    switch (pick_ne(a.E, nc)) {
        case 1: return launch_multi(gru_bwd_multi_kernel<1>, &a, nc, s, "gru_bwd_multi_kernel");
        case 2: return launch_multi(gru_bwd_multi_kernel<2>, &a, nc, s, "gru_bwd_multi_kernel");
        case 3: return launch_multi(gru_bwd_multi_kernel<3>, &a, nc, s, "gru_bwd_multi_kernel");
        default: return launch_multi(gru_bwd_multi_kernel<4>, &a, nc, s, "gru_bwd_multi_kernel");
    }
}

size_t fwd_smem(int H) { const int HU = H / CS; return (size_t)(3 * HU * H + 2 * H + 3 * HU + HU) * sizeof(float); }
size_t bwd_smem(int H) { const int HU = H / CS; return (size_t)(3 * HU * H + 3 * HU + 2 * CS * HU + H) * sizeof(float); }

template <typename K>
int launch_cluster(K kernel, const void* args_struct, int threads, int E, size_t smem, cudaStream_t s, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess && smem > 0) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * E);
    cfg.blockDim = dim3(threads);
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

int g_reg_kernels = 1;   // 0 (ppd_gru_set_mode(2)): use the generic shared-memory cluster kernels even at H = 512
int g_multi_clusters = 0; // resident clusters the persistent kernels assume (0: ask the occupancy API); ppd_gru_set_mode(100 + n)
int g_multi_min_e = 9;   // H = 512: E >= this runs the persistent interleaved-env BACKWARD kernel (ppd_gru_set_mode(3): every E; (4): never)

// Return 0 on success, -1 if the cluster path does not apply (caller falls back to gru.cu), >0 on CUDA error.
int gru_forward_cluster(const float* gi, const float* h0, const float* masks, const float* w_hh, const float* b_hh,
                        int T, int E, int H, float* hs, float* h_last, float* sr, float* sz, float* sn, float* sghn,
                        cudaStream_t s) {
    // One cluster per env, no dependency between clusters: any number of envs runs as waves of 148 / 16 = 9 clusters.
    // At H = 512 that beats the grid-cooperative kernel by ~3x even at E = 128 (14 waves x 0.49 ms against 20.8 ms per
    // 512-step minibatch), so the register kernels take every E; the shared-memory cluster variant stays for E <= 8.
    if ((E > 8 && !(H == kH && g_reg_kernels)) || E > 4095 || H % CS != 0 || H / CS > kThreads || fwd_smem(H) > 227 * 1024) return -1;
    FwdArgs a{gi, h0, masks, w_hh, b_hh, hs, h_last, sr, sz, sn, sghn, T, E, H};
    if (H == kH && g_reg_kernels) return launch_cluster(gru_fwd_cluster512_kernel, &a, kFwdThreads, E, 0, s, "gru_fwd_cluster512_kernel");
    if (H == 512) return launch_cluster(gru_fwd_cluster_kernel<16>, &a, kThreads, E, fwd_smem(H), s, "gru_fwd_cluster_kernel");
    return launch_cluster(gru_fwd_cluster_kernel<0>, &a, kThreads, E, fwd_smem(H), s, "gru_fwd_cluster_kernel");
}

int gru_backward_cluster(const float* dhs, const float* masks, const float* w_hh, const float* h0, const float* hs,
                         const float* sr, const float* sz, const float* sn, const float* sghn, int T, int E, int H,
                         float* dgi, float* dghn, float* dh0, cudaStream_t s) {
    if ((E > 8 && !(H == kH && g_reg_kernels)) || E > 4095 || H % CS != 0 || H / CS > kThreads || bwd_smem(H) > 227 * 1024) return -1;
    BwdArgs a{dhs, masks, w_hh, h0, hs, sr, sz, sn, sghn, dgi, dghn, dh0, T, E, H};
    if (H == kH && g_reg_kernels && E >= g_multi_min_e) { const int rc = multi_backward(a, s); if (rc >= 0) return rc; }
    if (H == kH && g_reg_kernels) return launch_cluster(gru_bwd_cluster512_kernel, &a, kH, E, 0, s, "gru_bwd_cluster512_kernel");
    return launch_cluster(gru_bwd_cluster_kernel, &a, kThreads, E, bwd_smem(H), s, "gru_bwd_cluster_kernel");
}

}  // namespace ppd


#ifdef PPD_GRU_TRACE
// Trace builds only (tools/probes/gru_trace.py)
extern "C" int ppd_gru_trace_read(unsigned* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_gru_trace, sizeof(unsigned) * 2 * 32 * 8);
}
#endif
