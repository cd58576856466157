// Returns / GAE as a single-pass chained affine scan, one env per lane.
// Replaces RolloutStorage.compute_returns (PKG/storage.py:82-121).  Compile with -fmad=false:
// the in-chunk replay keeps the reference's operation order (separate mul / add roundings).
//
// Layout: every field is time-major [T(+1), N]; for fixed t the N envs are contiguous, so a
// warp reading lane = env issues one 128-byte request per (field, t).
//
// Work split: CTA = 32 envs x one time segment of kWarps*kSteps steps; warp w owns kSteps
// consecutive steps.  The recurrence X_t = a_t X_{t+1} + c_t is an affine map per step, so a
// run of steps composes to X_out = P * X_in + Q:
//   phase A  each lane loads its kSteps x {r, V, m, (b)} into registers (all loads issued up
//            front -> ~3*kSteps requests in flight per thread), evaluates its chunk with zero
//            incoming carry, publishes its (P, Q) to shared memory; warp 0 folds the kWarps maps
//            into the segment's (P, Q) and publishes that to global memory (value + epoch in
//            one 8-byte store)
//   look-back the carry entering the segment is the fold of the maps of all LATER segments of
//            the same env block; each warp polls a contiguous range of them (they were
//            scheduled earlier because segments are ordered latest-first by block index, and
//            their maps do not depend on any carry, so there is no serial chain between CTAs)
//   phase B  fold the in-CTA maps to get the carry entering this warp's chunk
//   phase C  replay the chunk from the true carry with the reference's exact op order and
//            write returns[t].
// Each input element is read from HBM exactly once and each output written once:
// 16 B/step (GAE), 20 B/step with bad_masks.
#include "ppd_common.cuh"
#include "tma_utils.cuh"

namespace {

constexpr int kSteps = 16;                // steps per warp chunk (held in registers)
constexpr int kHeaderBytes = 128;
int g_warps = 8;                          // warps per CTA: 4, 8 or 16 (ppd_compute_returns_set_tuning)
int g_min_blocks = 3;                     // __launch_bounds__ min blocks per SM for the 8-warp variant: 3 or 4
// The persistent TMA variant is kept selectable but is NOT the default: measured on B200 at 4096 x 2048 it
// reaches 40 us (2 CTAs/SM, 2-deep ring) against 36.9 us for the register kernel below -- per-item latency
// (look-back + dependent fold + replay), not memory-level parallelism, is what bounds both.
int g_tma_mode = 0;                       // 0 = never (default), 1 = auto for >= 1M steps, 2 = always
int g_tma_stages = 2;                     // ring depth per CTA
int g_tma_ctas = 2;                       // persistent CTAs per SM

// Workspace header.  The workspace must be zero-filled once when it is allocated; every launch
// leaves it ready for the next one (the last CTA to finish bumps the epoch that tags published maps),
// so no memset is enqueued per call.
struct Header {
    unsigned unused;
    unsigned done;     // CTAs that finished
    unsigned epoch;    // number of completed launches; flags of the running launch carry epoch + 1
};

// A segment's affine map is published per lane as two 8-byte words {P, epoch} and {Q, epoch}:
// an aligned 8-byte store is single-copy atomic, so a reader that sees the epoch sees the value
// (no separate flag, no fence, one round trip per look-back step).
__device__ __forceinline__ void publish(uint2* slot, float P, float Q, unsigned epoch) {
    uint4 v = make_uint4(__float_as_uint(P), epoch, __float_as_uint(Q), epoch);
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(slot), "r"(v.x), "r"(v.y) : "memory");
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(slot + 1), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void wait_map(const uint2* slot, unsigned epoch, float& P, float& Q) {
    uint2 a, b;
    do {
        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(a.x), "=r"(a.y) : "l"(slot) : "memory");
        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(b.x), "=r"(b.y) : "l"(slot + 1) : "memory");
    } while (a.y != epoch || b.y != epoch);
    P = __uint_as_float(a.x);
    Q = __uint_as_float(b.x);
}

// Loads of one warp-chunk (kSteps steps of one env per lane, all issued before any use) and phase A:
// the chunk evaluated with zero incoming carry -> affine map (P, Q); per-step values kept for the replay.
template <bool GAE, bool PROPER, bool FULL>
__device__ __forceinline__ void load_and_fold(const float* __restrict__ rewards, const float* __restrict__ value_preds,
                                              const float* __restrict__ masks, const float* __restrict__ bad_masks,
                                              int T, int N, int n, int t0, float nv, float g, float gl, bool live,
                                              float (&f0)[kSteps], float (&f1)[kSteps], float (&f2)[kSteps],
                                              float (&f3)[kSteps], float& P, float& Q) {
    float r_[kSteps], v_[kSteps + 1], m_[kSteps], b_[kSteps];
    if (FULL) {
        const size_t o0 = (size_t)t0 * N + n;
        const float* pr = rewards + o0;
        const float* pm = masks + o0 + N;
        const float* pb = PROPER ? bad_masks + o0 + N : nullptr;
        const float* pv = value_preds + o0;
#pragma unroll
        for (int i = 0; i < kSteps; ++i) {
            r_[i] = __ldg(pr + (size_t)i * N);
            m_[i] = __ldg(pm + (size_t)i * N);                      // m_{t+1}
            if (PROPER) b_[i] = __ldg(pb + (size_t)i * N);          // b_{t+1}
            if (GAE || PROPER) v_[i] = pv[(size_t)i * N];
        }
        if (GAE) v_[kSteps] = (t0 + kSteps >= T) ? nv : pv[(size_t)kSteps * N];
    } else {
#pragma unroll
        for (int i = 0; i < kSteps; ++i) {
            const int t = t0 + i;
            const bool ok = live && t >= 0;
            const size_t o = (size_t)(ok ? t : 0) * N + (live ? n : 0);
            r_[i] = ok ? __ldg(rewards + o) : 0.f;
            m_[i] = ok ? __ldg(masks + o + N) : 1.f;
            if (PROPER) b_[i] = ok ? __ldg(bad_masks + o + N) : 1.f;
            if (GAE || PROPER) v_[i] = ok ? value_preds[o] : 0.f;
        }
        if (GAE) {
            const int tt = t0 + kSteps;
            v_[kSteps] = (tt >= T) ? nv : ((live && tt >= 0) ? value_preds[(size_t)tt * N + n] : 0.f);
        }
    }
    P = 1.f; Q = 0.f;
#pragma unroll
    for (int i = kSteps - 1; i >= 0; --i) {
        const bool in = FULL || (t0 + i) >= 0;
        if (GAE) {
            const float delta = (r_[i] + (g * v_[i + 1]) * m_[i]) - v_[i];   // storage.py:93-95
            const float coef = gl * m_[i];                                    // storage.py:96-97
            f0[i] = delta; f1[i] = coef; f2[i] = v_[i];
            if (PROPER) f3[i] = b_[i];
            if (in) {
                Q = delta + coef * Q;
                P = coef * P;
                if (PROPER) { Q = Q * b_[i]; P = P * b_[i]; }
            }
        } else {
            f0[i] = r_[i]; f1[i] = m_[i];
            if (PROPER) { f2[i] = v_[i]; f3[i] = b_[i]; }
            if (in) {
                Q = (Q * g) * m_[i] + r_[i];
                P = (P * g) * m_[i];
                if (PROPER) { Q = Q * b_[i] + (1.f - b_[i]) * v_[i]; P = P * b_[i]; }
            }
        }
    }
}

// Phase C: replay of the chunk from the true carry x, in the reference's exact operation order.
template <bool GAE, bool PROPER, bool FULL>
__device__ __forceinline__ void replay_and_store(float* __restrict__ returns, int N, int n, int t0, float g, bool live,
                                                 const float (&f0)[kSteps], const float (&f1)[kSteps],
                                                 const float (&f2)[kSteps], const float (&f3)[kSteps], float x) {
    float* po = returns + (size_t)(FULL ? t0 : 0) * N + n;
#pragma unroll
    for (int i = kSteps - 1; i >= 0; --i) {
        const int t = t0 + i;
        if (!FULL && t < 0) continue;
        float out;
        if (GAE) {
            x = f0[i] + f1[i] * x;                     // gae = delta + gamma*lambda*m*gae
            if (PROPER) x = x * f3[i];                 // gae = gae * bad_mask   (storage.py:98)
            out = x + f2[i];                           // returns = gae + V_t
        } else {
            x = (x * g) * f1[i] + f0[i];               // storage.py:120-121
            if (PROPER) x = x * f3[i] + (1.f - f3[i]) * f2[i];   // storage.py:104-105
            out = x;
        }
        if (FULL) __stcs(po + (size_t)i * N, out);
        else if (live) __stcs(returns + (size_t)t * N + n, out);
    }
}

template <bool GAE, bool PROPER, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
returns_scan_kernel(const float* __restrict__ rewards, float* __restrict__ value_preds,
                    const float* __restrict__ masks, const float* __restrict__ bad_masks,
                    float* __restrict__ returns, const float* __restrict__ next_value,
                    int T, int N, float g, float gl, int nblk, int nseg,
                    Header* __restrict__ hdr, uint2* __restrict__ seg_pq) {
    __shared__ float sP[kWarps][32];
    __shared__ float sQ[kWarps][32];
    __shared__ float lP[kWarps][32];
    __shared__ float lQ[kWarps][32];
    // Segments are ordered latest-first by block index; like CUB's decoupled look-back this relies on the
    // hardware dispatching CTAs in block-index order (a CTA only ever waits for lower-indexed CTAs).  The epoch
    // load is independent of the data loads below, so its latency is hidden behind them.
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&hdr->epoch) + 1u;
    const int seg = blockIdx.x / nblk;        // 0 = latest segment in time
    const int blk = blockIdx.x - seg * nblk;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int n = blk * 32 + lane;
    const bool live = n < N;
    const float nv = live ? next_value[n] : 0.f;
    constexpr int kSeg = kSteps * kWarps;     // steps per CTA
    const int t_hi = T - seg * kSeg;          // exclusive upper step of this segment
    if (seg == 0 && w == 0 && live) {
        if (GAE) value_preds[(size_t)T * N + n] = nv;   // storage.py:90,108
        else     returns[(size_t)T * N + n] = nv;       // storage.py:101,118
    }

    const int t0 = t_hi - (w + 1) * kSteps;   // first step of this warp's chunk (may be < 0)
    // FULL chunk: every step in range and every lane a real env -> no per-element predicates or clamps
    const bool full = (t0 >= 0) && (blk * 32 + 32 <= N);
    float f0[kSteps], f1[kSteps], f2[kSteps], f3[kSteps];
    float P, Q;
    if (full) load_and_fold<GAE, PROPER, true>(rewards, value_preds, masks, bad_masks, T, N, n, t0, nv, g, gl, true, f0, f1, f2, f3, P, Q);
    else      load_and_fold<GAE, PROPER, false>(rewards, value_preds, masks, bad_masks, T, N, n, t0, nv, g, gl, live, f0, f1, f2, f3, P, Q);
    sP[w][lane] = P;
    sQ[w][lane] = Q;
    __syncthreads();
    // ---- publish this segment's map (warp 0): fold of the kWarps chunk maps, latest chunk first
    if (w == 0 && seg + 1 < nseg) {
        float Ps = 1.f, Qs = 0.f;
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            Qs = sP[ww][lane] * Qs + sQ[ww][lane];
            Ps = sP[ww][lane] * Ps;
        }
        publish(seg_pq + (((size_t)blk * nseg + seg) * 32 + lane) * 2, Ps, Qs, epoch);
    }
    // ---- look-back: maps of the later segments 0..seg-1, split over the warps in contiguous
    //      ranges (composition is associative, not commutative), then folded in order
    {
        const int per = (seg + kWarps - 1) / kWarps;
        const int s_begin = min(seg, w * per), s_end = min(seg, s_begin + per);
        float Pw = 1.f, Qw = 0.f;
        for (int s = s_begin; s < s_end; ++s) {
            float Pm, Qm;
            wait_map(seg_pq + (((size_t)blk * nseg + s) * 32 + lane) * 2, epoch, Pm, Qm);
            Qw = Pm * Qw + Qm;
            Pw = Pm * Pw;
        }
        lP[w][lane] = Pw;
        lQ[w][lane] = Qw;
    }
    __syncthreads();
    float x = GAE ? 0.f : nv;   // X at T: gae accumulator (storage.py:91,109) or returns[T]
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) x = lP[ww][lane] * x + lQ[ww][lane];
    // ---- phase B: carry entering this warp's chunk
    for (int ww = 0; ww < w; ++ww) x = sP[ww][lane] * x + sQ[ww][lane];
    // ---- phase C: replay with the reference's exact operation order
    if (full) replay_and_store<GAE, PROPER, true>(returns, N, n, t0, g, true, f0, f1, f2, f3, x);
    else      replay_and_store<GAE, PROPER, false>(returns, N, n, t0, g, live, f0, f1, f2, f3, x);
    // ---- last CTA out re-arms the workspace for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&hdr->done, 1u);
        if (prev == gridDim.x - 1) {
            hdr->done = 0;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(&hdr->epoch) = epoch;
        }
    }
}

// =====================================================================================================
// Persistent TMA variant (large rollouts): one CTA per SM walks the (segment, env-block) items in
// latest-first order; the r / V / m (/ b) tiles of the next kTmaStages items are already in flight as
// cp.async.bulk.tensor.2d loads into a shared-memory ring (mbarrier expect_tx), so HBM streams
// continuously while the current item is folded, looked back and replayed.  Same maths, same carries.
// =====================================================================================================
constexpr int kTW = 8;                       // warps per CTA
constexpr int kTSeg = kSteps * kTW;          // 128 steps per item
constexpr int kMaxTmaStages = 4;
constexpr int kTileRows = kTSeg;             // r, m, b tiles: 128 rows x 32 envs
constexpr int kVRows = kTSeg + 1;            // V tile: 129 rows (V_t and V_{t+1})
constexpr int kVRowsPad = kTSeg + 2;

struct TmaArgs {
    float* value_preds; float* returns; const float* next_value;
    int T, N; float g, gl; int nblk, nseg; Header* hdr; uint2* seg_pq; int stages;
};

template <bool GAE, bool PROPER>
__global__ void __launch_bounds__(kTW * 32, 2)
returns_scan_tma_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmB, const TmaArgs a) {
    namespace tma = ppd::tma;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxTmaStages];
    __shared__ float sP[kTW][32], sQ[kTW][32], lP[kTW][32], lQ[kTW][32];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    constexpr uint32_t kTileBytes = kTileRows * 32 * 4, kVBytes = kVRows * 32 * 4, kVBytesPad = kVRowsPad * 32 * 4;
    constexpr uint32_t kStageBytes = kTileBytes + kVBytesPad + kTileBytes + (PROPER ? kTileBytes : 0);
    constexpr uint32_t kTxBytes = kTileBytes + ((GAE || PROPER) ? kVBytes : 0) + kTileBytes + (PROPER ? kTileBytes : 0);
    const int S = a.stages;
    const int T = a.T, N = a.N, nblk = a.nblk, nseg = a.nseg;
    const int items = nblk * nseg;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&a.hdr->epoch) + 1u;

    auto issue = [&](int k_local) {
        const int item = blockIdx.x + k_local * gridDim.x;
        if (item >= items) return;
        const int seg = item / nblk, blk = item - seg * nblk;
        const int t_lo = T - (seg + 1) * kTSeg;           // may be negative: those rows are zero-filled
        const int st = k_local % S;
        uint8_t* base = smem + (size_t)st * kStageBytes;
        tma::mbar_expect_tx(&full_bar[st], kTxBytes);
        tma::load_2d(&tmR, &full_bar[st], base, blk * 32, t_lo);
        if (GAE || PROPER) tma::load_2d(&tmV, &full_bar[st], base + kTileBytes, blk * 32, t_lo);
        tma::load_2d(&tmM, &full_bar[st], base + kTileBytes + kVBytesPad, blk * 32, t_lo + 1);     // m_{t+1}
        if (PROPER) tma::load_2d(&tmB, &full_bar[st], base + 2 * kTileBytes + kVBytesPad, blk * 32, t_lo + 1);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) tma::mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmM) : "memory");
        for (int k = 0; k < S; ++k) issue(k);
    }
    __syncthreads();

    float nv_next = 0.f;
    if ((int)blockIdx.x < items) {
        const int n0 = ((int)blockIdx.x % nblk) * 32 + lane;
        nv_next = (n0 < N) ? __ldg(a.next_value + n0) : 0.f;
    }
    for (int k = 0;; ++k) {
        const int item = blockIdx.x + k * gridDim.x;
        if (item >= items) break;
        const int seg = item / nblk, blk = item - seg * nblk;
        const int n = blk * 32 + lane;
        const bool live = n < N;
        const float nv = nv_next;
        {   // next item's bootstrap value: its latency hides behind this item
            const int item2 = item + gridDim.x;
            if (item2 < items) {
                const int n2 = (item2 % nblk) * 32 + lane;
                nv_next = (n2 < N) ? __ldg(a.next_value + n2) : 0.f;
            }
        }
        const int t_hi = T - seg * kTSeg;
        const int t_lo = t_hi - kTSeg;
        if (seg == 0 && w == 0 && live) {
            if (GAE) a.value_preds[(size_t)T * N + n] = nv;
            else     a.returns[(size_t)T * N + n] = nv;
        }
        // ---- look-back first: it does not need this item's data (which is still landing)
        {
            const int per = (seg + kTW - 1) / kTW;
            const int s_begin = min(seg, w * per), s_end = min(seg, s_begin + per);
            float Pw = 1.f, Qw = 0.f;
            for (int s = s_begin; s < s_end; ++s) {
                float Pm, Qm;
                wait_map(a.seg_pq + (((size_t)blk * nseg + s) * 32 + lane) * 2, epoch, Pm, Qm);
                Qw = Pm * Qw + Qm;
                Pw = Pm * Pw;
            }
            lP[w][lane] = Pw;
            lQ[w][lane] = Qw;
        }
        // ---- this warp's 16 steps from the landed tiles
        const int st = k % S;
        tma::mbar_wait(&full_bar[st], (uint32_t)(k / S) & 1u);
        const float* tR = reinterpret_cast<const float*>(smem + (size_t)st * kStageBytes);
        const float* tV = tR + kTileRows * 32;
        const float* tM = tV + kVRowsPad * 32;
        const float* tB = tM + kTileRows * 32;
        const int r0 = kTSeg - (w + 1) * kSteps;          // first local row of this warp's chunk
        const int t0 = t_lo + r0;
        float f0[kSteps], f1[kSteps], f2[kSteps], f3[kSteps];
        float P = 1.f, Q = 0.f;
        {
            float r_[kSteps], v_[kSteps + 1], m_[kSteps], b_[kSteps];
#pragma unroll
            for (int i = 0; i < kSteps; ++i) {
                r_[i] = tR[(r0 + i) * 32 + lane];
                m_[i] = tM[(r0 + i) * 32 + lane];
                if (PROPER) b_[i] = tB[(r0 + i) * 32 + lane];
                if (GAE || PROPER) v_[i] = tV[(r0 + i) * 32 + lane];
            }
            if (GAE) v_[kSteps] = (t0 + kSteps >= T) ? nv : tV[(r0 + kSteps) * 32 + lane];
#pragma unroll
            for (int i = kSteps - 1; i >= 0; --i) {
                const bool in = (t0 + i) >= 0;
                if (GAE) {
                    const float delta = (r_[i] + (a.g * v_[i + 1]) * m_[i]) - v_[i];
                    const float coef = a.gl * m_[i];
                    f0[i] = delta; f1[i] = coef; f2[i] = v_[i];
                    if (PROPER) f3[i] = b_[i];
                    if (in) {
                        Q = delta + coef * Q;
                        P = coef * P;
                        if (PROPER) { Q = Q * b_[i]; P = P * b_[i]; }
                    }
                } else {
                    f0[i] = r_[i]; f1[i] = m_[i];
                    if (PROPER) { f2[i] = v_[i]; f3[i] = b_[i]; }
                    if (in) {
                        Q = (Q * a.g) * m_[i] + r_[i];
                        P = (P * a.g) * m_[i];
                        if (PROPER) { Q = Q * b_[i] + (1.f - b_[i]) * v_[i]; P = P * b_[i]; }
                    }
                }
            }
        }
        sP[w][lane] = P;
        sQ[w][lane] = Q;
        // generic-proxy reads of the stage above, async-proxy (TMA) refill below: write-after-read across proxies needs this fence
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();       // every warp has copied its rows out of the stage; sP/sQ/lP/lQ are complete
        if (threadIdx.x == 0) issue(k + S);                // refill this ring slot
        if (w == 0 && seg + 1 < nseg) {
            float Ps = 1.f, Qs = 0.f;
#pragma unroll
            for (int ww = 0; ww < kTW; ++ww) {
                Qs = sP[ww][lane] * Qs + sQ[ww][lane];
                Ps = sP[ww][lane] * Ps;
            }
            publish(a.seg_pq + (((size_t)blk * nseg + seg) * 32 + lane) * 2, Ps, Qs, epoch);
        }
        float x = GAE ? 0.f : nv;
#pragma unroll
        for (int ww = 0; ww < kTW; ++ww) x = lP[ww][lane] * x + lQ[ww][lane];
        for (int ww = 0; ww < w; ++ww) x = sP[ww][lane] * x + sQ[ww][lane];
        replay_and_store<GAE, PROPER, false>(a.returns, N, n, t0, a.g, live, f0, f1, f2, f3, x);
        __syncthreads();       // sP/sQ/lP/lQ are free for the next item
    }
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&a.hdr->done, 1u);
        if (prev == gridDim.x - 1) {
            a.hdr->done = 0;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(&a.hdr->epoch) = epoch;
        }
    }
}

struct Plan { int nblk, nseg; size_t total; };
Plan plan(int T, int N, int warps) {
    Plan p;
    const int kSeg = kSteps * warps;
    p.nblk = (N + 31) / 32;
    p.nseg = (T + kSeg - 1) / kSeg;
    p.total = kHeaderBytes + (size_t)p.nblk * p.nseg * 32 * 2 * sizeof(uint2);
    return p;
}

}  // namespace

extern "C" size_t ppd_compute_returns_workspace(int T, int N) {
    if (T <= 0 || N <= 0) return 0;
    return plan(T, N, 4).total;          // the smallest segment size needs the most carry slots
}

extern "C" void ppd_compute_returns_set_tuning(int warps, int min_blocks) {
    if (warps == 4 || warps == 8 || warps == 16) g_warps = warps;
    if (min_blocks == 3 || min_blocks == 4) g_min_blocks = min_blocks;
    if (warps >= 100 && warps < 200) g_tma_mode = warps - 100;      // 100 = never TMA, 101 = auto, 102 = always
    if (warps >= 200) { g_tma_stages = (warps - 200) / 10; g_tma_ctas = (warps - 200) % 10; }   // 2SC: stages S, CTAs/SM C
}

namespace {
template <bool GAE, bool PROPER>
int launch_tma(const CUtensorMap& mR, const CUtensorMap& mV, const CUtensorMap& mM, const CUtensorMap& mB, const TmaArgs& a,
               int grid, cudaStream_t s) {
    constexpr size_t tile = kTileRows * 32 * 4, vpad = kVRowsPad * 32 * 4;
    constexpr size_t stage = tile + vpad + tile + (PROPER ? tile : 0);
    const size_t smem = (size_t)a.stages * stage + 128;
    cudaError_t e = cudaFuncSetAttribute(returns_scan_tma_kernel<GAE, PROPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ppd::set_error("ppd_compute_returns: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    returns_scan_tma_kernel<GAE, PROPER><<<grid, kTW * 32, smem, s>>>(mR, mV, mM, mB, a);
    return ppd::launch_status("returns_scan_tma_kernel");
}
}  // namespace

extern "C" int ppd_compute_returns(const float* rewards, float* value_preds, const float* masks,
                                   const float* bad_masks, float* returns, const float* next_value,
                                   int T, int N, double gamma, double gae_lambda, int use_gae,
                                   int use_proper_time_limits, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    PPD_REQUIRE(rewards && value_preds && masks && returns && next_value && workspace, "null pointer");
    PPD_REQUIRE(!use_proper_time_limits || bad_masks, "bad_masks required with use_proper_time_limits");
    PPD_REQUIRE(T > 0 && N > 0, "T and N must be positive");
    const Plan p = plan(T, N, g_warps);
    if (workspace_bytes < p.total) {
        ppd::set_error("ppd_compute_returns: workspace too small");
        return PPD_EWORKSPACE;
    }
    PPD_REQUIRE((uintptr_t)workspace % 16 == 0, "workspace must be 16-byte aligned");
    PPD_REQUIRE((int64_t)p.nblk * p.nseg <= 0x7fffffffLL, "grid too large");
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    cudaStream_t s = ppd::as_stream(stream);
    char* ws = reinterpret_cast<char*>(workspace);
    Header* hdr = reinterpret_cast<Header*>(ws);
    uint2* pq = reinterpret_cast<uint2*>(ws + kHeaderBytes);
    // ---- persistent TMA kernel for large rollouts (row stride must be a multiple of 16 bytes)
    const bool tma_ok = (N % 4 == 0) && (((uintptr_t)rewards | (uintptr_t)value_preds | (uintptr_t)masks |
                                           (uintptr_t)(use_proper_time_limits ? bad_masks : masks)) % 16 == 0);
    if (tma_ok && (g_tma_mode == 2 || (g_tma_mode == 1 && (int64_t)T * N >= (1 << 20)))) {
        const Plan pt = plan(T, N, kTW);
        CUtensorMap mR, mV, mM, mB;
        const float* bm = use_proper_time_limits ? bad_masks : masks;
        if (ppd::tma::make_map_2d(&mR, rewards, T, N, N, 32, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE) &&
            ppd::tma::make_map_2d(&mV, value_preds, T + 1, N, N, 32, kVRows, CU_TENSOR_MAP_SWIZZLE_NONE) &&
            ppd::tma::make_map_2d(&mM, masks, T + 1, N, N, 32, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE) &&
            ppd::tma::make_map_2d(&mB, bm, T + 1, N, N, 32, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE)) {
            const size_t tile = kTileRows * 32 * 4, vpad = kVRowsPad * 32 * 4;
            const size_t stage = tile + vpad + tile + (use_proper_time_limits ? tile : 0);
            int ctas = g_tma_ctas < 1 ? 1 : (g_tma_ctas > 2 ? 2 : g_tma_ctas);
            int stages = g_tma_stages < 1 ? 1 : (g_tma_stages > kMaxTmaStages ? kMaxTmaStages : g_tma_stages);
            while (stages > 1 && (size_t)ctas * (stages * stage + 2048) > 226 * 1024) --stages;
            while (ctas > 1 && (size_t)ctas * (stages * stage + 2048) > 226 * 1024) --ctas;
            TmaArgs a{value_preds, returns, next_value, T, N, g, gl, pt.nblk, pt.nseg, hdr, pq, stages};
            int grid_t = pt.nblk * pt.nseg;
            if (grid_t > ctas * ppd::kNumSMs) grid_t = ctas * ppd::kNumSMs;
            if (use_gae) return use_proper_time_limits ? launch_tma<true, true>(mR, mV, mM, mB, a, grid_t, s)
                                                       : launch_tma<true, false>(mR, mV, mM, mB, a, grid_t, s);
            return use_proper_time_limits ? launch_tma<false, true>(mR, mV, mM, mB, a, grid_t, s)
                                          : launch_tma<false, false>(mR, mV, mM, mB, a, grid_t, s);
        }
    }
    dim3 grid((unsigned)(p.nblk * p.nseg)), block(g_warps * 32);
#define PPD_LAUNCH2(G, P, W, MB) \
    returns_scan_kernel<G, P, W, MB><<<grid, block, 0, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, \
                                                            T, N, g, gl, p.nblk, p.nseg, hdr, pq)
#define PPD_LAUNCH(G, P)                                          \
    do {                                                          \
        if (g_warps == 4) PPD_LAUNCH2(G, P, 4, 6);                \
        else if (g_warps == 16) PPD_LAUNCH2(G, P, 16, 1);         \
        else if (g_min_blocks == 4) PPD_LAUNCH2(G, P, 8, 4);      \
        else PPD_LAUNCH2(G, P, 8, 3);                             \
    } while (0)
    if (use_gae) { if (use_proper_time_limits) PPD_LAUNCH(true, true); else PPD_LAUNCH(true, false); }
    else         { if (use_proper_time_limits) PPD_LAUNCH(false, true); else PPD_LAUNCH(false, false); }
#undef PPD_LAUNCH
#undef PPD_LAUNCH2
    return ppd::launch_status("ppd_compute_returns");
}
