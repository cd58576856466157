// Returns / GAE as a chunked affine scan, one env per lane.
// Replaces RolloutStorage.compute_returns (PKG/storage.py:82-121).  Compile with -fmad=false:
// the in-chunk replay keeps the reference's operation order (separate mul / add roundings).
//
// Layout: every field is time-major [T(+1), N]; for fixed t the N envs are contiguous, so a
// warp reading lane = env issues one 128-byte request per (field, t).
//
// Work split: CTA = 32 envs x kWarps warps.  Time is consumed from the end in super-chunks of
// kWarps*kSteps steps; inside a super-chunk warp w owns kSteps consecutive steps.
//   phase A  each lane loads its kSteps x {r, V, m, (b)} into registers (all loads issued up
//            front -> kSteps*3 requests in flight per thread), evaluates its chunk with zero
//            incoming carry and publishes the chunk's affine map (P, Q): X_out = P*X_in + Q
//   phase B  every lane walks the <= kWarps maps of its env column in order to get the carry
//            entering its own chunk (and the carry leaving the super-chunk)
//   phase C  the chunk is replayed from the true carry with the reference's exact op order
//            and returns[t] is written.
// Each input element is read from HBM exactly once and each output written once:
// 16 B/step (GAE), 20 B/step with bad_masks.
#include "ppd_common.cuh"

namespace {

constexpr int kSteps = 16;
constexpr int kWarps = 16;

template <bool GAE, bool PROPER>
__global__ void __launch_bounds__(kWarps * 32)
returns_scan_kernel(const float* __restrict__ rewards, float* __restrict__ value_preds,
                    const float* __restrict__ masks, const float* __restrict__ bad_masks,
                    float* __restrict__ returns, const float* __restrict__ next_value,
                    int T, int N, float g, float gl) {
    __shared__ float sP[kWarps][32];
    __shared__ float sQ[kWarps][32];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const bool live = n < N;
    const float nv = live ? next_value[n] : 0.f;
    float carry = GAE ? 0.f : nv;   // X at t_hi: gae accumulator (storage.py:91,109) or returns[T]
    if (w == 0 && live) {
        if (GAE) value_preds[(size_t)T * N + n] = nv;   // storage.py:90,108
        else     returns[(size_t)T * N + n] = nv;       // storage.py:101,118
    }

    for (int t_hi = T; t_hi > 0; t_hi -= kWarps * kSteps) {
        const int t0 = t_hi - (w + 1) * kSteps;   // first step of this warp's chunk (may be < 0)
        // per-step registers kept from phase A to phase C
        float f0[kSteps], f1[kSteps], f2[kSteps], f3[kSteps];
        // ---- loads (all independent, issued before any use)
        float r_[kSteps], v_[kSteps + 1], m_[kSteps], b_[kSteps];
#pragma unroll
        for (int i = 0; i < kSteps; ++i) {
            const int t = t0 + i;
            const bool ok = live && t >= 0;
            const size_t o = (size_t)(ok ? t : 0) * N + (live ? n : 0);
            r_[i] = ok ? __ldg(rewards + o) : 0.f;
            m_[i] = ok ? __ldg(masks + o + N) : 1.f;            // m_{t+1}
            if (PROPER) b_[i] = ok ? __ldg(bad_masks + o + N) : 1.f;   // b_{t+1}
            if (GAE || PROPER) v_[i] = ok ? value_preds[o] : 0.f;
        }
        if (GAE) {
            const int tt = t0 + kSteps;   // V_{t+1} of the chunk's last step
            v_[kSteps] = (tt >= T) ? nv : ((live && tt >= 0) ? value_preds[(size_t)tt * N + n] : 0.f);
        }
        // ---- phase A: chunk with zero carry
        float P = 1.f, Q = 0.f;
#pragma unroll
        for (int i = kSteps - 1; i >= 0; --i) {
            const bool in = (t0 + i) >= 0;
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
        __syncthreads();   // previous super-chunk's phase B readers are done with sP/sQ
        sP[w][lane] = P;
        sQ[w][lane] = Q;
        __syncthreads();
        // ---- phase B: carry entering this warp's chunk, and leaving the super-chunk
        float x = carry, mine = carry;
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            if (ww == w) mine = x;
            x = sP[ww][lane] * x + sQ[ww][lane];
        }
        carry = x;
        // ---- phase C: replay with the reference's exact operation order
        x = mine;
#pragma unroll
        for (int i = kSteps - 1; i >= 0; --i) {
            const int t = t0 + i;
            if (t < 0) continue;
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
            if (live) __stcs(returns + (size_t)t * N + n, out);
        }
    }
}

}  // namespace

extern "C" int ppd_compute_returns(const float* rewards, float* value_preds, const float* masks,
                                   const float* bad_masks, float* returns, const float* next_value,
                                   int T, int N, double gamma, double gae_lambda, int use_gae,
                                   int use_proper_time_limits, void* stream) {
    PPD_REQUIRE(rewards && value_preds && masks && returns && next_value, "null pointer");
    PPD_REQUIRE(!use_proper_time_limits || bad_masks, "bad_masks required with use_proper_time_limits");
    PPD_REQUIRE(T > 0 && N > 0, "T and N must be positive");
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    dim3 grid((N + 31) / 32), block(kWarps * 32);
    cudaStream_t s = ppd::as_stream(stream);
#define PPD_LAUNCH(G, P) \
    returns_scan_kernel<G, P><<<grid, block, 0, s>>>(rewards, value_preds, masks, bad_masks, returns, next_value, T, N, g, gl)
    if (use_gae) { if (use_proper_time_limits) PPD_LAUNCH(true, true); else PPD_LAUNCH(true, false); }
    else         { if (use_proper_time_limits) PPD_LAUNCH(false, true); else PPD_LAUNCH(false, false); }
#undef PPD_LAUNCH
    return ppd::launch_status("ppd_compute_returns");
}
