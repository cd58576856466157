// Minibatch gathers: feed_forward_generator (PKG/storage.py:123-160) and
// recurrent_generator (PKG/storage.py:162-223) as one permutation-gather kernel.
//
// Storage is time-major: row r = t*N + n of every field.  One CTA copies one chunk of one
// output row; the big field (obs, C*84*84 fp32 = 28 KB per channel plane) moves as 16-byte
// vectors with kUnroll independent loads in flight per thread, streaming past L1.  The CTA
// with chunk index 0 also moves the row's small fields (vector obs, hidden state, action,
// value, return, mask, old log-prob, advantage).  HBM traffic = 2 x row bytes per sample.
// Compile with -fmad=false (fused advantage normalisation mirrors ppo.py:35-37 op order).
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 7;                         // 256*7 = 1792 float4 >= 1764 = one 84x84 plane
constexpr int kChunkVec = kThreads * kUnroll;      // float4 per CTA
constexpr int kChunkScalar = kThreads * kUnroll;   // floats per CTA on the unaligned path

struct Args {
    ppd_gather_desc d;
    const int64_t* perm;
    int64_t perm_off;
    int64_t rows;
    int T, N, E;          // E == 0 -> feed-forward mode
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const Args a) {
    const int64_t i = blockIdx.x;                  // output row
    int64_t src;                                   // source row in [0, T*N)
    int64_t t = 0, j = 0;
    if (a.E == 0) {
        src = __ldg(a.perm + a.perm_off + i);                          // storage.py:142-143
    } else {
        t = i / a.E;
        j = i - t * a.E;
        src = t * a.N + __ldg(a.perm + a.perm_off + j);                // storage.py:182-183, 197, 212
    }
    const ppd_gather_desc& d = a.d;

    if (d.obs) {
        if (VEC) {
            const int64_t nvec = d.obs_row >> 2;
            const float4* s = reinterpret_cast<const float4*>(d.obs + src * d.obs_row);
            float4* o = reinterpret_cast<float4*>(d.obs_out + i * d.obs_row);
            const int64_t base = (int64_t)blockIdx.y * kChunkVec + threadIdx.x;
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t k = base + (int64_t)u * kThreads;
                if (k < nvec) v[u] = ppd::ldg_stream(s + k);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t k = base + (int64_t)u * kThreads;
                if (k < nvec) ppd::stg_stream(o + k, v[u]);
            }
        } else {
            const float* s = d.obs + src * d.obs_row;
            float* o = d.obs_out + i * d.obs_row;
            const int64_t base = (int64_t)blockIdx.y * kChunkScalar + threadIdx.x;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t k = base + (int64_t)u * kThreads;
                if (k < d.obs_row) o[k] = __ldg(s + k);
            }
        }
    }
    if (blockIdx.y != 0) return;

    // ---- small fields of this row
    if (d.vobs && d.vobs_row > 0) {
        const int64_t ld = d.vobs_out_ld ? d.vobs_out_ld : d.vobs_row;
        for (int64_t k = threadIdx.x; k < d.vobs_row; k += kThreads)
            d.vobs_out[i * ld + k] = __ldg(d.vobs + src * d.vobs_row + k);
    }
    if (d.hxs) {
        if (a.E == 0) {
            for (int64_t k = threadIdx.x; k < d.hxs_row; k += kThreads)
                d.hxs_out[i * d.hxs_row + k] = __ldg(d.hxs + src * d.hxs_row + k);
        } else if (t == 0) {                                           // storage.py:186,208-209
            for (int64_t k = threadIdx.x; k < d.hxs_row; k += kThreads)
                d.hxs_out[j * d.hxs_row + k] = __ldg(d.hxs + src * d.hxs_row + k);
        }
    }
    if (d.actions) {
        for (int64_t k = threadIdx.x; k < d.actions_row; k += kThreads)
            d.actions_out[i * d.actions_row + k] = __ldg(d.actions + src * d.actions_row + k);
    }
    if (threadIdx.x == 0) {
        float v = 0.f, r = 0.f;
        if (d.value_preds) { v = __ldg(d.value_preds + src); if (d.value_preds_out) d.value_preds_out[i] = v; }
        if (d.returns)     { r = __ldg(d.returns + src);     if (d.returns_out) d.returns_out[i] = r; }
        if (d.masks)  d.masks_out[i] = __ldg(d.masks + src);
        if (d.logp)   d.logp_out[i] = __ldg(d.logp + src);
        if (d.adv_out) {
            if (d.adv) d.adv_out[i] = __ldg(d.adv + src);
            else if (d.adv_stats) d.adv_out[i] = __fdiv_rn((r - v) - __ldg(d.adv_stats), __ldg(d.adv_stats + 1));
        }
    }
}

int launch(const ppd_gather_desc* d, const int64_t* perm, int64_t off, int64_t rows, int T, int N, int E,
           void* stream, const char* what) {
    if (!d || !perm) { ppd::set_error("%s: null pointer", what); return PPD_EINVAL; }
    if (rows <= 0 || T <= 0 || N <= 0) { ppd::set_error("%s: sizes must be positive", what); return PPD_EINVAL; }
    if (rows > 0x7fffffffLL) { ppd::set_error("%s: too many rows", what); return PPD_EINVAL; }
    if ((d->obs != nullptr) != (d->obs_out != nullptr) || (d->vobs != nullptr) != (d->vobs_out != nullptr) ||
        (d->hxs != nullptr) != (d->hxs_out != nullptr) || (d->actions != nullptr) != (d->actions_out != nullptr) ||
        (d->masks != nullptr) != (d->masks_out != nullptr) || (d->logp != nullptr) != (d->logp_out != nullptr)) {
        ppd::set_error("%s: source and destination of a field must both be set or both be NULL", what);
        return PPD_EINVAL;
    }
    if (d->adv_out && !d->adv && !(d->adv_stats && d->returns && d->value_preds)) {
        ppd::set_error("%s: adv_out needs adv, or adv_stats with returns and value_preds", what);
        return PPD_EINVAL;
    }
    Args a;
    a.d = *d;
    a.perm = perm; a.perm_off = off; a.rows = rows; a.T = T; a.N = N; a.E = E;
    const bool vec = d->obs && (d->obs_row % 4 == 0) && ((uintptr_t)d->obs % 16 == 0) && ((uintptr_t)d->obs_out % 16 == 0);
    int64_t chunks = 1;
    if (d->obs) chunks = vec ? ((d->obs_row / 4 + kChunkVec - 1) / kChunkVec) : ((d->obs_row + kChunkScalar - 1) / kChunkScalar);
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) { ppd::set_error("%s: obs row too large", what); return PPD_EINVAL; }
    dim3 grid((unsigned)rows, (unsigned)chunks);
    cudaStream_t s = ppd::as_stream(stream);
    if (vec) gather_rows_kernel<true><<<grid, kThreads, 0, s>>>(a);
    else     gather_rows_kernel<false><<<grid, kThreads, 0, s>>>(a);
    return ppd::launch_status(what);
}

}  // namespace

extern "C" int ppd_gather_feed_forward(const ppd_gather_desc* d, const int64_t* perm, int64_t mb_start,
                                       int64_t rows, int T, int N, void* stream) {
    return launch(d, perm, mb_start, rows, T, N, 0, stream, "ppd_gather_feed_forward");
}

extern "C" int ppd_gather_recurrent(const ppd_gather_desc* d, const int64_t* env_perm, int64_t env_start,
                                    int E, int T, int N, void* stream) {
    if (E <= 0) { ppd::set_error("ppd_gather_recurrent: E must be positive"); return PPD_EINVAL; }
    return launch(d, env_perm, env_start, (int64_t)T * E, T, N, E, stream, "ppd_gather_recurrent");
}
