// Advantage statistics and normalisation (PKG/algo/ppo.py:35-37).  Compile with -fmad=false.
//   adv = returns - value_preds (fp32, as the reference); moments accumulated in float64.
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4 * ppd::kNumSMs;

__global__ void __launch_bounds__(kThreads)
adv_moments_partial(const float* __restrict__ ret, const float* __restrict__ val, int64_t n,
                    double* __restrict__ partial) {
    __shared__ double scratch[32];
    double s = 0.0, q = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const float a = __ldg(ret + i) - __ldg(val + i);
        s += (double)a;
        q += (double)a * (double)a;
    }
    s = ppd::block_sum(s, scratch);
    q = ppd::block_sum(q, scratch);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = q;
    }
}

// Fixed-order reduction of the per-block partials -> deterministic moments.
__global__ void __launch_bounds__(kThreads)
adv_moments_final(const double* __restrict__ partial, int nblocks, int64_t n, double* __restrict__ moments) {
    __shared__ double scratch[32];
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) {
        s += partial[2 * i];
        q += partial[2 * i + 1];
    }
    s = ppd::block_sum(s, scratch);
    q = ppd::block_sum(q, scratch);
    if (threadIdx.x == 0) {
        moments[0] = s;
        moments[1] = q;
        moments[2] = (double)n;
    }
}

__global__ void adv_finalize(const double* __restrict__ moments, float* __restrict__ stats) {
    const double s = moments[0], q = moments[1], c = moments[2];
    const double mean = s / c;
    double var = (q - s * mean) / (c - 1.0);      // unbiased (torch.std default), ppo.py:37
    if (!(var > 0.0)) var = (c > 1.0) ? 0.0 : nan("");
    const float stdf = (float)sqrt(var);
    stats[0] = (float)mean;
    stats[1] = stdf + 1e-5f;                       // fp32 add, as `advantages.std() + 1e-5`
}

__global__ void __launch_bounds__(kThreads)
adv_normalize(const float* __restrict__ ret, const float* __restrict__ val, int64_t n,
              const float* __restrict__ stats, float* __restrict__ out) {
    const float mean = stats[0], denom = stats[1];
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const float a = __ldg(ret + i) - __ldg(val + i);
        out[i] = __fdiv_rn(a - mean, denom);
    }
}

int blocks_for(int64_t n) {
    int64_t b = (n + kThreads * 4 - 1) / (kThreads * 4);
    if (b < 1) b = 1;
    if (b > kMaxBlocks) b = kMaxBlocks;
    return (int)b;
}

}  // namespace

extern "C" size_t ppd_advantage_moments_workspace(int64_t n) { return (size_t)2 * blocks_for(n) * sizeof(double); }

extern "C" int ppd_advantage_moments(const float* returns, const float* value_preds, int64_t n,
                                     double* moments, void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(returns && value_preds && moments && workspace, "null pointer");
    PPD_REQUIRE(n > 0, "n must be positive");
    const int nb = blocks_for(n);
    if (workspace_bytes < (size_t)2 * nb * sizeof(double)) {
        ppd::set_error("ppd_advantage_moments: workspace too small");
        return PPD_EWORKSPACE;
    }
    cudaStream_t s = ppd::as_stream(stream);
    adv_moments_partial<<<nb, kThreads, 0, s>>>(returns, value_preds, n, (double*)workspace);
    int rc = ppd::launch_status("adv_moments_partial");
    if (rc) return rc;
    adv_moments_final<<<1, kThreads, 0, s>>>((const double*)workspace, nb, n, moments);
    return ppd::launch_status("adv_moments_final");
}

extern "C" int ppd_advantage_finalize(const double* moments, float* stats, void* stream) {
    PPD_REQUIRE(moments && stats, "null pointer");
    adv_finalize<<<1, 1, 0, ppd::as_stream(stream)>>>(moments, stats);
    return ppd::launch_status("adv_finalize");
}

extern "C" int ppd_advantage_normalize(const float* returns, const float* value_preds, int64_t n,
                                       const float* stats, float* adv_out, void* stream) {
    PPD_REQUIRE(returns && value_preds && stats && adv_out, "null pointer");
    PPD_REQUIRE(n > 0, "n must be positive");
    adv_normalize<<<blocks_for(n), kThreads, 0, ppd::as_stream(stream)>>>(returns, value_preds, n, stats, adv_out);
    return ppd::launch_status("adv_normalize");
}
