// Running observation normalisation: VecNormalize._obfilt (PKG/envs.py:208-217) on top of the
// baselines RunningMeanStd update (Chan et al. parallel moments, float64; third party, unpinned --
// see oracle/running_mean_std.py).  One thread per feature f; for fixed env n the features are
// contiguous, so every pass over n is a coalesced row read.  The batch (N x F fp32) is read
// twice for the moments (second pass is an L2 hit for Obstacle-Tower sizes: 32 x 21168 x 4 B =
// 2.7 MB) and once more for the normalise pass; stats are read and written once.
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
obs_rms_kernel(const float* __restrict__ obs, int N, int64_t F, double* __restrict__ mean,
               double* __restrict__ var, double count, int update, double epsilon, double clipob,
               float* __restrict__ out) {
    const int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (f >= F) return;
    double mu = mean[f], vr = var[f];
    if (update) {
        double s = 0.0;
        for (int n = 0; n < N; ++n) s += (double)__ldg(obs + (int64_t)n * F + f);
        const double bm = s / N;                                   // x.mean(axis=0)
        double q = 0.0;
        for (int n = 0; n < N; ++n) {
            const double d = (double)__ldg(obs + (int64_t)n * F + f) - bm;
            q += d * d;
        }
        const double bv = q / N;                                   // x.var(axis=0)  (ddof = 0)
        const double delta = bm - mu;
        const double tot = count + (double)N;
        const double new_mean = mu + delta * (double)N / tot;
        const double m2 = vr * count + bv * (double)N + delta * delta * count * (double)N / tot;
        mu = new_mean;
        vr = m2 / tot;
        mean[f] = mu;
        var[f] = vr;
    }
    if (out) {
        const double sd = sqrt(vr + epsilon);
        for (int n = 0; n < N; ++n) {
            double y = ((double)__ldg(obs + (int64_t)n * F + f) - mu) / sd;
            y = fmin(fmax(y, -clipob), clipob);
            out[(int64_t)n * F + f] = (float)y;
        }
    }
}

}  // namespace

extern "C" int ppd_obs_rms_update_normalize(const float* obs, int N, int64_t F, double* mean, double* var,
                                            double count_host, int update, double epsilon, double clipob,
                                            float* out, void* stream) {
    PPD_REQUIRE(obs && mean && var, "null pointer");
    PPD_REQUIRE(N > 0 && F > 0, "N and F must be positive");
    const int64_t blocks = (F + kThreads - 1) / kThreads;
    PPD_REQUIRE(blocks <= 0x7fffffffLL, "F too large");
    obs_rms_kernel<<<(unsigned)blocks, kThreads, 0, ppd::as_stream(stream)>>>(obs, N, F, mean, var, count_host,
                                                                             update, epsilon, clipob, out);
    return ppd::launch_status("obs_rms_kernel");
}
