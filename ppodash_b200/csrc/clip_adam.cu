// Global-norm gradient clip + Adam over one flat fp32 buffer (PKG/algo/ppo.py:82-84:
// nn.utils.clip_grad_norm_ then torch.optim.Adam.step).  Compile with -fmad=false so the update
// rounds like torch's unfused elementwise sequence.
//   pass 1  per-block sum of squares (fp32 per thread, fp64 across the block)       reads g
//   pass 2  fixed-order reduction of the partials -> total_norm, clip coefficient (1 CTA)
//   pass 3  Adam: reads g, p, m, v; writes p, m, v                                   28 B/param
#include <cooperative_groups.h>

#include "ppd_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kVec = 4;
constexpr int kMaxBlocks = 1024;

struct Scalars {       // lives at the start of the workspace
    float clip_coef;
    float total_norm;
};

int norm_blocks(int64_t n) {
    int64_t b = (n + (int64_t)kThreads * kVec * 4 - 1) / ((int64_t)kThreads * kVec * 4);
    if (b < 1) b = 1;
    if (b > kMaxBlocks) b = kMaxBlocks;
    return (int)b;
}

__global__ void __launch_bounds__(kThreads)
sqnorm_partial(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
    __shared__ double scratch[32];
    float acc = 0.f;
    const int64_t nvec = n / kVec;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
        const float4 x = __ldg(g4 + i);
        acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    if (blockIdx.x == 0) {
        for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    }
    const double s = ppd::block_sum((double)acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads)
norm_final(const double* __restrict__ partial, int nblocks, float max_norm, Scalars* __restrict__ sc,
           float* __restrict__ grad_norm_out, const float* __restrict__ loss_in, float* __restrict__ loss_acc) {
    __shared__ double scratch[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) s += partial[i];
    s = ppd::block_sum(s, scratch);
    if (threadIdx.x == 0) {
        const float total = (float)sqrt(s);
        float coef = 1.f;
        if (max_norm > 0.f) coef = fminf(__fdiv_rn(max_norm, total + 1e-6f), 1.0f);   // clip_grad_norm_
        sc->clip_coef = coef;
        sc->total_norm = total;
        if (grad_norm_out) *grad_norm_out = total;
        if (loss_in && loss_acc) {
            loss_acc[0] += loss_in[0];
            loss_acc[1] += loss_in[1];
            loss_acc[2] += loss_in[2];
        }
    }
}

struct AdamConst {
    float one_minus_b1, b2, one_minus_b2, neg_step_size, bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float coef, const AdamConst& c) {
    g = g * coef;                                       // clip_grad_norm_: grad.mul_(clip_coef_clamped)
    m = m + c.one_minus_b1 * (g - m);                   // exp_avg.lerp_(grad, 1 - beta1)
    v = v * c.b2;                                       // exp_avg_sq.mul_(beta2)
    v = v + (c.one_minus_b2 * g) * g;                   //   .addcmul_(grad, grad, value=1 - beta2)
    const float denom = __fdiv_rn(sqrtf(v), c.bc2_sqrt) + c.eps;
    p = p + c.neg_step_size * __fdiv_rn(m, denom);      // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, const Scalars* __restrict__ sc, AdamConst c) {
    const float coef = sc->clip_coef;
    const int64_t nvec = n / kVec;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        adam_one(pp.x, gg.x, mm.x, vv.x, coef, c);
        adam_one(pp.y, gg.y, mm.y, vv.y, coef, c);
        adam_one(pp.z, gg.z, mm.z, vv.z, coef, c);
        adam_one(pp.w, gg.w, mm.w, vv.w, coef, c);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0) {
        for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += kThreads) adam_one(p[i], g[i], m[i], v[i], coef, c);
    }
}

// Single cooperative launch for parameter sets that fit in L2 (the policy has 2.5 M parameters = 10 MB per
// array): pass 1 (sum of squares) -> grid barrier -> every CTA folds the per-CTA partials in the same fixed
// order -> clip coefficient -> Adam.  One launch instead of three; the gradient is re-read from L2.
__global__ void __launch_bounds__(kThreads)
clip_adam_fused_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       int64_t n, double* __restrict__ partial, Scalars* __restrict__ sc, float max_norm,
                       float* __restrict__ grad_norm_out, const float* __restrict__ loss_in, float* __restrict__ loss_acc,
                       AdamConst c) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double scratch[32];
    __shared__ float s_coef;
    const int64_t nvec = n / kVec;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
        const float4 x = __ldg(g4 + i);
        acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    if (blockIdx.x == 0)
        for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    double s = ppd::block_sum((double)acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    grid.sync();
    s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) s += __ldcg(partial + i);
    s = ppd::block_sum(s, scratch);
    if (threadIdx.x == 0) {
        const float total = (float)sqrt(s);
        float coef = 1.f;
        if (max_norm > 0.f) coef = fminf(__fdiv_rn(max_norm, total + 1e-6f), 1.0f);
        s_coef = coef;
        if (blockIdx.x == 0) {
            sc->clip_coef = coef;
            sc->total_norm = total;
            if (grad_norm_out) *grad_norm_out = total;
            if (loss_in && loss_acc) {
                loss_acc[0] += loss_in[0];
                loss_acc[1] += loss_in[1];
                loss_acc[2] += loss_in[2];
            }
        }
    }
    __syncthreads();
    const float coef = s_coef;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        adam_one(pp.x, gg.x, mm.x, vv.x, coef, c);
        adam_one(pp.y, gg.y, mm.y, vv.y, coef, c);
        adam_one(pp.z, gg.z, mm.z, vv.z, coef, c);
        adam_one(pp.w, gg.w, mm.w, vv.w, coef, c);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0)
        for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += kThreads) adam_one(p[i], g[i], m[i], v[i], coef, c);
}

// RMSprop (torch.optim.RMSprop, centered=False, momentum=0, weight_decay=0 -- what PKG/algo/a2c_acktr.py:30-31 constructs):
//   square_avg.mul_(alpha).addcmul_(grad, grad, value=1 - alpha); avg = square_avg.sqrt().add_(eps); param.addcdiv_(grad, avg, value=-lr)
__device__ __forceinline__ void rmsprop_one(float& p, float g, float& sq, float coef, float alpha, float one_minus_alpha, float eps, float neg_lr) {
    g = g * coef;
    sq = sq * alpha;
    sq = sq + (one_minus_alpha * g) * g;
    const float avg = sqrtf(sq) + eps;
    p = p + neg_lr * __fdiv_rn(g, avg);
}

__global__ void __launch_bounds__(kThreads)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, int64_t n, const Scalars* __restrict__ sc,
               float alpha, float one_minus_alpha, float eps, float neg_lr) {
    const float coef = sc->clip_coef;
    const int64_t nvec = n / kVec;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(sq);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
        float4 pp = p4[i], ss = s4[i];
        const float4 gg = __ldg(g4 + i);
        rmsprop_one(pp.x, gg.x, ss.x, coef, alpha, one_minus_alpha, eps, neg_lr);
        rmsprop_one(pp.y, gg.y, ss.y, coef, alpha, one_minus_alpha, eps, neg_lr);
        rmsprop_one(pp.z, gg.z, ss.z, coef, alpha, one_minus_alpha, eps, neg_lr);
        rmsprop_one(pp.w, gg.w, ss.w, coef, alpha, one_minus_alpha, eps, neg_lr);
        p4[i] = pp; s4[i] = ss;
    }
    if (blockIdx.x == 0)
        for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += kThreads) rmsprop_one(p[i], g[i], sq[i], coef, alpha, one_minus_alpha, eps, neg_lr);
}

constexpr int64_t kFusedMaxParams = 4 << 20;     // 4 arrays x 16 MB stay L2-resident
int g_fused = 1;

size_t ws_bytes(int64_t n) { (void)n; return 256 + (size_t)kMaxBlocks * sizeof(double); }

}  // namespace

extern "C" size_t ppd_clip_adam_workspace(int64_t n) { return ws_bytes(n > 0 ? n : 1); }

extern "C" void ppd_clip_adam_set_fused(int fused) { g_fused = fused; }

extern "C" int ppd_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                  int64_t n, int64_t step, double lr, double beta1, double beta2, double eps,
                                  double max_norm, float* grad_norm_out, const float* loss_in, float* loss_acc,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(params && grads && exp_avg && exp_avg_sq && workspace, "null pointer");
    PPD_REQUIRE(n > 0 && step >= 1, "n must be positive and step >= 1");
    PPD_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                "buffers must be 16-byte aligned");
    if (workspace_bytes < ws_bytes(n)) {
        ppd::set_error("ppd_clip_adam_step: workspace too small");
        return PPD_EWORKSPACE;
    }
    cudaStream_t s = ppd::as_stream(stream);
    Scalars* sc = reinterpret_cast<Scalars*>(workspace);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    // scalar prefactors in double, as torch's python-side Adam does (_single_tensor_adam)
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    AdamConst c;
    c.one_minus_b1 = (float)(1.0 - beta1);
    c.b2 = (float)beta2;
    c.one_minus_b2 = (float)(1.0 - beta2);
    c.neg_step_size = (float)(-(lr / bc1));
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = (float)eps;
    float mn = (float)max_norm;
    if (g_fused && n <= kFusedMaxParams) {
        static int capacity = 0;
        if (!capacity) {
            int per_sm = 0, dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clip_adam_fused_kernel, kThreads, 0) == cudaSuccess)
                capacity = per_sm * sms;
            cudaGetLastError();
        }
        int64_t want = (n / kVec + kThreads - 1) / kThreads;
        int grid = (int)(want < 1 ? 1 : want);
        int cap = capacity < 4 * ppd::kNumSMs ? capacity : 4 * ppd::kNumSMs;
        if (cap > kMaxBlocks) cap = kMaxBlocks;
        if (grid > cap) grid = cap;
        if (grid >= 1) {
            void* kargs[] = {&params, &grads, &exp_avg, &exp_avg_sq, &n, &partial, &sc, &mn,
                             &grad_norm_out, &loss_in, &loss_acc, &c};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)clip_adam_fused_kernel, dim3(grid), dim3(kThreads), kargs, 0, s);
            if (e == cudaSuccess) return ppd::launch_status("clip_adam_fused_kernel");
            cudaGetLastError();        // fall through to the three-kernel path
        }
    }
    const int nb = norm_blocks(n);
    sqnorm_partial<<<nb, kThreads, 0, s>>>(grads, n, partial);
    int rc = ppd::launch_status("sqnorm_partial");
    if (rc) return rc;
    norm_final<<<1, kThreads, 0, s>>>(partial, nb, mn, sc, grad_norm_out, loss_in, loss_acc);
    rc = ppd::launch_status("norm_final");
    if (rc) return rc;
    int64_t blocks = (n / kVec + kThreads - 1) / kThreads;
    if (blocks < 1) blocks = 1;
    if (blocks > 8 * ppd::kNumSMs) blocks = 8 * ppd::kNumSMs;
    adam_kernel<<<(int)blocks, kThreads, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, sc, c);
    return ppd::launch_status("adam_kernel");
}

extern "C" int ppd_clip_rmsprop_step(float* params, const float* grads, float* square_avg, int64_t n, double lr, double alpha,
                                     double eps, double max_norm, float* grad_norm_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    PPD_REQUIRE(params && grads && square_avg && workspace, "null pointer");
    PPD_REQUIRE(n > 0, "n must be positive");
    PPD_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)square_avg) % 16 == 0, "buffers must be 16-byte aligned");
    if (workspace_bytes < ws_bytes(n)) {
        ppd::set_error("ppd_clip_rmsprop_step: workspace too small");
        return PPD_EWORKSPACE;
    }
    cudaStream_t s = ppd::as_stream(stream);
    Scalars* sc = reinterpret_cast<Scalars*>(workspace);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    const int nb = norm_blocks(n);
    sqnorm_partial<<<nb, kThreads, 0, s>>>(grads, n, partial);
    int rc = ppd::launch_status("sqnorm_partial");
    if (rc) return rc;
    norm_final<<<1, kThreads, 0, s>>>(partial, nb, (float)max_norm, sc, grad_norm_out, nullptr, nullptr);
    rc = ppd::launch_status("norm_final");
    if (rc) return rc;
    int64_t blocks = (n / kVec + kThreads - 1) / kThreads;
    if (blocks < 1) blocks = 1;
    if (blocks > 8 * ppd::kNumSMs) blocks = 8 * ppd::kNumSMs;
    rmsprop_kernel<<<(int)blocks, kThreads, 0, s>>>(params, grads, square_avg, n, sc, (float)alpha, (float)(1.0 - alpha), (float)eps, (float)(-lr));
    return ppd::launch_status("rmsprop_kernel");
}
