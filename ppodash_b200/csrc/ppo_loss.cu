// Fused PPO loss forward + backward (PKG/algo/ppo.py:61-81) with the Categorical head's
// log-softmax / log-prob / entropy (PKG/distributions.py:23-25,66-68; torch.distributions.Categorical).
// Compile with -fmad=false: v_clip = V_old + clamp(v - V_old) must round exactly as the
// reference's two ops do, because which branch of max(e1, e2) wins inside the clip range is
// decided by that rounding (closed-form gradients).
//
// One thread per row; z row = [logit_0 .. logit_{A-1}, value].  Gradient of
//   loss = vcoef * 0.5*mean(max(e1,e2)) - mean(min(s1,s2)) - ecoef * mean(H)
// w.r.t. z is written in the same layout.  Means are over `global_rows`.
#include <float.h>

#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 128;

struct RowStats {
    float lse;      // logsumexp of the logits
    float ent;      // entropy
    float logp;     // log-prob of the taken action
};

__device__ __forceinline__ RowStats row_softmax_stats(const float* __restrict__ zr, int A, int64_t action) {
    float mx = -FLT_MAX;
    for (int j = 0; j < A; ++j) mx = fmaxf(mx, zr[j]);
    float se = 0.f;
    for (int j = 0; j < A; ++j) se += expf(zr[j] - mx);
    RowStats s;
    s.lse = mx + logf(se);
    float h = 0.f;
    for (int j = 0; j < A; ++j) {
        const float lp = fmaxf(zr[j] - s.lse, -FLT_MAX);   // Categorical.entropy clamps logits at finfo.min
        h += lp * expf(lp);
    }
    s.ent = -h;
    s.logp = (action >= 0 && action < A) ? (zr[action] - s.lse) : nanf("");
    return s;
}

__global__ void __launch_bounds__(kThreads)
ppo_loss_kernel(const float* __restrict__ z, int ldz, int A, const int64_t* __restrict__ actions,
                const float* __restrict__ old_logp, const float* __restrict__ adv,
                const float* __restrict__ old_values, const float* __restrict__ returns,
                int64_t B, float inv_rows, float clip, float vcoef, float ecoef, int clipped_vloss,
                float* __restrict__ dz, float* __restrict__ logp_out, float* __restrict__ ent_out,
                float* __restrict__ partial) {
    __shared__ float scratch[32];
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    float t_v = 0.f, t_a = 0.f, t_e = 0.f;
    if (i < B) {
        const float* zr = z + i * ldz;
        float* dzr = dz + i * ldz;
        const int64_t a = actions[i];
        const RowStats st = row_softmax_stats(zr, A, a);
        // ---- policy term (ppo.py:61-66)
        const float adv_i = adv[i];
        const float ratio = expf(st.logp - old_logp[i]);
        const float s1 = ratio * adv_i;
        const float rc = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
        const float s2 = rc * adv_i;
        t_a = fminf(s1, s2);
        // d(-mean min)/d logp_a : gradient flows through s1 when s1 <= s2 (ties included)
        const float g_lp = (s1 <= s2) ? (-(adv_i * ratio) * inv_rows) : 0.f;
        // ---- value term (ppo.py:68-77)
        const float v = zr[A];
        const float R = returns[i];
        float g_v;
        if (clipped_vloss) {
            const float vo = old_values[i];
            const float dv = v - vo;
            const float vclip = vo + fminf(fmaxf(dv, -clip), clip);
            const float d1 = v - R, d2 = vclip - R;
            const float e1 = d1 * d1, e2 = d2 * d2;
            t_v = fmaxf(e1, e2);
            const float in = (dv >= -clip && dv <= clip) ? 1.f : 0.f;   // clamp passes gradient at the bounds
            if (e1 > e2) g_v = d1;
            else if (e2 > e1) g_v = d2 * in;
            else g_v = 0.5f * d1 + 0.5f * (d2 * in);
        } else {
            const float d1 = R - v;                                     // ppo.py:77
            t_v = d1 * d1;
            g_v = v - R;
        }
        dzr[A] = (vcoef * inv_rows) * g_v;
        // ---- entropy term + softmax backward
        t_e = st.ent;
        const float ce = ecoef * inv_rows;
        for (int j = 0; j < A; ++j) {
            const float lp = zr[j] - st.lse;
            const float p = expf(lp);
            const float onehot = (j == a) ? 1.f : 0.f;
            // dL/dl_j = g_lp*(1[j=a] - p_j) - ce * dH/dl_j ,  dH/dl_j = -p_j (lp_j + H)
            dzr[j] = g_lp * (onehot - p) + ce * (p * (lp + st.ent));
        }
        if (logp_out) logp_out[i] = st.logp;
        if (ent_out) ent_out[i] = st.ent;
    }
    t_v = ppd::block_sum(t_v, scratch);
    t_a = ppd::block_sum(t_a, scratch);
    t_e = ppd::block_sum(t_e, scratch);
    if (threadIdx.x == 0) {
        partial[3 * blockIdx.x + 0] = t_v;
        partial[3 * blockIdx.x + 1] = t_a;
        partial[3 * blockIdx.x + 2] = t_e;
    }
}

// A2C loss forward + backward (PKG/algo/a2c_acktr.py:49-52,71-72): advantages = returns - values;
//   value_loss = mean(adv^2), action_loss = -mean(adv.detach() * logp), loss = vcoef * value_loss + action_loss - ecoef * mean(H).
// partial[] holds { sum adv^2, sum adv * logp, sum H } per block (ppo_loss_final negates the middle one and scales).
__global__ void __launch_bounds__(kThreads)
a2c_loss_kernel(const float* __restrict__ z, int ldz, int A, const int64_t* __restrict__ actions, const float* __restrict__ returns,
                int64_t B, float inv_rows, float vcoef, float ecoef, float* __restrict__ dz, float* __restrict__ partial) {
    __shared__ float scratch[32];
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    float t_v = 0.f, t_a = 0.f, t_e = 0.f;
    if (i < B) {
        const float* zr = z + i * ldz;
        float* dzr = dz + i * ldz;
        const int64_t a = actions[i];
        const RowStats st = row_softmax_stats(zr, A, a);
        const float adv = returns[i] - zr[A];
        t_v = adv * adv;
        t_a = adv * st.logp;
        t_e = st.ent;
        dzr[A] = (vcoef * inv_rows) * (-2.0f * adv);                 // d mean(adv^2) / dv
        const float g_lp = -adv * inv_rows;                           // d(-mean(adv.detach() * logp)) / dlogp_a
        const float ce = ecoef * inv_rows;
        for (int j = 0; j < A; ++j) {
            const float lp = zr[j] - st.lse;
            const float p = expf(lp);
            const float onehot = (j == a) ? 1.f : 0.f;
            dzr[j] = g_lp * (onehot - p) + ce * (p * (lp + st.ent));
        }
    }
    t_v = ppd::block_sum(t_v, scratch);
    t_a = ppd::block_sum(t_a, scratch);
    t_e = ppd::block_sum(t_e, scratch);
    if (threadIdx.x == 0) {
        partial[3 * blockIdx.x + 0] = t_v;
        partial[3 * blockIdx.x + 1] = t_a;
        partial[3 * blockIdx.x + 2] = t_e;
    }
}

__global__ void __launch_bounds__(256)
ppo_loss_final(const float* __restrict__ partial, int nblocks, float inv_rows, double vscale, float* __restrict__ loss_out) {
    __shared__ double scratch[32];
    double v = 0, a = 0, e = 0;
    for (int i = threadIdx.x; i < nblocks; i += 256) {
        v += partial[3 * i];
        a += partial[3 * i + 1];
        e += partial[3 * i + 2];
    }
    v = ppd::block_sum(v, scratch);
    a = ppd::block_sum(a, scratch);
    e = ppd::block_sum(e, scratch);
    if (threadIdx.x == 0) {
        loss_out[0] = (float)(vscale * v * inv_rows);
        loss_out[1] = (float)(-a * inv_rows);
        loss_out[2] = (float)(e * inv_rows);
    }
}

__global__ void __launch_bounds__(kThreads)
categorical_eval_kernel(const float* __restrict__ z, int ldz, int A, const int64_t* __restrict__ actions,
                        int64_t B, float* __restrict__ logp_out, float* __restrict__ ent_out,
                        int64_t* __restrict__ mode_out, float* __restrict__ probs_out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= B) return;
    const float* zr = z + i * ldz;
    int64_t a = actions ? actions[i] : 0;
    if (mode_out) {                       // FixedCategorical.mode: probs.argmax (first maximum)
        int best = 0;
        float bv = zr[0];
        for (int j = 1; j < A; ++j) if (zr[j] > bv) { bv = zr[j]; best = j; }
        mode_out[i] = best;
        if (!actions) a = best;
    }
    const RowStats st = row_softmax_stats(zr, A, a);
    if (logp_out) logp_out[i] = st.logp;
    if (ent_out) ent_out[i] = st.ent;
    if (probs_out) for (int j = 0; j < A; ++j) probs_out[i * A + j] = expf(zr[j] - st.lse);
}

int nblocks_for(int64_t B) { return (int)((B + kThreads - 1) / kThreads); }

}  // namespace

extern "C" size_t ppd_ppo_loss_workspace(int64_t B) { return (size_t)3 * nblocks_for(B > 0 ? B : 1) * sizeof(float); }

extern "C" int ppd_ppo_loss_fwd_bwd(const float* z, int ldz, int A, const int64_t* actions,
                                    const float* old_logp, const float* adv, const float* old_values,
                                    const float* returns, int64_t B, int64_t global_rows,
                                    float clip_param, float value_coef, float entropy_coef,
                                    int use_clipped_value_loss, float* dz, float* logp_out,
                                    float* entropy_out, float* loss_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(z && actions && old_logp && adv && returns && dz && loss_out && workspace, "null pointer");
    PPD_REQUIRE(!use_clipped_value_loss || old_values, "old_values required for the clipped value loss");
    PPD_REQUIRE(B > 0 && global_rows >= B && A > 0 && ldz >= A + 1, "bad sizes");
    const int nb = nblocks_for(B);
    if (workspace_bytes < (size_t)3 * nb * sizeof(float)) {
        ppd::set_error("ppd_ppo_loss_fwd_bwd: workspace too small");
        return PPD_EWORKSPACE;
    }
    const float inv_rows = (float)(1.0 / (double)global_rows);
    cudaStream_t s = ppd::as_stream(stream);
    ppo_loss_kernel<<<nb, kThreads, 0, s>>>(z, ldz, A, actions, old_logp, adv, old_values, returns, B, inv_rows,
                                            clip_param, value_coef, entropy_coef, use_clipped_value_loss, dz,
                                            logp_out, entropy_out, (float*)workspace);
    int rc = ppd::launch_status("ppo_loss_kernel");
    if (rc) return rc;
    ppo_loss_final<<<1, 256, 0, s>>>((const float*)workspace, nb, inv_rows, 0.5, loss_out);
    return ppd::launch_status("ppo_loss_final");
}

extern "C" int ppd_a2c_loss_fwd_bwd(const float* z, int ldz, int A, const int64_t* actions, const float* returns, int64_t B,
                                    int64_t global_rows, float value_coef, float entropy_coef, float* dz, float* loss_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(z && actions && returns && dz && loss_out && workspace, "null pointer");
    PPD_REQUIRE(B > 0 && global_rows >= B && A > 0 && ldz >= A + 1, "bad sizes");
    const int nb = nblocks_for(B);
    if (workspace_bytes < (size_t)3 * nb * sizeof(float)) {
        ppd::set_error("ppd_a2c_loss_fwd_bwd: workspace too small");
        return PPD_EWORKSPACE;
    }
    const float inv_rows = (float)(1.0 / (double)global_rows);
    cudaStream_t s = ppd::as_stream(stream);
    a2c_loss_kernel<<<nb, kThreads, 0, s>>>(z, ldz, A, actions, returns, B, inv_rows, value_coef, entropy_coef, dz, (float*)workspace);
    int rc = ppd::launch_status("a2c_loss_kernel");
    if (rc) return rc;
    ppo_loss_final<<<1, 256, 0, s>>>((const float*)workspace, nb, inv_rows, 1.0, loss_out);
    return ppd::launch_status("ppo_loss_final");
}

extern "C" int ppd_categorical_eval(const float* z, int ldz, int A, const int64_t* actions, int64_t B,
                                    float* logp_out, float* entropy_out, int64_t* mode_out, float* probs_out,
                                    void* stream) {
    PPD_REQUIRE(z && (actions || mode_out), "null pointer");
    PPD_REQUIRE(B > 0 && A > 0 && ldz >= A, "bad sizes");
    categorical_eval_kernel<<<nblocks_for(B), kThreads, 0, ppd::as_stream(stream)>>>(z, ldz, A, actions, B, logp_out,
                                                                                    entropy_out, mode_out, probs_out);
    return ppd::launch_status("categorical_eval_kernel");
}
