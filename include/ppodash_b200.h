/*
 * ppodash_b200 -- C ABI of the B200 (sm_100a) kernels behind the PPO-Dash training hot path.
 *
 * The reference (Sohojoe/ppo-dash) is pure Python/PyTorch and has no FFI of its own; the
 * drop-in boundary is its Python class API.  Each entry point below replaces
 * the torch-op sequence of one reference function; the Python classes in ppodash_b200/ call
 * them through ctypes (INTEGRATION.md shows the binding).  Paths are relative to
 *   PKG = ppo-dash-training/pytorch-a2c-ppo-acktr-gail/a2c_ppo_acktr
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch's current stream); calls only enqueue
 *     work: they never allocate, never synchronise, never touch the host copy of the data;
 *   - return value: 0 on success, a positive cudaError_t value if CUDA reported an error,
 *     a negative PPD_E* value for bad arguments; ppd_last_error() returns a message;
 *   - all arithmetic is fp32 unless stated; int64 is used for actions and permutations
 *     (PKG/storage.py:26, torch.randperm).
 */
#ifndef PPODASH_B200_H
#define PPODASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPD_ABI_VERSION 1
#define PPD_EINVAL (-1)    /* bad argument (null pointer, negative size, unsupported shape) */
#define PPD_EWORKSPACE (-2) /* workspace too small */

int ppd_abi_version(void);
const char* ppd_last_error(void);
/* Number of this library's kernels launched by the calling thread since the last reset. */
int64_t ppd_launch_count(void);
void ppd_reset_launch_count(void);
/* Strided host->device upload on `stream`: `rows` rows of `row_bytes` bytes, source pitch `src_pitch`, destination pitch `dst_pitch`
 * (cudaMemcpy2DAsync).  Used by RolloutStorage.upload_from to move ONE env's observations -- rows of a [T+1, N, ...] tensor, N rows
 * apart -- from pinned host memory, so that the upload of a rollout can be ordered by the first epoch's env permutation and overlap
 * the update (the reference's `rollouts.to(device)`, PKG/storage.py:34-46, is one blocking copy of everything). */
int ppd_upload_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows, void* stream);

/* ---------------------------------------------------------------------------------------
 * Returns / GAE                          replaces RolloutStorage.compute_returns, PKG/storage.py:82-121
 * rewards [T,N]; value_preds, masks, bad_masks, returns [T+1,N]; next_value [N].
 * One env per lane, time split into per-warp chunks that are combined as a single-pass chained
 * affine scan X_t = a_t X_{t+1} + c_t; inside a chunk the reference's operation order is replayed.
 * `workspace` (>= ppd_compute_returns_workspace bytes, 16-byte aligned) holds the inter-CTA carries;
 * it must be zero-filled once when allocated and may then be reused by every later call on the
 * same stream (each launch re-arms it).
 * Side effects as in the reference: use_gae -> value_preds[T] = next_value (storage.py:90,108),
 * returns[T] untouched; otherwise returns[T] = next_value (storage.py:101,118).
 * gamma*gae_lambda is formed in double and rounded once, as Python does in the reference.
 */
size_t ppd_compute_returns_workspace(int T, int N);
/* Tuning: warps per CTA (4, 8 or 16; 16 steps each) and, for 8 warps, resident CTAs per SM (3 or 4).
 * warps = 100 / 101 / 102 selects never / auto / always for the persistent TMA-staged variant (default never:
 * it measured slower than the register kernel); warps = 200 + 10*S + C sets its ring depth S and CTAs/SM C. */
void ppd_compute_returns_set_tuning(int warps, int min_blocks);
int ppd_compute_returns(const float* rewards, float* value_preds, const float* masks,
                        const float* bad_masks, float* returns, const float* next_value,
                        int T, int N, double gamma, double gae_lambda,
                        int use_gae, int use_proper_time_limits,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Advantage statistics                    replaces PKG/algo/ppo.py:35-37
 * moments[3] (double, device) = { sum(adv), sum(adv^2), count } with adv = returns - value_preds
 * over the first n = T*N elements.  Split in two so that a data-parallel caller can
 * all-reduce the three doubles between the calls.
 * stats[2] (float, device) = { mean, unbiased_std + 1e-5 }.
 */
size_t ppd_advantage_moments_workspace(int64_t n);
int ppd_advantage_moments(const float* returns, const float* value_preds, int64_t n,
                          double* moments, void* workspace, size_t workspace_bytes, void* stream);
int ppd_advantage_finalize(const double* moments, float* stats, void* stream);
/* adv_out[i] = (returns[i] - value_preds[i] - mean) / (std + 1e-5) */
int ppd_advantage_normalize(const float* returns, const float* value_preds, int64_t n,
                            const float* stats, float* adv_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Minibatch gathers        replace RolloutStorage.feed_forward_generator (PKG/storage.py:123-160)
 *                          and RolloutStorage.recurrent_generator   (PKG/storage.py:162-223)
 * Source fields are the time-major storage buffers ([T(+1), N, ...], row r = t*N + n).
 *  feed-forward: output row i  <- source row perm[mb_start + i],        i in [0, rows)
 *  recurrent   : output row t*E + j <- source row t*N + perm[env_start + j]; hidden state only
 *                at t = 0 -> hxs_out [E, hxs_row]
 * Any (src,dst) pair may be NULL to skip that field.  Advantages: if `adv` is non-NULL it is
 * gathered like the other fields; otherwise, if `adv_stats` is non-NULL, the normalised
 * advantage (returns - value_preds - stats[0]) / stats[1] is computed on the fly.
 * `*_out_ld` is the output row stride in elements (0 = dense), so a caller can gather the
 * vector obs straight into a padded feature matrix.
 */
typedef struct ppd_gather_desc {
    const float* obs;          float* obs_out;          int64_t obs_row;   /* elems per row: C*H*W */
    const float* vobs;         float* vobs_out;         int64_t vobs_row;  int64_t vobs_out_ld;
    const float* hxs;          float* hxs_out;          int64_t hxs_row;
    const int64_t* actions;    int64_t* actions_out;    int64_t actions_row;
    const float* value_preds;  float* value_preds_out;
    const float* returns;      float* returns_out;
    const float* masks;        float* masks_out;
    const float* logp;         float* logp_out;
    const float* adv;          float* adv_out;
    const float* adv_stats;
} ppd_gather_desc;

int ppd_gather_feed_forward(const ppd_gather_desc* d, const int64_t* perm, int64_t mb_start,
                            int64_t rows, int T, int N, void* stream);
int ppd_gather_recurrent(const ppd_gather_desc* d, const int64_t* env_perm, int64_t env_start,
                         int E, int T, int N, void* stream);

/* ---------------------------------------------------------------------------------------
 * uint8 observation frames: normalise-on-read, device frame stack, minibatch gather      
 * replaces, bit-exactly, the env-side float pipeline of the reference:
 *   NormalizeWrapper.observation   ppo-dash-study/013_.../sohojoe_wrappers.py:871-885   (obs - mean) / std  or  obs / 255, float64
 *   TransposeImage.observation     ppo-dash-training/.../make_env.py:155-160            HWC -> CHW
 *   VecPyTorch.step_wait           make_env.py:105-113                                  torch.from_numpy(obs).float()
 *   VecPyTorchFrameStack.step_wait make_env.py:39-46                                    shift, zero on episode end, append
 * followed by the obs part of the minibatch generators (PKG/storage.py:144-146,171-197).
 * The storage keeps every frame ONCE as uint8, CHW:  frames [T + nstack, N, C, H*W]; storage slot t's newest frame is frame
 * t + nstack - 1, so slot t's stack entry j (oldest first) is frame t + j.  age [T+1, N] (uint8) = number of earlier frames of the
 * same episode, saturating at nstack - 1 (NULL: always nstack - 1, i.e. no episode boundaries).  Output element:
 *     out = float32( (float64(u8) - mean[c,y,x]) / divisor )    IEEE float64, one rounding to float32 (= numpy, then .float())
 * or 0.0f for a stack entry that predates the episode.  mean: float64 [C, H*W] (the reference's (H,W,C) file transposed) or NULL
 * (= 0: obs / 255 with divisor 255; raw values with divisor 1).  C*H*W must be a multiple of 16; frames, mean, out 16-byte aligned.
 *   expand                 out [N, nstack*C, H*W]       <- storage slot t           (what Policy.act / get_value are fed)
 *   gather_feed_forward    out row i                    <- (t, n) = divmod(perm[mb_start + i], N)
 *   gather_recurrent       out row t*E + j              <- (t, env_perm[env_start + j])
 */
typedef struct ppd_obs_u8_desc {
    const uint8_t* frames;
    const uint8_t* age;
    const double* mean;
    double divisor;
    int C;
    int64_t HW;
    int nstack;
    int multiply_exact;      /* 1: (u8 - mean) * (1 / divisor) was certified bit-identical to the division for this (mean, divisor) */
} ppd_obs_u8_desc;
/* bad[0] (device int) = 0 if float32((u - mean[p]) * (1 / divisor)) == float32((u - mean[p]) / divisor) for ALL u in 0..255 and all
 * n pixels p (mean NULL: m = 0), else 1.  Run once per (mean, divisor); a certified pair lets the kernels below replace the float64
 * division of every element by one multiply without giving up bit-exactness. */
int ppd_obs_u8_certify(const double* mean, int64_t n, double divisor, int* bad, void* stream);
int ppd_obs_u8_expand(const ppd_obs_u8_desc* d, int64_t t, int N, float* out, void* stream);
int ppd_gather_obs_u8_feed_forward(const ppd_obs_u8_desc* d, const int64_t* perm, int64_t mb_start, int64_t rows, int T, int N,
                                   float* out, void* stream);
int ppd_gather_obs_u8_recurrent(const ppd_obs_u8_desc* d, const int64_t* env_perm, int64_t env_start, int E, int T, int N,
                                float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused PPO loss, forward + backward      replaces PKG/algo/ppo.py:61-81 and the Categorical
 *                                         log-prob / entropy of PKG/distributions.py:23-25,66-68
 * z [B, ldz]: columns 0..A-1 = action logits, column A = value prediction (the heads GEMM
 * writes both).  Per row: log-softmax, log-prob of the taken action, entropy, ratio, clipped
 * surrogate, clipped value loss; and the closed-form gradient dz [B, ldz] of
 *   loss = value_loss*value_coef + action_loss - entropy*entropy_coef
 * with every mean taken over `global_rows` rows (= B on one GPU; sum of B over ranks when the
 * minibatch is sharded, so that summing dz-derived gradients over ranks gives the global mean).
 * loss_out[3] (device) = { value_loss, action_loss, dist_entropy } contributions of these B rows
 * (already divided by global_rows).  logp_out / entropy_out may be NULL.
 */
size_t ppd_ppo_loss_workspace(int64_t B);
int ppd_ppo_loss_fwd_bwd(const float* z, int ldz, int A, const int64_t* actions,
                         const float* old_logp, const float* adv, const float* old_values,
                         const float* returns, int64_t B, int64_t global_rows,
                         float clip_param, float value_coef, float entropy_coef,
                         int use_clipped_value_loss,
                         float* dz, float* logp_out, float* entropy_out, float* loss_out,
                         void* workspace, size_t workspace_bytes, void* stream);
/* A2C loss, forward + backward          replaces PKG/algo/a2c_acktr.py:49-52,71-72
 *   advantages = returns - values; value_loss = mean(adv^2); action_loss = -mean(adv.detach() * log_prob);
 *   dz = gradient of value_loss*value_coef + action_loss - entropy*entropy_coef; loss_out[3] = { value_loss, action_loss,
 *   dist_entropy } contributions of these B rows (means over global_rows).  Same z layout and workspace as the PPO loss. */
int ppd_a2c_loss_fwd_bwd(const float* z, int ldz, int A, const int64_t* actions, const float* returns, int64_t B,
                         int64_t global_rows, float value_coef, float entropy_coef, float* dz, float* loss_out,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Forward only: log-prob of `actions`, per-row entropy and (if mode_out) the arg-max action.
 * Used by Policy.act / evaluate_actions (PKG/model.py:54-79). */
int ppd_categorical_eval(const float* z, int ldz, int A, const int64_t* actions, int64_t B,
                         float* logp_out, float* entropy_out, int64_t* mode_out, float* probs_out,
                         void* stream);

/* ---------------------------------------------------------------------------------------
 * Gradient-norm clip + Adam over one flat buffer     replaces PKG/algo/ppo.py:82-84
 *   total_norm = ||grads||_2 ; coef = min(1, max_norm / (total_norm + 1e-6)) (skipped if max_norm <= 0)
 *   torch.optim.Adam (no amsgrad / weight decay): m = lerp(m, g, 1-b1); v = v*b2 + (1-b2) g^2;
 *   p -= (lr / (1-b1^step)) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * `step` is the 1-based optimiser step.  grad_norm_out (device float, may be NULL) receives total_norm.
 * If loss_in/loss_acc are non-NULL, loss_acc[0..2] += loss_in[0..2] (running sums over minibatches,
 * read once per update instead of three .item() syncs per minibatch, ppo.py:86-88).
 */
size_t ppd_clip_adam_workspace(int64_t n);
/* 1 (default): parameter sets of up to 4 Mi floats run as ONE cooperative launch (norm -> grid barrier ->
 * Adam, gradient re-read from L2); 0: always the three-kernel path. */
void ppd_clip_adam_set_fused(int fused);
int ppd_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                       int64_t n, int64_t step, double lr, double beta1, double beta2, double eps,
                       double max_norm, float* grad_norm_out, const float* loss_in, float* loss_acc,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Gradient-norm clip + RMSprop step          replaces nn.utils.clip_grad_norm_ + optim.RMSprop.step of PKG/algo/a2c_acktr.py:30-31,
 * 74-78 (centered = False, momentum = 0, weight_decay = 0): square_avg = alpha*square_avg + (1-alpha)*g*g;
 * p -= lr * g / (sqrt(square_avg) + eps), g already scaled by min(1, max_norm / (||g|| + 1e-6)).  Workspace as for Adam. */
int ppd_clip_rmsprop_step(float* params, const float* grads, float* square_avg, int64_t n, double lr, double alpha, double eps,
                          double max_norm, float* grad_norm_out, void* workspace, size_t workspace_bytes, void* stream);


/* ---------------------------------------------------------------------------------------
 * Running observation normalisation     replaces VecNormalize._obfilt (PKG/envs.py:208-217) +
 *                                       baselines RunningMeanStd.update (third party, unpinned)
 * obs [N, F] fp32; mean, var [F] float64 (device, updated in place when update != 0);
 * count_host = running count BEFORE this batch (the caller adds N afterwards).
 * out [N, F] = clip((obs - mean) / sqrt(var + epsilon), -clipob, clipob) using the UPDATED moments.
 */
int ppd_obs_rms_update_normalize(const float* obs, int N, int64_t F, double* mean, double* var,
                                 double count_host, int update, double epsilon, double clipob,
                                 float* out, void* stream);

/* =======================================================================================
 * Policy network (PKG/model.py).  "fp32" mode: SIMT fp32 kernels, used for the 1e-5 parity gate.
 * ======================================================================================= */

/* ---------------------------------------------------------------------------------------
 * GEMM family                 replaces nn.Conv2d / nn.Linear / GRU input projection forward and
 *                             backward (PKG/model.py:176-180,186-188,90; PKG/distributions.py:64)
 *   C[i,j] (+)= sum_kk OpA(i,kk) * OpB(j,kk)
 *   OpA(i,kk) = a_kmajor ? A[i*lda + kk] : A[kk*lda + i];  OpB likewise with b_kmajor / ldb.
 * Epilogue (in this order): + bias[j]; ReLU if relu; zero where mask[i*ldm + j] <= 0 (ReLU
 * backward against the saved activation); C += result if accumulate, else C = result.
 * Long contractions are split over CTAs and reduced in a fixed order (deterministic).
 */
typedef struct ppd_gemm_args {
    const float* A; int64_t lda; int a_kmajor;
    const float* B; int64_t ldb; int b_kmajor;
    float* C; int64_t ldc;
    int64_t I, J, KK;
    const float* bias;
    const float* mask; int64_t ldm;
    int relu; int accumulate;
} ppd_gemm_args;
size_t ppd_sgemm_workspace(int64_t I, int64_t J, int64_t KK);
int ppd_sgemm(const ppd_gemm_args* g, void* workspace, size_t workspace_bytes, void* stream);
/* Same contraction on the 5th-generation tensor cores: tcgen05.mma kind::tf32 (fp32 operands read as
 * TF32, fp32 accumulation in tensor memory), TMA-fed 4-stage shared-memory ring.  Operands must be
 * 16-byte aligned with leading dimensions that are multiples of 4 floats (ppd_tc_gemm_supported).
 * flags:
 *   PPD_TC_TRANSPOSE_OUT  store the result transposed (C is then [J, I] with row stride ldc; bias is
 *                         still indexed by j and mask uses C's layout) -- puts the wide dimension of a
 *                         weight gradient on the 128-row MMA axis;
 *   PPD_TC_SPLIT3         "3xTF32": every operand tile x is split in shared memory into hi = TF32(x)
 *                         (round to nearest) and lo = TF32(x - hi), and A_lo*B_hi + A_hi*B_lo + A_hi*B_hi is
 *                         accumulated, which restores fp32-level accuracy (~2^-23 per product, zero-mean)
 *                         on the tensor cores.  Without it ("tf32" mode) results agree with fp32 to ~1e-3. */
#define PPD_TC_TRANSPOSE_OUT 1
#define PPD_TC_SPLIT3 2
size_t ppd_tc_gemm_workspace(int64_t I, int64_t J, int64_t KK);
int ppd_tc_gemm_supported(const ppd_gemm_args* g);
int ppd_tc_gemm(const ppd_gemm_args* g, int flags, void* workspace, size_t workspace_bytes, void* stream);
typedef struct ppd_conv_geom { int B, H, W, C, kh, kw, stride; } ppd_conv_geom;   /* input tensor [B,H,W,C] and the filter */
/* 3xTF32 with a PRE-SPLIT B operand (weights): g->B holds hi = TF32(B), b_lo the residuals TF32(B - hi), same layout
 * and leading dimension; ppd_split_tf32 produces both (n a multiple of 4, 16-byte aligned).  The kernel then neither
 * re-splits B per tile nor fences shared memory for it: the weights of a minibatch are split once, after the optimiser step. */
int ppd_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
int ppd_tc_gemm_bsplit(const ppd_gemm_args* g, const float* b_lo, int flags, void* workspace, size_t workspace_bytes,
                       void* stream);
/* Implicit-GEMM NHWC convolution, 3xTF32 (fp32-level accuracy), no im2col / col2im matrices in HBM: the persistent tcgen05
 * kernel loads the patch rows of a tile straight from the activation tensor with 4-D TMA boxes (an overlapping-stride
 * im2col VIEW).  geom describes the INPUT tensor x / dx [B,H,W,C] and the filter; weights are [Cout, (ky,kx,c)] pre-split
 * (ppd_split_tf32).  Requires kw*C % 32 == 0 (forward), kh == kw, kh | H | W multiples of the stride and Cout % 32 == 0
 * (dgrad), C and Cout in {32, 64} on the output side.
 *   fwd  : out[B*OH*OW, Cout] = (ReLU)(conv(x) + bias)                      replaces nn.Conv2d forward, PKG/model.py:177-178
 *   dgrad: dx[B,H,W,C] = (act_mask > 0) * conv_transpose(dy[B,OH,OW,Cout])  gather form over the stride's parity classes:
 *          every dx element is written once (no atomics, no memset, ReLU backward fused); deterministic. */
int ppd_conv_fwd_nhwc(const float* x, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                      const float* bias, int relu, float* out, void* stream);
int ppd_conv_dgrad_nhwc(const float* dy, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                        const float* act_mask, float* dx, void* stream);
/* Forward over NCHW observations x [B,C,H,W] (geom: same fields, C = channels), weights [Cout, (c,ky,kx)], NHWC output
 * [B*OH*OW, Cout]: the kernel stages raw image rows and expands the 8-wide windows on the way to registers
 * (replaces main.0 = Conv2d(C, 32, 8, stride 4), PKG/model.py:177).  kw == 8, stride == 4, W % 4 == 0. */
int ppd_conv_fwd_nchw(const float* x, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                      const float* bias, int relu, float* out, void* stream);
/*   wgrad: dW[Cout, K] (+)= sum over output pixels of dy[pixel, cout] * patch(x)[pixel, k], contraction split over the SMs and
 *          reduced in a fixed order.  nchw = 0: x is NHWC, patch / weight order (ky,kx,c), kw*C % 64 == 0;
 *          nchw = 1: x is NCHW (the observations), patch order (c,ky,kx), 8x8 filter (5-D TMA view, no channel padding). */
size_t ppd_conv_wgrad_workspace(const ppd_conv_geom* geom, int Cout);
int ppd_conv_wgrad(const float* x, const ppd_conv_geom* geom, int nchw, const float* dy, int Cout, float* dW, int accumulate,
                   void* workspace, size_t workspace_bytes, void* stream);
/* dgrad of an NHWC convolution with col2im fused into the epilogue: the product dY[M,N] W[N,(ky,kx,c)] is not
 * stored but scatter-added (red.global.add.v4.f32) into dx[B,H,W,C], which the caller has zeroed; follow with
 * ppd_relu_mask.  Replaces the dcols round trip through HBM (write + col2im read).  g->C = dx, g->ldc ignored.
 * Summation order of the overlapping taps is not fixed (fp32 atomics); the "fp32" mode keeps the deterministic
 * ppd_sgemm + ppd_col2im_nhwc pair. */
int ppd_tc_gemm_col2im(const ppd_gemm_args* g, const ppd_conv_geom* geom, int flags, void* stream);
/* x[i] = act[i] > 0 ? x[i] : 0 */
int ppd_relu_mask(float* x, const float* act, int64_t n, void* stream);
/* Tuning / A-B switches (defaults in brackets): 0 / [1] two co-resident CTAs for the narrow tiles of the non-persistent kernel;
 * 32 / 64 / 128 / 256 / [-1] force its tile width; 2 / [3] its TMEM-A mode; 4 / [5] non-persistent / persistent kernel for 3xTF32;
 * 6 / [7] implicit convolutions with one im2col box per row and k-block / with the input staged once per tile; 8 / [9] streamed /
 * shared-memory-resident weight tiles in the convolutions; 1000 + n: at most n CTAs for the persistent kernel ([1000] = one per SM) --
 * a data-parallel caller lowers it while an NCCL all-reduce is in flight so that both kernels are resident instead of the collective
 * delaying a statically scheduled CTA. */
void ppd_tc_gemm_set_option(int option);
/* out[j] (+)= sum_i X[i*ld + j]  (bias gradients) */
size_t ppd_colsum_workspace(int64_t I, int64_t J);
int ppd_colsum(const float* X, int64_t ld, int64_t I, int64_t J, float* out, int accumulate,
               void* workspace, size_t workspace_bytes, void* stream);
/* Up to 8 column sums in two launches: all bias gradients of a minibatch at once. */
typedef struct ppd_colsum_seg { const float* X; int64_t ld, I, J; float* out; int accumulate; } ppd_colsum_seg;
size_t ppd_colsum_multi_workspace(const ppd_colsum_seg* segs, int n);
int ppd_colsum_multi(const ppd_colsum_seg* segs, int n, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Convolution lowering (conv = im2col + GEMM; activations NHWC between the convolutions).
 *   im2col_nchw: x [B,C,H,W] -> cols [B*OH*OW, ld], patch index (c,ky,kx)   (conv1 reads the obs)
 *   im2col_nhwc: x [B,H,W,C] -> cols [B*OH*OW, ld], patch index (ky,kx,c)
 *   col2im_nhwc: transpose of im2col_nhwc (gather form, no atomics); if act_mask != NULL the
 *                result is zeroed where act_mask[b,y,x,c] <= 0 (ReLU backward fused)
 *   batched_transpose: y[b,c,r] = x[b,r,c]  (NHWC <-> NCHW of the last conv activation, so the
 *                flatten order equals the reference's, PKG/model.py:10-12)
 */
int ppd_im2col_nchw(const float* x, int B, int C, int H, int W, int kh, int kw, int stride,
                    float* cols, int64_t ld, void* stream);
int ppd_im2col_nhwc(const float* x, int B, int H, int W, int C, int kh, int kw, int stride,
                    float* cols, int64_t ld, void* stream);
int ppd_col2im_nhwc(const float* dcols, int64_t ld, int B, int H, int W, int C, int kh, int kw, int stride,
                    const float* act_mask, float* dx, void* stream);
int ppd_batched_transpose(const float* x, int64_t B, int R, int Cc, float* y, void* stream);

/* ---------------------------------------------------------------------------------------
 * GRU with mask reset            replaces NNBase._forward_gru, PKG/model.py:111-166
 *   h_t = GRU(x_t, h_{t-1} * m_t), gate order r,z,n; gi = x W_ih^T + b_ih is computed by the caller
 *   for all T steps (one GEMM).  Rows are time-major: row = t*E + e.
 * forward : gi [T*E,3H], h0 [E,H], masks [T*E], w_hh [3H,H], b_hh [3H] -> hs [T*E,H] (all hidden
 *           states = the GRU output), h_last [E,H] (may be NULL); save_* [T*E,H] (r, z, n and
 *           W_hn h + b_hn) are needed by backward, pass NULL for inference.
 * backward: dhs [T*E,H] = dL/d hs -> dgi [T*E,3H] (gradient wrt gi) and dghn [T*E,H] (gradient wrt
 *           the hidden-side n pre-activation; the r,z hidden-side gradients equal dgi's); dh0
 *           [E,H] may be NULL.  Weight gradients follow as GEMMs over all T*E rows.
 * One persistent cooperative launch runs all T steps (one grid barrier per step).
 * masked_prev: hm[t*E+e,:] = (t ? hs[(t-1)*E+e,:] : h0[e,:]) * masks[t*E+e]  (operand of dW_hh).
 */
int ppd_gru_forward(const float* gi, const float* h0, const float* masks, const float* w_hh,
                    const float* b_hh, int T, int E, int H, float* hs, float* h_last,
                    float* save_r, float* save_z, float* save_n, float* save_ghn, void* stream);
int ppd_gru_backward(const float* dhs, const float* masks, const float* w_hh, const float* h0,
                     const float* hs, const float* save_r, const float* save_z, const float* save_n,
                     const float* save_ghn, int T, int E, int H, float* dgi, float* dghn, float* dh0,
                     void* stream);
int ppd_gru_masked_prev(const float* hs, const float* h0, const float* masks, int T, int E, int H,
                        float* hm, void* stream);
/* Kernel selection for the two calls above: 0 (default) = thread-block-cluster / DSMEM kernels when H % 16 == 0 (one 16-CTA
 * cluster per env, one mbarrier handshake per step; at H = 512 the W_hh slice lives in registers and any E runs as waves of
 * clusters, with the BACKWARD pass of E >= 9 envs on persistent clusters that interleave up to four envs each; generic H: W_hh
 * slices in shared memory, E <= 8), else the grid-cooperative kernels; 1 = always grid-cooperative; 2 = cluster kernels with the
 * W_hh slice in shared memory even at H = 512 (the generic-H variant, kept selectable for testing); 3 = the interleaved-env
 * backward kernel for every E; 4 = one cluster per env for every E; 100 + n = the interleaved kernel assumes n resident clusters
 * (0 = ask the occupancy API; tuning / tests). */
void ppd_gru_set_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* PPODASH_B200_H */
