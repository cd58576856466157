// uint8 observation frames: normalise-on-read, device frame stack, minibatch gather.
//
// The reference keeps float32, already normalised and already stacked observations in RolloutStorage: the env-side wrappers turn
// every uint8 frame into  (frame - mean) / std  (or frame / 255) in float64 (NormalizeWrapper.observation,
// ppo-dash-study/013_.../sohojoe_wrappers.py:871-885), transpose it to CHW (TransposeImage, make_env.py:124-160), cast it to float32
// (VecPyTorch, make_env.py:84-86,105-113) and, in the frame-stacking studies, shift it into a [N, nstack*C, H, W] buffer that is
// zeroed for an env whose episode just ended (VecPyTorchFrameStack.step_wait, make_env.py:39-46).  That is 4 bytes per pixel and
// nstack copies of every frame in HBM and over PCIe.  Here the storage holds each uint8 frame ONCE; the float32 value is produced
// when a frame is read -- by the minibatch gather (training) or by the per-step expand (act / get_value):
//
//     out = float32( (float64(u8) - mean[c, y, x]) / divisor )         IEEE float64 subtract and divide, one rounding to float32:
//                                                                      bit-identical to numpy followed by torch's .float()
//     stack slot j of storage slot t (oldest first) = frame t - (nstack-1-j) if it belongs to the same episode, else 0.0f
//
// `age[t, n]` = number of earlier frames of the current episode (saturating at nstack-1), maintained by RolloutStorage.insert.
// Bytes per gathered sample: C*HW read + nstack*C*HW*4 written (the float32 gather reads and writes nstack*C*HW*4).
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 256;

struct Args {
    ppd_obs_u8_desc d;
    const int64_t* perm;
    int64_t perm_off;
    int64_t t_fixed;      // mode 2: storage slot
    int64_t rows;
    int T, N, E;
    int mode;             // 0 feed-forward (row <- perm), 1 recurrent (row t*E+j <- (t, perm[j])), 2 one storage slot (row n <- (t_fixed, n))
    float* out;
};

// float32( (float64(u) - m) / div ) without a float64 division per element.
//   x = u - m is exact (u is a small integer).  q' = x * (1/div) differs from the correctly rounded quotient q = fl64(x / div) by at
//   most 2 ulp64, so fl32(q') == fl32(q) unless q' lies within a few ulp64 of a MIDPOINT between two adjacent floats.  There are only
//   256 * C*H*W possible (u, pixel) pairs for a given (mean, divisor): certify_kernel checks ALL of them once (ppd_obs_u8_certify,
//   called when the storage is created); if every pair rounds like the division, the multiply kernel is bit-exact for every frame
//   that can ever be stored, otherwise the division kernel is used.  (ObtRetro-v6 mean / std: all 5 419 008 pairs exact.)
//   double(u) comes from the 2^52 trick: no integer->double conversion instruction (those run at a quarter of the fp64 rate).
template <bool DIV>
__device__ __forceinline__ float norm1(unsigned u, double m, double div, double rinv) {
    const double ud = __hiloint2double(0x43300000, (int)u) - 4503599627370496.0;      // (2^52 + u) - 2^52 = u, exact
    const double x = ud - m;
    return (float)(DIV ? x / div : x * rinv);
}

__global__ void __launch_bounds__(256) certify_kernel(const double* __restrict__ mean, int64_t n, double div, int* __restrict__ bad) {
    const double rinv = 1.0 / div;
    int mine = 0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) {
        const double m = mean ? mean[p] : 0.0;
        for (unsigned u = 0; u < 256; ++u)
            mine |= __float_as_uint(norm1<false>(u, m, div, rinv)) != __float_as_uint(norm1<true>(u, m, div, rinv));
    }
    if (mine) atomicOr(bad, 1);
}

constexpr int kRows = 4;        // output rows per CTA: a thread's 16 mean values are loaded once and reused for all of them
constexpr int kWords = 4;       // 4-byte words of a frame per thread and row, kThreads words apart: a warp reads 128 and writes 512 contiguous bytes

template <bool HAS_MEAN, bool DIV>
__global__ void __launch_bounds__(kThreads) obs_u8_kernel(const Args a) {
    const ppd_obs_u8_desc& d = a.d;
    const int j = blockIdx.y;                                       // stack slot, oldest first
    const int64_t row = (int64_t)d.C * d.HW;                         // bytes per frame
    const int64_t nwords = row >> 2;
    const int64_t w0 = (int64_t)blockIdx.z * (kThreads * kWords) + threadIdx.x;
    const double rinv = 1.0 / d.divisor;
    double m[kWords][4];
    if (HAS_MEAN) {
#pragma unroll
        for (int k = 0; k < kWords; ++k) {
            const int64_t w = w0 + k * kThreads;
            if (w < nwords) {
                const double2 a0 = __ldg(reinterpret_cast<const double2*>(d.mean + (w << 2)));
                const double2 a1 = __ldg(reinterpret_cast<const double2*>(d.mean + (w << 2)) + 1);
                m[k][0] = a0.x; m[k][1] = a0.y; m[k][2] = a1.x; m[k][3] = a1.y;
            }
        }
    }
    const int back = d.nstack - 1 - j;
    const int64_t i0 = (int64_t)blockIdx.x * kRows;
    unsigned q[kRows][kWords];
    bool valid[kRows];
    // every load of the CTA's rows first (kRows * kWords independent loads in flight per thread), then the arithmetic
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        valid[r] = false;
        const int64_t i = i0 + r;
        if (i < a.rows) {
            int64_t t, n;
            if (a.mode == 0) {
                const int64_t p = __ldg(a.perm + a.perm_off + i);
                t = p / a.N; n = p - t * a.N;
            } else if (a.mode == 1) {
                t = i / a.E;
                n = __ldg(a.perm + a.perm_off + (i - t * a.E));
            } else {
                t = a.t_fixed; n = i;
            }
            const int age = d.age ? (int)__ldg(d.age + t * a.N + n) : d.nstack - 1;
            valid[r] = back <= age;
            if (valid[r]) {
                const unsigned* src = reinterpret_cast<const unsigned*>(d.frames + ((t + j) * a.N + n) * row);     // frame t + j
#pragma unroll
                for (int k = 0; k < kWords; ++k) {
                    const int64_t w = w0 + k * kThreads;
                    if (w < nwords) q[r][k] = ppd::ldg_stream_u32(src + w);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int64_t i = i0 + r;
        if (i >= a.rows) break;
        float4* dst = reinterpret_cast<float4*>(a.out + (i * d.nstack + j) * row);
#pragma unroll
        for (int k = 0; k < kWords; ++k) {
            const int64_t w = w0 + k * kThreads;
            if (w >= nwords) continue;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[r]) {
                const unsigned x = q[r][k];
                o.x = norm1<DIV>(x & 0xffu, HAS_MEAN ? m[k][0] : 0.0, d.divisor, rinv);
                o.y = norm1<DIV>((x >> 8) & 0xffu, HAS_MEAN ? m[k][1] : 0.0, d.divisor, rinv);
                o.z = norm1<DIV>((x >> 16) & 0xffu, HAS_MEAN ? m[k][2] : 0.0, d.divisor, rinv);
                o.w = norm1<DIV>(x >> 24, HAS_MEAN ? m[k][3] : 0.0, d.divisor, rinv);
            }
            ppd::stg_stream(dst + w, o);
        }
    }
}

int launch(const ppd_obs_u8_desc* d, const int64_t* perm, int64_t off, int64_t rows, int64_t t_fixed, int T, int N, int E, int mode,
           float* out, void* stream, const char* what) {
    if (!d || !d->frames || !out) { ppd::set_error("%s: null pointer", what); return PPD_EINVAL; }
    if (mode != 2 && !perm) { ppd::set_error("%s: null permutation", what); return PPD_EINVAL; }
    if (rows <= 0 || N <= 0 || d->C <= 0 || d->HW <= 0 || d->nstack <= 0 || d->nstack > 65535) { ppd::set_error("%s: sizes must be positive", what); return PPD_EINVAL; }
    if (rows > 0x7fffffffLL) { ppd::set_error("%s: too many rows", what); return PPD_EINVAL; }
    if (!(d->divisor != 0.0)) { ppd::set_error("%s: divisor must be non-zero", what); return PPD_EINVAL; }
    const int64_t row = (int64_t)d->C * d->HW;
    if (((uintptr_t)d->frames & 15) || ((uintptr_t)out & 15) || ((uintptr_t)d->mean & 15)) {
        ppd::set_error("%s: frames, mean and out must be 16-byte aligned", what);
        return PPD_EINVAL;
    }
    if (row & 15) { ppd::set_error("%s: C*H*W must be a multiple of 16 (84 x 84 frames are)", what); return PPD_EINVAL; }
    Args a;
    a.d = *d; a.perm = perm; a.perm_off = off; a.t_fixed = t_fixed; a.rows = rows; a.T = T; a.N = N; a.E = E; a.mode = mode; a.out = out;
    const int64_t nwords = row >> 2;
    const int64_t z = (nwords + kThreads * kWords - 1) / (kThreads * kWords);      // 5292 words of a 3 x 84 x 84 frame: 6 CTAs
    if (z > 65535) { ppd::set_error("%s: frame too large", what); return PPD_EINVAL; }
    dim3 grid((unsigned)((rows + kRows - 1) / kRows), (unsigned)d->nstack, (unsigned)z);
    cudaStream_t s = ppd::as_stream(stream);
    const bool div = !d->multiply_exact;
    if (d->mean) { if (div) obs_u8_kernel<true, true><<<grid, kThreads, 0, s>>>(a); else obs_u8_kernel<true, false><<<grid, kThreads, 0, s>>>(a); }
    else         { if (div) obs_u8_kernel<false, true><<<grid, kThreads, 0, s>>>(a); else obs_u8_kernel<false, false><<<grid, kThreads, 0, s>>>(a); }
    return ppd::launch_status(what);
}

}  // namespace

extern "C" int ppd_obs_u8_certify(const double* mean, int64_t n, double divisor, int* bad, void* stream) {
    if (!bad || n <= 0 || !(divisor != 0.0)) { ppd::set_error("ppd_obs_u8_certify: bad arguments"); return PPD_EINVAL; }
    int64_t nb = (n + 255) / 256;
    if (nb > 4 * ppd::kNumSMs) nb = 4 * ppd::kNumSMs;
    cudaStream_t s = ppd::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), s);
    if (e != cudaSuccess) { ppd::set_error("ppd_obs_u8_certify: %s", cudaGetErrorString(e)); return (int)e; }
    certify_kernel<<<(unsigned)nb, 256, 0, s>>>(mean, n, divisor, bad);
    return ppd::launch_status("ppd_obs_u8_certify");
}

extern "C" int ppd_obs_u8_expand(const ppd_obs_u8_desc* d, int64_t t, int N, float* out, void* stream) {
    if (t < 0) { ppd::set_error("ppd_obs_u8_expand: negative slot"); return PPD_EINVAL; }
    return launch(d, nullptr, 0, N, t, 0, N, 0, 2, out, stream, "ppd_obs_u8_expand");
}

extern "C" int ppd_gather_obs_u8_feed_forward(const ppd_obs_u8_desc* d, const int64_t* perm, int64_t mb_start, int64_t rows, int T, int N,
                                              float* out, void* stream) {
    return launch(d, perm, mb_start, rows, 0, T, N, 0, 0, out, stream, "ppd_gather_obs_u8_feed_forward");
}

extern "C" int ppd_gather_obs_u8_recurrent(const ppd_obs_u8_desc* d, const int64_t* env_perm, int64_t env_start, int E, int T, int N,
                                           float* out, void* stream) {
    if (E <= 0 || T <= 0) { ppd::set_error("ppd_gather_obs_u8_recurrent: E and T must be positive"); return PPD_EINVAL; }
    return launch(d, env_perm, env_start, (int64_t)T * E, 0, T, N, E, 1, out, stream, "ppd_gather_obs_u8_recurrent");
}
