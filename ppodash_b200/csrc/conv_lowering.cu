// Convolution lowering for the Nature-CNN trunk (PKG/model.py:176-180): conv = im2col + GEMM.
//
// Activations between the convolutions are kept NHWC ([B,H,W,C]) because that is exactly the
// row-major [B*H*W, C] matrix the GEMM writes; the reduction index of a patch is then ordered
// (ky, kx, c), for which each (output pixel, ky) segment is kw*C contiguous floats -> pure
// 16-byte vector copies.  The first convolution reads the rollout obs, which the reference
// stores NCHW ([B,C,84,84]); its patches are ordered (c, ky, kx) = the natural layout of
// conv1.weight.  The last activation is transposed per sample to NCHW so that the flatten
// order matches the reference's `Flatten` (c*49 + y*7 + x) and main.7.weight is used as is.
#include "ppd_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
im2col_nchw_kernel(const float* __restrict__ x, int B, int C, int H, int W, int kh, int kw, int stride,
                   int OH, int OW, float* __restrict__ cols, int64_t ld, int vec) {
    const int K = C * kh * kw;
    const int K4 = K >> 2;
    const int64_t total = (int64_t)B * OH * OW * K4;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
        const int64_t m = idx / K4;
        const int k = (int)(idx - m * K4) << 2;
        const int ox = (int)(m % OW);
        const int64_t t = m / OW;
        const int oy = (int)(t % OH);
        const int64_t b = t / OH;
        float4 v;
        if (vec) {     // kw % 4 == 0: the 4 taps share (c, ky) and are contiguous and 16-byte aligned
            const int kx = k % kw, ky = (k / kw) % kh, c = k / (kw * kh);
            const float* p = x + (((b * C + c) * H + (oy * stride + ky)) * (int64_t)W + ox * stride + kx);
            v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u;
                const int kx = kk % kw, ky = (kk / kw) % kh, c = kk / (kw * kh);
                e[u] = __ldg(x + (((b * C + c) * H + (oy * stride + ky)) * (int64_t)W + ox * stride + kx));
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4*>(cols + m * ld + k) = v;
    }
}

__global__ void __launch_bounds__(kThreads)
im2col_nhwc_kernel(const float* __restrict__ x, int B, int H, int W, int C, int kh, int kw, int stride,
                   int OH, int OW, float* __restrict__ cols, int64_t ld) {
    const int seg = kw * C;            // contiguous floats per (pixel, ky)
    const int K4 = (kh * seg) >> 2;
    const int64_t total = (int64_t)B * OH * OW * K4;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
        const int64_t m = idx / K4;
        const int k = (int)(idx - m * K4) << 2;
        const int ox = (int)(m % OW);
        const int64_t t = m / OW;
        const int oy = (int)(t % OH);
        const int64_t b = t / OH;
        const int ky = k / seg, r = k - ky * seg;
        const float* p = x + (((b * H + (oy * stride + ky)) * (int64_t)W + ox * stride) * C + r);
        *reinterpret_cast<float4*>(cols + m * ld + k) = __ldg(reinterpret_cast<const float4*>(p));
    }
}

// Row-structured im2col: blockIdx.x = (b, oy) output row, threadIdx.x = float4 index inside a patch (so the patch
// decomposition is done once per thread, no divisions in the loop), threadIdx.y strides over ox.  For one ox the
// threads of a row write one contiguous patch (K*4 bytes) -> coalesced 16-byte stores; 32-bit index arithmetic.
__global__ void __launch_bounds__(256)
im2col_nchw_rows_kernel(const float* __restrict__ x, int C, int H, int W, int kh, int kw, int stride,
                        int OH, int OW, float* __restrict__ cols, int64_t ld) {
    const int k = threadIdx.x << 2;                       // first of 4 consecutive taps (same c, ky)
    const int kx = k % kw, ky = (k / kw) % kh, c = k / (kw * kh);
    const int b = blockIdx.x / OH, oy = blockIdx.x - b * OH;
    const float* src = x + (((size_t)b * C + c) * H + (oy * stride + ky)) * (size_t)W + kx;
    float* dst = cols + ((size_t)blockIdx.x * OW) * ld + k;
    for (int ox = threadIdx.y; ox < OW; ox += blockDim.y) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + ox * stride));
        __stcs(reinterpret_cast<float4*>(dst + (size_t)ox * ld), v);
    }
}

__global__ void __launch_bounds__(256)
im2col_nhwc_rows_kernel(const float* __restrict__ x, int H, int W, int C, int kh, int kw, int stride,
                        int OH, int OW, float* __restrict__ cols, int64_t ld) {
    const int seg = kw * C;                               // contiguous floats per (pixel, ky)
    const int K4 = (kh * seg) >> 2;
    const int b = blockIdx.x / OH, oy = blockIdx.x - b * OH;
    for (int k4 = threadIdx.x; k4 < K4; k4 += blockDim.x) {
        const int k = k4 << 2;
        const int ky = k / seg, r = k - ky * seg;
        const float* src = x + (((size_t)b * H + (oy * stride + ky)) * W) * (size_t)C + r;
        float* dst = cols + ((size_t)blockIdx.x * OW) * ld + k;
        for (int ox = threadIdx.y; ox < OW; ox += blockDim.y) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)ox * stride * C));
            __stcs(reinterpret_cast<float4*>(dst + (size_t)ox * ld), v);
        }
    }
}

// dx[b,y,x,c] = sum over patches containing the pixel of dcols[(b,oy,ox), (ky,kx,c)], times (act > 0).
__global__ void __launch_bounds__(kThreads)
col2im_nhwc_kernel(const float* __restrict__ dcols, int64_t ld, int B, int H, int W, int C, int kh, int kw,
                   int stride, int OH, int OW, const float* __restrict__ act, float* __restrict__ dx) {
    const int C4 = C >> 2;
    const int64_t total = (int64_t)B * H * W * C4;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
        const int c = (int)(idx % C4) << 2;
        int64_t t = idx / C4;
        const int xx = (int)(t % W); t /= W;
        const int yy = (int)(t % H);
        const int64_t b = t / H;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < kh; ++ky) {
            const int ry = yy - ky;
            if (ry < 0 || ry % stride) continue;
            const int oy = ry / stride;
            if (oy >= OH) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int rx = xx - kx;
                if (rx < 0 || rx % stride) continue;
                const int ox = rx / stride;
                if (ox >= OW) continue;
                const int64_t m = (b * OH + oy) * OW + ox;
                const float4 v = __ldg(reinterpret_cast<const float4*>(dcols + m * ld + (ky * kw + kx) * C + c));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        const int64_t o = ((b * H + yy) * (int64_t)W + xx) * C + c;
        if (act) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(act + o));
            s.x = a.x > 0.f ? s.x : 0.f; s.y = a.y > 0.f ? s.y : 0.f;
            s.z = a.z > 0.f ? s.z : 0.f; s.w = a.w > 0.f ? s.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + o) = s;
    }
}

// y[b, c, r] = x[b, r, c]   (R x Cc per batch entry, staged through shared memory)
__global__ void __launch_bounds__(kThreads)
batched_transpose_kernel(const float* __restrict__ x, int R, int Cc, float* __restrict__ y) {
    extern __shared__ float tile[];
    const int64_t b = blockIdx.x;
    const int n = R * Cc;
    const float* xb = x + b * n;
    float* yb = y + b * n;
    for (int i = threadIdx.x; i < n; i += kThreads) tile[i + i / 32] = xb[i];     // pad 1 per 32 against bank conflicts
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += kThreads) {
        const int c = o / R, r = o - c * R;
        const int i = r * Cc + c;
        yb[o] = tile[i + i / 32];
    }
}

// x = act > 0 ? x : 0   (ReLU backward applied after a scatter-added col2im)
__global__ void __launch_bounds__(kThreads)
relu_mask_kernel(float* __restrict__ x, const float* __restrict__ act, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        float4 v = reinterpret_cast<float4*>(x)[i];
        const float4 a = __ldg(reinterpret_cast<const float4*>(act) + i);
        v.x = a.x > 0.f ? v.x : 0.f; v.y = a.y > 0.f ? v.y : 0.f; v.z = a.z > 0.f ? v.z : 0.f; v.w = a.w > 0.f ? v.w : 0.f;
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

int grid_for(int64_t total) {
    int64_t b = (total + kThreads - 1) / kThreads;
    const int64_t cap = 16 * (int64_t)ppd::kNumSMs;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" int ppd_im2col_nchw(const float* x, int B, int C, int H, int W, int kh, int kw, int stride,
                               float* cols, int64_t ld, void* stream) {
    PPD_REQUIRE(x && cols, "null pointer");
    PPD_REQUIRE(B > 0 && C > 0 && H >= kh && W >= kw && kh > 0 && kw > 0 && stride > 0, "bad sizes");
    const int K = C * kh * kw;
    PPD_REQUIRE(K % 4 == 0 && ld >= K && ld % 4 == 0 && (uintptr_t)cols % 16 == 0, "patch size and ld must be multiples of 4");
    const int OH = (H - kh) / stride + 1, OW = (W - kw) / stride + 1;
    const int vec = (kw % 4 == 0) && (stride % 4 == 0) && (W % 4 == 0) && ((uintptr_t)x % 16 == 0);
    const int64_t total = (int64_t)B * OH * OW * (K / 4);
    if (vec && K / 4 <= 256 && (int64_t)B * OH <= 0x7fffffffLL) {
        const int tx = K / 4;
        int ty = 256 / tx;
        if (ty > OW) ty = OW;
        if (ty < 1) ty = 1;
        im2col_nchw_rows_kernel<<<(unsigned)(B * OH), dim3(tx, ty), 0, ppd::as_stream(stream)>>>(x, C, H, W, kh, kw, stride,
                                                                                                 OH, OW, cols, ld);
        return ppd::launch_status("im2col_nchw_rows_kernel");
    }
    im2col_nchw_kernel<<<grid_for(total), kThreads, 0, ppd::as_stream(stream)>>>(x, B, C, H, W, kh, kw, stride, OH, OW,
                                                                               cols, ld, vec);
    return ppd::launch_status("im2col_nchw_kernel");
}

extern "C" int ppd_im2col_nhwc(const float* x, int B, int H, int W, int C, int kh, int kw, int stride,
                               float* cols, int64_t ld, void* stream) {
    PPD_REQUIRE(x && cols, "null pointer");
    PPD_REQUIRE(B > 0 && C > 0 && H >= kh && W >= kw && kh > 0 && kw > 0 && stride > 0, "bad sizes");
    PPD_REQUIRE(C % 4 == 0 && ld >= kh * kw * C && ld % 4 == 0, "channels and ld must be multiples of 4");
    PPD_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)cols % 16 == 0, "buffers must be 16-byte aligned");
    const int OH = (H - kh) / stride + 1, OW = (W - kw) / stride + 1;
    const int64_t total = (int64_t)B * OH * OW * (kh * kw * C / 4);
    if ((int64_t)B * OH <= 0x7fffffffLL) {
        const int K4 = kh * kw * C / 4;
        const int tx = K4 < 128 ? K4 : 128;
        int ty = 256 / tx;
        if (ty > OW) ty = OW;
        if (ty < 1) ty = 1;
        im2col_nhwc_rows_kernel<<<(unsigned)(B * OH), dim3(tx, ty), 0, ppd::as_stream(stream)>>>(x, H, W, C, kh, kw, stride,
                                                                                                 OH, OW, cols, ld);
        return ppd::launch_status("im2col_nhwc_rows_kernel");
    }
    im2col_nhwc_kernel<<<grid_for(total), kThreads, 0, ppd::as_stream(stream)>>>(x, B, H, W, C, kh, kw, stride, OH, OW,
                                                                               cols, ld);
    return ppd::launch_status("im2col_nhwc_kernel");
}

extern "C" int ppd_col2im_nhwc(const float* dcols, int64_t ld, int B, int H, int W, int C, int kh, int kw, int stride,
                               const float* act_mask, float* dx, void* stream) {
    PPD_REQUIRE(dcols && dx, "null pointer");
    PPD_REQUIRE(B > 0 && C > 0 && H >= kh && W >= kw && kh > 0 && kw > 0 && stride > 0, "bad sizes");
    PPD_REQUIRE(C % 4 == 0 && ld >= kh * kw * C && ld % 4 == 0, "channels and ld must be multiples of 4");
    PPD_REQUIRE((uintptr_t)dcols % 16 == 0 && (uintptr_t)dx % 16 == 0 && (uintptr_t)act_mask % 16 == 0,
                "buffers must be 16-byte aligned");
    const int OH = (H - kh) / stride + 1, OW = (W - kw) / stride + 1;
    const int64_t total = (int64_t)B * H * W * (C / 4);
    col2im_nhwc_kernel<<<grid_for(total), kThreads, 0, ppd::as_stream(stream)>>>(dcols, ld, B, H, W, C, kh, kw, stride,
                                                                               OH, OW, act_mask, dx);
    return ppd::launch_status("col2im_nhwc_kernel");
}

extern "C" int ppd_batched_transpose(const float* x, int64_t B, int R, int Cc, float* y, void* stream) {
    PPD_REQUIRE(x && y && x != y, "null or aliased pointer");
    PPD_REQUIRE(B > 0 && R > 0 && Cc > 0 && B <= 0x7fffffffLL, "bad sizes");
    const int n = R * Cc;
    const size_t smem = (size_t)(n + n / 32 + 1) * sizeof(float);
    PPD_REQUIRE(smem <= 48 * 1024, "matrix too large for the shared-memory transpose");
    batched_transpose_kernel<<<(unsigned)B, kThreads, smem, ppd::as_stream(stream)>>>(x, R, Cc, y);
    return ppd::launch_status("batched_transpose_kernel");
}

extern "C" int ppd_relu_mask(float* x, const float* act, int64_t n, void* stream) {
    PPD_REQUIRE(x && act, "null pointer");
    PPD_REQUIRE(n > 0 && n % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)act % 16 == 0, "n must be a multiple of 4, buffers 16-byte aligned");
    relu_mask_kernel<<<grid_for(n / 4), kThreads, 0, ppd::as_stream(stream)>>>(x, act, n / 4);
    return ppd::launch_status("relu_mask_kernel");
}
