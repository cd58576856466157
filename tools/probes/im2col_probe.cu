// Does a 5-D tiled tensor map with OVERLAPPING strides (an implicit im2col view of an NCHW image) load correctly?
//   im2col_probe <variant>   1: dims {kx, ky, ox, oy, cb} sw128   2: same, no swizzle   3: strides ascending {kx, ox, ky, oy, cb}
//                            4: 3-D {kx, ky, cb}                   5: 4-D {kx, ky, ox, cb}
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__global__ void probe(const __grid_constant__ CUtensorMap tm, int nd, int c0, int c1, int c2, int c3, int c4, uint32_t bytes, float* out) {
    __shared__ __align__(1024) float tile[32 * 32];
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) tile[i] = -1.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        if (nd == 5)
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
        else if (nd == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 1;
    EncodeTiledFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    const int CB = 6;
    float* hi = (float*)malloc(CB * 7056 * 4);
    for (int i = 0; i < CB * 7056; ++i) hi[i] = (float)i;
    float *di, *o; CK(cudaMalloc(&di, CB * 7056 * 4)); CK(cudaMemcpy(di, hi, CB * 7056 * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&o, 32 * 32 * 4));
    CUtensorMap tm; CUresult r;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    int nd = 5; uint32_t bytes = 2560;
    int c[5] = {0, 4, 0, 7, 5};           // kx0, ky0, ox0, oy, cb
    CUtensorMapSwizzle sw = variant == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B;
    if (variant == 1 || variant == 2) {
        cuuint64_t dims[5] = {8, 8, 20, 20, CB}; cuuint64_t str[4] = {84 * 4, 4 * 4, 4 * 84 * 4, 7056 * 4}; cuuint32_t box[5] = {8, 4, 20, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, di, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (variant == 3) {
        cuuint64_t dims[5] = {8, 20, 8, 20, CB}; cuuint64_t str[4] = {4 * 4, 84 * 4, 4 * 84 * 4, 7056 * 4}; cuuint32_t box[5] = {8, 20, 4, 1, 1};
        c[1] = 0; c[2] = 4;
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, di, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (variant == 4) {
        cuuint64_t dims[3] = {8, 8, CB}; cuuint64_t str[2] = {84 * 4, 7056 * 4}; cuuint32_t box[3] = {8, 4, 1};
        nd = 3; bytes = 128; c[2] = 5;
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, di, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {8, 8, 20, CB}; cuuint64_t str[3] = {84 * 4, 4 * 4, 7056 * 4}; cuuint32_t box[4] = {8, 4, 20, 1};
        nd = 4; c[3] = 5;
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, di, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r) { printf("variant %d: encode failed %d\n", variant, (int)r); return 1; }
    probe<<<1, 128>>>(tm, nd, c[0], c[1], c[2], c[3], c[4], bytes, o);
    CK(cudaDeviceSynchronize());
    float h[32 * 32]; CK(cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost));
    printf("variant %d loaded; first rows:\n", variant);
    for (int rr = 0; rr < 3; ++rr) { for (int cc = 0; cc < 32; ++cc) printf("%g ", h[rr * 32 + cc]); printf("\n"); }
    if (variant == 1 || variant == 2) {
        int ok = 1;
        for (int ox = 0; ox < 20; ++ox) for (int ky = 0; ky < 4; ++ky) for (int kx = 0; kx < 8; ++kx) {
            const float want = (float)(5 * 7056 + (7 * 4 + 4 + ky) * 84 + ox * 4 + kx);
            const int k = ky * 8 + kx, ch = k / 4;
            const int cc = variant == 2 ? ch : (ch ^ (ox & 7));
            if (h[ox * 32 + cc * 4 + (k & 3)] != want) ok = 0;
        }
        printf("rows = pixels of 32 patch floats (kx fastest, then ky): %s\n", ok ? "YES" : "no");
    }
    return 0;
}
