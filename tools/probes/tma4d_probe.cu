// Cost of 4-D TMA boxes over an NHWC activation [B,20,20,32] as used by the implicit convolutions: swizzled 128-byte rows vs
// un-swizzled 256-byte rows, overlapping (im2col view) vs plain strides, one box per op, ops issued back to back by one thread.
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
__global__ void __launch_bounds__(32, 1) probe(const __grid_constant__ CUtensorMap tm, int nops, uint32_t box_bytes, int B, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) {
        long long tot = 0, tissue = 0;
        const int rounds = 64;
        for (int r = 0; r < rounds; ++r) {
            const int b = (blockIdx.x * rounds + r) % B;
            long long t0 = clock64();
            mbar_expect_tx(&bar, box_bytes * nops);
            for (int g = 0; g < nops; ++g)
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem + g * ((box_bytes + 127) & ~127u))), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(g), "r"(b) : "memory");
            long long t1 = clock64();
            mbar_wait(&bar, r & 1);
            long long t2 = clock64();
            tissue += t1 - t0; tot += t2 - t0;
        }
        if (blockIdx.x == 0) { out[0] = tissue / rounds; out[1] = tot / rounds; }
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    EncodeTiledFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    const int B = 2048;
    float* d; CK(cudaMalloc(&d, (size_t)B * 20 * 20 * 32 * 4)); CK(cudaMemset(d, 0, (size_t)B * 20 * 20 * 32 * 4));
    long long* o; CK(cudaMalloc(&o, 16));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    struct Cfg { const char* name; cuuint64_t d0, d1; cuuint64_t s1; cuuint32_t b0, b1; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; };
    // dims {d0 floats, d1 pixels (stride s1 bytes), 20 rows (stride 2560), B}
    Cfg cfgs[] = {
        {"im2col view  {128,9 @256B} box {32,9}  sw128", 128, 9, 256, 32, 9, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"im2col view  {128,9 @256B} box {64,9}  none ", 128, 9, 256, 64, 9, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"im2col view  {128,9 @256B} box {64,9}  none  L2-none", 128, 9, 256, 64, 9, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE},
        {"im2col view  {128,9 @256B} box {32,9}  none ", 128, 9, 256, 32, 9, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"im2col view  {128,9 @256B} box {128,9} none ", 128, 9, 256, 128, 9, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"plain view   {32,20 @128B} box {32,18} none ", 32, 20, 128, 32, 18, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"plain view   {64,10 @256B} box {64,9}  none ", 64, 10, 256, 64, 9, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"plain view   {32,20 @128B} box {32,18} sw128", 32, 20, 128, 32, 18, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
    };
    for (auto& c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {c.d0, c.d1, 20, (cuuint64_t)B}; cuuint64_t str[3] = {c.s1, 2560, 51200}; cuuint32_t box[4] = {c.b0, c.b1, 1, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("%-55s encode failed %d\n", c.name, (int)r); continue; }
        const uint32_t bb = c.b0 * c.b1 * 4;
        for (int nops : {1, 6}) {
            for (int grid : {1, 148}) {
                probe<<<grid, 32, 64 * 1024>>>(tm, nops, bb, B, o);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost));
                printf("%-55s ops=%d grid=%3d: issue %6lld clk, until landed %7lld clk  (%u B/op)\n", c.name, nops, grid, h[0], h[1], bb);
            }
        }
    }
    return 0;
}
