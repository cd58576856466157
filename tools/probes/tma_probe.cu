// TMA streaming-throughput probe: how fast can one CTA per SM pull a [rows x cols] fp32 matrix through TMA boxes of a given shape?
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(const void* src, uint64_t* bar, void* dst, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: 2-D tensor boxes; mode 1: 1-D bulk copies of box_bytes contiguous bytes
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, const float* base, int mode, int stages, uint32_t box_bytes,
                                               int box_cols, int box_rows, int ncol_tiles, int64_t ntiles, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t stride = (box_bytes + 1023) & ~1023u;
    if (threadIdx.x == 0) {
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const uint32_t s = it % stages;
            mbar_wait(&empty[s], ((it / stages) & 1u) ^ 1u);
            mbar_expect_tx(&full[s], box_bytes);
            if (mode == 0) {
                const int ct = (int)(t % ncol_tiles);
                const int64_t rt = t / ncol_tiles;
                tma_load_2d(&tm, &full[s], smem + s * stride, ct * box_cols, (int)(rt * box_rows));
            } else {
                bulk_load_1d(reinterpret_cast<const uint8_t*>(base) + t * (int64_t)box_bytes, &full[s], smem + s * stride, box_bytes);
            }
        }
    } else if (threadIdx.x == 32) {
        uint32_t it = 0;
        float acc = 0.f;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const uint32_t s = it % stages;
            mbar_wait(&full[s], (it / stages) & 1u);
            acc += *reinterpret_cast<volatile float*>(smem + s * stride);
            mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) *sink = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    const int64_t rows = 819200, cols = 192;          // the conv1 im2col matrix of a 2048-sample minibatch (629 MB)
    float* d; float* sink;
    CK(cudaMalloc(&d, rows * cols * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 0, rows * cols * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Cfg { const char* name; int mode; int bc, br; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; int stages; int grid; int64_t view_cols; };
    Cfg cfgs[] = {
        {"2d box 32x128 sw128 L2-256 s4 g148", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 4, 148, 192},
        {"2d box 32x128 sw128 L2-256 s8 g148", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 8, 148, 192},
        {"2d box 32x128 sw128 L2-128 s8 g148", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 8, 148, 192},
        {"2d box 32x128 sw128 L2-none s8 g148", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, 8, 148, 192},
        {"2d box 32x128 none  L2-256 s8 g148", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 8, 148, 192},
        {"2d box 32x256 sw128 L2-256 s4 g148", 0, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 4, 148, 192},
        {"2d box 32x64  sw128 L2-256 s16 g148", 0, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 16, 148, 192},
        {"2d box 192x32 none L2-256 s6 g148 (full rows)", 0, 192, 32, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 6, 148, 192},
        {"2d box 192x64 none L2-256 s4 g148 (full rows)", 0, 192, 64, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 4, 148, 192},
        {"2d box 32x128 sw128 as [rows*6 x 32] contiguous s8", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 8, 148, 32},
        {"2d box 256x16 none as [rows*192/256 x 256] contiguous s8", 0, 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 8, 148, 256},
        {"1d bulk 16 KB s8 g148", 1, 4096, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 8, 148, 192},
        {"1d bulk 16 KB s8 g296", 1, 4096, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4, 296, 192},
        {"1d bulk 4 KB s16 g148", 1, 1024, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 16, 148, 192},
        {"2d box 32x128 sw128 L2-256 s4 g296", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 4, 296, 192},
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& c : cfgs) {
        const int64_t vcols = c.view_cols, vrows = rows * cols / vcols;
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)vcols, (cuuint64_t)vrows};
        cuuint64_t strides[1] = {(cuuint64_t)vcols * 4};
        cuuint32_t box[2] = {(cuuint32_t)(c.mode ? 32 : c.bc), (cuuint32_t)(c.mode ? 8 : c.br)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-60s encode failed %d\n", c.name, (int)r); continue; }
        const uint32_t box_bytes = c.mode ? (uint32_t)c.bc * 4 : (uint32_t)c.bc * c.br * 4;
        const int nct = c.mode ? 1 : (int)(vcols / c.bc);
        const int64_t ntiles = c.mode ? rows * cols * 4 / box_bytes : nct * (vrows / c.br);
        const size_t smem = (size_t)c.stages * ((box_bytes + 1023) & ~1023u) + 1024;
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            probe<<<c.grid, 64, smem>>>(tm, d, c.mode, c.stages, box_bytes, c.bc, c.br, nct, ntiles, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("%-60s %8.1f us  %7.1f GB/s\n", c.name, best * 1e3, (double)ntiles * box_bytes / best / 1e6);
    }
    return 0;
}
