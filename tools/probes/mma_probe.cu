// tcgen05.mma kind::tf32 cost per instruction: N in {16..256}, M in {128, 64}, A from shared memory or tensor memory,
// one or two accumulators.      nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(128, 1) probe(int n, int a_tmem, int nacc, int reps, int m, int uniform, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_slot;
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // provably warp-uniform
    if (uniform && warp_u == 0) {
        // whole warp runs the loop (uniform control flow -> descriptors live in uniform registers); one elected lane issues
        const uint32_t idesc = make_idesc(m, n);
        const uint32_t sa = smem_u32(smem), sb = sa + 16384;
        long long t0 = clock64();
        const uint64_t db0 = make_desc(sb, 0, 1024, 2), da0 = make_desc(sa, 0, 1024, 2);
        for (int r = 0; r < reps; r += 8) {
            const uint32_t d = tb + (uint32_t)(((r >> 3) & (nacc - 1)) * 256);
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    // descriptor start address advances by 32 B per k-step: +2 in the low word
                    if (a_tmem) umma_ts(d, tb + 480 + (u & 3) * 8, db0 + (uint64_t)((u & 3) * 2), idesc, (r + u) >= nacc * 8);
                    else        umma_ss(d, da0 + (uint64_t)((u & 3) * 2), db0 + (uint64_t)((u & 3) * 2), idesc, (r + u) >= nacc * 8);
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (!uniform && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(m, n);
        const uint32_t sa = smem_u32(smem), sb = sa + 16384;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tb + (uint32_t)((r % nacc) * 256);
            const uint64_t db = make_desc(sb + (r & 3) * 32, 0, 1024, 2);
            if (a_tmem) umma_ts(d, tb + 480 + (r & 3) * 8, db, idesc, r >= nacc);
            else        umma_ss(d, make_desc(sa + (r & 3) * 32, 0, 1024, 2), db, idesc, r >= nacc);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main() {
    long long* d; CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int reps = 512;
    for (int uniform = 0; uniform < 2; ++uniform)
    for (int m : {128})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int nacc = 1; nacc <= 2; ++nacc)
            for (int n : {32, 64, 128, 256}) {
                long long h[2];
                for (int k = 0; k < 2; ++k) {
                    probe<<<1, 128, 50 * 1024>>>(n, a_tmem, nacc, reps, m, uniform, d);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                printf("uniform=%d M=%3d A=%s acc=%d N=%3d: issue %6.1f clk/mma, complete %6.1f clk/mma  (floor %d)\n", uniform, m, a_tmem ? "tmem" : "smem", nacc, n,
                       (double)h[0] / reps, (double)h[1] / reps, 128 * n / 256);
            }
    return 0;
}
