// Cost of the per-k-block hand-off instructions of the MMA-issuer warp, in isolation: a whole warp loops over
//   try_wait on an already-completed mbarrier | tcgen05.fence::after_thread_sync | elect | tcgen05.commit (x1, x2) | __syncwarp
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o handoff_probe handoff_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// WHAT bits: 1 first ready try_wait, 2 second try_wait, 4 tcgen05.fence::after_thread_sync, 8 elect_one() section,
//            16 one tcgen05.commit inside it, 32 a second commit, 64 __syncwarp() after it, 128 lane-0 section instead of elect
template <int WHAT>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long* out) {
    __shared__ __align__(8) uint64_t done_bar, done_bar2, sink_bar[4];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1); mbar_init(&done_bar2, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&sink_bar[i], 1u << 20);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive(&done_bar); mbar_arrive(&done_bar2);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
            if (WHAT & 1) mbar_wait(&done_bar, 0);
            if (WHAT & 2) mbar_wait(&done_bar2, 0);
            if (WHAT & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (WHAT & 8) {
                if (elect_one()) {
                    if (WHAT & 16) umma_commit(&sink_bar[r & 3]);
                    if (WHAT & 32) umma_commit(&sink_bar[(r + 1) & 3]);
                }
            }
            if (WHAT & 128) {
                if (lane == 0) {
                    if (WHAT & 16) umma_commit(&sink_bar[r & 3]);
                    if (WHAT & 32) umma_commit(&sink_bar[(r + 1) & 3]);
                }
            }
            if (WHAT & 64) __syncwarp();
        }
        long long t1 = clock64();
        if (lane == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(32u) : "memory");
}

template <int WHAT>
static void run(const char* name, long long* d) {
    const int reps = 4000;
    long long h = 0;
    for (int k = 0; k < 2; ++k) { probe<WHAT><<<1, 128>>>(reps, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("%-60s %7.1f clk / iteration\n", name, (double)h / reps);
}


// The transform <-> MMA-issuer ring of tca_gemm_kernel with nothing in it: GROUPS groups of four warps take alternate k-blocks,
// each waits ta_empty[it % TA] (a tcgen05.commit of the issuer), optionally stores 64 columns to TMEM, and arrives on
// ta_full; the issuer warp waits ta_full, (optionally issues nothing), and commits to ta_empty.  Clocks per k-block.
template <int GROUPS, int TA, int STTM>
__global__ void __launch_bounds__(64 + 128 * GROUPS, 1) chain(int nkb, long long* out) {
    __shared__ __align__(8) uint64_t ta_full[TA], ta_empty[TA];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TA; ++i) { mbar_init(&ta_full[i], 4); mbar_init(&ta_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    long long t0 = clock64();
    if (warp == 1) {
        for (int it = 0; it < nkb; ++it) {
            const uint32_t ts = it % TA;
            mbar_wait(&ta_full[ts], (it / TA) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) umma_commit(&ta_empty[ts]);
            __syncwarp();
        }
        // drain: the last commits
        for (int i = 0; i < TA && i < nkb; ++i) { const int it = nkb - 1 - i; mbar_wait(&ta_empty[it % TA], (it / TA) & 1u); }
        long long t1 = clock64();
        if (lane == 0) out[0] = t1 - t0;
    } else if (warp >= 2) {
        const int grp = (warp - 2) >> 2, q = warp & 3;
        uint32_t v[16];
        for (int c = 0; c < 16; ++c) v[c] = lane + c;
        for (int it = grp; it < nkb; it += GROUPS) {
            const uint32_t ts = it % TA;
            mbar_wait(&ta_empty[ts], ((it / TA) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (STTM) {
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ts * 64u;
                for (int j = 0; j < 4; ++j)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                 ::"r"(ta + 16u * j), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&ta_full[ts]);
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int GROUPS, int TA, int STTM>
static void run_chain(long long* d) {
    const int nkb = 4096;
    long long h = 0;
    for (int k = 0; k < 2; ++k) { chain<GROUPS, TA, STTM><<<1, 64 + 128 * GROUPS>>>(nkb, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("ring: %d group(s), %d TMEM stages, %s: %7.1f clk / k-block\n", GROUPS, TA, STTM ? "4 x tcgen05.st.x16 + wait::st" : "no TMEM stores", (double)h / nkb);
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    run<0>("empty loop", d);
    run<1>("1 ready try_wait", d);
    run<3>("2 ready try_waits (different barriers)", d);
    run<4>("fence::after", d);
    run<8>("elect section (empty)", d);
    run<8 | 64>("elect section + syncwarp", d);
    run<8 | 16>("elect { commit }", d);
    run<8 | 16 | 32>("elect { 2 commits }", d);
    run<128 | 16>("lane0 { commit }", d);
    run<128 | 16 | 32>("lane0 { 2 commits }", d);
    run<1 | 4 | 8 | 16>("wait, fence, elect { commit }", d);
    run<3 | 4 | 8 | 16 | 32>("2 waits, fence, elect { 2 commits }", d);
    run<3 | 4 | 8 | 16 | 32 | 64>("2 waits, fence, elect { 2 commits }, syncwarp", d);
    run_chain<1, 1, 0>(d); run_chain<1, 2, 0>(d); run_chain<1, 4, 0>(d);
    run_chain<2, 2, 0>(d); run_chain<2, 4, 0>(d); run_chain<2, 8, 0>(d); run_chain<3, 6, 0>(d);
    run_chain<1, 4, 1>(d); run_chain<2, 4, 1>(d); run_chain<2, 8, 1>(d); run_chain<3, 6, 1>(d);
    return 0;
}
