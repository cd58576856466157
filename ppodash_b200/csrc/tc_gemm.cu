// tcgen05 (5th-gen tensor core) GEMM for the policy network's dense contractions, TF32 inputs with
// fp32 accumulation in tensor memory.  Same problem statement as sgemm.cu:
//   C[i,j] (+)= sum_kk OpA(i,kk) * OpB(j,kk),  operands row-major fp32 in HBM, either contraction-
//   contiguous ("k-major": OpA(i,kk) = A[i*lda + kk]) or output-contiguous ("mn-major": A[kk*lda + i]).
// fp32 operands are fed to the tensor cores as they are (kind::tf32 reads the top 19 bits), so there
// are no conversion passes and no second copy of weights or activations.
//
// Structure (one 128 x BLOCK_N output tile per CTA, optional split of the contraction over gridDim.z):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d loads of 128B-swizzled tiles into a 4-stage
//            shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1   MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BLOCK_N,
//            K=8 per instruction, 4 per 32-wide k-block) from shared-memory descriptors into a TMEM
//            accumulator; tcgen05.commit releases ring slots / signals the epilogue
//   warps 2-9 (3xTF32 mode) split every landed tile into hi / lo TF32 parts in shared memory; then the
//            epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias / ReLU / ReLU-backward mask /
//            accumulate, vectorised global stores (or transposed / scatter-added / raw split-K partials)
// Shared-memory tile layouts are the canonical UMMA ones (cute/atom/mma_traits_sm100.hpp):
//   k-major : rows of 32 floats (128 B), 8-row 1024 B swizzle atoms, SBO = 1024 B; one TMA box
//             {32 k, rows}; successive MMAs advance the descriptor start address by 32 B
//   mn-major: 32-bit mn-major operands must use the 128B swizzle with 32-byte atomicity
//             (UMMA SWIZZLE_128B_BASE32B / TMA SWIZZLE_128B_ATOM_32B): 512 B atoms of [4 k][32 mn]; one
//             TMA box {32 mn, 32 k} per 32 output rows, atoms along mn LBO = 4096 B apart, along k
//             SBO = 512 B apart; successive MMAs (8 k each) advance by 1024 B
#include <cuda.h>

#include "ppd_common.cuh"
#include "tca_gemm.cuh"

namespace {

constexpr int kMaxStages = 4;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;   // warps 2-9: operand split during the main loop, then the epilogue
constexpr int BM = 128;            // UMMA M
constexpr int BK = 32;             // floats per k-block (one 128-byte swizzle row)
constexpr int kTmemCols = 256;

struct Args {
    float* C; int64_t ldc;
    int64_t I, J, KK;
    const float* bias; const float* mask; int64_t ldm;
    int relu, accumulate, transpose_out;
    int block_n, a_mn, b_mn;
    int split3;                // 3xTF32: also multiply the low-order residuals (fp32-level accuracy)
    int a_tmem;                // split3 only: A tiles go smem -> registers (hi/lo split) -> tensor memory; the MMAs read A from TMEM
    int stages;
    int64_t kk_per_split;
    float* partial;
    // fused col2im (dgrad of an NHWC convolution): column j = (ky, kx, c), row i = (b, oy, ox);
    // the tile is scatter-added into dx[b, oy*stride+ky, ox*stride+kx, c] instead of being stored
    int scatter, OH, OW, kw, cstride, Cin, Hin, Win;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row, one 32-bit column per contraction element), B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version 1 <<46 | layout type <<61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
constexpr uint32_t kLayoutSw128 = 2;        // UMMA::LayoutType::SWIZZLE_128B          (k-major tiles)
constexpr uint32_t kLayoutSw128Base32 = 1;  // UMMA::LayoutType::SWIZZLE_128B_BASE32B  (mn-major 32-bit tiles)

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b_format TF32 (2) @7/@10,
// a_major @15, b_major @16, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (uint32_t)(a_mn ? 1 : 0) << 15;
    d |= (uint32_t)(b_mn ? 1 : 0) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t ready_bar[kMaxStages];     // split3: residual tiles written
    __shared__ __align__(8) uint64_t ta_empty_bar[2];           // a_tmem: TMEM A stage consumed by its MMAs
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    // 1024-byte aligned ring: [stage][A tile 16 KB | B tile block_n*128 B | (split3: A residual | B residual)]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int bn = a.block_n;
    const int kStages = a.stages;
    const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)bn * BK * 4;
    const uint32_t tx_bytes = a_bytes + b_bytes;
    // a_tmem: [A raw | B hi | B lo]; split3 in shared memory: [A hi | B hi | A lo | B lo]; else [A | B]
    const uint32_t stage_bytes = a.a_tmem ? a_bytes + 2 * b_bytes : (a.split3 ? 2 * tx_bytes : tx_bytes);
    const uint32_t b_lo_off = a.a_tmem ? b_bytes : tx_bytes;     // B residual tile relative to the B tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = (int64_t)blockIdx.y * BM;
    const int64_t j0 = (int64_t)blockIdx.x * bn;
    const int64_t kk_begin = (int64_t)blockIdx.z * a.kk_per_split;
    const int64_t kk_end = min(a.KK, kk_begin + a.kk_per_split);
    const int nkb = (int)((kk_end - kk_begin + BK - 1) / BK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&ready_bar[s], kEpiThreads);
        }
        mbar_init(&ta_empty_bar[0], 1);
        mbar_init(&ta_empty_bar[1], 1);
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                mbar_expect_tx(&full_bar[s], tx_bytes);
                const int kk = (int)(kk_begin + (int64_t)kb * BK);
                if (!a.a_mn) {
                    tma_load_2d(&tmA, &full_bar[s], sa, kk, (int)i0);                         // box {32 k, 128 rows}
                } else if (a.a_tmem) {
                    tma_load_2d(&tmA, &full_bar[s], sa, (int)i0, kk);                         // box {128 rows(i), 32 k}, no swizzle
                } else {
                    for (int q = 0; q < BM / 32; ++q)                                          // box {32 rows(i), 32 k}
                        tma_load_2d(&tmA, &full_bar[s], sa + q * 4096, (int)i0 + 32 * q, kk);
                }
                if (!a.b_mn) {
                    tma_load_2d(&tmB, &full_bar[s], sb, kk, (int)j0);                         // box {32 k, block_n rows}
                } else {
                    for (int q = 0; q < bn / 32; ++q)
                        tma_load_2d(&tmB, &full_bar[s], sb + q * 4096, (int)j0 + 32 * q, kk);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer
            const uint32_t idesc = make_idesc(bn, a.a_mn, a.b_mn);
            const uint32_t idesc_ts = make_idesc(bn, 0, a.b_mn);      // A from tensor memory is always k-major
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(a.split3 ? &ready_bar[s] : &full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t sb = sa + a_bytes;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint32_t oa = a.a_mn ? k * 1024 : k * 32, ob = a.b_mn ? k * 1024 : k * 32;
                    const uint64_t da = a.a_mn ? make_desc(sa + oa, 4096, 512, kLayoutSw128Base32)
                                               : make_desc(sa + oa, 0, 1024, kLayoutSw128);
                    const uint64_t db = a.b_mn ? make_desc(sb + ob, 4096, 512, kLayoutSw128Base32)
                                               : make_desc(sb + ob, 0, 1024, kLayoutSw128);
                    if (a.a_tmem) {
                        // A hi / lo of this k-block sit in TMEM columns [ta, ta+32) / [ta+32, ta+64), 8 per MMA
                        const uint32_t ta = tmem_base + 128u + (uint32_t)(kb & 1) * 64u + (uint32_t)k * 8u;
                        const uint64_t dbl = a.b_mn ? make_desc(sb + b_lo_off + ob, 4096, 512, kLayoutSw128Base32)
                                                    : make_desc(sb + b_lo_off + ob, 0, 1024, kLayoutSw128);
                        umma_tf32_ts(tmem_base, ta + 32u, db, idesc_ts, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32_ts(tmem_base, ta, dbl, idesc_ts, 1u);
                        umma_tf32_ts(tmem_base, ta, db, idesc_ts, 1u);
                    } else if (a.split3) {
                        // x = hi + lo with hi = the 19 bits the tensor core reads; add the small terms first
                        const uint64_t dal = a.a_mn ? make_desc(sa + tx_bytes + oa, 4096, 512, kLayoutSw128Base32)
                                                    : make_desc(sa + tx_bytes + oa, 0, 1024, kLayoutSw128);
                        const uint64_t dbl = a.b_mn ? make_desc(sb + tx_bytes + ob, 4096, 512, kLayoutSw128Base32)
                                                    : make_desc(sb + tx_bytes + ob, 0, 1024, kLayoutSw128);
                        umma_tf32(tmem_base, dal, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_base, da, dbl, idesc, 1u);
                        umma_tf32(tmem_base, da, db, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);          // ring slot free once these MMAs have read it
                if (a.a_tmem) umma_commit(&ta_empty_bar[kb & 1]);
            }
            umma_commit(&tmem_full_bar);             // accumulator complete
        }
    } else {
        // ================= (split3) hi = round-to-nearest TF32 of x (written back in place), lo = TF32(x - hi)
        // written next to it at the same (swizzled) offsets; both are then handed to the async proxy.
        // Rounding (instead of the tensor core's truncation of the raw fp32 bits) keeps the neglected
        // lo*lo term and the rounding of lo zero-mean, so long sums with cancellation (weight gradients)
        // do not pick up a bias.
        if (a.a_tmem) {
            // A: each thread owns one tile row (TMEM lane) and half of the 32 k columns; warps 2-5 take k 0..15,
            // warps 6-9 k 16..31 (a warp may only touch TMEM lanes [32*(warp%4), +32)).  B is split in place.
            const int et = threadIdx.x - 64;
            const int r = (warp & 3) * 32 + lane, h = (warp - 2) >> 2;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                const uint8_t* sa = smem + (size_t)s * stage_bytes;
                float x[16], hi[16], lo[16];
                if (!a.a_mn) {
                    // 128B-swizzled rows of 32 floats: 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
                    const float4* rowp = reinterpret_cast<const float4*>(sa + r * 128);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = rowp[(4 * h + c) ^ (r & 7)];
                        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                    }
                } else {
                    const float* colp = reinterpret_cast<const float*>(sa) + (16 * h) * BM + r;   // [k][128 rows]
#pragma unroll
                    for (int c = 0; c < 16; ++c) x[c] = colp[c * BM];
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) split_tf32(x[c], hi[c], lo[c]);
                mbar_wait(&ta_empty_bar[kb & 1], ((uint32_t)(kb >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)(kb & 1) * 64u + 16u * h;
                tmem_st16(ta, hi);
                tmem_st16(ta + 32u, lo);
                float4* src = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes);
                float4* dst = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes + b_bytes);
                const int nvec = (int)(b_bytes >> 4);
                for (int v = et; v < nvec; v += kEpiThreads) {
                    const float4 xb = src[v];
                    float4 hb, rb;
                    split_tf32(xb.x, hb.x, rb.x);
                    split_tf32(xb.y, hb.y, rb.y);
                    split_tf32(xb.z, hb.z, rb.z);
                    split_tf32(xb.w, hb.w, rb.w);
                    src[v] = hb;
                    dst[v] = rb;
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
            }
        } else if (a.split3) {
            const int et = threadIdx.x - 64;                         // 0..255
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                float4* src = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
                float4* dst = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + tx_bytes);
                const int nvec = (int)(tx_bytes >> 4);
                for (int v = et; v < nvec; v += kEpiThreads) {
                    const float4 x = src[v];
                    float4 h, r;
                    split_tf32(x.x, h.x, r.x);
                    split_tf32(x.y, h.y, r.y);
                    split_tf32(x.z, h.z, r.z);
                    split_tf32(x.w, h.w, r.w);
                    src[v] = h;
                    dst[v] = r;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
            }
        }
        // ================= epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
        // two warps share each TMEM lane quarter (a warp may only touch lanes [32*(warp%4), +32)): warps 2-5 take
        // the even 32-column chunks, warps 6-9 the odd ones
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t i = i0 + q * 32 + lane;
        for (int c0 = 32 * half; c0 < bn; c0 += 64) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (nkb == 0) {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0.f;
            }
            const int64_t jb = j0 + c0;
            if (a.partial) {
                if (i < a.I) {
                    float* P = a.partial + ((int64_t)blockIdx.z * a.I + i) * a.J;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (jb + c < a.J) P[jb + c] = v[c];
                }
                continue;
            }
            if (a.bias) {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] += (jb + c < a.J) ? __ldg(a.bias + jb + c) : 0.f;
            }
            if (a.relu) {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
            }
            if (a.scatter) {
                // 32 consecutive columns = 32 consecutive input channels of one filter tap (Cin is a multiple of 32)
                if (i < a.I && jb < a.J) {
                    const int tap = (int)(jb / a.Cin), c0 = (int)(jb - (int64_t)tap * a.Cin);
                    const int ky = tap / a.kw, kx = tap - ky * a.kw;
                    const int64_t b = i / ((int64_t)a.OH * a.OW);
                    const int rem = (int)(i - b * a.OH * a.OW);
                    const int oy = rem / a.OW, ox = rem - oy * a.OW;
                    float* dst = a.C + (((b * a.Hin + (oy * a.cstride + ky)) * (int64_t)a.Win + (ox * a.cstride + kx)) * a.Cin + c0);
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[c]), "f"(v[c + 1]),
                                     "f"(v[c + 2]), "f"(v[c + 3])
                                     : "memory");
                }
            } else if (a.transpose_out) {
                // C is [J, I] row-major: for fixed column the 32 lanes write 32 consecutive floats
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (i < a.I && jb + c < a.J) {
                        float* p = a.C + (jb + c) * a.ldc + i;
                        float x = v[c];
                        if (a.mask) x = (__ldg(a.mask + (jb + c) * a.ldm + i) > 0.f) ? x : 0.f;
                        *p = a.accumulate ? (*p + x) : x;
                    }
                }
            } else if (i < a.I) {
                float* crow = a.C + i * a.ldc + jb;
                const float* mrow = a.mask ? a.mask + i * a.ldm + jb : nullptr;
                const bool vec = (jb + 31 < a.J) && ((a.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) &&
                                 (!mrow || (((a.ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)));
                if (vec) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        float4 x = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                        if (mrow) {
                            const float4 m = __ldg(reinterpret_cast<const float4*>(mrow + c));
                            x.x = m.x > 0.f ? x.x : 0.f; x.y = m.y > 0.f ? x.y : 0.f;
                            x.z = m.z > 0.f ? x.z : 0.f; x.w = m.w > 0.f ? x.w : 0.f;
                        }
                        float4* p = reinterpret_cast<float4*>(crow + c);
                        if (a.accumulate) { const float4 o = *p; x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
                        *p = x;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (jb + c < a.J) {
                            float x = v[c];
                            if (mrow) x = (__ldg(mrow + c) > 0.f) ? x : 0.f;
                            crow[c] = a.accumulate ? (crow[c] + x) : x;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
}

__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t I, int64_t J, float* __restrict__ C,
                        int64_t ldc, const float* __restrict__ bias, const float* __restrict__ mask, int64_t ldm,
                        int relu, int accumulate, int transpose_out) {
    const int64_t n = I * J;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        // four independent accumulators, eight loads in flight: the kernel is a latency chain over `splits` (dozens), not bandwidth;
        // the order of the sum is fixed (deterministic), only its association differs from a serial sum
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* p = partial + e;
        int z = 0;
#pragma unroll 2
        for (; z + 4 <= splits; z += 4) {
            a0 += __ldg(p + (int64_t)z * n); a1 += __ldg(p + (int64_t)(z + 1) * n);
            a2 += __ldg(p + (int64_t)(z + 2) * n); a3 += __ldg(p + (int64_t)(z + 3) * n);
        }
        for (; z < splits; ++z) a0 += __ldg(p + (int64_t)z * n);
        float v = (a0 + a1) + (a2 + a3);
        const int64_t i = e / J, j = e - i * J;
        if (bias) v += __ldg(bias + j);
        if (relu) v = fmaxf(v, 0.f);
        const int64_t co = transpose_out ? (j * ldc + i) : (i * ldc + j);
        if (mask) v = (__ldg(mask + (transpose_out ? (j * ldm + i) : (i * ldm + j))) > 0.f) ? v : 0.f;
        C[co] = accumulate ? (C[co] + v) : v;
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map over a row-major matrix [rows, cols] (cols contiguous, row stride ld floats),
// box = {box_cols, box_rows}, 128-byte swizzle, out-of-bounds elements read as zero.
int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
             CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { ppd::set_error("ppd_tc_gemm: cuTensorMapEncodeTiled not available"); return PPD_EINVAL; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ppd::set_error("ppd_tc_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return PPD_EINVAL; }
    return 0;
}

int g_two_ctas = 1;
int g_persistent = 1;     // 3xTF32 products run on the persistent TMEM-A kernel (tca_gemm.cu); 0 = the kernel in this file
int g_a_tmem = 1;         // 3xTF32: stage the A operand through registers into tensor memory (0 = split it in shared memory)
int g_force_bn = 0;       // tuning: 0 = heuristic, else 32/64/128/256 where it divides the problem sensibly
struct Plan { int bn; int64_t gx, gy; int splits; int64_t kk_per_split; size_t ws; };

int pick_bn(int64_t I, int64_t J, bool split3, bool a_kmajor) {
    if (g_force_bn && J > g_force_bn / 2) return g_force_bn;
    if (J <= 32) return 32;
    if (J <= 64) return 64;
    // 3xTF32 forward / dgrad products of a small minibatch (FC, GRU input projection: 16 row tiles): 64-wide tiles
    // give twice the CTAs, two of which fit per SM (48 KB stages) -- measured 2x faster than 128-wide on B200
    if (split3 && a_kmajor && ((I + BM - 1) / BM) * ((J + 127) / 128) < 4 * ppd::kNumSMs) return 64;
    if (J <= 128) return 128;
    if (!split3 && J <= 256 && J > 192) return 256;
    return 128;
}

Plan make_plan(int64_t I, int64_t J, int64_t KK, size_t ws_avail, bool limit, bool split3 = false, bool a_kmajor = false) {
    Plan p;
    p.bn = pick_bn(I, J, split3, a_kmajor);
    p.gx = (J + p.bn - 1) / p.bn;
    p.gy = (I + BM - 1) / BM;
    const int64_t tiles = p.gx * p.gy;
    int64_t splits = 1;
    const int64_t target = 2 * ppd::kNumSMs;
    if (tiles < target && KK >= 16 * BK) {
        splits = (target + tiles - 1) / tiles;
        const int64_t max_splits = KK / (8 * BK);
        if (splits > max_splits) splits = max_splits;
        if (splits > 1024) splits = 1024;
        if (splits < 1) splits = 1;
    }
    if (limit) {
        while (splits > 1 && (size_t)splits * I * J * sizeof(float) > ws_avail) --splits;
    }
    int64_t per = (KK + splits - 1) / splits;
    per = (per + BK - 1) / BK * BK;
    splits = (KK + per - 1) / per;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;
    p.kk_per_split = per;
    p.ws = splits > 1 ? (size_t)splits * I * J * sizeof(float) : 0;
    return p;
}

}  // namespace

extern "C" void ppd_tc_gemm_set_option(int v) {
    if (v == 0 || v == 1) g_two_ctas = v;
    else if (v == 2 || v == 3) g_a_tmem = v - 2;
    else if (v == 4 || v == 5) g_persistent = v - 4;
    else if (v == 6 || v == 7) ppd::tca::g_conv_resident = v - 6;
    else if (v == 8 || v == 9) ppd::tca::g_b_resident = v - 8;
    else if (v == 32 || v == 64 || v == 128 || v == 256) g_force_bn = v;
    else if (v == -1) g_force_bn = 0;
    else if (v >= 1000 && v <= 1000 + ppd::kNumSMs) ppd::tca::g_max_ctas = (v == 1000) ? ppd::kNumSMs : v - 1000;
}

extern "C" size_t ppd_tc_gemm_workspace(int64_t I, int64_t J, int64_t KK) {
    if (I <= 0 || J <= 0 || KK <= 0) return 0;
    const size_t a = make_plan(I, J, KK, 0, false, false, false).ws, b = make_plan(I, J, KK, 0, false, true, true).ws;
    const size_t c = ppd::tca::make_plan(I, J, KK, 0, false).ws;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);      // upper bound over the kernels' tile choices
}

// 1 if ppd_tc_gemm can run this problem (alignment of the operands for TMA), else 0.
extern "C" int ppd_tc_gemm_supported(const ppd_gemm_args* g) {
    if (!g || !g->A || !g->B || !g->C) return 0;
    if (g->I <= 0 || g->J <= 0 || g->KK <= 0) return 0;
    if ((g->lda & 3) || (g->ldb & 3)) return 0;
    if (((uintptr_t)g->A & 15) || ((uintptr_t)g->B & 15)) return 0;
    if (g->I > 0x7fffffffLL || g->J > 0x7fffffffLL || g->KK > 0x7fffffffLL) return 0;
    return 1;
}

namespace {
int tc_gemm_impl(const ppd_gemm_args* g, int flags, const ppd_conv_geom* geom, void* workspace, size_t workspace_bytes,
                 void* stream, const float* b_lo = nullptr);
}

extern "C" int ppd_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    PPD_REQUIRE(x && hi && lo && n >= 0 && (n & 3) == 0, "n must be a multiple of 4");
    PPD_REQUIRE(!(((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15), "pointers must be 16-byte aligned");
    return ppd::tca::split_operand(x, hi, lo, n, ppd::as_stream(stream));
}

extern "C" int ppd_conv_fwd_nhwc(const float* x, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                                 const float* bias, int relu, float* out, void* stream) {
    PPD_REQUIRE(x && geom && w_hi && w_lo && out, "null pointer");
    return ppd::tca::conv_forward(x, geom, Cout, w_hi, w_lo, bias, relu, out, ppd::as_stream(stream));
}

extern "C" int ppd_conv_fwd_nchw(const float* x, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                                 const float* bias, int relu, float* out, void* stream) {
    PPD_REQUIRE(x && geom && w_hi && w_lo && out, "null pointer");
    return ppd::tca::conv_forward_nchw(x, geom, Cout, w_hi, w_lo, bias, relu, out, ppd::as_stream(stream));
}

extern "C" int ppd_conv_dgrad_nhwc(const float* dy, const ppd_conv_geom* geom, int Cout, const float* w_hi, const float* w_lo,
                                   const float* act_mask, float* dx, void* stream) {
    PPD_REQUIRE(dy && geom && w_hi && w_lo && dx, "null pointer");
    return ppd::tca::conv_dgrad(dy, geom, Cout, w_hi, w_lo, act_mask, dx, ppd::as_stream(stream));
}

extern "C" size_t ppd_conv_wgrad_workspace(const ppd_conv_geom* geom, int Cout) {
    return geom ? ppd::tca::conv_wgrad_workspace(geom, Cout) : 0;
}

extern "C" int ppd_conv_wgrad(const float* x, const ppd_conv_geom* geom, int nchw, const float* dy, int Cout, float* dW, int accumulate,
                              void* workspace, size_t workspace_bytes, void* stream) {
    PPD_REQUIRE(x && geom && dy && dW, "null pointer");
    cudaStream_t ts = ppd::as_stream(stream);
    int splits = 0;
    int rc = ppd::tca::conv_wgrad(x, geom, nchw, dy, Cout, dW, accumulate, workspace, workspace_bytes, ts, &splits);
    if (rc) return rc;
    const int64_t K = (int64_t)geom->kh * geom->kw * geom->C;
    int64_t nb = (K * Cout + 255) / 256;
    if (nb > 4 * ppd::kNumSMs) nb = 4 * ppd::kNumSMs;
    // partials are [split][K, Cout]; dW is [Cout, K]: reduce and store transposed
    tc_splitk_reduce_kernel<<<(unsigned)nb, 256, 0, ts>>>(reinterpret_cast<float*>(workspace), splits, K, Cout, dW, K, nullptr, nullptr, 0,
                                                          0, accumulate, 1);
    return ppd::launch_status("tc_splitk_reduce_kernel");
}

extern "C" int ppd_tc_gemm_bsplit(const ppd_gemm_args* g, const float* b_lo, int flags, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    PPD_REQUIRE(g && b_lo && (flags & PPD_TC_SPLIT3), "pre-split B operands are a 3xTF32 feature");
    PPD_REQUIRE(!((uintptr_t)b_lo & 15), "b_lo must be 16-byte aligned");
    return tc_gemm_impl(g, flags, nullptr, workspace, workspace_bytes, stream, b_lo);
}

extern "C" int ppd_tc_gemm(const ppd_gemm_args* g, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    return tc_gemm_impl(g, flags, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ppd_tc_gemm_col2im(const ppd_gemm_args* g, const ppd_conv_geom* geom, int flags, void* stream) {
    PPD_REQUIRE(g && geom, "null pointer");
    PPD_REQUIRE(!(flags & PPD_TC_TRANSPOSE_OUT) && !g->bias && !g->mask && !g->relu, "plain product only");
    PPD_REQUIRE(geom->C % 32 == 0 && geom->kh > 0 && geom->kw > 0 && geom->stride > 0, "channels must be a multiple of 32");
    const int OH = (geom->H - geom->kh) / geom->stride + 1, OW = (geom->W - geom->kw) / geom->stride + 1;
    PPD_REQUIRE(g->I == (int64_t)geom->B * OH * OW && g->J == (int64_t)geom->kh * geom->kw * geom->C, "GEMM shape does not match the geometry");
    PPD_REQUIRE(((uintptr_t)g->C & 15) == 0, "dx must be 16-byte aligned");
    return tc_gemm_impl(g, flags, geom, nullptr, 0, stream);
}

namespace {
int tc_gemm_impl(const ppd_gemm_args* g, int flags, const ppd_conv_geom* geom, void* workspace, size_t workspace_bytes,
                 void* stream, const float* b_lo) {
    const int transpose_out = flags & PPD_TC_TRANSPOSE_OUT;
    const int split3 = (flags & PPD_TC_SPLIT3) ? 1 : 0;
    PPD_REQUIRE(ppd_tc_gemm_supported(g), "operands must be 16-byte aligned with leading dimensions that are multiples of 4");
    PPD_REQUIRE(geom || (transpose_out ? g->ldc >= g->I : g->ldc >= g->J), "bad ldc");
    PPD_REQUIRE(!b_lo || (split3 && !geom), "pre-split B needs the persistent 3xTF32 kernel");
    if (split3 && (g_persistent || b_lo) && !geom) {
        ppd::tca::Plan tp;
        cudaStream_t ts = ppd::as_stream(stream);
        int trc = ppd::tca::launch(g, transpose_out, b_lo, workspace, workspace_bytes, ts, &tp);
        if (trc || tp.splits == 1) return trc;
        int64_t nb = (g->I * g->J + 255) / 256;
        if (nb > 4 * ppd::kNumSMs) nb = 4 * ppd::kNumSMs;
        tc_splitk_reduce_kernel<<<(unsigned)nb, 256, 0, ts>>>(reinterpret_cast<float*>(workspace), tp.splits, g->I, g->J, g->C,
                                                              g->ldc, g->bias, g->mask, g->ldm, g->relu, g->accumulate, transpose_out);
        return ppd::launch_status("tc_splitk_reduce_kernel");
    }
    Plan p = make_plan(g->I, g->J, g->KK, workspace ? workspace_bytes : 0, true, split3, g->a_kmajor != 0);
    if (geom && p.splits > 1) {           // the scatter epilogue adds complete products only
        p.splits = 1;
        p.kk_per_split = (g->KK + BK - 1) / BK * BK;
    }
    PPD_REQUIRE(p.gy <= 65535 && p.splits <= 65535, "grid too large");
    CUtensorMap tmA, tmB;
    int rc;
    // k-major operand: matrix [rows = I or J, cols = KK]; mn-major: matrix [rows = KK, cols = I or J]
    const int a_tmem = (split3 && g_a_tmem && p.bn <= 128) ? 1 : 0;
    if (g->a_kmajor) rc = make_map(&tmA, g->A, g->I, g->KK, g->lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    else if (a_tmem) rc = make_map(&tmA, g->A, g->KK, g->I, g->lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    else             rc = make_map(&tmA, g->A, g->KK, g->I, g->lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (g->b_kmajor) rc = make_map(&tmB, g->B, g->J, g->KK, g->ldb, BK, p.bn, CU_TENSOR_MAP_SWIZZLE_128B);
    else             rc = make_map(&tmB, g->B, g->KK, g->J, g->ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    Args a;
    a.C = g->C; a.ldc = g->ldc; a.I = g->I; a.J = g->J; a.KK = g->KK;
    a.bias = g->bias; a.mask = g->mask; a.ldm = g->ldm; a.relu = g->relu; a.accumulate = g->accumulate;
    a.transpose_out = transpose_out;
    a.block_n = p.bn; a.a_mn = g->a_kmajor ? 0 : 1; a.b_mn = g->b_kmajor ? 0 : 1;
    a.kk_per_split = p.kk_per_split;
    a.partial = p.splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
    a.split3 = split3;
    a.a_tmem = a_tmem;
    a.scatter = geom ? 1 : 0;
    if (geom) {
        a.OH = (geom->H - geom->kh) / geom->stride + 1; a.OW = (geom->W - geom->kw) / geom->stride + 1;
        a.kw = geom->kw; a.cstride = geom->stride; a.Cin = geom->C; a.Hin = geom->H; a.Win = geom->W;
    } else {
        a.OH = a.OW = a.kw = a.cstride = a.Cin = a.Hin = a.Win = 0;
    }
    const size_t stage = a_tmem ? (size_t)BM * BK * 4 + 2 * (size_t)p.bn * BK * 4
                                : (size_t)(split3 ? 2 : 1) * (BM * BK * 4 + (size_t)p.bn * BK * 4);
    // Ring depth: if two CTAs (2 x 256 TMEM columns) can be co-resident with at least a 2-deep ring each, size
    // the ring for that -- one CTA's prologue / epilogue then hides behind the other's main loop (the kernel
    // is not persistent); otherwise give the single CTA as deep a ring as fits.
    int stages = (int)((110 * 1024) / stage);
    if (stages < 2 || !g_two_ctas) stages = (int)((196 * 1024) / stage);
    if (stages > kMaxStages) stages = kMaxStages;
    a.stages = stages;
    const size_t smem = (size_t)stages * stage + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { ppd::set_error("ppd_tc_gemm: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
        attr_set = true;
    }
    cudaStream_t s = ppd::as_stream(stream);
    dim3 grid((unsigned)p.gx, (unsigned)p.gy, (unsigned)p.splits);
    tc_gemm_kernel<<<grid, kThreads, smem, s>>>(tmA, tmB, a);
    rc = ppd::launch_status("tc_gemm_kernel");
    if (rc || p.splits == 1) return rc;
    int64_t nb = (g->I * g->J + 255) / 256;
    if (nb > 4 * ppd::kNumSMs) nb = 4 * ppd::kNumSMs;
    tc_splitk_reduce_kernel<<<(unsigned)nb, 256, 0, s>>>(a.partial, p.splits, g->I, g->J, g->C, g->ldc, g->bias, g->mask,
                                                         g->ldm, g->relu, g->accumulate, transpose_out);
    return ppd::launch_status("tc_splitk_reduce_kernel");
}
}  // namespace
