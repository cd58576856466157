// Persistent 3xTF32 tcgen05 GEMM with the A operand staged through tensor memory.
//
//   C[i,j] (+)= sum_kk OpA(i,kk) * OpB(j,kk)   (problem statement of sgemm.cu / tc_gemm.cu)
//
// fp32-level accuracy on the tensor cores needs every operand split into hi = TF32(x) and lo = TF32(x - hi)
// and three products (A_lo B_hi + A_hi B_lo + A_hi B_hi).  Doing that split in shared memory (tc_gemm.cu)
// costs 7 shared-memory passes over every A tile (TMA write, split read, two split writes, three MMA
// reads).  Here the A tile makes two: TMA writes it, the transform warps read it into registers, split it
// there and store hi / lo into TENSOR MEMORY (tcgen05.st); the MMAs take A from TMEM
// (tcgen05.mma ... [d], [a], b-desc) and only the (small) B tile from shared memory.  Because the tile
// passes through registers, row-major and transposed A tiles, and the patch tiles of an implicit-GEMM
// convolution, all land in the same TMEM layout (lane = output row, one 32-bit column per contraction
// element); no operand-layout constraints reach the TMA side.
//
// One CTA per SM, persistent over (m tile, n tile, k split) work items; 16 warps:
//   warp 0      TMA producer of A (ring of kSA x 16 KB: deep enough to cover the HBM latency at full bandwidth)
//   warp 1      MMA issuer; owns the TMEM allocation (512 columns: 2 tiles x [main | correction] accumulators, 2 A slots x 128 = two k-blocks each)
//   warp 2      TMA producer of B (ring of [hi | lo] tiles), independent of the A ring
//   warps 4-11  transform: smem A tile -> registers -> hi/lo -> TMEM A stage (and the split of B in shared
//               memory when the caller has no pre-split copy of it)
//   warps 12-15 epilogue: tcgen05.ld of the finished accumulator while the next tile's MMAs fill the other; the tile leaves through a
//               shared-memory staging tile and TMA bulk tensor stores where the output is dense (a thread owns a ROW of the tile, and
//               row-per-thread global stores cost the LSU one cache line per lane: measured as a ~1000-clock bubble per tile in
//               the transform warps and the issuer), else through 256-bit row stores
// Every hand-off is an mbarrier; tcgen05.commit releases B stages, TMEM A stages and accumulators.  k-blocks are handed from the
// transform groups to the issuer in pairs (one 128-column TMEM slot per group).  Launches are programmatic dependent launches.
#include <cuda.h>
#include <stdlib.h>

#include "ppd_common.cuh"
#include "tca_gemm.cuh"

namespace ppd {
namespace tca {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef PPD_TCA_TRACE
// debug builds: a wait that does not complete within ~2^26 polls reports itself and traps instead of hanging the GPU
__device__ __noinline__ void mbar_wait_timeout(uint64_t* bar, uint32_t parity) {
    printf("tca_gemm: mbarrier wait timed out: block %d thread %d barrier smem offset %u parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
           smem_u32(bar), parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t n = 0;; ++n) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (n > (1u << 24)) mbar_wait_timeout(bar, parity);
    }
}
#else
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
#endif
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 1-D bulk copy of `bytes` contiguous bytes (16-byte aligned on both sides)
__device__ __forceinline__ void bulk_load_1d(const void* src, uint64_t* bar, uint32_t dst, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
#ifdef PPD_ST32
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
        : "memory");
}
#endif
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

// hi = x rounded to TF32 (10 mantissa bits, round half away in magnitude), lo = (x - hi) rounded likewise.
// Integer rounding on the bit pattern: 5 instructions per element (cvt.rna.tf32.f32 expands to ~8 on sm_100a).
// x - hi is exact; Inf stays Inf, the largest finite values round to Inf as round-to-nearest does.
// Used for the B operand (weights, split once per forward by split_tf32_kernel).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    const float r = x - __uint_as_float(hi);
    lo = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
// The A operand is split by the transform warps once per k-block and tile -- 32 elements per thread, and the transform
// warps' instruction issue is what bounds the kernel (ncu: 163 of ~350 instructions per warp and k-block were the 5-instruction
// split).  Two instructions: hi = x with the low 13 mantissa bits cleared (exactly representable in TF32), lo = x - hi, exact,
// stored with all its bits -- the tensor core reads the TF32 part of it.  |x - hi - tf32(lo)| <= 2^-20 |x| (2^-22 with the
// rounding split): still far below the 1e-5 gate, and the [hi | lo] pair never meets the rounding split of B in one product term.
__device__ __forceinline__ void split_a(float x, uint32_t& hi, uint32_t& lo) {
#ifdef PPD_SPLIT_ROUND
    split_tf32(x, hi, lo);
#else
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
#endif
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
constexpr uint32_t kLayoutSw128 = 2;        // k-major tiles
constexpr uint32_t kLayoutSw128Base32 = 1;  // mn-major 32-bit tiles

// c_format F32 @4, a/b_format TF32 @7/@10, a_major @15 (0: A from TMEM is k-major), b_major @16, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(int n, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (uint32_t)(b_mn ? 1 : 0) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

struct Item { int64_t i0, j0, kk_begin; int nkb; int z; int cls, seg0, nvalid, kb0; };

// x / d for 0 <= x < 2^24, d >= 1, through the float reciprocal.  Role loops are warp-uniform code, and an integer division
// by a run-time divisor there is expanded on the UNIFORM datapath (no reciprocal unit, ~10-clock dependent ops): measured
// ~2000 clocks per division-laden TMA issue.  I2F / MUFU.RCP / F2I run on the vector pipes in ~50.
__device__ __forceinline__ int fdiv(int x, int d) {
    int q = __float2int_rz(__int2float_rn(x) * __frcp_rn(__int2float_rn(d)));
    const int r = x - q * d;
    if (r < 0) --q;
    else if (r >= d) ++q;
    return q;
}
// stage index + phase bit of a ring with a run-time number of stages
struct Ring {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// Pipeline timeline of CTA 0 (debug builds only: -DPPD_TCA_TRACE): trace[(it * 16 + slot)] = clock64()
#ifdef PPD_TCA_TRACE2
// low-perturbation timeline: clock64 stamps go to SHARED memory (CS2R + STS), CTA 0 dumps them after its last tile
__device__ unsigned g_trace2[24 * 8 + 8];     // + entry / exit / first issue / first epilogue stamps
#define TR2(it_, slot_) do { if (blockIdx.x == 0 && lane == 0 && (it_) >= 32u && (it_) < 56u) tr2[((it_) - 32u) * 8u + (slot_)] = (unsigned)clock64(); } while (0)
#else
#define TR2(it_, slot_) do { } while (0)
#endif
#ifdef PPD_TCA_TRACE
__device__ long long* g_trace = nullptr;
#define TCA_TRACE(it_, slot_) do { if (g_trace && blockIdx.x == 0 && (it_) < 256 && lane == 0) g_trace[(it_) * 16 + (slot_)] = clock64(); } while (0)
#define TCA_TRACE1(it_, slot_) do { if (g_trace && blockIdx.x == 0 && (it_) < 256) g_trace[(it_) * 16 + (slot_)] = clock64(); } while (0)
#else
#define TCA_TRACE(it_, slot_) do { } while (0)
#define TCA_TRACE1(it_, slot_) do { } while (0)
#endif

template <int MODE>
__device__ __forceinline__ Item decode(const Args& a, int w) {
    constexpr int mode = MODE == 5 ? 3 : MODE;
    Item it;
    it.kb0 = 0;
    if (mode == 3) {
        it.z = fdiv(w, a.num_m);
        const int mt = w - it.z * a.num_m;
        it.i0 = (int64_t)mt * BM; it.j0 = 0; it.kk_begin = 0;
        it.kb0 = it.z * a.conv.kbps;
        it.nkb = min(a.conv.kbps, a.conv.total_kb - it.kb0);
        it.cls = mt; it.seg0 = 0; it.nvalid = 0;
        return it;
    }
    if (mode) {
        it.cls = fdiv(w, a.conv.ntile_class);
        it.seg0 = (w - it.cls * a.conv.ntile_class) * a.conv.nseg;
        it.nvalid = min(a.conv.nseg, a.conv.nseg_class - it.seg0);
        it.i0 = 0; it.j0 = 0; it.kk_begin = 0; it.z = 0;
        it.nkb = a.conv.nkb;
        return it;
    }
    it.cls = it.seg0 = it.nvalid = 0;
    const int r = fdiv(w, a.num_n);
    const int n = w - r * a.num_n;
    it.z = fdiv(r, a.num_m);
    const int m = r - it.z * a.num_m;
    it.i0 = (int64_t)m * BM;
    it.j0 = (int64_t)n * a.bn;
    it.kk_begin = (int64_t)it.z * a.kk_per_split;
    const int64_t kk_end = min(a.KK, it.kk_begin + a.kk_per_split);
    it.nkb = (int)((kk_end - it.kk_begin + BK - 1) / BK);
    return it;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Roles branch on a warp index the compiler can prove warp-uniform (shfl), and every role loop is executed by the
// whole warp with one ELECTED lane issuing the TMA / MMA / commit instructions.  That keeps tensor-map coordinates,
// descriptors and barrier addresses in uniform registers: issued from a divergent `lane == 0` branch, each
// tcgen05.mma cost ~185 clocks (seven R2UR moves per instruction) instead of ~33 (measured, tools/probes/mma_probe.cu).
template <int MODE>       // ConvA::mode, as a compile-time constant (5 = mode 3 with raw rows): one lean kernel per operand layout
__global__ void __launch_bounds__(kThreads, 1)
tca_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC, const Args a) {
    constexpr int mode = MODE == 5 ? 3 : MODE;
    constexpr bool raw = MODE == 5;
    // Weight gradients (modes 3 / 5) of the N = 32 layers: the activation B tile is split into [hi | lo] by the four EPILOGUE warps,
    // which otherwise idle for the ~170-350 k-blocks of a work item; everywhere else a non-pre-split B is split by the transform warps.
    constexpr bool epi_splits_b = (MODE == 3 || MODE == 5);
    const bool esb = epi_splits_b && !a.b_presplit && a.bn <= 32;      // measured: -13 % for the N = 32 layers, +2 % at N = 64
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[kSA], empty_a[kSA], full_b[kMaxSB], empty_b[kMaxSB];
    __shared__ __align__(8) uint64_t ta_full[kTA], ta_empty[kTA], acc_full[2], acc_empty[2];
    __shared__ __align__(8) uint64_t bs_full[kMaxSB];   // weight gradients: B stage split into [hi | lo] by the epilogue warps
    __shared__ uint32_t tmem_base_slot;
#ifdef PPD_TCA_TRACE2
    __shared__ unsigned tr2[24 * 8 + 8];
    for (int i = threadIdx.x; i < 24 * 8 + 8; i += blockDim.x) tr2[i] = 0;
    if (threadIdx.x == 0) tr2[24 * 8] = (unsigned)clock64();
#endif
    __shared__ int tap_line[kMaxTaps];                  // modes 6 / 7: line offset of k-block kb inside the tile's stage

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int bn = a.bn;
    const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)bn * BK * 4;
    uint8_t* smemA = smem;                               // kSA stages of one 128 x 32 fp32 tile
    uint8_t* smemB = smem + (a.a_region_bytes ? a.a_region_bytes : kSA * a_bytes);      // kSB stages of [B hi | B lo]
    const uint32_t kSB = (uint32_t)a.sb_stages;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

#ifdef PPD_TCA_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("tca_gemm barriers: full_a %u empty_a %u full_b %u empty_b %u ta_full %u ta_empty %u acc_full %u acc_empty %u | mode %d items %d\n",
               smem_u32(full_a), smem_u32(empty_a), smem_u32(full_b), smem_u32(empty_b), smem_u32(ta_full), smem_u32(ta_empty),
               smem_u32(acc_full), smem_u32(acc_empty), mode, a.total_items);
#endif
    if (threadIdx.x == 0) {
        // convolution stages are filled by TWO producer warps (each arms the barrier for its own boxes)
        for (int s = 0; s < kSA; ++s) { mbar_init(&full_a[s], mode ? kAProd : 1); mbar_init(&empty_a[s], (mode == 6 || mode == 7) ? kXformWarps : 4); }
        for (int s = 0; s < kMaxSB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); mbar_init(&bs_full[s], kEpiWarps); }
        for (int s = 0; s < kTA; ++s) { mbar_init(&ta_full[s], 4); mbar_init(&ta_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (a.b_presplit) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
        if (a.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    }
    if ((mode == 6 || mode == 7) && (int)threadIdx.x >= 64 && (int)threadIdx.x - 64 < a.conv.nkb) {
        // k-block kb = (channel plane, filter column, filter row), planes fastest.  Forward: tap (ky, kx) is lw*ky + kx lines
        // further; dgrad: tap (dky, dkx) reads dY row i - dky, pixel j - dkx.
        const ConvA& cv = a.conv;
        const int kb = (int)threadIdx.x - 64, fw = mode == 7 ? cv.T : cv.KW;
        const int lw = mode == 7 ? cv.segw + cv.T - 1 : cv.Win;
        const int k_p = kb % cv.planes, tap = kb / cv.planes, k_kx = tap % fw, k_ky = tap / fw;
        tap_line[kb] = k_p * cv.nrows_max * lw + (mode == 7 ? -(k_ky * lw + k_kx) : (k_ky * lw + k_kx));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    // programmatic dependent launch: let the next kernel of the stream start its own prologue as SMs free up, and wait here for the
    // previous kernel (and its memory) before the first global access of any role
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0 && mode == 0) {
        // ================= TMA producer, A tiles of a plain GEMM: one box per stage
        uint32_t it = 0;
        for (int w = blockIdx.x; w < a.total_items; w += gridDim.x) {
            const Item t = decode<MODE>(a, w);
            for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                const int kk = (int)(t.kk_begin + (int64_t)kb * BK);
                const uint32_t s = it % kSA;
                mbar_wait(&empty_a[s], ((it / kSA) & 1u) ^ 1u);
                if (elect_one()) {
                    TCA_TRACE1(it, 0);
                    uint8_t* sa = smemA + s * a_bytes;
                    mbar_expect_tx(&full_a[s], a_bytes);
                    if (!a.a_mn) tma_load_2d(&tmA, &full_a[s], sa, kk, (int)t.i0);     // box {32 k, 128 rows}, 128B swizzle
                    else         tma_load_2d(&tmA, &full_a[s], sa, (int)t.i0, kk);     // box {128 rows, 32 k}, no swizzle
                }
                __syncwarp();
            }
        }
    } else if ((warp == 0 || warp == 3 || warp >= kFirstExtra) && mode) {
        // ================= TMA producers of a convolution: a stage is several boxes, shared between warp 0 and warp 3, each
        // walking its boxes from one elected thread (~55 clocks per box in a uniform-datapath loop; every lane issuing its
        // own box costs ~200: the compiler serialises them through R2UR broadcasts).  The k loop carries NO division: the
        // position inside the filter and the (sample, row) of the first box advance as counters -- a run-time division on
        // this path (even the float-reciprocal one) costs the producer hundreds of clocks per k-block, and the producer's own
        // loop time, not the TMA engine or the transform warps, was what bounded every implicit convolution.
        const ConvA& cv = a.conv;
        const int pw = warp == 0 ? 0 : (warp == 3 ? 1 : warp - kFirstExtra + 2);       // producer index 0 .. kAProd-1
        uint32_t it = 0;
        if (mode == 6 || mode == 7) {
            // tile-resident raw input: one stage per TILE (two stages).  Forward (6): image rows s*oy_a .. s*oy_b + KH - 1 of every
            // run of output rows inside one sample; dgrad (7): dY rows i_a - T + 1 .. i_b of every run of dx rows, T - 1 pixels of
            // left halo (out-of-bounds rows / pixels read as zero).  Per channel plane, plain NHWC boxes {32 channels, lw pixels};
            // rows alternate between the two producer warps.  Nothing is loaded per k-block.
            const bool isd = mode == 7;
            const int lw = isd ? cv.segw + cv.T - 1 : cv.Win;                 // lines per staged row
            const uint32_t row_bytes = (uint32_t)lw * 128u;
            uint32_t tile_it = 0;
            for (int w = blockIdx.x; w < a.total_items; w += gridDim.x, ++tile_it) {
                const Item t = decode<MODE>(a, w);
                const uint32_t st = tile_it & 1u;
                mbar_wait(&empty_a[st], ((tile_it >> 1) & 1u) ^ 1u);
                const int b0 = fdiv(t.seg0, cv.rows_per_img);
                const int oy0 = t.seg0 - b0 * cv.rows_per_img;
                if (elect_one()) {
                    // pass 1: rows of this producer
                    int mine = 0;
                    {
                        int left = t.nvalid, oy = oy0, slot = 0;
                        while (left > 0) {
                            const int run = min(left, cv.rows_per_img - oy);
                            const int nr = isd ? run + cv.T - 1 : cv.s * (run - 1) + cv.KH;
                            mine += (nr + ((slot & 1) == pw ? 1 : 0)) >> 1;           // rows with (slot + y) & 1 == pw
                            slot += nr; left -= run; oy = 0;
                        }
                    }
                    mbar_expect_tx(&full_a[st], (uint32_t)(mine * cv.planes) * row_bytes);
                    const uint32_t base = smem_u32(smemA) + st * cv.tile_stage_bytes;
                    int left = t.nvalid, oy = oy0, slot = 0, b = b0;
                    while (left > 0) {
                        const int run = min(left, cv.rows_per_img - oy);
                        const int nr = isd ? run + cv.T - 1 : cv.s * (run - 1) + cv.KH;
                        const int y0 = isd ? oy - (cv.T - 1) : cv.s * oy, x0 = isd ? -(cv.T - 1) : 0;
                        for (int y = ((slot & 1) == pw) ? 0 : 1; y < nr; y += 2)
                            for (int p2 = 0; p2 < cv.planes; ++p2)
                                tma_load_4d(&tmA, &full_a[st], base + (uint32_t)((p2 * cv.nrows_max + slot + y)) * row_bytes, p2 * 32, x0,
                                            y0 + y, b);
                        slot += nr; left -= run; oy = 0; ++b;
                    }
                }
                __syncwarp();
            }
        } else if (mode == 3 && raw) {
            // weight gradient over NCHW observations, raw rows: a k-block is 32 consecutive pixels of the flattened (b, oy, ox)
            // grid = at most 3 output rows; output row (b, oy) of channel c needs image rows s*oy .. s*oy + kh - 1, one
            // CONTIGUOUS range -> one 1-D bulk copy per (output row, channel).  (The 5-D im2col view of the same data is 80 runs
            // of 32 bytes per box, 480 per k-block: TMA-engine-bound.)  Stage = [channel half][output row 0..2][kh rows][W].
            const uint32_t row_bytes = (uint32_t)(cv.KW * cv.Win) * 4u;          // kh == kw image rows
            const int OW = cv.spr;                                                // pixels per output row
            for (int w = blockIdx.x; w < a.total_items; w += gridDim.x) {
                const Item t = decode<MODE>(a, w);
                const int ch0 = 2 * t.cls;
                const int nhalf = min(2, cv.nchunks - ch0);
                if (pw >= nhalf) {                                                 // nothing to load: only arrive
                    for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                        const uint32_t s = it % kSA;
                        mbar_wait(&empty_a[s], ((it / kSA) & 1u) ^ 1u);
                        if (elect_one()) mbar_expect_tx(&full_a[s], 0u);
                        __syncwarp();
                    }
                    continue;
                }
                const int ch = ch0 + pw;
                int P0 = t.kb0 * 32;
                const int R = fdiv(P0, OW);                                       // first output row (flattened b*OH + oy) and pixel
                int ox0 = P0 - R * OW;
                const int b = fdiv(R, cv.rows_per_img);
                int oy = R - b * cv.rows_per_img;
                // image rows of output row (b, oy), channel ch: a running pointer, no multiplications in the k loop
                const int row_step = cv.s * cv.Win;
                const size_t wrap = (size_t)cv.C * cv.Hin * cv.Win - (size_t)cv.rows_per_img * row_step;
                const float* src0 = a.a_ptr + ((size_t)(b * cv.C + ch) * cv.Hin + (size_t)oy * cv.s) * cv.Win;
                const bool leader = elect_one();
                for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                    const uint32_t s = it % kSA;
                    mbar_wait(&empty_a[s], ((it / kSA) & 1u) ^ 1u);
                    const int e = ox0 + min(32, cv.total_seg - P0);
                    const int nrows = 1 + (e > OW ? 1 : 0) + (e > 2 * OW ? 1 : 0);      // a k-block is at most 3 output rows
                    if (leader) {
                        TCA_TRACE1(it, pw ? 10 : 0);
                        mbar_expect_tx(&full_a[s], (uint32_t)nrows * row_bytes);
                        uint32_t dst = smem_u32(smemA + s * a_bytes) + (uint32_t)pw * 3u * row_bytes;
                        const float* src = src0;
                        int oy2 = oy;
                        for (int r = 0; r < nrows; ++r, dst += row_bytes) {
                            bulk_load_1d(src, &full_a[s], dst, row_bytes);
                            src += row_step;
                            if (++oy2 == cv.rows_per_img) { oy2 = 0; src += wrap; }
                        }
                    }
                    // advance 32 pixels
                    P0 += 32;
                    ox0 += 32;
                    while (ox0 >= OW) { ox0 -= OW; src0 += row_step; if (++oy == cv.rows_per_img) { oy = 0; src0 += wrap; } }
                    __syncwarp();
                }
            }
        } else if (mode == 3) {
            // One walk over the segments of a k-block per producer: every lane advances the running position (sample b, output
            // row oy, row segment sub) in uniform registers, the elected lane issues the boxes that are this producer's.  (A
            // separate skip / issue / catch-up walk cost ~245 uniform-datapath instructions per k-block: the A producers, not
            // TMA or the transform warps, bounded the weight-gradient kernels.)
            const uint32_t seg_bytes = (uint32_t)cv.segw * 256u;
            for (int w = blockIdx.x; w < a.total_items; w += gridDim.x) {
                const Item t = decode<MODE>(a, w);
                const int ch0 = 2 * t.cls;
                const int nhalf = min(2, cv.nchunks - ch0);
                // two chunk halves: producers (pw & 1) take one half each, and with four producers (pw >> 1) splits the segments;
                // one half: the segments are split between all producers
                const int hh = nhalf == 2 ? (pw & 1) : 0;
                const int part = nhalf == 2 ? (pw >> 1) : pw, nparts = nhalf == 2 ? kAProd / 2 : kAProd;
                const int ch = ch0 + hh;
                const int ky = cv.nchw ? 0 : fdiv(ch, cv.cpr);
                const int j0 = (ch - ky * cv.cpr) * 64;
                int seg = t.kb0 * cv.nseg;
                const int rowidx0 = fdiv(seg, cv.spr);
                int b = fdiv(rowidx0, cv.rows_per_img);
                int oy = rowidx0 - b * cv.rows_per_img, sub = seg - rowidx0 * cv.spr;
                int x = sub * cv.segw, y = oy * cv.s + ky, cb = b * cv.C + ch;
                const bool leader = elect_one();
                for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                    const uint32_t s = it % kSA;
                    mbar_wait(&empty_a[s], ((it / kSA) & 1u) ^ 1u);
                    const int nvalid = min(cv.nseg, cv.total_seg - seg);
                    const int per = nparts == 1 ? nvalid : (nparts == 2 ? (nvalid + 1) >> 1 : (nvalid + 3) >> 2);
                    const int g_lo = min(nvalid, part * per), g_hi = min(nvalid, g_lo + per);
                    if (leader) {
                        TCA_TRACE1(it, pw ? 10 : 0);
                        mbar_expect_tx(&full_a[s], (uint32_t)(g_hi - g_lo) * seg_bytes);
                    }
                    uint32_t dst = smem_u32(smemA + s * a_bytes) + hh * 8192u + g_lo * seg_bytes;
                    for (int g = 0; g < nvalid; ++g) {
                        if (leader && g >= g_lo && g < g_hi) {
                            if (cv.nchw) tma_load_5d(&tmA, &full_a[s], dst, 0, 0, x, oy, cb);
                            else         tma_load_4d(&tmA, &full_a[s], dst, j0, x, y, b);
                            dst += seg_bytes;
                        }
                        x += cv.segw;
                        if (++sub == cv.spr) {
                            sub = 0; x = 0; y += cv.s;
                            if (++oy == cv.rows_per_img) { oy = 0; y = ky; ++b; cb += cv.C; }
                        }
                    }
                    seg += nvalid;
                    __syncwarp();
                }
            }
        } else {
            const int kyg = 32 / max(cv.KW, 1);
            const uint32_t seg_bytes = mode == 4 ? (uint32_t)(kyg * cv.Win) * 4u : (uint32_t)cv.segw * 128u;
            const uint32_t seg_pitch = seg_bytes;
            for (int w = blockIdx.x; w < a.total_items; w += gridDim.x) {
                const Item t = decode<MODE>(a, w);
                // modes 1, 2, 4: the segments of a tile are consecutive rows (b, row); the two producer warps take half each
                const int per = (t.nvalid + kAProd - 1) / kAProd;
                const int lo_op = min(t.nvalid, pw * per), hi_op = min(t.nvalid, lo_op + per);
                const int sg0 = t.seg0 + lo_op;
                const int b0 = fdiv(sg0, cv.rows_per_img);
                const int row0 = sg0 - b0 * cv.rows_per_img;
                // position inside the filter, advanced once per k-block:
                //   mode 1: (c0 = float offset in the filter row, ky)   mode 2: (c0 = 32-channel chunk, dkx, dky)   mode 4: (kyq, channel)
                int k0 = 0, k1 = 0, k2 = 0;
                for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                    const uint32_t s = it % kSA;
                    mbar_wait(&empty_a[s], ((it / kSA) & 1u) ^ 1u);
                    if (elect_one()) {
                        TCA_TRACE1(it, pw ? 10 : 0);
                        mbar_expect_tx(&full_a[s], (uint32_t)(hi_op - lo_op) * seg_bytes);
                        uint32_t dst = smem_u32(smemA + s * a_bytes) + lo_op * seg_pitch;
                        int row = row0, b = b0;
                        // one tight loop per mode: c = the coordinate that moves with the row, the rest is fixed for the stage
                        if (mode == 1) {
                            const int c0 = k0 * 32;
                            int y = row * cv.s + k1;
                            for (int g = lo_op; g < hi_op; ++g, dst += seg_pitch) {
                                tma_load_4d(&tmA, &full_a[s], dst, c0, 0, y, b);
                                y += cv.s;
                                if (++row == cv.rows_per_img) { row = 0; ++b; y = k1; }
                            }
                        } else if (mode == 2) {
                            const int c0 = k0 * 32, x0 = -k1;
                            int y = row - k2;
                            for (int g = lo_op; g < hi_op; ++g, dst += seg_pitch) {
                                tma_load_4d(&tmA, &full_a[s], dst, c0, x0, y, b);
                                ++y;
                                if (++row == cv.rows_per_img) { row = 0; ++b; y = -k2; }
                            }
                        } else {
                            // mode 4: output row oy needs the kyg = s image rows s*oy + k0*kyg .. +kyg-1 of channel k1, so the rows of
                            // CONSECUTIVE output rows are one contiguous range of the image: one 1-D bulk copy per run inside a sample
                            int left = hi_op - lo_op;
                            while (left > 0) {
                                const int run = min(left, cv.rows_per_img - row);
                                const float* src = a.a_ptr + ((size_t)(b * cv.C + k1) * cv.Hin + row * cv.s + k0 * kyg) * cv.Win;
                                bulk_load_1d(src, &full_a[s], dst, (uint32_t)run * seg_bytes);
                                dst += (uint32_t)run * seg_bytes; left -= run; row = 0; ++b;
                            }
                        }
                    }
                    if (mode == 2) { if (++k0 == cv.kpk) { k0 = 0; if (++k1 == cv.T) { k1 = 0; ++k2; } } }
                    else              { if (++k0 == cv.kpk) { k0 = 0; ++k1; } }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 2) {
        // ================= TMA producer, B tiles
        // Resident B (convolutions whose weight tiles of ALL k-blocks fit): stage kb holds k-block kb for as long as the tile class
        // (dgrad: the parity class that selects the filter taps) does not change -- the weights were otherwise re-streamed through
        // TMA for every tile, as many bytes as the patch stream itself (conv1 forward: 328 MB of weights for 173 MB of pixels).
        uint32_t it = 0, epoch = 0;
        int loaded_cls = -1;
        Ring rb;
        for (int w = blockIdx.x; w < a.total_items; w += gridDim.x) {
            const Item t = decode<MODE>(a, w);
            const bool reload = !a.b_resident || t.cls != loaded_cls;
            // dgrad: (channel chunk, dkx, dky) of the k-block as counters, parity class of the tile
            int bk0 = 0, bk1 = 0, bk2 = 0;
            const bool dgrad = mode == 2 || mode == 7;
            const int bpy = dgrad ? fdiv(t.cls, a.conv.s) : 0, bpx = dgrad ? t.cls - bpy * a.conv.s : 0;
            for (int kb = 0; kb < t.nkb; ++kb, ++it, rb.next(kSB)) {
                const int kk = (int)(t.kk_begin + (int64_t)kb * BK);
                const uint32_t s = a.b_resident ? (uint32_t)kb : rb.s;
                if (reload) mbar_wait(&empty_b[s], a.b_resident ? ((epoch & 1u) ^ 1u) : (rb.ph ^ 1u));
                if (reload && elect_one()) {
                    TCA_TRACE1(it, 1);
                    mbar_expect_tx(&full_b[s], a.b_presplit ? 2 * b_bytes : b_bytes);
                    uint8_t* sb = smemB + s * 2 * b_bytes;
                    if (mode == 3) {
                        // dY [pixels, Cout]: the 32 pixel rows starting at the k-block's first pixel (rows past its last valid
                        // pixel belong to the next k-block: finite values that meet the zeroed pad slots of A)
                        const int p0 = (t.kb0 + kb) * a.conv.nseg * a.conv.segw;
                        for (int q = 0; q < bn / 32; ++q) tma_load_2d(&tmB, &full_b[s], sb + q * 4096, 32 * q, p0);
                    } else if (mode == 2 || mode == 7) {
                        // W [Cout, (ky, kx, c)]: box {32 c, 32 couts} of filter tap (ky, kx) = (py + s dky, px + s dkx)
                        const ConvA& cv = a.conv;
                        const int col = ((bpy + cv.s * bk2) * cv.KW + (bpx + cv.s * bk1)) * cv.Cin;
                        for (int q = 0; q < bn / 32; ++q) {
                            tma_load_2d(&tmB, &full_b[s], sb + q * 4096, col + 32 * q, bk0 * 32);
                            tma_load_2d(&tmBlo, &full_b[s], sb + b_bytes + q * 4096, col + 32 * q, bk0 * 32);
                        }
                    } else if (!a.b_mn) {
                        tma_load_2d(&tmB, &full_b[s], sb, kk, (int)t.j0);               // box {32 k, bn rows}
                        if (a.b_presplit) tma_load_2d(&tmBlo, &full_b[s], sb + b_bytes, kk, (int)t.j0);
                    } else {
                        for (int q = 0; q < bn / 32; ++q) {                             // box {32 cols(j), 32 k}
                            tma_load_2d(&tmB, &full_b[s], sb + q * 4096, (int)t.j0 + 32 * q, kk);
                            if (a.b_presplit) tma_load_2d(&tmBlo, &full_b[s], sb + b_bytes + q * 4096, (int)t.j0 + 32 * q, kk);
                        }
                    }
                }
                if (dgrad) { if (++bk0 == a.conv.kpk) { bk0 = 0; if (++bk1 == a.conv.T) { bk1 = 0; ++bk2; } } }
                __syncwarp();
            }
            if (a.b_resident && reload) { loaded_cls = t.cls; ++epoch; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer.  Two accumulators per tile: main = A_hi B_hi in columns [0, bn) and the
        // corrections A_hi B_lo + A_lo B_hi in columns [bn, 2bn).  The B stage holds [hi rows | lo rows] back to back,
        // so ONE N = 2bn MMA multiplies A_hi with both; a second N = bn MMA adds A_lo B_hi: two instructions per
        // k-step instead of three, and the truncating tcgen05 accumulation sees a third of the steps per accumulator.
        const uint32_t idesc2 = make_idesc(2 * bn, a.b_mn), idesc1 = make_idesc(bn, a.b_mn);
        const uint64_t bdesc0 = a.b_mn ? make_desc(smem_u32(smemB), 4096, 512, kLayoutSw128Base32)
                                       : make_desc(smem_u32(smemB), 0, 1024, kLayoutSw128);
        const uint32_t kstep = a.b_mn ? (1024u >> 4) : (32u >> 4);          // descriptor start-address step per 8 k
        uint32_t it = 0, tile_it = 0, epoch = 0, pit = 0;
        int loaded_cls = -1;
        Ring rb;
        for (int w = blockIdx.x; w < a.total_items; w += gridDim.x, ++tile_it) {
#ifdef PPD_TCA_TRACE2
            if (blockIdx.x == 0 && lane == 0 && tile_it == 7) tr2[24 * 8 + 6] = (unsigned)clock64();
#endif
            const Item t = decode<MODE>(a, w);
            const bool reload = !a.b_resident || t.cls != loaded_cls;
            // resident B: its stages are released only when this CTA's NEXT tile needs other weights
            const int wn = w + gridDim.x;
            const bool release = !a.b_resident || (wn < a.total_items && decode<MODE>(a, wn).cls != t.cls);
            const uint32_t acc = tile_it & 1u;
#ifdef PPD_TCA_TRACE2
            if (blockIdx.x == 0 && lane == 0 && tile_it == 7) tr2[24 * 8 + 7] = (unsigned)clock64();
#endif
            mbar_wait(&acc_empty[acc], ((tile_it >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + kAccCol0 + acc * kAccStride;
            // k-blocks are handed over in PAIRS (one tensor-memory slot of 128 columns = two k-blocks, filled by ONE transform group):
            // one ta_full wait, one fence, one election and one ta_empty commit per 16 MMAs.  Every mbarrier wait costs this warp
            // 150-250 clocks even when its phase is long complete, and the issuer's loop is one of the two serial chains a k-block
            // passes through.
            const int npairs = (t.nkb + 1) >> 1;
            for (int pp = 0; pp < npairs; ++pp, ++pit) {
                const int nh = min(2, t.nkb - 2 * pp);
                const uint32_t slot = pit % kTP;
#ifdef PPD_TCA_TRACE2
                if (blockIdx.x == 0 && lane == 0 && pit == 0) tr2[24 * 8 + 2] = (unsigned)clock64();
#endif
                TR2(it, 5);
                mbar_wait(&ta_full[slot], (pit / kTP) & 1u);
                uint32_t sh[2];
                for (int h = 0; h < nh; ++h) {
                    const int kb = 2 * pp + h;
                    const uint32_t s = a.b_resident ? (uint32_t)kb : rb.s;
                    sh[h] = s;
                    if (a.b_presplit && reload) mbar_wait(&full_b[s], a.b_resident ? (epoch & 1u) : rb.ph);
                    if (esb) mbar_wait(&bs_full[s], rb.ph);
                    rb.next(kSB);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                TR2(it, 6);
                if (elect_one()) {
                    TCA_TRACE1(it, 8);
                    for (int h = 0; h < nh; ++h) {
                        const int kb = 2 * pp + h;
                        const uint64_t db = bdesc0 + (uint64_t)((sh[h] * 2u * b_bytes) >> 4);
                        const uint32_t ta = tmem_base + kTaCol0 + slot * 128u + (uint32_t)h * 64u;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
#ifndef PPD_ABL_NOMMA
                            umma_tf32_ts(d, ta + k * 8u, db + (uint64_t)(k * kstep), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
#ifndef PPD_ABL_ONEMMA
                            umma_tf32_ts(d + (uint32_t)bn, ta + 32u + k * 8u, db + (uint64_t)(k * kstep), idesc1, 1u);
#endif
#endif
                        }
                        if (release) umma_commit(&empty_b[sh[h]]);
                    }
                    umma_commit(&ta_empty[slot]);
                    TCA_TRACE1(it, 9);
                }
                __syncwarp();
                TR2(it, 7);
                it += (uint32_t)nh;
            }
            if (elect_one()) umma_commit(&acc_full[acc]);
            __syncwarp();
            if (a.b_resident && reload) { loaded_cls = t.cls; ++epoch; }
        }
    } else if (warp == 3) {
        // (second A producer in convolution modes, handled above; idle for plain GEMMs)
    } else if (warp < 4 + kXformWarps) {
        // ================= transform: two groups of four warps take alternate k-blocks (two k-blocks in flight hide
        // the shared-memory / TMEM / barrier latencies of one another); thread = one tile row, all 32 contraction columns
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int grp = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int gt = (threadIdx.x - 128) & 127;     // thread index within the group
        const uint32_t smemA_u = smem_u32(smemA), smemB_u = smem_u32(smemB);
        uint32_t it = 0, tile_it = 0, pit = 0;
        int turn = 0;                                  // pit % kGroups
        Ring rb;
        // per-thread constants of the convolution modes (row r of the tile = pixel ox of tile row g): no division in the loops
        const int g_r = mode ? r / max(a.conv.segw, 1) : 0, ox_r = mode ? r - g_r * a.conv.segw : 0;
        const uint32_t row4 = (uint32_t)a.conv.Win * 4u;
        const uint32_t off4 = mode == 4 ? (uint32_t)(g_r * (32 / max(a.conv.KW, 1)) * a.conv.Win + ox_r * a.conv.s) * 4u : 0u;
        for (int w = blockIdx.x; w < a.total_items; w += gridDim.x, ++tile_it) {
#ifdef PPD_TCA_TRACE2
            if (blockIdx.x == 0 && lane == 0 && q == 0 && grp == 0 && tile_it == 7) tr2[24 * 8 + 4] = (unsigned)clock64();
#endif
            const Item t = decode<MODE>(a, w);
#ifdef PPD_TCA_TRACE2
            if (blockIdx.x == 0 && lane == 0 && q == 0 && grp == 0 && tile_it == 7) tr2[24 * 8 + 5] = (unsigned)clock64();
#endif
            const bool ok4 = g_r < t.nvalid;
            // mode 6 (tile-resident raw input): this thread's pixel is line `line0` of the stage for tap (0, 0); wait for the stage once
            uint32_t line0 = 0, stage6 = 0;
            const bool res6 = mode == 6, res7 = mode == 7, resident = res6 || res7;
            const int lw = res7 ? a.conv.segw + a.conv.T - 1 : a.conv.Win;              // lines per staged row
            if (resident) {
                const ConvA& cv = a.conv;
                const int g = g_r, ox = ox_r;
                if (res7) line0 = (uint32_t)((cv.T - 1) * lw + cv.T - 1);       // rows past the tile: any in-range line (the dgrad taps subtract)
                if (g < t.nvalid) {
                    const int b0 = fdiv(t.seg0, cv.rows_per_img);
                    int oy = t.seg0 - b0 * cv.rows_per_img, run_slot = 0, run_a = oy;
                    for (int gg = 0; gg < g; ++gg)
                        if (++oy == cv.rows_per_img) {
                            run_slot += res7 ? (cv.rows_per_img - run_a) + cv.T - 1 : cv.s * (cv.rows_per_img - 1 - run_a) + cv.KH;
                            run_a = 0; oy = 0;
                        }
                    line0 = res7 ? (uint32_t)((run_slot + (oy - run_a) + cv.T - 1) * lw + ox + cv.T - 1)
                                 : (uint32_t)((run_slot + cv.s * (oy - run_a)) * lw + ox * cv.s);
                }
                stage6 = smemA_u + (tile_it & 1u) * cv.tile_stage_bytes;
                mbar_wait(&full_a[tile_it & 1u], (tile_it >> 1) & 1u);
            }
            // k-blocks are taken in PAIRS: one tensor-memory slot (128 columns), one ta_empty wait, one wait::st / fence / arrive and one
            // trip around this loop per TWO k-blocks -- the fixed latencies of a hand-off, not the split arithmetic, are what a transform
            // group's time per k-block is made of
            const int npairs = (t.nkb + 1) >> 1;
            for (int pp = 0; pp < npairs; ++pp, ++pit, turn = (turn + 1 == kGroups) ? 0 : turn + 1) {
                const int nh = min(2, t.nkb - 2 * pp);
                if (turn != grp) {
                    it += (uint32_t)nh;
                    for (int h = 0; h < nh; ++h) rb.next(kSB);
                    continue;
                }
                const uint32_t slot = pit % kTP;
                for (int h = 0; h < nh; ++h, ++it, rb.next(kSB)) {
                const int kb = 2 * pp + h;
                const uint32_t s = it % kSA;
                if (q == 0) TCA_TRACE(it, 2);
                if (q == 0) TR2(it, 0);
                if (!resident) mbar_wait(&full_a[s], (it / kSA) & 1u);
                if (q == 0) TR2(it, 1);
                if (q == 0) TCA_TRACE(it, 3);
                const uint32_t sa = smemA_u + s * a_bytes;
                float x[32];
#ifdef PPD_ABL_NOLDS
                for (int c = 0; c < 32; ++c) x[c] = __uint_as_float(sa + c);
                if (sa == 0xffffffffu)
#endif
                if (resident) {
                    // forward: tap (ky, kx) is lw*ky + kx lines further; dgrad: tap (dky, dkx) reads dY row i - dky, pixel j - dkx
                    const uint32_t line = line0 + (uint32_t)tap_line[kb];
                    const uint32_t la = stage6 + line * 128u;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = lds128(la + (((uint32_t)c ^ (line & 7u)) << 4));
                        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                    }
                } else if (mode == 3 && raw) {
                    // raw rows [half][output row][ky][W]: patch element (ky, kx) of pixel slot p sits at row ky, float s*ox_p + kx
                    const ConvA& cv = a.conv;
                    const int OW = cv.spr;
                    const int P0 = (t.kb0 + kb) * 32;
                    const int npx = min(32, cv.total_seg - P0);
                    const bool chunk_ok = 2 * t.cls + (r >> 6) < cv.nchunks;
                    const uint32_t row_bytes = (uint32_t)(cv.KW * cv.Win) * 4u;
                    const uint32_t base = sa + (uint32_t)(r >> 6) * 3u * row_bytes + (uint32_t)(((r >> 3) & 7) * cv.Win + (r & 7)) * 4u;
                    const int ox0 = P0 - fdiv(P0, OW) * OW;
                    // Offset of pixel slot c inside the stage (the same for every thread): lane c computes it once, the
                    // unrolled loop below fetches it with one shuffle -- 3 instructions per element instead of ~10 of a
                    // running (ox, row) counter (20 % of the kernel's instructions, ncu).  A k-block is at most 3 output rows.
                    const int p = ox0 + lane, prow = (p >= OW ? 1 : 0) + (p >= 2 * OW ? 1 : 0);
                    const uint32_t myoff = (uint32_t)prow * row_bytes + (uint32_t)((p - prow * OW) * cv.s) * 4u;
                    if (!chunk_ok) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) x[c] = 0.f;
                    } else if (npx == 32) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) x[c] = lds32(base + __shfl_sync(0xffffffffu, myoff, c));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const uint32_t o = __shfl_sync(0xffffffffu, myoff, c);
                            x[c] = c < npx ? lds32(base + o) : 0.f;
                        }
                    }
                } else if (mode == 3) {
                    // [half][pixel slot][64 patch floats]; slots past the valid pixels and chunks past K are zero
                    const ConvA& cv = a.conv;
                    const int seg0 = (t.kb0 + kb) * cv.nseg;
                    const int npx = min(cv.nseg, cv.total_seg - seg0) * cv.segw;
                    const bool chunk_ok = 2 * t.cls + (r >> 6) < cv.nchunks;
                    const uint32_t base = sa + (uint32_t)(r >> 6) * 8192u + (uint32_t)(r & 63) * 4u;
                    if (chunk_ok && npx >= 32) {             // the common case: no predicates
#pragma unroll
                        for (int c = 0; c < 32; ++c) x[c] = lds32(base + c * 256);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) x[c] = (chunk_ok && c < npx) ? lds32(base + c * 256) : 0.f;
                    }
                } else if (mode == 4) {
                    // stage = [output row g][kyg = 4 image rows][W]; this thread's pixel starts at float off4 of the stage
                    const uint32_t base = sa + off4;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // float4 index c: filter row c / 2 (kw = 8: two float4 per row)
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok4) v = lds128(base + (uint32_t)(c >> 1) * row4 + (uint32_t)(c & 1) * 16u);
                        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                    }
                } else if (!a.a_mn) {
                    // 128B-swizzled rows of 32 floats: 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = lds128(sa + r * 128 + ((c ^ (r & 7)) << 4));
                        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) x[c] = lds32(sa + (c * BM + r) * 4);                // [k][128 rows]
                }
                uint32_t hi[32], lo[32];
#pragma unroll
#ifdef PPD_ABL_NOSPLIT
                for (int c = 0; c < 32; ++c) { hi[c] = __float_as_uint(x[c]); lo[c] = 0u; }
#else
                for (int c = 0; c < 32; ++c) split_a(x[c], hi[c], lo[c]);
#endif
                // The stage was read through the generic proxy (ld.shared) and will be re-written through the async proxy (TMA / bulk
                // copy): that write-after-read needs a proxy fence before the release.  Without it the 1-D bulk copies of the NCHW
                // forward overtook in-flight reads once a CTA re-used a stage (second tile onwards): ~1e-3 of the tiles of a
                // 2048-sample conv1 came out wrong, non-deterministically (found by tests/test_gpu_production_size.py).
                if (!resident) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // (resident stages: once per tile, below)
                __syncwarp();
                if (lane == 0 && !resident) mbar_arrive(&empty_a[s]);          // the tile is in registers: slot back to the producer
                if (q == 0) TCA_TRACE(it, 4);
                if (q == 0) TR2(it, 2);
                if (h == 0) {
                    mbar_wait(&ta_empty[slot], ((pit / kTP) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                if (q == 0) TR2(it, 3);
                if (q == 0) TCA_TRACE(it, 5);
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kTaCol0 + slot * 128u + (uint32_t)h * 64u;
#if defined(PPD_ABL_NOSTTM)
                if (hi[0] == 0x12345678u) tmem_st16(ta, hi);
#elif defined(PPD_ST32)
                tmem_st32(ta, hi);
                tmem_st32(ta + 32u, lo);
#else
                tmem_st16(ta, hi);
                tmem_st16(ta + 16u, hi + 16);
                tmem_st16(ta + 32u, lo);
                tmem_st16(ta + 48u, lo + 16);
#endif
                if (q == 0) TCA_TRACE(it, 6);
                if (!a.b_presplit && !esb) {
                    const uint32_t sbs = rb.s;
                    mbar_wait(&full_b[sbs], rb.ph);
                    const uint32_t src = smemB_u + sbs * 2 * b_bytes;
                    const int nvec = (int)(b_bytes >> 4);
#ifdef PPD_ABL_NOBSPLIT
                    if (nvec < 0)
#endif
                    for (int v = gt; v < nvec; v += 128) {
                        const float4 xb = lds128(src + (v << 4));
                        uint4 hb, rb;
                        split_a(xb.x, hb.x, rb.x);
                        split_a(xb.y, hb.y, rb.y);
                        split_a(xb.z, hb.z, rb.z);
                        split_a(xb.w, hb.w, rb.w);
                        sts128(src + (v << 4), hb);
                        sts128(src + b_bytes + (v << 4), rb);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                }       // k-blocks of the pair
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&ta_full[slot]);
                if (q == 0) TR2(it - 1, 4);
                if (q == 0) TCA_TRACE(it - 1, 7);
            }
            if (resident) {                            // every read of this warp from the tile's stage is done
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic reads before the async-proxy refill
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_a[tile_it & 1u]);
            }
        }
    } else {
        // ================= epilogue: warp w may touch TMEM lanes [32*(w%4), +32); result = main + correction accumulator
        const int q = warp & 3;
        uint32_t tile_it = 0;
        Ring rbs;                                       // B ring position (weight gradients: these warps split the B stages)
        const uint32_t smemB_e = smem_u32(smemB);
        for (int w = blockIdx.x; w < a.total_items; w += gridDim.x, ++tile_it) {
            const Item t = decode<MODE>(a, w);
            const uint32_t acc = tile_it & 1u;
            if (esb) {
                const int nvec = (int)(b_bytes >> 4);
                for (int kb = 0; kb < t.nkb; ++kb, rbs.next(kSB)) {
                    mbar_wait(&full_b[rbs.s], rbs.ph);
                    const uint32_t src = smemB_e + rbs.s * 2 * b_bytes;
                    for (int v = q * 32 + lane; v < nvec; v += 32 * kEpiWarps) {
                        const float4 xb = lds128(src + (v << 4));
                        uint4 hb, lb;
                        split_a(xb.x, hb.x, lb.x);
                        split_a(xb.y, hb.y, lb.y);
                        split_a(xb.z, hb.z, lb.z);
                        split_a(xb.w, hb.w, lb.w);
                        sts128(src + (v << 4), hb);
                        sts128(src + b_bytes + (v << 4), lb);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bs_full[rbs.s]);
                }
            }
            mbar_wait(&acc_full[acc], (tile_it >> 1) & 1u);
#ifdef PPD_TCA_TRACE2
            if (blockIdx.x == 0 && lane == 0 && q == 0 && tile_it == 0) tr2[24 * 8 + 3] = (unsigned)clock64();
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (a.tma_store) {
                // the staging tile is free again once the previous tile's bulk store has READ it (issued by warp 12's lane 0)
                if (q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const int64_t i = t.i0 + q * 32 + lane;
            bool row_ok = i < a.I;
            int64_t crow_off = i * a.ldc + t.j0, mrow_off = i * a.ldm + t.j0;
            if (mode == 1 || mode == 2 || mode == 4 || mode == 6 || mode == 7) {
                const ConvA& cv = a.conv;
                const int r = q * 32 + lane;
                row_ok = r < t.nvalid * cv.segw;
                if (mode != 2 && mode != 7) {
                    crow_off = ((int64_t)t.seg0 * cv.segw + r) * a.ldc;            // output pixels of whole rows are contiguous
                } else {
                    const int g = r / cv.segw, j = r - g * cv.segw;
                    const int sg = t.seg0 + g;
                    const int b = sg / cv.rows_per_img, ii = sg - b * cv.rows_per_img;
                    const int py = t.cls / cv.s, px = t.cls - py * cv.s;
                    crow_off = (((int64_t)b * cv.Hin + (cv.s * ii + py)) * cv.Win + (cv.s * j + px)) * cv.Cin;
                }
                mrow_off = crow_off;
            }
            for (int c0 = 0; c0 < bn; c0 += 32) {
                float v[32];
                {
                    float u[32];
                    const uint32_t tad = tmem_base + ((uint32_t)(q * 32) << 16) + kAccCol0 + acc * kAccStride + (uint32_t)c0;
#ifdef PPD_ABL_NOLDTM
                    for (int c = 0; c < 32; ++c) { u[c] = 0.f; v[c] = __uint_as_float(tad + c); }     // timing experiment: no TMEM reads
#else
                    tmem_ld32(tad + (uint32_t)bn, u);
                    tmem_ld32(tad, v);
#endif
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] += u[c];
                }
                if (c0 + 32 >= bn) {
                    // last read of this accumulator: hand it back before the global stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                const int64_t jb = t.j0 + c0;
#ifdef PPD_ABL_NOEPI
                if (v[0] != 123.456f) continue;          // timing experiment: no epilogue math / stores
#endif
                if (a.partial) {
                    if (i < a.I) {
                        float* P = a.partial + ((int64_t)t.z * a.I + i) * a.J;
                        if (jb + 31 < a.J && (a.J & 3) == 0) {
#pragma unroll
                            for (int c = 0; c < 32; c += 4)
                                *reinterpret_cast<float4*>(P + jb + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (jb + c < a.J) P[jb + c] = v[c];
                        }
                    }
                    continue;
                }
                if (a.bias) {
                    if (jb + 32 <= a.J && ((reinterpret_cast<uintptr_t>(a.bias + jb) & 15) == 0)) {
                        const float4* b4 = reinterpret_cast<const float4*>(a.bias + jb);     // the same 128 bytes for every lane: broadcast loads
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 bb = __ldg(b4 + c);
                            v[4 * c] += bb.x; v[4 * c + 1] += bb.y; v[4 * c + 2] += bb.z; v[4 * c + 3] += bb.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] += (jb + c < a.J) ? __ldg(a.bias + jb + c) : 0.f;
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
                }
                if (a.tma_store) {
                    if (a.mask && row_ok) {                 // dgrad: ReLU mask of the activation this gradient flows into (same rows as dx)
                        const float* mrow = a.mask + mrow_off + c0;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            float m[8];
                            asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=f"(m[0]), "=f"(m[1]), "=f"(m[2]), "=f"(m[3]), "=f"(m[4]), "=f"(m[5]), "=f"(m[6]), "=f"(m[7])
                                         : "l"(mrow + c));
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[c + e] = m[e] > 0.f ? v[c + e] : 0.f;
                        }
                    }
                    // Row r of the tile -> 128-byte row r of the staging tile of this 32-column block, 16-byte chunk c at position
                    // c ^ (r & 7) (the 128B swizzle of the output tensor map: conflict-free for the 8 lanes of a store phase).  The
                    // global write is ONE bulk tensor store per block: no LSU line-per-lane traffic at all.
                    const int r = q * 32 + lane;
                    const uint32_t srow = smem_u32(smem) + a.stage_off + (uint32_t)(c0 >> 5) * (BM * 128u) + (uint32_t)r * 128u;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        sts128(srow + (((uint32_t)c ^ ((uint32_t)r & 7u)) << 4),
                               make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3])));
                    if (c0 + 32 >= bn) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic writes -> async-proxy read
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (q == 0 && lane == 0) {
                            for (int cb = 0; cb < bn; cb += 32) {
                                const uint32_t src = smem_u32(smem) + a.stage_off + (uint32_t)(cb >> 5) * (BM * 128u);
                                if (mode == 0)
                                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC), "r"(src),
                                                 "r"((int)t.j0 + cb), "r"((int)t.i0)
                                                 : "memory");
                                else if (mode == 2 || mode == 7) {
                                    // dx seen as {C, j (stride s pixels), segment (stride s rows), px, py}: the tile's pixels of one parity class
                                    const int py = t.cls / a.conv.s, px = t.cls - py * a.conv.s;
                                    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tmC),
                                                 "r"(src), "r"(cb), "r"(0), "r"(t.seg0), "r"(px), "r"(py)
                                                 : "memory");
                                } else
                                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmC), "r"(src),
                                                 "r"(cb), "r"(0), "r"(t.seg0)
                                                 : "memory");
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                } else if (a.transpose_out) {
                    // C is [J, I] row-major: for a fixed column the 32 lanes write 32 consecutive floats
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (i < a.I && jb + c < a.J) {
                            float* p = a.C + (jb + c) * a.ldc + i;
                            float x = v[c];
                            if (a.mask) x = (__ldg(a.mask + (jb + c) * a.ldm + i) > 0.f) ? x : 0.f;
                            *p = a.accumulate ? (*p + x) : x;
                        }
                    }
                } else if (row_ok) {
                    float* crow = a.C + crow_off + c0;
                    const float* mrow = a.mask ? a.mask + mrow_off + c0 : nullptr;
                    const bool vec = (jb + 31 < a.J) && ((a.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) &&
                                     (!mrow || (((a.ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)));
                    const bool vec8 = vec && !a.accumulate && ((reinterpret_cast<uintptr_t>(crow) & 31) == 0) &&
                                      (!mrow || ((reinterpret_cast<uintptr_t>(mrow) & 31) == 0));
                    if (vec8) {
                        // A thread owns a row, so every store instruction of a warp touches 32 different cache lines: the LSU handles
                        // one line per clock, and those clocks are taken from the transform warps' shared-memory reads and tcgen05.st
                        // (clock64 stamps: a ~1000-clock bubble in the transform AND the issuer once per tile, gone with the stores
                        // removed).  256-bit accesses (sm_100) halve the instruction count: 4 x 32 lines per row block instead of 8 x 32.
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            if (mrow) {
                                float m[8];
                                asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                             : "=f"(m[0]), "=f"(m[1]), "=f"(m[2]), "=f"(m[3]), "=f"(m[4]), "=f"(m[5]), "=f"(m[6]), "=f"(m[7])
                                             : "l"(mrow + c));
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[c + e] = m[e] > 0.f ? v[c + e] : 0.f;
                            }
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(crow + c), "f"(v[c]), "f"(v[c + 1]),
                                         "f"(v[c + 2]), "f"(v[c + 3]), "f"(v[c + 4]), "f"(v[c + 5]), "f"(v[c + 6]), "f"(v[c + 7])
                                         : "memory");
                        }
                    } else if (vec) {
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
    }
    // the last bulk store must have completed before the CTA's shared memory goes away (issued by the first epilogue warp's lane 0)
    if (a.tma_store && warp == 4 + kXformWarps && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
#ifdef PPD_TCA_TRACE2
    if (blockIdx.x == 0) if (threadIdx.x == 0) tr2[24 * 8 + 1] = (unsigned)clock64();
    __syncthreads();
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 24 * 8 + 8; i += blockDim.x) g_trace2[i] = tr2[i];
#endif
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// PPD_TMA_STORE=0: row-per-thread global stores in every epilogue (A/B timing)
static int tma_store_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PPD_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
// PPD_PDL=0 launches without the programmatic-serialisation attribute (A/B timing)
static int pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PPD_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
// One instantiation per operand layout (ConvA::mode; 5 = weight gradient over raw rows).
template <int MODE>
cudaError_t launch_mode(int grid, size_t smem, cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo,
                        const CUtensorMap& tmC, const Args& a) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tca_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // Programmatic dependent launch: this kernel's CTAs may start (barrier init, TMEM allocation, tensor-map prefetch, work decode)
    // while the previous kernel of the stream drains; griddepcontrol.wait in the kernel orders every global access behind it.
    // 126.9 -> 124.8 ms per PPO update (1088 launches of this kernel), losses bit-identical.  The same treatment of the ~11 000 small
    // launches of an update (reductions, transposes, SIMT GEMMs, loss, gathers) was measured too and made it 0.6 ms SLOWER: their
    // early-scheduled CTAs only occupy SMs that the side streams' kernels could have used.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // ... but not while SMs are being left to an in-flight collective (g_max_ctas below the SM count): an early-scheduled dependent's
    // CTAs would sit on exactly those SMs, spinning in griddepcontrol.wait, and the NCCL kernel would wait for them
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_enabled() && g_max_ctas >= kNumSMs) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tca_gemm_kernel<MODE>, tmA, tmB, tmBlo, tmC, a);
}
cudaError_t launch_kernel(int grid, size_t smem, cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo,
                          const Args& a, const CUtensorMap* tmC_opt = nullptr) {
    const CUtensorMap& tmC = tmC_opt ? *tmC_opt : tmA;          // (unused unless a.tma_store)
    switch (a.conv.mode == 3 && a.conv.raw ? 5 : a.conv.mode) {
        case 0: return launch_mode<0>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 1: return launch_mode<1>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 2: return launch_mode<2>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 3: return launch_mode<3>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 4: return launch_mode<4>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 5: return launch_mode<5>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 6: return launch_mode<6>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
        case 7: return launch_mode<7>(grid, smem, s, tmA, tmB, tmBlo, tmC, a);
    }
    return cudaErrorInvalidValue;
}

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

}  // namespace

#ifdef PPD_TCA_TRACE2
extern "C" int ppd_tca_trace2_read(unsigned* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_trace2, sizeof(unsigned) * (24 * 8 + 8));
}
#endif
#ifdef PPD_TCA_TRACE
extern "C" int ppd_tca_set_trace(void* dev_ptr) {
    return (int)cudaMemcpyToSymbol(g_trace, &dev_ptr, sizeof(void*));
}
#endif

namespace {
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, uint4* __restrict__ hi, uint4* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = x[i];
        uint4 h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z);
        split_tf32(v.w, h.w, l.w);
        hi[i] = h;
        lo[i] = l;
    }
}
}  // namespace

int split_operand(const float* x, float* hi, float* lo, int64_t n, cudaStream_t s) {
    const int64_t n4 = n / 4;
    int64_t nb = (n4 + 255) / 256;
    if (nb > 8 * kNumSMs) nb = 8 * kNumSMs;
    if (nb < 1) nb = 1;
    split_tf32_kernel<<<(unsigned)nb, 256, 0, s>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint4*>(hi),
                                                   reinterpret_cast<uint4*>(lo), n4);
    return launch_status("split_tf32_kernel");
}

int make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("tca_gemm: cuTensorMapEncodeTiled not available"); return PPD_EINVAL; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tca_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return PPD_EINVAL; }
    return 0;
}

int g_max_ctas = kNumSMs;     // CTAs of the persistent kernel (one per SM); lowered while a collective needs SMs of its own
int g_b_resident = 1;         // convolutions keep the weight tiles of all k-blocks in shared memory when they fit
int g_conv_resident = 1;      // forward NHWC convolutions stage their input once per tile (ConvA mode 6); 0 = one im2col box per row and k-block (mode 1)
// N-D fp32 tensor map (dims / strides innermost first; strides in bytes for dims 1..nd-1), OOB elements read as zero.
static int make_map_nd(CUtensorMap* m, const float* base, int nd, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("tca_gemm: cuTensorMapEncodeTiled not available"); return PPD_EINVAL; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)nd, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tca_gemm: cuTensorMapEncodeTiled (%d-d) failed (%d)", nd, (int)r); return PPD_EINVAL; }
    return 0;
}

static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo, Args& a, cudaStream_t s) {
    a.a_mn = 0; a.b_presplit = 1; a.transpose_out = 0; a.accumulate = 0; a.partial = nullptr;
    a.num_m = 1; a.num_n = 1; a.splits = 1; a.kk_per_split = a.KK; a.ldm = a.ldc;
    const size_t a_region = a.a_region_bytes ? a.a_region_bytes : (size_t)kSA * BM * BK * 4;
    int sb_stages = (int)((kSmemBudget - 1024 - a_region) / ((size_t)2 * a.bn * BK * 4));
    a.b_resident = (g_b_resident && a.conv.nkb >= 2 && a.conv.nkb <= kMaxSB && a.conv.nkb <= sb_stages) ? 1 : 0;      // all weight tiles fit: keep them
    if (a.b_resident) sb_stages = a.conv.nkb;
    if (sb_stages > 8 && !a.b_resident) sb_stages = 8;
    PPD_REQUIRE(sb_stages >= 2, "shared memory: no room for the B ring");
    // Forward convolutions write their tiles with TMA: the epilogue stages the finished tile (128 rows x 32 columns per block) in shared
    // memory and ONE bulk tensor store per block moves it -- the output tensor seen as {Cout, OW pixels, output rows}, box = the
    // tile's whole output rows, so the rows past the tile's last valid one are never touched and the tensor bound clips the last tile.
    CUtensorMap tmC;
    const size_t stage_bytes = (size_t)(a.bn / 32) * BM * 128;
    a.tma_store = 0;
    const bool dgrad_mode = a.conv.mode == 2 || a.conv.mode == 7;
    if (tma_store_enabled() && (a.conv.mode == 1 || a.conv.mode == 4 || a.conv.mode == 6 || dgrad_mode) && a.ldc == a.J && a.J % 32 == 0) {
        const size_t b_stage = (size_t)2 * a.bn * BK * 4;
        if (!a.b_resident)
            while (sb_stages > 4 && a_region + sb_stages * b_stage + stage_bytes + 1024 > kSmemBudget) --sb_stages;
        if (a_region + sb_stages * b_stage + stage_bytes + 1024 <= kSmemBudget) {
            const ConvA& cv = a.conv;
            int rc;
            if (dgrad_mode) {
                // dx [B, Hin, Win, C] with Hin = s Hq, Win = s Wq: pixel (segment sg = b Hq + i, j) of parity class (py, px) is row
                // s sg + py, pixel s j + px -- {C, j, sg, px, py} with strides {1, s C, s Win C, C, Win C}
                const cuuint64_t C = (cuuint64_t)a.J, st = (cuuint64_t)cv.s;
                cuuint64_t dims[5] = {C, (cuuint64_t)cv.segw, (cuuint64_t)cv.nseg_class, st, st};
                cuuint64_t str[4] = {st * C * 4, st * (cuuint64_t)cv.Win * C * 4, C * 4, (cuuint64_t)cv.Win * C * 4};
                cuuint32_t box[5] = {32, (cuuint32_t)cv.segw, (cuuint32_t)cv.nseg, 1, 1};
                rc = make_map_nd(&tmC, a.C, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
            } else {
                cuuint64_t dims[3] = {(cuuint64_t)a.J, (cuuint64_t)cv.segw, (cuuint64_t)(a.I / cv.segw)};
                cuuint64_t str[2] = {(cuuint64_t)a.J * 4, (cuuint64_t)cv.segw * a.J * 4};
                cuuint32_t box[3] = {32, (cuuint32_t)cv.segw, (cuuint32_t)cv.nseg};
                rc = make_map_nd(&tmC, a.C, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
            }
            if (rc) return rc;
            a.tma_store = 1;
            a.stage_off = (uint32_t)(a_region + sb_stages * b_stage);
        }
    }
    a.sb_stages = sb_stages;
    const size_t smem = a_region + (size_t)sb_stages * 2 * a.bn * BK * 4 + (a.tma_store ? stage_bytes : 0) + 1024;
    const int grid = a.total_items < g_max_ctas ? a.total_items : g_max_ctas;
    cudaError_t e = launch_kernel(grid, smem, s, tmA, tmB, tmBlo, a, a.tma_store ? &tmC : nullptr);
    if (e != cudaSuccess) { set_error("tca conv: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return launch_status("tca_gemm_kernel(conv)");
}

int conv_forward(const float* x, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* bias, int relu,
                 float* out, cudaStream_t s) {
    const int OH = (g->H - g->kh) / g->stride + 1, OW = (g->W - g->kw) / g->stride + 1;
    const int rowf = g->kw * g->C;                                   // contiguous floats of one filter row in NHWC
    PPD_REQUIRE(rowf % 32 == 0 && (Cout == 32 || Cout == 64) && OW >= 1 && OW <= BM && OH >= 1, "unsupported convolution shape");
    PPD_REQUIRE(!(((uintptr_t)x | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out) & 15), "pointers must be 16-byte aligned");
    const int64_t K = (int64_t)g->kh * rowf;
    CUtensorMap tmA, tmB, tmBlo;
    cuuint64_t dims[4] = {(cuuint64_t)rowf, (cuuint64_t)OW, (cuuint64_t)g->H, (cuuint64_t)g->B};
    cuuint64_t str[3] = {(cuuint64_t)g->stride * g->C * 4, (cuuint64_t)g->W * g->C * 4, (cuuint64_t)g->H * g->W * g->C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)OW, 1, 1};
    int rc = make_map_nd(&tmA, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if ((rc = make_map_2d(&tmB, w_hi, Cout, K, K, BK, Cout, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map_2d(&tmBlo, w_lo, Cout, K, K, BK, Cout, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    Args a = {};
    a.C = out; a.ldc = Cout; a.I = (int64_t)g->B * OH * OW; a.J = Cout; a.KK = K;
    a.bias = bias; a.mask = nullptr; a.relu = relu;
    a.bn = Cout; a.b_mn = 0;
    ConvA& cv = a.conv;
    cv.mode = 1; cv.segw = OW; cv.nseg = BM / OW; cv.nseg_class = g->B * OH;
    cv.ntile_class = (cv.nseg_class + cv.nseg - 1) / cv.nseg;
    cv.rows_per_img = OH; cv.s = g->stride; cv.kpk = rowf / 32; cv.T = 0; cv.KW = g->kw; cv.Cin = g->C; cv.Hin = g->H; cv.Win = g->W;
    cv.nkb = g->kh * cv.kpk;
    a.total_items = cv.ntile_class;
    PPD_REQUIRE(cv.nseg <= 32, "output width not supported (too many TMA boxes per tile)");
    // Tile-resident raw input (mode 6) when two stages of it and a B ring fit: every input element goes through TMA once per tile
    if (g_conv_resident && g->C % 32 == 0) {
        const int planes = g->C / 32;
        // two stages of the rows of a tile + a B ring must fit (a smaller tile is not worth it: more tiles, each with all its k-blocks)
        int nseg = cv.nseg, max_runs = 0, nrows_max = 0;
        size_t stage = 0;
        for (; nseg >= 1 && nseg >= cv.nseg; --nseg) {
            max_runs = 1 + (nseg - 1 + OH - 1) / OH;
            nrows_max = g->stride * nseg + (g->kh > g->stride ? (g->kh - g->stride) * max_runs : 0) + 1;
            stage = (((size_t)planes * nrows_max * g->W * 128) + 1023) & ~(size_t)1023;
            // ... with room for a B ring of at least four stages (or all k-blocks): with the two left over for the 4x4x32 -> 64
            // layer the weight loads sit on the critical path and the per-k-block boxes of mode 1 are faster (76 vs 82 us, same-box A/B)
            const size_t b_stages_wanted = cv.nkb < 4 ? cv.nkb : 4;
            if (2 * stage + b_stages_wanted * (size_t)2 * Cout * BK * 4 + 1024 <= kSmemBudget) break;
        }
        if (cv.nkb <= kMaxTaps && nseg >= 1 && nseg >= cv.nseg) {
            cv.nseg = nseg;
            cv.ntile_class = (cv.nseg_class + cv.nseg - 1) / cv.nseg;
            a.total_items = cv.ntile_class;
            cuuint64_t d4[4] = {(cuuint64_t)g->C, (cuuint64_t)g->W, (cuuint64_t)g->H, (cuuint64_t)g->B};
            cuuint64_t s4[3] = {(cuuint64_t)g->C * 4, (cuuint64_t)g->W * g->C * 4, (cuuint64_t)g->H * g->W * g->C * 4};
            cuuint32_t b4[4] = {32, (cuuint32_t)g->W, 1, 1};
            if ((rc = make_map_nd(&tmA, x, 4, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            cv.mode = 6; cv.KH = g->kh; cv.planes = planes; cv.nrows_max = nrows_max; cv.tile_stage_bytes = (uint32_t)stage;
            a.a_region_bytes = (uint32_t)(2 * stage);
        }
    }
    return launch_conv(tmA, tmB, tmBlo, a, s);
}

int conv_forward_nchw(const float* x, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* bias,
                      int relu, float* out, cudaStream_t s) {
    const int OH = (g->H - g->kh) / g->stride + 1, OW = (g->W - g->kw) / g->stride + 1;
    PPD_REQUIRE(g->kw == 8 && g->kh % 4 == 0 && g->stride == 4 && g->W % 4 == 0, "NCHW forward supports 8-wide filters with stride 4");
    PPD_REQUIRE((Cout == 32 || Cout == 64) && OW >= 1 && OW <= BM && OH >= 1, "unsupported convolution shape");
    PPD_REQUIRE(!(((uintptr_t)x | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out) & 15), "pointers must be 16-byte aligned");
    const int kyg = 32 / g->kw;
    const int64_t K = (int64_t)g->C * g->kh * g->kw;
    Args a = {};
    ConvA& cv = a.conv;
    cv.mode = 4; cv.segw = OW; cv.nseg = BM / OW;
    while ((size_t)cv.nseg * kyg * g->W * 4 > (size_t)BM * BK * 4) --cv.nseg;       // the raw rows of a stage must fit its 16 KB       // the raw rows of a stage must fit its 16 KB
    cv.nseg_class = g->B * OH;
    cv.ntile_class = (cv.nseg_class + cv.nseg - 1) / cv.nseg;
    cv.rows_per_img = OH; cv.s = g->stride; cv.kpk = g->kh / kyg; cv.KW = g->kw; cv.Cin = g->C; cv.C = g->C; cv.Hin = g->H; cv.Win = g->W;
    cv.nkb = g->C * cv.kpk;
    CUtensorMap tmA, tmB, tmBlo;
    int rc = make_map_2d(&tmA, x, (int64_t)g->B * g->C * g->H, g->W, g->W, g->W, kyg, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    if ((rc = make_map_2d(&tmB, w_hi, Cout, K, K, BK, Cout, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map_2d(&tmBlo, w_lo, Cout, K, K, BK, Cout, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    a.C = out; a.ldc = Cout; a.I = (int64_t)g->B * OH * OW; a.J = Cout; a.KK = K;
    a.bias = bias; a.mask = nullptr; a.relu = relu;
    a.bn = Cout; a.b_mn = 0;
    a.a_ptr = x;
    a.total_items = cv.ntile_class;
    PPD_REQUIRE(cv.nseg <= 32, "output width not supported (too many TMA boxes per tile)");
    return launch_conv(tmA, tmB, tmBlo, a, s);
}

int conv_dgrad(const float* dy, const ppd_conv_geom* g, int Cout, const float* w_hi, const float* w_lo, const float* act_mask,
               float* dx, cudaStream_t s) {
    const int st = g->stride;
    const int OH = (g->H - g->kh) / st + 1, OW = (g->W - g->kw) / st + 1;
    PPD_REQUIRE(g->kh == g->kw && g->kh % st == 0 && g->H % st == 0 && g->W % st == 0, "filter and input sizes must be multiples of the stride");
    PPD_REQUIRE(Cout % 32 == 0 && (g->C == 32 || g->C == 64) && g->W / st <= BM, "unsupported convolution shape");
    PPD_REQUIRE(!(((uintptr_t)dy | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)dx | (uintptr_t)act_mask) & 15), "pointers must be 16-byte aligned");
    const int Hq = g->H / st, Wq = g->W / st;
    const int64_t K = (int64_t)g->kh * g->kw * g->C;                 // columns of the weight matrix [Cout, (ky, kx, c)]
    CUtensorMap tmA, tmB, tmBlo;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)g->B};
    cuuint64_t str[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)OW * Cout * 4, (cuuint64_t)OH * OW * Cout * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)Wq, 1, 1};
    int rc = make_map_nd(&tmA, dy, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if ((rc = make_map_2d(&tmB, w_hi, Cout, K, K, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map_2d(&tmBlo, w_lo, Cout, K, K, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    Args a = {};
    a.C = dx; a.ldc = g->C; a.I = (int64_t)g->B * g->H * g->W; a.J = g->C; a.KK = (int64_t)(g->kh / st) * (g->kw / st) * Cout;
    a.bias = nullptr; a.mask = act_mask; a.relu = 0;
    a.bn = g->C; a.b_mn = 1;
    ConvA& cv = a.conv;
    cv.mode = 2; cv.segw = Wq; cv.nseg = BM / Wq; cv.nseg_class = g->B * Hq;
    cv.ntile_class = (cv.nseg_class + cv.nseg - 1) / cv.nseg;
    cv.rows_per_img = Hq; cv.s = st; cv.kpk = Cout / 32; cv.T = g->kh / st; cv.KW = g->kw; cv.Cin = g->C; cv.Hin = g->H; cv.Win = g->W;
    cv.nkb = cv.T * cv.T * cv.kpk;
    a.total_items = st * st * cv.ntile_class;
    PPD_REQUIRE(cv.nseg <= 32, "input width not supported (too many TMA boxes per tile)");
    // Tile-resident dY (mode 7): every dY element goes through TMA once per tile instead of once per tap that reads it
    if (g_conv_resident) {
        const int planes = Cout / 32, lw = Wq + cv.T - 1;
        const int max_runs = 1 + (cv.nseg - 1 + Hq - 1) / Hq;
        const int nrows_max = cv.nseg + (cv.T - 1) * max_runs + 1;
        const size_t stage = (((size_t)planes * nrows_max * lw * 128) + 1023) & ~(size_t)1023;
        if (lw <= 256 && cv.nkb <= kMaxTaps && 2 * stage + 2 * (size_t)2 * g->C * BK * 4 + 1024 <= kSmemBudget) {
            cuuint32_t bx[4] = {32, (cuuint32_t)lw, 1, 1};
            if ((rc = make_map_nd(&tmA, dy, 4, dims, str, bx, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            cv.mode = 7; cv.planes = planes; cv.nrows_max = nrows_max; cv.tile_stage_bytes = (uint32_t)stage;
            a.a_region_bytes = (uint32_t)(2 * stage);
        }
    }
    return launch_conv(tmA, tmB, tmBlo, a, s);
}

static void wgrad_plan(const ppd_conv_geom* g, int Cout, int nchw, ConvA& cv, int& num_m, int& splits) {
    const int OH = (g->H - g->kh) / g->stride + 1, OW = (g->W - g->kw) / g->stride + 1;
    // divisor of OW that fills most of the 32 pixel slots with at most 8 segments (16 TMA boxes) per k-block; ties: the larger
    int segw = 0, best = -1;
    for (int d = 1; d <= OW && d <= 32; ++d) {
        if (OW % d || 32 / d > 8) continue;
        const int cover = 32 / d * d;
        if (cover >= best) { best = cover; segw = d; }
    }
    if (!segw) segw = OW <= 32 ? OW : 1;         // (conv_wgrad rejects shapes that would need more than 32 boxes)
    cv = ConvA{};
    cv.mode = 3; cv.segw = segw; cv.nseg = 32 / segw; cv.spr = OW / segw; cv.rows_per_img = OH; cv.s = g->stride;
    const int K = g->kh * g->kw * g->C;
    cv.nchunks = K / 64; cv.nchw = nchw; cv.C = g->C; cv.cpr = nchw ? 1 : g->kw * g->C / 64;
    // NCHW 8x8 filters: raw-row staging (1-D bulk copies, windows expanded on the register read) when 32 pixels span <= 3 output rows
    cv.raw = (nchw && g->kh == 8 && g->kw == 8 && OW >= 16 && (size_t)6 * g->kw * g->W * 4 <= (size_t)BM * BK * 4 && g->stride % 4 == 0) ? 1 : 0;
    cv.KW = g->kw; cv.Win = g->W; cv.Hin = g->H;
    if (cv.raw) { cv.segw = 1; cv.nseg = 32; cv.spr = OW; }          // "segments" are single pixels: k-block = 32 consecutive pixels
    cv.total_seg = g->B * OH * cv.spr;
    cv.total_kb = (cv.total_seg + cv.nseg - 1) / cv.nseg;
    num_m = (cv.nchunks + 1) / 2;
    splits = g_max_ctas / num_m;            // one wave of the CTAs the kernel may use
    if (splits > cv.total_kb / 8) splits = cv.total_kb / 8;
    if (splits < 1) splits = 1;
    // tcgen05 accumulates with truncation, so the error of an accumulator grows with the length of its chain: a 2048-sample
    // conv1 (819 200 pixels = 25 600 k-blocks over 74 splits = 346 k-blocks per chain) was 2.7e-5 of the gradient's scale off on
    // random-sign data (gate 2e-5; 4e-5 at C = 12).  Chains are cut at kMaxChainKb k-blocks: whole multiples of the one-wave split count,
    // so that the SMs stay evenly loaded; the extra partial tiles are a few MB.
    {
        const int chain = (cv.total_kb + splits - 1) / splits;
        const int k = (chain + kMaxChainKb - 1) / kMaxChainKb;
        if (k > 1) splits *= k;
    }
    cv.kbps = (cv.total_kb + splits - 1) / splits;
    splits = (cv.total_kb + cv.kbps - 1) / cv.kbps;
    (void)Cout;
}

size_t conv_wgrad_workspace(const ppd_conv_geom* g, int Cout) {
    ConvA cv; int num_m, s0, s1;
    wgrad_plan(g, Cout, 0, cv, num_m, s0);
    wgrad_plan(g, Cout, 1, cv, num_m, s1);
    return (size_t)(s0 > s1 ? s0 : s1) * g->kh * g->kw * g->C * Cout * sizeof(float);
}

int conv_wgrad(const float* x, const ppd_conv_geom* g, int nchw, const float* dy, int Cout, float* dW, int accumulate,
               void* workspace, size_t workspace_bytes, cudaStream_t s, int* splits_out) {
    const int OH = (g->H - g->kh) / g->stride + 1, OW = (g->W - g->kw) / g->stride + 1;
    const int K = g->kh * g->kw * g->C;
    PPD_REQUIRE(K % 64 == 0 && (Cout == 32 || Cout == 64) && OH >= 1 && OW >= 1, "unsupported convolution shape");
    PPD_REQUIRE(nchw ? (g->kh * g->kw == 64 && g->kw * 4 % 16 == 0 && g->stride * 4 % 16 == 0 && g->W * 4 % 16 == 0)
                     : (g->kw * g->C % 64 == 0), "unsupported patch layout");
    PPD_REQUIRE(!(((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dW | (uintptr_t)workspace) & 15), "pointers must be 16-byte aligned");
    Args a = {};
    int num_m, splits;
    wgrad_plan(g, Cout, nchw, a.conv, num_m, splits);
    PPD_REQUIRE(a.conv.raw || a.conv.nseg * 2 <= 32, "output width not supported (too many TMA boxes per k-block)");
    PPD_REQUIRE(workspace && workspace_bytes >= (size_t)splits * K * Cout * sizeof(float), "workspace too small (ppd_conv_wgrad_workspace)");
    const ConvA& cv = a.conv;
    CUtensorMap tmA, tmB;
    int rc;
    if (nchw) {
        cuuint64_t dims[5] = {(cuuint64_t)g->kw, (cuuint64_t)g->kh, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)g->B * g->C};
        cuuint64_t str[4] = {(cuuint64_t)g->W * 4, (cuuint64_t)g->stride * 4, (cuuint64_t)g->stride * g->W * 4, (cuuint64_t)g->H * g->W * 4};
        cuuint32_t box[5] = {(cuuint32_t)g->kw, (cuuint32_t)g->kh, (cuuint32_t)cv.segw, 1, 1};
        rc = make_map_nd(&tmA, x, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)g->kw * g->C, (cuuint64_t)OW, (cuuint64_t)g->H, (cuuint64_t)g->B};
        cuuint64_t str[3] = {(cuuint64_t)g->stride * g->C * 4, (cuuint64_t)g->W * g->C * 4, (cuuint64_t)g->H * g->W * g->C * 4};
        cuuint32_t box[4] = {64, (cuuint32_t)cv.segw, 1, 1};
        rc = make_map_nd(&tmA, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    if (rc) return rc;
    const int64_t M = (int64_t)g->B * OH * OW;
    if ((rc = make_map_2d(&tmB, dy, M, Cout, Cout, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    a.C = nullptr; a.ldc = 0; a.I = K; a.J = Cout; a.KK = M;
    a.a_ptr = x;
    a.bn = Cout; a.a_mn = 1; a.b_mn = 1; a.b_presplit = 0;
    a.num_m = num_m; a.num_n = 1; a.splits = splits; a.kk_per_split = 0;
    a.partial = reinterpret_cast<float*>(workspace);
    a.total_items = num_m * splits;
    int sb_stages = (int)((kSmemBudget - 1024 - (size_t)kSA * BM * BK * 4) / ((size_t)2 * a.bn * BK * 4));
    if (sb_stages > 8) sb_stages = 8;
    a.sb_stages = sb_stages;
    const size_t smem = (size_t)kSA * BM * BK * 4 + (size_t)sb_stages * 2 * a.bn * BK * 4 + 1024;
    const int grid = a.total_items < g_max_ctas ? a.total_items : g_max_ctas;
    cudaError_t e = launch_kernel(grid, smem, s, tmA, tmB, tmB, a);
    if (e != cudaSuccess) { set_error("tca conv: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    if (splits_out) *splits_out = splits;
    (void)accumulate; (void)dW;
    return launch_status("tca_gemm_kernel(wgrad)");
}

Plan make_plan(int64_t I, int64_t J, int64_t KK, size_t ws_avail, bool limit) {
    Plan p;
    p.bn = J <= 32 ? 32 : 64;          // the MMAs run at N = 2 bn ([hi | lo] of B side by side)
    p.num_n = (int)((J + p.bn - 1) / p.bn);
    p.num_m = (int)((I + BM - 1) / BM);
    const int64_t tiles = (int64_t)p.num_n * p.num_m;
    const int64_t nkb = (KK + BK - 1) / BK;
    int64_t splits = 1;
    if (tiles < g_max_ctas) {
        splits = g_max_ctas / tiles;                // fill the SMs once; every split keeps >= 4 k-blocks
        if (splits > nkb / 4) splits = nkb / 4;
        if (splits < 1) splits = 1;
    }
    // tcgen05 accumulates with truncation: with ONE accumulator a chain of KK/8 x 3 accumulate steps lost ~1.2e-5 (relative) at
    // KK ~ 1500.  The main / correction accumulator pair sees a third of the steps per accumulator (4e-6 at KK = 1568, inside the
    // 1e-5 gate), so chains are only cut beyond 64 k-blocks (2048), where the output is small enough for the partials to be cheap.
    if (KK > 2048 && tiles <= 2 * g_max_ctas && splits < (nkb + 63) / 64) splits = (nkb + 63) / 64;
    if (limit) {
        while (splits > 1 && (size_t)splits * I * J * sizeof(float) > ws_avail) --splits;
    }
    int64_t per = (nkb + splits - 1) / splits * BK;
    splits = (KK + per - 1) / per;
    p.splits = (int)splits;
    p.kk_per_split = per;
    p.ws = splits > 1 ? (size_t)splits * I * J * sizeof(float) : 0;
    return p;
}

int launch(const ppd_gemm_args* g, int transpose_out, const float* b_lo, void* workspace, size_t workspace_bytes,
           cudaStream_t s, Plan* plan_out) {
    Plan p = make_plan(g->I, g->J, g->KK, workspace ? workspace_bytes : 0, true);
    CUtensorMap tmA, tmB, tmBlo;
    int rc;
    if (g->a_kmajor) rc = make_map_2d(&tmA, g->A, g->I, g->KK, g->lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    else             rc = make_map_2d(&tmA, g->A, g->KK, g->I, g->lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    for (int k = 0; k < (b_lo ? 2 : 1); ++k) {
        const float* base = k ? b_lo : g->B;
        CUtensorMap* m = k ? &tmBlo : &tmB;
        if (g->b_kmajor) rc = make_map_2d(m, base, g->J, g->KK, g->ldb, BK, p.bn, CU_TENSOR_MAP_SWIZZLE_128B);
        else             rc = make_map_2d(m, base, g->KK, g->J, g->ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    if (!b_lo) tmBlo = tmB;
    Args a = {};
    a.C = g->C; a.ldc = g->ldc; a.I = g->I; a.J = g->J; a.KK = g->KK;
    a.bias = g->bias; a.mask = g->mask; a.ldm = g->ldm; a.relu = g->relu; a.accumulate = g->accumulate;
    a.transpose_out = transpose_out;
    a.bn = p.bn; a.a_mn = g->a_kmajor ? 0 : 1; a.b_mn = g->b_kmajor ? 0 : 1; a.b_presplit = b_lo ? 1 : 0;
    a.num_m = p.num_m; a.num_n = p.num_n; a.splits = p.splits; a.kk_per_split = p.kk_per_split;
    a.partial = p.splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
    a.total_items = p.num_m * p.num_n * p.splits;
    int sb_stages = (int)((kSmemBudget - 1024 - (size_t)kSA * BM * BK * 4) / ((size_t)2 * p.bn * BK * 4));
    if (sb_stages > 8) sb_stages = 8;
    // dense, unmasked, un-split outputs (FC and GRU-projection forward) leave through TMA like the forward convolutions
    CUtensorMap tmC;
    const size_t a_region = (size_t)kSA * BM * BK * 4, b_stage = (size_t)2 * p.bn * BK * 4, stage_bytes = (size_t)(p.bn / 32) * BM * 128;
    a.tma_store = 0;
    if (tma_store_enabled() && !transpose_out && p.splits == 1 && !g->accumulate && !g->mask && g->ldc == g->J && g->J % 32 == 0 &&
        !((uintptr_t)g->C & 15)) {
        while (sb_stages > 4 && a_region + sb_stages * b_stage + stage_bytes + 1024 > kSmemBudget) --sb_stages;
        if (a_region + sb_stages * b_stage + stage_bytes + 1024 <= kSmemBudget) {
            if ((rc = make_map_2d(&tmC, g->C, g->I, g->J, g->ldc, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            a.tma_store = 1;
            a.stage_off = (uint32_t)(a_region + sb_stages * b_stage);
        }
    }
    a.sb_stages = sb_stages;
    const size_t smem = a_region + (size_t)sb_stages * b_stage + (a.tma_store ? stage_bytes : 0) + 1024;
    const int grid = a.total_items < g_max_ctas ? a.total_items : g_max_ctas;
    {
        cudaError_t e = launch_kernel(grid, smem, s, tmA, tmB, tmBlo, a, a.tma_store ? &tmC : nullptr);
        if (e != cudaSuccess) { set_error("tca_gemm: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    }
    if (plan_out) *plan_out = p;
    return launch_status("tca_gemm_kernel");
}

}  // namespace tca
}  // namespace ppd
