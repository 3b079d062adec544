// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), plus the shared-memory operand
// layout used by every tensor-core GEMM in this library.
//
// Operand layout ("blocked", no swizzle).  A matrix X[R][Cc] of 32-bit elements (R % 8 == 0,
// Cc % 4 == 0) is stored as 8-row x 4-column "core matrices" of 128 contiguous bytes:
//     byte_off(r, c) = (r / 8) * RS + (c / 4) * CS + (r % 8) * 16 + (c % 4) * 4
// With UMMA's SWIZZLE_NONE canonical layouts (cute/atom/mma_traits_sm100.hpp, make_umma_desc):
//   * used K-major  (rows = M or N index, cols = K):  LBO = CS, SBO = RS; one K=8 step = 2 chunks,
//     the next step starts 2*CS further.
//   * used MN-major (rows = K index, cols = M or N):  SBO = CS, LBO = RS; one K=8 step = one row
//     group, the next step starts RS further.
// So one copy of an activation / weight matrix in shared memory serves the forward GEMM (K-major),
// the weight-gradient GEMM (MN-major, reduction over rows) and the input-gradient GEMM.
//
// 3xTF32: x = hi + lo with hi = x & 0xFFFFE000 (exact in tf32) and lo = x - hi (exact in fp32, the
// tensor core keeps its top 11 significant bits); a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with
// relative error ~2^-21, i.e. fp32-faithful results from kind::tf32 MMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__host__ __device__ __forceinline__ constexpr uint32_t blk_off(uint32_t r, uint32_t c, uint32_t RS, uint32_t CS) {
    return (r >> 3) * RS + (c >> 2) * CS + (r & 7u) * 16u + (c & 3u) * 4u;
}

// ---- descriptors ------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // version_ = 1 (Blackwell)
    return d;
}
// 32-bit instruction descriptor for kind::tf32, f32 accumulate
__host__ __device__ __forceinline__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)            // c_format  = F32
         | (2u << 7)            // a_format  = TF32
         | (2u << 10)           // b_format  = TF32
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with bf16 operands, f32 accumulate (K = 16 per instruction)
__host__ __device__ __forceinline__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core's reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// TS mode: the A operand ([M = 128 lanes] x [K = 16 bf16 = 8 packed 32-bit columns], K-major only) comes from
// tensor memory instead of shared memory -- an M128 x N64 x K16 MMA then reads 2 KB (B) instead of 6 KB from smem.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 8 consecutive 32-bit columns of the calling thread's TMEM lane (thread t of warp w <-> lane 32 * (w % 4) + t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of warp w reads lane 32*(w%4)+t, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {     // one 32-bit column of the calling thread's lane
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(mbar)), "r"(parity) : "memory");
}

// transaction-count arrive + 1-D bulk copy global -> shared (TMA engine, completes on the mbarrier)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
// orders generic-proxy accesses (any state space) against later async-proxy accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- 3xTF32 split -----------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// D (+)= A*B with fp32-faithful accuracy: three kind::tf32 MMAs per K=8 step over `ksteps` steps.
// a_hi/a_lo/b_hi/b_lo: shared addresses of the split operands (same layout), *_step: bytes to the
// next K step.  Issued by one thread.
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                            uint32_t a_step, uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo,
                                            uint32_t b_step, uint32_t idesc, int ksteps, bool accumulate_first) {
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ah = smem_desc(a_hi + k * a_step, a_lbo, a_sbo), al = smem_desc(a_lo + k * a_step, a_lbo, a_sbo);
        const uint64_t bh = smem_desc(b_hi + k * b_step, b_lbo, b_sbo), bl = smem_desc(b_lo + k * b_step, b_lbo, b_sbo);
        mma_tf32(d_tmem, al, bh, idesc, acc);   // small terms first
        mma_tf32(d_tmem, ah, bl, idesc, 1u);
        mma_tf32(d_tmem, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

// One lane of a converged warp (elect.sync).  MMA issue code must be WARP-UNIFORM: run it in a
// whole warp and predicate the tcgen05 instructions with this, so that descriptors stay in uniform
// registers (a divergent `if (threadIdx.x == 0)` makes ptxas wrap every UTCHMMA in an election loop).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// Descriptor halves: lo32 = (addr>>4) | (LBO>>4)<<16 ; hi32 = (SBO>>4) | version(1)<<14.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// Warp-level version of gemm_bf16x3 with a compile-time K-step count: call from ALL lanes of one
// warp; one elected lane issues the 6*KSTEPS MMAs.  `pieces` = 3 (full) or 1 (operand is exact in bf16).
// FULL = false: three products only -- A0 B0 + A1 B0 + A0 B1 (two pieces per operand, relative error ~2^-16 per term
// instead of ~2^-22): used for the weight-gradient GEMMs, whose results are held to 2e-4 and which are paced by the
// per-instruction cost of their M = 64 shared-memory-operand MMAs, not by flops.
template <int KSTEPS, int A_PIECES = 3, int B_PIECES = 3, bool FULL = true>
__device__ __forceinline__ void gemm_bf16x3_warp(uint32_t d_tmem, uint32_t a0, uint32_t a_part, uint32_t a_lbo, uint32_t a_sbo,
                                                 uint32_t a_step, uint32_t b0, uint32_t b_part, uint32_t b_lbo, uint32_t b_sbo,
                                                 uint32_t b_step, uint32_t idesc) {
    const uint32_t ahi = desc_hi(a_sbo), bhi = desc_hi(b_sbo);
    uint32_t alo[3], blo[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        alo[p] = desc_lo(a0 + (p < A_PIECES ? p : 0) * a_part, a_lbo);
        blo[p] = desc_lo(b0 + (p < B_PIECES ? p : 0) * b_part, b_lbo);
    }
    const uint32_t astep = a_step >> 4, bstep = b_step >> 4;
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
            uint64_t A[3], B[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                A[p] = desc_pack(alo[p] + k * astep, ahi);
                B[p] = desc_pack(blo[p] + k * bstep, bhi);
            }
            bool first = (k == 0);
            // smallest terms first; pairs (i, j) with i + j <= 2
            if (FULL && A_PIECES == 3) { mma_bf16(d_tmem, A[2], B[0], idesc, first ? 0u : 1u); first = false; }
            if (FULL && B_PIECES == 3) { mma_bf16(d_tmem, A[0], B[2], idesc, first ? 0u : 1u); first = false; }
            if (FULL && A_PIECES == 3 && B_PIECES == 3) { mma_bf16(d_tmem, A[1], B[1], idesc, first ? 0u : 1u); first = false; }
            if (A_PIECES == 3) { mma_bf16(d_tmem, A[1], B[0], idesc, first ? 0u : 1u); first = false; }
            if (B_PIECES == 3) { mma_bf16(d_tmem, A[0], B[1], idesc, first ? 0u : 1u); first = false; }
            mma_bf16(d_tmem, A[0], B[0], idesc, first ? 0u : 1u);
        }
    }
    __syncwarp();
}

// Same six-product scheme with the A pieces in TMEM: piece p at columns a_tmem + p * a_part_cols, K step k at
// + 8 k columns (2 bf16 per column).  WARP-LEVEL like gemm_bf16x3_warp.
template <int KSTEPS>
__device__ __forceinline__ void gemm_bf16x3_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint32_t a_part_cols, uint32_t b0,
                                                    uint32_t b_part, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step,
                                                    uint32_t idesc) {
    const uint32_t bhi = desc_hi(b_sbo);
    uint32_t blo[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) blo[p] = desc_lo(b0 + p * b_part, b_lbo);
    const uint32_t bstep = b_step >> 4;
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
            uint32_t A[3];
            uint64_t B[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                A[p] = a_tmem + p * a_part_cols + 8u * k;
                B[p] = desc_pack(blo[p] + k * bstep, bhi);
            }
            mma_bf16_ts(d_tmem, A[2], B[0], idesc, k == 0 ? 0u : 1u);
            mma_bf16_ts(d_tmem, A[0], B[2], idesc, 1u);
            mma_bf16_ts(d_tmem, A[1], B[1], idesc, 1u);
            mma_bf16_ts(d_tmem, A[1], B[0], idesc, 1u);
            mma_bf16_ts(d_tmem, A[0], B[1], idesc, 1u);
            mma_bf16_ts(d_tmem, A[0], B[0], idesc, 1u);
        }
    }
    __syncwarp();
}

// fp32-faithful product from three bf16 pieces per operand (x = b0 + b1 + b2, 24 bits): keep the
// six partial products of weight >= 2^-16; part p of an operand lives `part_bytes` after part p-1.
__device__ __forceinline__ void gemm_bf16x3(uint32_t d_tmem, uint32_t a0, uint32_t a_part, uint32_t a_lbo, uint32_t a_sbo,
                                            uint32_t a_step, uint32_t b0, uint32_t b_part, uint32_t b_lbo, uint32_t b_sbo,
                                            uint32_t b_step, uint32_t idesc, int ksteps, bool accumulate_first) {
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int k = 0; k < ksteps; ++k) {
        uint64_t A[3], B[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            A[p] = smem_desc(a0 + p * a_part + k * a_step, a_lbo, a_sbo);
            B[p] = smem_desc(b0 + p * b_part + k * b_step, b_lbo, b_sbo);
        }
        mma_bf16(d_tmem, A[2], B[0], idesc, acc);   // smallest terms first
        mma_bf16(d_tmem, A[0], B[2], idesc, 1u);
        mma_bf16(d_tmem, A[1], B[1], idesc, 1u);
        mma_bf16(d_tmem, A[1], B[0], idesc, 1u);
        mma_bf16(d_tmem, A[0], B[1], idesc, 1u);
        mma_bf16(d_tmem, A[0], B[0], idesc, 1u);
        acc = 1u;
    }
}

}  // namespace umma
