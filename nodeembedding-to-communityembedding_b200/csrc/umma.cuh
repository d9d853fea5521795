// umma.cuh -- sm_100a building blocks for the community-GEMM kernels (o3 grouped step, fused pass, GMM E-step):
// tcgen05.mma (kind::tf32, operands in shared memory, accumulator in TMEM), TMEM allocation and tcgen05.ld, mbarriers,
// bulk (TMA engine) global->shared copies, and the 128-byte-swizzled K-major operand layout.
//
// The one contraction of the ComEmb hot path is  Y[128 x n] = S_c^(T) [128 x 128] . D[128 x n]  with S_c one community's
// inverse covariance and the n columns the (x - mu_c) vectors of rows that belong to community c.  It is run as
// 3xTF32: every fp32 operand is split as v = hi + lo with hi = tf32(v) and lo = v - hi (exact in fp32), and
//      Y = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi          (fp32 accumulation in TMEM)
// which drops only the lo.lo term (2^-22 relative) and the rounding of lo to TF32 (2^-11 of a 2^-11 term), i.e. the
// result carries fp32-level accuracy although it runs on the tensor cores.
//
// Operand layout (both operands K-major, SWIZZLE_128B -- the canonical layout of cute's Layout_K_SW128_Atom<tf32>):
//   a [rows x 128] fp32 operand is stored as 4 "K-atoms" of 32 elements (128 bytes) per row:
//       byte offset(row, k) = (k / 32) * rows * 128            K-atom block
//                           + row * 128                          (8-row groups are 1024 bytes apart: SBO = 1024)
//                           + (((k % 32) / 4) ^ (row % 8)) * 16   16-byte chunk, XOR-swizzled with the row
//                           + (k % 4) * 4
//   Blocks must be 1024-byte aligned (the hardware applies the XOR to address bits [4,7) ^ [7,10)).  One tcgen05.mma of
//   kind::tf32 consumes K = 8 elements (32 bytes): K-step s of a K-atom starts 32*s bytes into the atom's rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

constexpr int KATOM = 32;             // fp32/tf32 elements per 128-byte swizzle row
constexpr int KSTEP = 8;              // K of one tcgen05.mma.kind::tf32
constexpr uint32_t SPIN_LIMIT = 1u << 28;  // bounded waits: a protocol bug must trap, not hang the GPU

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) inside an operand image of `rows` rows (see the header comment)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int rows, int row, int k) {
    return (uint32_t)((k >> 5) * rows * 128 + row * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + ((k & 3) << 2));
}

__device__ __forceinline__ float tf32_round(float x) {  // round-to-nearest TF32, returned as an fp32 bit pattern
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- mbarrier ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > SPIN_LIMIT) asm volatile("trap;");
}

// ---- TMA engine: 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) --------------------------
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------------------------
// one full warp; ncols: power of two >= 32.  The base address is written to *smem_out.
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_out, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid; taddr = lane_base<<16 | column)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---- tcgen05.mma ---------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups `sbo_bytes` apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused by swizzled K-major), [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor: D fp32, A and B TF32, both K-major, M = 128, N = n (multiple of 16, <= 256)
__device__ __forceinline__ uint32_t idesc_tf32_m128(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues on behalf of the CTA
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all tcgen05.mma issued so far by this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The 3xTF32 product of a resident A (hi and lo images, 128 rows each) with a B tile of n rows (hi and lo images):
// 16 K-steps x {hi.hi, hi.lo, lo.hi}.  Small terms first so that they are not absorbed by a large partial sum.
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             int n_rows_b, int n) {
    const uint32_t idesc = idesc_tf32_m128(n);
    bool acc = false;
#pragma unroll 1
    for (int pass = 0; pass < 3; pass++) {
        const uint32_t a0 = pass == 0 ? a_lo : a_hi;
        const uint32_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < 128 / KSTEP; ks++) {
            const uint32_t ka = (uint32_t)((ks >> 2) * 128 * 128 + (ks & 3) * 32);       // A image: 128 rows per K-atom
            const uint32_t kb = (uint32_t)((ks >> 2) * n_rows_b * 128 + (ks & 3) * 32);  // B image: n_rows_b rows
            mma_tf32(d_tmem, smem_desc_sw128(a0 + ka, 1024), smem_desc_sw128(b0 + kb, 1024), idesc, acc);
            acc = true;
        }
    }
}

}  // namespace umma
