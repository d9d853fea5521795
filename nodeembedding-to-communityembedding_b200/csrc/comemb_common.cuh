// comemb_common.cuh -- shared device helpers of the ComEmb B200 hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/comemb_b200.h"

#define EXP_TABLE_SIZE 1000  // pyx:89
#define MAX_EXP_F 6.0f       // pyx:90
#define MAX_SENTENCE_LEN 10000  // pyx:18
#define LCG_MUL 25214903917ULL  // pyx:134
#define LCG_MASK 281474976710655ULL  // pyx:121
#define FULL 0xffffffffu

// The sigmoid table of pyx:92-95 / 531-533 lives in a per-device global buffer uploaded by comemb_init(); kernels
// receive its pointer and stage it into shared memory.  nullptr until comemb_init() ran on the current device.
const float *comemb_lut_device();

#define CUDA_TRY(x)                        \
    do {                                   \
        cudaError_t _e = (x);              \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

// ---- LCG of pyx:133-134 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t lcg_next(uint64_t x) { return (x * LCG_MUL + 11ULL) & LCG_MASK; }

// x advanced n steps in O(log n): composition of affine maps mod 2^48.
__device__ __forceinline__ uint64_t lcg_skip(uint64_t x, uint64_t n) {
    uint64_t a = LCG_MUL, c = 11ULL, A = 1ULL, C = 0ULL;
    while (n) {
        if (n & 1ULL) {
            A = A * a;
            C = C * a + c;
        }
        c = (a + 1ULL) * c;
        a = a * a;
        n >>= 1;
    }
    return (A * x + C) & LCG_MASK;
}

// x advanced by k = 0..NEG steps as affine maps with compile-time constants: lane k of a warp jumps straight to the
// state its sample is drawn from.
template <int NEG>
struct LcgJump {  // x_{n+k} = A_k x_n + C_k  (mod 2^48)
    uint64_t A[NEG + 1], C[NEG + 1];
    __host__ __device__ constexpr LcgJump() : A{}, C{} {
        uint64_t a = 1, c = 0;
        for (int k = 0; k <= NEG; k++) {
            A[k] = a & LCG_MASK;
            C[k] = c & LCG_MASK;
            c = (c * LCG_MUL + 11ULL) & LCG_MASK;
            a = (a * LCG_MUL) & LCG_MASK;
        }
    }
};

// (next_random >> 16) % table_len  (pyx:133).  The state has 48 bits, so the dividend is a 32-bit value.
struct TableMod {
    uint64_t len;    // table_len
    uint64_t magic;  // ceil(2^64 / len) for len < 2^32 (Lemire fastmod), 0 otherwise
};
static inline TableMod make_table_mod(uint64_t len) {
    TableMod m;
    m.len = len;
    m.magic = (len > 1 && len <= 0xFFFFFFFFULL) ? (0xFFFFFFFFFFFFFFFFULL / len + 1ULL) : 0ULL;
    return m;
}
__device__ __forceinline__ uint64_t table_slot(uint64_t next_random, const TableMod &m) {
    uint64_t a = next_random >> 16;  // < 2^32
    if (m.magic) return __umul64hi(m.magic * a, m.len);
    return m.len == 1 ? 0ULL : a;  // len >= 2^32 > a
}

// sigma lookup of pyx:143: EXP_TABLE[(int)((f + 6.0) * 83)] with the double arithmetic of the generated C.
__device__ __forceinline__ int lut_index(float f) { return __double2int_rz(((double)f + 6.0) * 83.0); }

__device__ __forceinline__ float warp_sum_xor(float v) {
    v += __shfl_xor_sync(FULL, v, 16);
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    return v;
}

// 8-slot transposed warp reduction.  Slot sums are formed by exactly the same lane pairings, level by level
// (xor 16, 8, 4, 2, 1), as eight independent xor-butterflies would use -- so every slot total is bit-identical to
// warp_sum_xor() of that slot -- but each level halves the number of live slots: 9 shuffles instead of 40.
// On return lane l holds the total of slot (l >> 2) & 7.
__device__ __forceinline__ float reduce8_transposed(float p0, float p1, float p2, float p3, float p4, float p5,
                                                    float p6, float p7, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const float q0 = (b4 ? p4 : p0) + __shfl_xor_sync(FULL, b4 ? p0 : p4, 16);
    const float q1 = (b4 ? p5 : p1) + __shfl_xor_sync(FULL, b4 ? p1 : p5, 16);
    const float q2 = (b4 ? p6 : p2) + __shfl_xor_sync(FULL, b4 ? p2 : p6, 16);
    const float q3 = (b4 ? p7 : p3) + __shfl_xor_sync(FULL, b4 ? p3 : p7, 16);
    const float r0 = (b3 ? q2 : q0) + __shfl_xor_sync(FULL, b3 ? q0 : q2, 8);
    const float r1 = (b3 ? q3 : q1) + __shfl_xor_sync(FULL, b3 ? q1 : q3, 8);
    float t = (b2 ? r1 : r0) + __shfl_xor_sync(FULL, b2 ? r0 : r1, 4);
    t += __shfl_xor_sync(FULL, t, 2);
    t += __shfl_xor_sync(FULL, t, 1);
    return t;
}

// p_i ends on the four lanes whose bits (b4,b3,b2) spell i: level 16 keeps by i>>2, level 8 by (i>>1)&1, level 4 by i&1
__host__ __device__ constexpr int lane_of_p(int i) {
    return ((i >> 2) << 4) | (((i >> 1) & 1) << 3) | ((i & 1) << 2);
}

// splitmix64: per-unit seed derivation for COMEMB_F_SEED_HASH and the Hogwild walker.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// red.global.add.v4.f32 (sm_90+): one 16-byte vector reduction, no return value.
__device__ __forceinline__ void red_add4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// L2 eviction-priority hints (createpolicy): embedding rows are the reused working set (evict_last), walk tokens and
// the sample table are streamed / sparsely reused (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldcg4_hint(const float *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st4_hint(float *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_f32_hint(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// ---- cp.async (LDGSTS): global -> shared without a register or a load scoreboard ------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
int comemb_check_init();  // COMEMB_E_NOINIT unless comemb_init() ran on the current device
// the calling thread's launch options (comemb_set_opts; thread-local, defaults all zero)
const comemb_opts_t &comemb_opts();
