// sgns_ordered.cu -- ORDERED mode of o1 / o2 (and the legacy fused pass): one warp replays the reference's
// sequential update stream, bit-for-bit.
//
// What "bit-for-bit" means here: the reference computes dot products through BLAS sdot and updates through BLAS
// saxpy (pyx:140, 146-149).  The golden vectors were produced with OpenBLAS' SkylakeX kernels; dot_refblas() below
// performs the same fp32 operations in the same association order (four 16-lane FMA accumulators per 64 elements,
// 16->8 fold, an optional 32-element block on 8-lane accumulators, ((a0+a1)+a2)+a3, 8->4 fold, two horizontal adds,
// <32-element tail in double) and, for FAST_VERSION 0 (pyx:536-541), the "float return register read as a double"
// reinterpretation.  saxpy is one FMA per element.  Every step is an IEEE-754 operation with a single defined
// rounding, so the device result equals the CPU result exactly.
//
// This is a correctness mode: a single warp, no parallelism across pairs (pair p+1 must see every write of pair p).
#include "comemb_common.cuh"

bool g_force_generic_ordered = false;  // tests: comemb_set_tuning(.., .., 900) routes size 128 to the generic kernels

namespace {

// ---- dot product in the reference's summation order ---------------------------------------------------------------------
// x, y: rows in global memory, previously written (possibly by other lanes of this warp) and made visible by
// __syncwarp().  Returns the value on every lane.
__device__ float dot_refblas(const float *x, const float *y, int n, bool quirk) {
    const int lane = threadIdx.x & 31;
    // accumulator q = 16k+l of the four 16-lane vectors lives on lane q%32, register q/32
    float a0 = 0.f, a1 = 0.f;
    const int n64 = n & ~63;
    int i = 0;
    for (; i < n64; i += 64) {
        a0 = fmaf(x[i + lane], y[i + lane], a0);
        a1 = fmaf(x[i + 32 + lane], y[i + 32 + lane], a1);
    }
    // fold 16 -> 8 lanes: acc_k[l] = a5_k[l] + a5_k[l+8]; valid on lanes with (lane%16) < 8
    a0 = a0 + __shfl_down_sync(FULL, a0, 8);
    a1 = a1 + __shfl_down_sync(FULL, a1, 8);
    // now: a0 on lanes 0..7 = acc_0, lanes 16..23 = acc_1; a1 on lanes 0..7 = acc_2, lanes 16..23 = acc_3
    const int n32 = n & ~31;
    if (i < n32) {  // at most one block of 32: acc_k[l] = fma(x[i+8k+l], y[i+8k+l], acc_k[l])
        const int l = lane & 7;
        const bool hi = (lane & 16) != 0;  // lanes 16..23 hold k=1 (a0) and k=3 (a1)
        if ((lane & 8) == 0) {
            const int e0 = i + (hi ? 8 : 0) + l;
            const int e1 = i + (hi ? 24 : 16) + l;
            a0 = fmaf(x[e0], y[e0], a0);
            a1 = fmaf(x[e1], y[e1], a1);
        }
        i += 32;
    }
    // v[l] = ((acc0[l] + acc1[l]) + acc2[l]) + acc3[l] on lanes 0..7
    float v = a0 + __shfl_down_sync(FULL, a0, 16);
    v = v + a1;
    v = v + __shfl_down_sync(FULL, a1, 16);
    // h[l] = v[l] + v[l+4], l < 4 ; my = (h0 + h1) + (h2 + h3)
    float h = v + __shfl_down_sync(FULL, v, 4);
    float p = h + __shfl_down_sync(FULL, h, 1);
    float my = p + __shfl_down_sync(FULL, p, 2);
    my = __shfl_sync(FULL, my, 0);
    // tail in double from fp32 products, then + (double)my
    double dot = 0.0;
    for (; i < n; i++) dot = __dadd_rn(dot, (double)__fmul_rn(y[i], x[i]));
    dot = __dadd_rn(dot, (double)my);
    float fl = __double2float_rn(dot);
    if (!quirk) return fl;
    // FAST_VERSION 0: caller reads xmm0 as a double = {upper half of the double intermediate, float bit pattern}
    double seen = __hiloint2double(__double2hiint(dot), __float_as_int(fl));
    return __double2float_rn(seen);
}

// y += a*x over n elements, one FMA each, lane-strided.
__device__ __forceinline__ void axpy_rows(int n, float a, const float *x, float *y) {
    for (int e = threadIdx.x & 31; e < n; e += 32) y[e] = fmaf(a, x[e], y[e]);
}

struct Sampler {
    const uint32_t *table;
    TableMod mod;
};

// fast_o2 (pyx:105-151).  work: shared memory [size].
__device__ uint64_t pair_o2(const Sampler &S, float *node, float *ctx, int size, uint32_t word_index,
                            uint32_t word2_index, float lr, float lambda, float *work, uint64_t next_random,
                            int negative, bool quirk, const float *lut) {
    const int lane = threadIdx.x & 31;
    float *row1 = node + (int64_t)word2_index * size;
    for (int e = lane; e < size; e += 32) work[e] = 0.f;  // pyx:126
    __syncwarp();
    for (int d = 0; d < negative + 1; d++) {
        uint32_t target;
        float label;
        if (d == 0) {
            target = word_index;
            label = 1.f;
        } else {
            target = S.table[table_slot(next_random, S.mod)];  // pyx:133
            next_random = lcg_next(next_random);                // pyx:134
            if (target == word_index) continue;                 // pyx:135-136
            label = 0.f;
        }
        float *row2 = ctx + (int64_t)target * size;
        float f = dot_refblas(row1, row2, size, quirk);  // pyx:140
        if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;  // pyx:141-142
        f = lut[lut_index(f)];                            // pyx:143
        float g = __fmul_rn(__fmul_rn(label - f, lr), lambda);  // pyx:144
        axpy_rows(size, g, row2, work);  // pyx:146
        axpy_rows(size, g, row1, row2);  // pyx:147
        __syncwarp();
    }
    for (int e = lane; e < size; e += 32) row1[e] = row1[e] + work[e];  // pyx:149 (saxpy with alpha=1)
    __syncwarp();
    return next_random;
}

// fast_o1 (pyx:205-249): targets are rows of the node table and are not updated.
__device__ uint64_t pair_o1(const Sampler &S, float *node, int size, uint32_t word_index, uint32_t word2_index,
                            float lr, float *work, uint64_t next_random, int negative, bool quirk,
                            const float *lut) {
    const int lane = threadIdx.x & 31;
    float *row1 = node + (int64_t)word2_index * size;
    for (int e = lane; e < size; e += 32) work[e] = 0.f;
    __syncwarp();
    for (int d = 0; d < negative + 1; d++) {
        uint32_t target;
        float label;
        if (d == 0) {
            target = word_index;
            label = 1.f;
        } else {
            target = S.table[table_slot(next_random, S.mod)];
            next_random = lcg_next(next_random);
            if (target == word_index) continue;
            label = 0.f;
        }
        const float *row2 = node + (int64_t)target * size;
        float f = dot_refblas(row1, row2, size, quirk);
        if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
        f = lut[lut_index(f)];
        float g = __fmul_rn(label - f, lr);  // pyx:243
        axpy_rows(size, g, row2, work);      // pyx:245
        __syncwarp();
    }
    for (int e = lane; e < size; e += 32) row1[e] = row1[e] + work[e];  // pyx:247
    __syncwarp();
    return next_random;
}

__global__ void __launch_bounds__(32) o2_ordered_kernel(float *node, float *ctx, int size, const uint32_t *walks,
                                                        const int64_t *walk_off, int64_t n_walks,
                                                        const uint64_t *seeds, uint64_t base_seed, Sampler S,
                                                        int window, int negative, float lr, float lambda, bool quirk,
                                                        int64_t *n_tokens, const float *g_exp_table) {
    extern __shared__ float smem[];
    float *lut = smem;                    // [1000]
    float *work = smem + EXP_TABLE_SIZE;  // [size]
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += 32) lut[e] = g_exp_table[e];
    __syncwarp();
    int64_t tokens = 0;
    for (int64_t w = 0; w < n_walks; w++) {  // context_embeddings.py:83-84: paths in order, one worker
        const uint32_t *path = walks + walk_off[w];
        int64_t len = walk_off[w + 1] - walk_off[w];
        if (len > MAX_SENTENCE_LEN) len = MAX_SENTENCE_LEN;  // pyx:480
        uint64_t next_random = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        for (int64_t i = 0; i < len; i++) {  // pyx:494
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            tokens++;
            int64_t j = i - window;
            if (j < 0) j = 0;
            int64_t k = i + window + 1;
            if (k > len) k = len;
            for (; j < k; j++) {  // pyx:503
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                next_random = pair_o2(S, node, ctx, size, wi, wj, lr, lambda, work, next_random, negative, quirk, lut);
            }
        }
    }
    if (n_tokens && threadIdx.x == 0) *n_tokens += tokens;
}

// ---- o2 ORDERED, size == 128, register-resident ---------------------------------------------------------------------------
// Same result as o2_ordered_kernel, several times faster: lane t keeps elements {t, t+32, t+64, t+96} of every row in
// registers (that is the element->accumulator mapping of the reference's sdot: accumulator q = e mod 64 lives on lane
// q mod 32), the NEG+1 rows of a pair are gathered together, the positive context row stays in registers across the
// centre's window, samples are fetched one pair ahead.  Every floating-point operation and its association order is
// the one of dot_refblas()/axpy_rows() above, so the tables stay bit-identical to the reference.
struct Row4 {
    float v0, v1, v2, v3;  // elements t, t+32, t+64, t+96
};
__device__ __forceinline__ Row4 ld_row4(const float *row, int lane) {
    Row4 r;
    r.v0 = row[lane]; r.v1 = row[lane + 32]; r.v2 = row[lane + 64]; r.v3 = row[lane + 96];
    return r;
}
__device__ __forceinline__ void st_row4(float *row, int lane, const Row4 &r) {
    row[lane] = r.v0; row[lane + 32] = r.v1; row[lane + 64] = r.v2; row[lane + 96] = r.v3;
}
// dot of two 128-element rows in the reference's order (see dot_refblas): value on every lane
__device__ __forceinline__ float dot128_refblas(const Row4 &x, const Row4 &y, bool quirk) {
    float a0 = fmaf(x.v2, y.v2, fmaf(x.v0, y.v0, 0.f));  // accumulator q = t      : elements t, t+64
    float a1 = fmaf(x.v3, y.v3, fmaf(x.v1, y.v1, 0.f));  // accumulator q = t + 32 : elements t+32, t+96
    a0 = a0 + __shfl_down_sync(FULL, a0, 8);
    a1 = a1 + __shfl_down_sync(FULL, a1, 8);
    float v = a0 + __shfl_down_sync(FULL, a0, 16);
    v = v + a1;
    v = v + __shfl_down_sync(FULL, a1, 16);
    const float h = v + __shfl_down_sync(FULL, v, 4);
    const float p = h + __shfl_down_sync(FULL, h, 1);
    float my = p + __shfl_down_sync(FULL, p, 2);
    my = __shfl_sync(FULL, my, 0);
    if (!quirk) return my;  // (float)(0.0 + (double)my) == my
    const double dot = (double)my;
    return __double2float_rn(__hiloint2double(__double2hiint(dot), __float_as_int(my)));
}
__device__ __forceinline__ void fma_row4(Row4 &y, float a, const Row4 &x) {
    y.v0 = fmaf(a, x.v0, y.v0); y.v1 = fmaf(a, x.v1, y.v1); y.v2 = fmaf(a, x.v2, y.v2); y.v3 = fmaf(a, x.v3, y.v3);
}

template <int NEG>
__global__ void __launch_bounds__(32)
    o2_ordered_d128_kernel(float *node, float *ctx, const uint32_t *walks, const int64_t *walk_off, int64_t n_walks,
                           const uint64_t *seeds, uint64_t base_seed, Sampler S, int window, float lr, float lambda,
                           bool quirk, int64_t *n_tokens, const float *g_exp_table) {
    constexpr int D = 128;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < EXP_TABLE_SIZE; e += 32) lut[e] = g_exp_table[e];
    __syncwarp();
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    int64_t tokens = 0;
    for (int64_t w = 0; w < n_walks; w++) {
        const uint32_t *path = walks + walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
        uint64_t rnd = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        uint32_t tnext = (lane < NEG) ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
        for (int i = 0; i < len; i++) {  // pyx:494
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            tokens++;
            float *pos_ptr = ctx + (int64_t)wi * D;
            Row4 cpos = ld_row4(pos_ptr, lane);
            const int j1 = min(len, i + window + 1);
            for (int j = max(0, i - window); j < j1; j++) {  // pyx:503
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1_ptr = node + (int64_t)wj * D;
                const Row4 x = ld_row4(row1_ptr, lane);
                const uint32_t tmine = tnext;
                tnext = (lane < NEG) ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                uint32_t t[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) t[k] = __shfl_sync(FULL, tmine, k);
                bool anydup = false;
#pragma unroll
                for (int k = 1; k < NEG; k++)
#pragma unroll
                    for (int a = 0; a < k; a++) anydup = anydup || (t[a] == t[k]);
                Row4 work = {0.f, 0.f, 0.f, 0.f};  // pyx:126
                {  // positive target, pyx:129-131
                    const float f = dot128_refblas(x, cpos, quirk);
                    if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                        const float g = __fmul_rn(__fmul_rn(1.f - lut[lut_index(f)], lr), lambda);
                        fma_row4(work, g, cpos);  // pyx:146
                        fma_row4(cpos, g, x);     // pyx:147 (registers; flushed when the centre ends)
                    }
                }
                if (!anydup) {
                    Row4 c[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) c[k] = ld_row4(ctx + (int64_t)t[k] * D, lane);
                    float f[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) f[k] = dot128_refblas(x, c[k], quirk);
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        if (t[k] == wi) continue;                                  // pyx:135-136
                        if (f[k] <= -MAX_EXP_F || f[k] >= MAX_EXP_F) continue;    // pyx:141-142
                        const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f[k])], lr), lambda);
                        fma_row4(work, g, c[k]);
                        fma_row4(c[k], g, x);
                        st_row4(ctx + (int64_t)t[k] * D, lane, c[k]);
                    }
                } else {  // equal samples inside one pair: strictly one after the other, re-reading the row
#pragma unroll 1
                    for (int k = 0; k < NEG; k++) {
                        const uint32_t tk = __shfl_sync(FULL, tmine, k);
                        if (tk == wi) continue;
                        float *cp = ctx + (int64_t)tk * D;
                        Row4 c = ld_row4(cp, lane);
                        const float f = dot128_refblas(x, c, quirk);
                        if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                        const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f)], lr), lambda);
                        fma_row4(work, g, c);
                        fma_row4(c, g, x);
                        st_row4(cp, lane, c);
                    }
                }
                Row4 nx = {x.v0 + work.v0, x.v1 + work.v1, x.v2 + work.v2, x.v3 + work.v3};  // pyx:149
                st_row4(row1_ptr, lane, nx);
            }
            st_row4(pos_ptr, lane, cpos);
        }
    }
    if (n_tokens && lane == 0) *n_tokens += tokens;
}

// ---- o1 ORDERED, size == 128, register-resident (same arithmetic as pair_o1 / dot_refblas, see o2_ordered_d128_kernel) -
template <int NEG>
__global__ void __launch_bounds__(32)
    o1_ordered_d128_kernel(float *node, const uint32_t *edges, int64_t n_edges, const uint64_t *seeds,
                           uint64_t base_seed, Sampler S, float lr, bool quirk, const float *g_exp_table) {
    constexpr int D = 128;
    constexpr LcgJump<2 * NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < EXP_TABLE_SIZE; e += 32) lut[e] = g_exp_table[e];
    __syncwarp();
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < 2 * NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    // one directed update (fast_o1, pyx:205-249): returns row x + work
    auto directed = [&](const Row4 &x, const Row4 &target, uint32_t word_index, const uint32_t (&tt)[NEG],
                        const Row4 (&c)[NEG]) -> Row4 {
        Row4 work = {0.f, 0.f, 0.f, 0.f};
        {
            const float f = dot128_refblas(x, target, quirk);
            if (f > -MAX_EXP_F && f < MAX_EXP_F) fma_row4(work, __fmul_rn(1.f - lut[lut_index(f)], lr), target);
        }
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            if (tt[k] == word_index) continue;  // pyx:234-235
            const float f = dot128_refblas(x, c[k], quirk);
            if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
            fma_row4(work, __fmul_rn(0.f - lut[lut_index(f)], lr), c[k]);  // pyx:243-245
        }
        Row4 nx = {x.v0 + work.v0, x.v1 + work.v1, x.v2 + work.v2, x.v3 + work.v3};  // pyx:247
        return nx;
    };
    for (int64_t q = 0; q < n_edges; q++) {  // node_embeddings.py:70-71
        const uint32_t e0 = edges[2 * q], e1 = edges[2 * q + 1];
        const uint64_t rnd = seeds ? seeds[q] : (splitmix64(base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        const uint32_t tmine = (lane < 2 * NEG) ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
        Row4 r0 = ld_row4(node + (int64_t)e0 * D, lane);
        Row4 r1 = ld_row4(node + (int64_t)e1 * D, lane);
        uint32_t ta[NEG], tb[NEG];
        Row4 ca[NEG], cb[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            ta[k] = __shfl_sync(FULL, tmine, k);
            tb[k] = __shfl_sync(FULL, tmine, NEG + k);
            ca[k] = ld_row4(node + (int64_t)ta[k] * D, lane);
            cb[k] = ld_row4(node + (int64_t)tb[k] * D, lane);
        }
        r0 = directed(r0, r1, e1, ta, ca);  // pyx:444 (targets are read-only: a sample equal to e0 is the old row)
        st_row4(node + (int64_t)e0 * D, lane, r0);
        if (e1 == e0) r1 = r0;
#pragma unroll
        for (int k = 0; k < NEG; k++)
            if (tb[k] == e0) cb[k] = r0;  // pyx:447 sees the updated row e0, also as a sample
        r1 = directed(r1, r0, e0, tb, cb);
        st_row4(node + (int64_t)e1 * D, lane, r1);
    }
}

__global__ void __launch_bounds__(32) o1_ordered_kernel(float *node, int size, const uint32_t *edges, int64_t n_edges,
                                                        const uint64_t *seeds, uint64_t base_seed, Sampler S,
                                                        int negative, float lr, bool quirk, const float *g_exp_table) {
    extern __shared__ float smem[];
    float *lut = smem;
    float *work = smem + EXP_TABLE_SIZE;
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += 32) lut[e] = g_exp_table[e];
    __syncwarp();
    for (int64_t q = 0; q < n_edges; q++) {  // node_embeddings.py:70-71
        const uint32_t e0 = edges[2 * q], e1 = edges[2 * q + 1];
        uint64_t next_random = seeds ? seeds[q] : (splitmix64(base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        next_random = pair_o1(S, node, size, e1, e0, lr, work, next_random, negative, quirk, lut);  // pyx:444
        next_random = pair_o1(S, node, size, e0, e1, lr, work, next_random, negative, quirk, lut);  // pyx:447
    }
}

// ---- legacy fused pass (stale train_sg; training_sdg_inner.c:1597-1905, 2520-2715, 2988-3740) ----------------------------
// smem: lut[1000] | work[size] | work_o3[size]
__global__ void __launch_bounds__(32)
    sg_fused_ordered_kernel(float *node, float *negemb, int size, const uint32_t *walks, const int64_t *walk_off,
                            int64_t n_walks, const int32_t *reduced_windows, const uint64_t *seeds,
                            uint64_t base_seed, Sampler S, const float *mu, const float *inv_cov, const float *pi, int K,
                            int window, int negative, float lr, float lambda1, float lambda2, int is_node_embedding,
                            bool quirk, const float *g_exp_table) {
    extern __shared__ float smem[];
    float *lut = smem;
    float *work = lut + EXP_TABLE_SIZE;
    float *work_o3 = work + size;
    const int lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += 32) lut[e] = g_exp_table[e];
    __syncwarp();
    const float clipv = __double2float_rn(__dmul_rn((double)lr, 0.1));  // c:2556 `_alpha = alpha * 0.1` (double product)
    const float nl2 = -lambda2;               // c:3132
    for (int64_t w = 0; w < n_walks; w++) {
        const uint32_t *path = walks + walk_off[w];
        const int32_t *rw = reduced_windows ? reduced_windows + walk_off[w] : nullptr;
        int64_t len = walk_off[w + 1] - walk_off[w];
        if (len > MAX_SENTENCE_LEN) len = MAX_SENTENCE_LEN;
        uint64_t next_random = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        for (int64_t i = 0; i < len; i++) {
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int r = rw ? rw[i] : 0;
            int64_t j = i - window + r;
            if (j < 0) j = 0;
            int64_t k = i + window + 1 - r;
            if (k > len) k = len;
            for (; j < k; j++) {
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1 = node + (int64_t)wj * size;
                // (1) o3 gradient of x_j from its current value; lanes own outputs a = lane, lane+32, ...
                for (int e = lane; e < size; e += 32) work_o3[e] = 0.f;
                if (nl2 != 0.f) {
                    for (int a = lane; a < size; a += 32) {
                        double acc = 0.0;
                        for (int c = 0; c < K; c++) {
                            const float p = pi[(int64_t)wj * K + c];
                            const float *Sm = inv_cov + (int64_t)c * size * size;
                            const float *m = mu + (int64_t)c * size;
                            double t = 0.0;
                            for (int b = 0; b < size; b++) {
                                const float df = row1[b] - m[b];
                                t = __dadd_rn(t, __dmul_rn((double)__fmul_rn(p, Sm[(int64_t)b * size + a]), (double)df));
                            }
                            acc = __dadd_rn(acc, t);
                        }
                        float v = __fmul_rn(nl2, __double2float_rn(acc));
                        work_o3[a] = v < -clipv ? -clipv : (v > clipv ? clipv : v);
                    }
                }
                // (2) SGNS pair
                for (int e = lane; e < size; e += 32) work[e] = 0.f;
                __syncwarp();
                for (int d = 0; d < negative + 1; d++) {
                    uint32_t target;
                    float label;
                    if (d == 0) {
                        target = wi;
                        label = 1.f;
                    } else {
                        target = S.table[table_slot(next_random, S.mod)];
                        next_random = lcg_next(next_random);
                        if (target == wi) continue;
                        label = 0.f;
                    }
                    float *row2 = negemb + (int64_t)target * size;
                    float f = dot_refblas(row1, row2, size, quirk);
                    if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                    f = lut[lut_index(f)];
                    const float g = __fmul_rn(label - f, lr);  // c:1813
                    const float gl = __fmul_rn(g, lambda1);    // c:1822
                    axpy_rows(size, g, row2, work);
                    if (!is_node_embedding) axpy_rows(size, gl, row1, row2);  // c:1840-1859
                    __syncwarp();
                }
                for (int e = lane; e < size; e += 32) {
                    float v = fmaf(lambda1, work[e], row1[e]);  // c:1870
                    row1[e] = v + work_o3[e];                    // c:3668
                }
                __syncwarp();
            }
        }
    }
}


// ---- Python-twin semantics (utils/embedding.py:15-98): explicit target lists, exact sigmoid, vectorised update ------------
// One pair = (node2 row, neg+1 target rows [positive first]) prepared by the host, which also draws the negatives
// with the twin's own rejection loop on np.random (embedding.py:48-52).  Per pair, as the numpy code does it:
//   fb_k  = expit(float32 dot(x, c_k))            all k from the OLD context rows            (:78-83)
//   gb_k  = (label_k - fb_k) * alpha              float64
//   work  = sum_k gb_k * c_k                      float64                                    (:58)
//   o3    = -clip(sum_c (pi*inv_cov[c]) @ (x - mu_c) * lambda2, +-0.1*alpha)   float32, inv_cov NOT transposed (:86-98)
//   ctx[t_k] = float32(c_k + gb_k * x)            computed from the gathered copy: duplicates -> last one wins (:67)
//   x     = float32(x + (lambda1*work + o3))                                                   (:70)
// smem: xs[size] | newx[size] | cold[(neg+1)*size]
__global__ void __launch_bounds__(32)
    sg_twin_kernel(float *node, float *ctxemb, int size, const uint32_t *pair_row, const uint32_t *targets,
                   int64_t n_pairs, int negative, double alpha, double lambda1, double lambda2, const float *mu,
                   const float *inv_cov, const float *pi, int K, int is_node_embedding) {
    extern __shared__ float smem[];
    float *xs = smem;
    float *newx = smem + size;
    float *cold = smem + 2 * size;
    const int lane = threadIdx.x & 31;
    const int T = negative + 1;  // <= 8 (checked by the launcher)
    for (int64_t p = 0; p < n_pairs; p++) {
        const uint32_t n2 = pair_row[p];
        const uint32_t *t = targets + p * T;
        float *x = node + (int64_t)n2 * size;
        __syncwarp();
        for (int e = lane; e < size; e += 32) xs[e] = x[e];
        for (int k = 0; k < T; k++)
            for (int e = lane; e < size; e += 32) cold[k * size + e] = ctxemb[(int64_t)t[k] * size + e];
        __syncwarp();
        double gb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            gb[k] = 0.0;
            if (k < T) {
                double acc = 0.0;
                for (int e = lane; e < size; e += 32) acc += (double)xs[e] * (double)cold[k * size + e];
                for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                const float f = (float)acc;                            // np.dot of float32 operands -> float32
                const float fb = 1.0f / (1.0f + expf(-f));            // scipy.special.expit on float32
                gb[k] = ((k == 0 ? 1.0 : 0.0) - (double)fb) * alpha;  // float64
            }
        }
        for (int e = lane; e < size; e += 32) {
            double work = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k < T) work += gb[k] * (double)cold[k * size + e];
            float o3 = 0.f;
            if (lambda2 > 0.0) {
                float grad = 0.f;
                for (int c = 0; c < K; c++) {
                    const float pw = pi[(int64_t)n2 * K + c];
                    const float *S = inv_cov + (int64_t)c * size * size + (int64_t)e * size;  // row e: NOT transposed
                    double sdot = 0.0;
                    for (int b = 0; b < size; b++)
                        sdot += (double)__fmul_rn(pw, S[b]) * (double)(xs[b] - mu[(int64_t)c * size + b]);
                    grad = grad + __fmul_rn((float)sdot, (float)lambda2);
                }
                const float lim = (float)(0.1 * alpha);
                o3 = -fminf(fmaxf(grad, -lim), lim);
            }
            newx[e] = (float)((double)xs[e] + (lambda1 * work + (double)o3));
            if (!is_node_embedding) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < T) ctxemb[(int64_t)t[k] * size + e] = (float)((double)cold[k * size + e] + gb[k] * (double)xs[e]);
            }
        }
        __syncwarp();
        for (int e = lane; e < size; e += 32) x[e] = newx[e];
    }
}

}  // namespace

// ---- launchers (called from capi.cu) -------------------------------------------------------------------------------------
int launch_o2_ordered(float *node, float *ctx, int size, const uint32_t *walks, const int64_t *walk_off, int64_t n_walks,
                      const uint64_t *seeds, uint64_t base_seed, const uint32_t *table, uint64_t table_len, int window,
                      int negative, float lr, float lambda, bool quirk, int64_t *n_tokens, cudaStream_t st) {
    Sampler S{table, make_table_mod(table_len)};
    if (size == 128 && !g_force_generic_ordered) {  // register-resident fast path, same bits
        switch (negative) {
#define COMEMB_CASE(N)                                                                                             \
    case N:                                                                                                        \
        o2_ordered_d128_kernel<N><<<1, 32, 0, st>>>(node, ctx, walks, walk_off, n_walks, seeds, base_seed, S, window, \
                                                    lr, lambda, quirk, n_tokens, comemb_lut_device());             \
        return (int)cudaGetLastError();
            COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            default: break;
        }
    }
    size_t smem = (EXP_TABLE_SIZE + (size_t)size) * sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(o2_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    o2_ordered_kernel<<<1, 32, smem, st>>>(node, ctx, size, walks, walk_off, n_walks, seeds, base_seed, S, window,
                                           negative, lr, lambda, quirk, n_tokens, comemb_lut_device());
    return (int)cudaGetLastError();
}

int launch_o1_ordered(float *node, int size, const uint32_t *edges, int64_t n_edges, const uint64_t *seeds,
                      uint64_t base_seed, const uint32_t *table, uint64_t table_len, int negative, float lr, bool quirk,
                      cudaStream_t st) {
    Sampler S{table, make_table_mod(table_len)};
    if (size == 128 && !g_force_generic_ordered) {  // register-resident fast path, same bits
        switch (negative) {
#define COMEMB_CASE(N)                                                                                              \
    case N:                                                                                                         \
        o1_ordered_d128_kernel<N><<<1, 32, 0, st>>>(node, edges, n_edges, seeds, base_seed, S, lr, quirk,           \
                                                    comemb_lut_device());                                           \
        return (int)cudaGetLastError();
            COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            default: break;
        }
    }
    size_t smem = (EXP_TABLE_SIZE + (size_t)size) * sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(o1_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    o1_ordered_kernel<<<1, 32, smem, st>>>(node, size, edges, n_edges, seeds, base_seed, S, negative, lr, quirk,
                                           comemb_lut_device());
    return (int)cudaGetLastError();
}

int launch_sg_fused_ordered(float *node, float *negemb, int size, const uint32_t *walks, const int64_t *walk_off,
                            int64_t n_walks, const int32_t *reduced_windows, const uint64_t *seeds, uint64_t base_seed,
                            const uint32_t *table, uint64_t table_len, const float *mu, const float *inv_cov,
                            const float *pi, int K, int window, int negative, float lr, float lambda1, float lambda2,
                            int is_node_embedding, bool quirk, cudaStream_t st) {
    Sampler S{table, make_table_mod(table_len)};
    size_t smem = (EXP_TABLE_SIZE + 2 * (size_t)size) * sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(sg_fused_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sg_fused_ordered_kernel<<<1, 32, smem, st>>>(node, negemb, size, walks, walk_off, n_walks, reduced_windows, seeds,
                                                 base_seed, S, mu, inv_cov, pi, K, window, negative, lr, lambda1,
                                                 lambda2, is_node_embedding, quirk, comemb_lut_device());
    return (int)cudaGetLastError();
}

int launch_sg_twin(float *node, float *ctxemb, int size, const uint32_t *pair_row, const uint32_t *targets,
                   int64_t n_pairs, int negative, double alpha, double lambda1, double lambda2, const float *mu,
                   const float *inv_cov, const float *pi, int K, int is_node_embedding, cudaStream_t st) {
    if (negative + 1 > 8) return COMEMB_E_UNSUPPORTED;
    if (n_pairs == 0) return 0;
    size_t smem = (size_t)(negative + 3) * size * sizeof(float);
    if (smem > 200 * 1024) return COMEMB_E_UNSUPPORTED;
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(sg_twin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sg_twin_kernel<<<1, 32, smem, st>>>(node, ctxemb, size, pair_row, targets, n_pairs, negative, alpha, lambda1, lambda2,
                                        mu, inv_cov, pi, K, is_node_embedding);
    return (int)cudaGetLastError();
}
