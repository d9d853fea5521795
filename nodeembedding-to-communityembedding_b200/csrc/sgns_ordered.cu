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
// This is a correctness mode: no parallelism across pairs (pair p+1 must see every write of pair p).  Kernels:
//   o2/o1_ordered_kernel, sg_fused_ordered_kernel, sg_twin_kernel   any size, one warp, rows in global/shared memory
//   o2/o1_ordered_d128_kernel                                       size 128, one warp, rows in registers
//   o2/o1_ordered_d128_pipe_kernel                                  + next pair's rows requested before the current one
//   o2_ordered_d128_team_kernel (default for o2, window <= 15)      scheduling warp + one worker warp per target row
// They all produce the same bits (tests/test_gpu_parity.py::test_o2_ordered_d128_kernel_variants_hazards).
#include "comemb_common.cuh"
#include "ordered_d128.cuh"
// comemb_opts().variant selects among the size-128 ORDERED kernels: default = one worker warp per target row (o2) /
// pipelined single warp (o1); ORDERED_PIPE = single warp, software-pipelined; ORDERED_PLAIN = single warp, plain;
// GENERIC = the any-size kernels.

namespace {
using namespace ordered;

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

// ---- o2 ORDERED, size == 128, register-resident AND software-pipelined -----------------------------------------------------
// Same arithmetic as o2_ordered_d128_kernel; the difference is when the loads are issued.  A single warp replaying a
// sequential stream is bound by memory latency (row gather -> dots -> scatter -> next gather), so the rows of pair p+1
// (its node row and, inside one centre's window, its NEG sampled context rows) are requested BEFORE pair p is computed
// and the samples of pair p+2 before that.  Exactness is kept by explicit hazard checks: a prefetched row that pair p
// writes (same walk token, or a sample equal to one of pair p's samples) is replaced by the register copy / re-read
// after pair p's stores, and nothing is prefetched across a centre change (the positive row lives in registers until
// then).  Requires node and ctx not to overlap (the launcher checks; otherwise the kernel above is used).
template <int NEG>
__global__ void __launch_bounds__(32)
    o2_ordered_d128_pipe_kernel(float *node, float *ctx, const uint32_t *walks, const int64_t *walk_off,
                                int64_t n_walks, const uint64_t *seeds, uint64_t base_seed, Sampler S, int window,
                                float lr, float lambda, bool quirk, int64_t *n_tokens, const float *g_exp_table) {
    constexpr int D = 128;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    extern __shared__ uint32_t path[];  // [MAX_SENTENCE_LEN] tokens of the current walk
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
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
        {  // tokens go to shared memory: a global token load would wait behind the row requests in flight
            const uint32_t *gpath = walks + walk_off[w];
            __syncwarp();
            int cnt = 0;
            for (int e = lane; e < len; e += 32) {
                const uint32_t tok = gpath[e];
                path[e] = tok;
                cnt += (tok != COMEMB_TOKEN_NONE);
            }
            __syncwarp();
            tokens += __reduce_add_sync(FULL, cnt);
        }
        uint64_t rnd = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        auto fetch = [&]() -> uint32_t {  // the next NEG samples of the stream, one per lane (pyx:133-134)
            const uint32_t v = (lane < NEG) ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
            rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
            return v;
        };
        // first valid (centre i, context j) at or after the candidate position, in the reference's order (pyx:494-507)
        auto seek = [&](int &i, int &j, uint32_t &wi, uint32_t &wj) -> bool {
            while (i < len) {
                if (wi != COMEMB_TOKEN_NONE) {
                    const int j1 = min(len, i + window + 1);
                    for (; j < j1; j++) {
                        if (j == i) continue;
                        wj = path[j];
                        if (wj != COMEMB_TOKEN_NONE) return true;
                    }
                }
                i++;
                if (i < len) {
                    wi = path[i];
                    j = max(0, i - window);
                }
            }
            return false;
        };
        int i = 0, j = 0;
        uint32_t wi = len > 0 ? path[0] : COMEMB_TOKEN_NONE, wj = 0;
        if (!seek(i, j, wi, wj)) continue;
        uint32_t tq_cur = fetch();
        uint32_t tq_nxt = fetch();
        Row4 cpos = ld_row4(ctx + (int64_t)wi * D, lane);
        Row4 x = ld_row4(node + (int64_t)wj * D, lane);
        uint32_t t[NEG];
        Row4 c[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            t[k] = __shfl_sync(FULL, tq_cur, k);
            c[k] = ld_row4(ctx + (int64_t)t[k] * D, lane);
        }
        for (;;) {
            // ---- look ahead: pair p+1, its rows, and the samples of pair p+2 ----
            int ni = i, nj = j + 1;
            uint32_t nwi = wi, nwj = 0;
            const bool nvalid = seek(ni, nj, nwi, nwj);
            const bool same = nvalid && ni == i;
            uint32_t tn[NEG];
            Row4 cn[NEG];
            Row4 xn = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < NEG; k++) tn[k] = __shfl_sync(FULL, tq_nxt, k);
            const uint32_t tq_new = fetch();
            if (nvalid) xn = ld_row4(node + (int64_t)nwj * D, lane);
            bool stale = !same;
#pragma unroll
            for (int k = 0; k < NEG; k++) cn[k] = c[k];
            if (same) {
#pragma unroll
                for (int k = 0; k < NEG; k++) cn[k] = ld_row4(ctx + (int64_t)tn[k] * D, lane);
#pragma unroll
                for (int k = 0; k < NEG; k++)
#pragma unroll
                    for (int m = 0; m < NEG; m++) stale = stale || (tn[k] == t[m]);
            }
            // ---- pair p (fast_o2, pyx:105-151) ----
            bool anydup = false;
#pragma unroll
            for (int k = 1; k < NEG; k++)
#pragma unroll
                for (int a = 0; a < k; a++) anydup = anydup || (t[a] == t[k]);
            Row4 work = {0.f, 0.f, 0.f, 0.f};  // pyx:126
            {
                const float f = dot128_refblas(x, cpos, quirk);
                if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                    const float g = __fmul_rn(__fmul_rn(1.f - lut[lut_index(f)], lr), lambda);
                    fma_row4(work, g, cpos);  // pyx:146
                    fma_row4(cpos, g, x);     // pyx:147
                }
            }
            if (!anydup) {
                float f[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) f[k] = dot128_refblas(x, c[k], quirk);
#pragma unroll
                for (int k = 0; k < NEG; k++) {
                    if (t[k] == wi) continue;                                // pyx:135-136
                    if (f[k] <= -MAX_EXP_F || f[k] >= MAX_EXP_F) continue;  // pyx:141-142
                    const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f[k])], lr), lambda);
                    fma_row4(work, g, c[k]);
                    fma_row4(c[k], g, x);
                    st_row4(ctx + (int64_t)t[k] * D, lane, c[k]);
                }
            } else {  // equal samples inside one pair: strictly one after the other, re-reading the row
#pragma unroll 1
                for (int k = 0; k < NEG; k++) {
                    const uint32_t tk = __shfl_sync(FULL, tq_cur, k);
                    if (tk == wi) continue;
                    float *cp = ctx + (int64_t)tk * D;
                    Row4 cc = ld_row4(cp, lane);
                    const float f = dot128_refblas(x, cc, quirk);
                    if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                    const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f)], lr), lambda);
                    fma_row4(work, g, cc);
                    fma_row4(cc, g, x);
                    st_row4(cp, lane, cc);
                }
            }
            const Row4 nx = {x.v0 + work.v0, x.v1 + work.v1, x.v2 + work.v2, x.v3 + work.v3};  // pyx:149
            st_row4(node + (int64_t)wj * D, lane, nx);
            if (!nvalid) {
                st_row4(ctx + (int64_t)wi * D, lane, cpos);
                break;
            }
            // ---- hazards, then rotate ----
            if (nwj == wj) xn = nx;
            if (!same) {
                st_row4(ctx + (int64_t)wi * D, lane, cpos);
                cpos = ld_row4(ctx + (int64_t)nwi * D, lane);
            }
            if (stale) {
#pragma unroll
                for (int k = 0; k < NEG; k++) cn[k] = ld_row4(ctx + (int64_t)tn[k] * D, lane);
            }
            x = xn;
#pragma unroll
            for (int k = 0; k < NEG; k++) {
                c[k] = cn[k];
                t[k] = tn[k];
            }
            tq_cur = tq_nxt;
            tq_nxt = tq_new;
            i = ni; j = nj; wi = nwi; wj = nwj;
        }
    }
    if (n_tokens && lane == 0) *n_tokens += tokens;
}

// ---- o2 ORDERED, size == 128: one warp PER TARGET ROW of a pair + a scheduling warp; strictly sequential over pairs ---------
// A single warp replaying the stream is bound by instruction latency (about 570 dependent-ish instructions per pair, no
// other warp to hide them).  Two things are taken off that chain:
//  * Everything that does not depend on the embeddings -- the order of the (centre, context) pairs, the LCG stream, the
//    sampler-table look-ups, and which rows of neighbouring pairs coincide -- is produced by a SCHEDULING warp, 32 pairs
//    at a time (one pair per lane), two chunks ahead of the workers, into a shared-memory ring of pair descriptors.
//  * The NEG+1 targets of one pair are independent of each other unless two samples coincide, so worker warp 0 owns the
//    centre's context row (registers, across the window) and worker r>0 owns sample r-1: phase A (all workers in
//    parallel) dot -> sigma -> g, publish (g, old target row) in shared memory, update + store the target row; one
//    barrier among the workers; phase B (every worker, redundantly, so that each has the new node row for forwarding)
//    accumulates `work` in the reference's order d = 0..NEG and forms x + work; worker 0 stores it.
// Rows are requested two pairs ahead with cp.async into per-warp shared-memory rings (register prefetch does not work
// at this depth: the load scoreboard is a counter, so touching the oldest request waits for the youngest).  A request
// for pair p+2 is issued after barrier p, which makes the stores of pairs <= p visible; a row that pair p+1 writes is
// not requested but read at the start of pair p+2 (hazard bit in the descriptor), node rows of the two pairs in flight
// are forwarded from registers.  A pair with two equal samples is executed by worker 0 alone, one target after the
// other.  Every floating-point operation and its order is the one of o2_ordered_d128_kernel: same bits.
// Requires window <= 15 (a centre's window fits the 32 lanes of the scheduling warp) and disjoint node / ctx tables.
template <int ID, int NTHREADS>
__device__ __forceinline__ void named_barrier() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
}
__device__ __forceinline__ Row4 lds_row4(const float *row, int lane) {
    Row4 v;
    v.v0 = row[lane]; v.v1 = row[lane + 32]; v.v2 = row[lane + 64]; v.v3 = row[lane + 96];
    return v;
}

constexpr uint32_t PD_VALID = 1u, PD_SERIAL = 2u, PD_LAST = 4u, PD_NEWC = 8u;  // pair descriptor flags
constexpr int PD_XSRC_SHIFT = 4;  // 2 bits: 0 = node row from the ring, 1 / 2 = x + work of pair p-1 / p-2
constexpr int PD_HZ_SHIFT = 8;    // bit 8+k: sample k's row is written by pair p-1 -> read it at the start of pair p
constexpr int TEAM_MAX_WINDOW = 15;

template <int NEG>
__global__ void __launch_bounds__(32 * (NEG + 2), 1)
    o2_ordered_d128_team_kernel(float *node, float *ctx, const uint32_t *walks, const int64_t *walk_off,
                                int64_t n_walks, const uint64_t *seeds, uint64_t base_seed, Sampler S, int window,
                                float lr, float lambda, bool quirk, int64_t *n_tokens, const float *g_exp_table) {
    constexpr int D = 128;
    constexpr int NT = NEG + 1;          // worker warps
    constexpr int NALL = 32 * (NT + 1);  // + the scheduling warp
    __shared__ float lut[EXP_TABLE_SIZE];
    __shared__ float4 stage[2][NT][32];  // old target rows of the pair (Row4 layout), or x + work in serial mode
    __shared__ __align__(16) float gs[2][8];   // g of target d (padded to 8: two 128-bit reads)
    __shared__ __align__(16) int used[2][8];   // target d took part (pyx:135-136, 141-142)
    __shared__ __align__(16) float xring[NT][3][D];  // per worker: node row of pairs p, p+1, p+2
    __shared__ __align__(16) float cring[NT][3][D];  // per worker: own sample row of pairs p, p+1, p+2
    // pair descriptors, ring of 3 chunks x 32 pairs, index q = 32 * chunk + slot
    __shared__ __align__(16) uint4 d_hdr[96];       // {flags, wi (centre row), wj (context row), -}
    __shared__ __align__(16) uint32_t d_t[96][8];   // the NEG samples of the pair
    __shared__ uint64_t JA[33], JC[33];  // LCG jump by NEG*k steps
    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += NALL) lut[e] = g_exp_table[e];

    if (r == NT) {
        // =========================== scheduling warp ===========================
        for (int k = lane; k < 33; k += 32) {
            uint64_t A = 1, C = 0;
            for (int q = 0; q < NEG * k; q++) {
                C = (C * LCG_MUL + 11ULL) & LCG_MASK;
                A = (A * LCG_MUL) & LCG_MASK;
            }
            JA[k] = A;
            JC[k] = C;
        }
        __syncwarp();
        int64_t w = -1, tokens = 0;
        const uint32_t *gpath = walks;
        int len = 0, i = 0, j_base = 0;
        uint64_t rnd_cur = 0;
        uint32_t wi_cur = COMEMB_TOKEN_NONE, pending = 0, full = 0;
        bool ended = false;
        int end_chunk = -1;
        auto produce = [&](int c) {
            const int buf = c % 3, q0 = 32 * buf, qprev = 32 * ((buf + 2) % 3);
            d_hdr[q0 + lane].x = 0u;
            __syncwarp();
            int filled = 0;
            while (filled < 32 && !ended) {
                if (pending == 0u) {  // next centre (pyx:494-501)
                    i++;
                    while (i >= len) {
                        w++;
                        if (w >= n_walks) {
                            ended = true;
                            break;
                        }
                        gpath = walks + walk_off[w];
                        len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
                        rnd_cur = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
                        i = 0;
                    }
                    if (ended) break;
                    wi_cur = gpath[i];
                    if (wi_cur == COMEMB_TOKEN_NONE) continue;
                    tokens++;
                    j_base = max(0, i - window);
                    const int jj = j_base + lane;
                    const bool ok = jj < min(len, i + window + 1) && jj != i && gpath[jj] != COMEMB_TOKEN_NONE;
                    full = __ballot_sync(FULL, ok);  // pyx:503-507: the centre's valid contexts, in order
                    pending = full;
                    if (full == 0u) continue;
                }
                const int take = min(__popc(pending), 32 - filled);
                const int rank = __popc(pending & ((1u << lane) - 1u));
                const bool mine = ((pending >> lane) & 1u) && rank < take;
                if (mine) {  // this lane emits pair `filled + rank`: context position j_base + lane
                    const int slot = filled + rank;
                    const uint32_t wj = gpath[j_base + lane];
                    uint32_t fl = PD_VALID;
                    if (lane == __ffs(full) - 1) fl |= PD_NEWC;
                    if (lane == 31 - __clz(full)) fl |= PD_LAST;
                    uint64_t rr = (JA[rank] * rnd_cur + JC[rank]) & LCG_MASK;  // pair's position in the walk's stream
#pragma unroll
                    for (int k = 0; k < NEG; k++) {  // pyx:133-134
                        d_t[q0 + slot][k] = S.table[table_slot(rr, S.mod)];
                        rr = lcg_next(rr);
                    }
                    d_hdr[q0 + slot] = make_uint4(fl, wi_cur, wj, 0u);
                }
                rnd_cur = (JA[take] * rnd_cur + JC[take]) & LCG_MASK;
                pending &= ~__ballot_sync(FULL, mine);
                filled += take;
            }
            if (ended && end_chunk < 0) end_chunk = c;  // this chunk holds the end marker (a descriptor without PD_VALID)
            __syncwarp();
            // relations with the two previous pairs (lane = slot)
            uint32_t fl = d_hdr[q0 + lane].x;
            if (fl & PD_VALID) {
                const int q1 = lane >= 1 ? q0 + lane - 1 : qprev + 31;      // pair p-1
                const int q2 = lane >= 2 ? q0 + lane - 2 : qprev + 30 + lane;  // pair p-2
                const bool has1 = c > 0 || lane >= 1, has2 = c > 0 || lane >= 2;
                uint32_t t[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) t[k] = d_t[q0 + lane][k];
                bool dup = false;
#pragma unroll
                for (int k = 1; k < NEG; k++)
#pragma unroll
                    for (int q = 0; q < k; q++) dup = dup || (t[q] == t[k]);
                if (dup) fl |= PD_SERIAL;
                const uint32_t wj = d_hdr[q0 + lane].z;
                if (has1 && wj == d_hdr[q1].z) fl |= 1u << PD_XSRC_SHIFT;
                else if (has2 && wj == d_hdr[q2].z) fl |= 2u << PD_XSRC_SHIFT;
                if (has1) {
                    const uint32_t pwi = d_hdr[q1].y;
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        bool hz = t[k] == pwi;  // the previous pair may end its centre and store that row
#pragma unroll
                        for (int q = 0; q < NEG; q++) hz = hz || (t[k] == d_t[q1][q]);
                        if (hz) fl |= 1u << (PD_HZ_SHIFT + k);
                    }
                }
                d_hdr[q0 + lane].x = fl;
            }
            __syncwarp();
        };
        i = 0; len = 0; i = -1;  // so that the first `i++ ... i >= len` opens walk 0
        produce(0);
        produce(1);
        named_barrier<0, NALL>();  // chunks 0 and 1 are ready
        for (int c = 0;; c++) {     // workers are in chunk c
            if (end_chunk >= 0 && end_chunk <= c) break;  // they stop inside this chunk: no further chunk barrier
            produce(c + 2);
            named_barrier<0, NALL>();  // workers move from chunk c to c+1
        }
        if (n_tokens && lane == 0) *n_tokens += tokens;
        return;
    }

    // =========================== worker warps ===========================
    const int ks = r > 0 ? r - 1 : 0;  // own target row: worker 0 -> ctx[wi] (cpos), worker r -> ctx[sample r-1]
    float *const xr = &xring[r][0][0], *const cr = &cring[r][0][0];
    const float *const node_l = node + 4 * lane, *const ctx_l = ctx + 4 * lane;
    auto request = [&](int q, int ring) {  // cp.async the rows of the pair described at q into ring slot `ring`
        const uint4 h = d_hdr[q];
        if (h.x & PD_VALID) {
            if (((h.x >> PD_XSRC_SHIFT) & 3u) == 0u) cp_async16(xr + ring * D + 4 * lane, node_l + (int64_t)h.z * D);
            if (r > 0 && !((h.x >> (PD_HZ_SHIFT + ks)) & 1u))
                cp_async16(cr + ring * D + 4 * lane, ctx_l + (int64_t)d_t[q][ks] * D);
        }
        cp_async_commit();  // exactly one group per pair (possibly empty)
    };
    named_barrier<0, NALL>();  // chunks 0 and 1 are ready (also: lut is loaded)
    request(0, 0);
    request(1, 1);
    Row4 cpos = {0.f, 0.f, 0.f, 0.f};
    Row4 nx1 = cpos, nx2 = cpos;  // x + work of pairs p-1, p-2
    int s = 0, p3 = 0, q = 0;  // q: descriptor index of pair p
    for (;;) {
        // ---- start of pair p: its requests have landed (at most the youngest group may be pending) ----
        cp_async_wait<1>();
        __syncwarp();
        const uint4 hdr = d_hdr[q];
        const uint32_t fl = hdr.x;
        if (!(fl & PD_VALID)) break;
        const uint32_t wi0 = hdr.y, wj0 = hdr.z, u0 = d_t[q][ks];
        const uint32_t xsrc = (fl >> PD_XSRC_SHIFT) & 3u;
        Row4 x0;
        if (xsrc == 1u) x0 = nx1;
        else if (xsrc == 2u) x0 = nx2;
        else x0 = lds_row4(xr + p3 * D, lane);
        Row4 c0;
        if (r == 0) {
            if (fl & PD_NEWC) cpos = ld_row4(ctx + (int64_t)wi0 * D, lane);  // after barrier p-1: everything visible
            c0 = cpos;
        } else {
            if ((fl >> (PD_HZ_SHIFT + ks)) & 1u) c0 = ld_row4(ctx + (int64_t)u0 * D, lane);
            else c0 = lds_row4(cr + p3 * D, lane);
        }
        const bool serial = (fl & PD_SERIAL) != 0u, last_of_centre = (fl & PD_LAST) != 0u;
        // ---- phase A ----
        Row4 nx;
        if (!serial) {
            bool use = !(r > 0 && u0 == wi0);  // pyx:135-136
            float g = 0.f;
            if (use) {
                const float f = dot128_refblas(x0, c0, quirk);
                use = f > -MAX_EXP_F && f < MAX_EXP_F;  // pyx:141-142
                if (use) g = __fmul_rn(__fmul_rn((r == 0 ? 1.f : 0.f) - lut[lut_index(f)], lr), lambda);
            }
            if (lane == 0) {
                gs[s][r] = g;
                used[s][r] = use ? 1 : 0;
            }
            if (use) {
                stage[s][r][lane] = make_float4(c0.v0, c0.v1, c0.v2, c0.v3);
                fma_row4(c0, g, x0);  // pyx:147
                if (r > 0) st_row4(ctx + (int64_t)u0 * D, lane, c0);
                else cpos = c0;
            }
            if (r == 0 && last_of_centre) st_row4(ctx + (int64_t)wi0 * D, lane, cpos);
        } else if (r == 0) {
            Row4 work = {0.f, 0.f, 0.f, 0.f};
            {
                const float f = dot128_refblas(x0, cpos, quirk);
                if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                    const float g = __fmul_rn(__fmul_rn(1.f - lut[lut_index(f)], lr), lambda);
                    fma_row4(work, g, cpos);
                    fma_row4(cpos, g, x0);
                }
            }
#pragma unroll 1
            for (int k = 0; k < NEG; k++) {
                const uint32_t tk = d_t[q][k];
                if (tk == wi0) continue;
                float *cp = ctx + (int64_t)tk * D;
                Row4 cc = ld_row4(cp, lane);
                const float f = dot128_refblas(x0, cc, quirk);
                if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f)], lr), lambda);
                fma_row4(work, g, cc);
                fma_row4(cc, g, x0);
                st_row4(cp, lane, cc);
            }
            nx.v0 = x0.v0 + work.v0; nx.v1 = x0.v1 + work.v1; nx.v2 = x0.v2 + work.v2; nx.v3 = x0.v3 + work.v3;
            stage[s][0][lane] = make_float4(nx.v0, nx.v1, nx.v2, nx.v3);
            if (last_of_centre) st_row4(ctx + (int64_t)wi0 * D, lane, cpos);
        }
        named_barrier<1, 32 * NT>();  // barrier p: target-row stores of pair p visible, stage[s] complete
        // ---- phase B: x + work, work accumulated in the reference's order (pyx:146, 149) ----
        if (!serial) {
            Row4 work = {0.f, 0.f, 0.f, 0.f};
            int un[8];
            float gn[8];
            *(int4 *)&un[0] = *(const int4 *)&used[s][0];
            *(float4 *)&gn[0] = *(const float4 *)&gs[s][0];
            if (NT > 4) {
                *(int4 *)&un[4] = *(const int4 *)&used[s][4];
                *(float4 *)&gn[4] = *(const float4 *)&gs[s][4];
            }
#pragma unroll
            for (int m = 0; m < NT; m++) {  // predicated, no branches: all shared loads issue together
                const bool on = un[m] != 0;
                const float g = gn[m];
                const float4 cc = stage[s][m][lane];  // stale data of an older pair when !on: not used
                work.v0 = on ? fmaf(g, cc.x, work.v0) : work.v0; work.v1 = on ? fmaf(g, cc.y, work.v1) : work.v1;
                work.v2 = on ? fmaf(g, cc.z, work.v2) : work.v2; work.v3 = on ? fmaf(g, cc.w, work.v3) : work.v3;
            }
            nx.v0 = x0.v0 + work.v0; nx.v1 = x0.v1 + work.v1; nx.v2 = x0.v2 + work.v2; nx.v3 = x0.v3 + work.v3;
        } else {
            const float4 q = stage[s][0][lane];
            nx.v0 = q.x; nx.v1 = q.y; nx.v2 = q.z; nx.v3 = q.w;
        }
        if (r == 0) st_row4(node + (int64_t)wj0 * D, lane, nx);
        // ---- requests for pair p+2 (stores of pairs <= p are visible; pair p+1's are not) ----
        request(q >= 94 ? q - 94 : q + 2, p3 == 0 ? 2 : p3 - 1);
        // ---- rotate ----
        nx2 = nx1; nx1 = nx;
        s ^= 1;
        p3 = p3 == 2 ? 0 : p3 + 1;
        if ((q & 31) == 31) named_barrier<0, NALL>();  // the scheduling warp has finished chunk c+2; chunk c's buffer is free
        q = q == 95 ? 0 : q + 1;
    }
    cp_async_wait<0>();
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

// ---- o1 ORDERED, size == 128, register-resident and software-pipelined (see o2_ordered_d128_pipe_kernel) -------------------
// The 2+2*NEG rows of edge q+1 are requested before edge q is computed, the samples of edge q+2 before that.  Edge q
// writes rows e0 and e1 only; a prefetched row with one of those indices is re-read after edge q's stores.
template <int NEG>
__global__ void __launch_bounds__(32)
    o1_ordered_d128_pipe_kernel(float *node, const uint32_t *edges, int64_t n_edges, const uint64_t *seeds,
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
    auto directed = [&](const Row4 &x, const Row4 &target, uint32_t word_index, const uint32_t (&tt)[NEG],
                        const Row4 (&c)[NEG]) -> Row4 {  // fast_o1, pyx:205-249
        Row4 work = {0.f, 0.f, 0.f, 0.f};
        {
            const float f = dot128_refblas(x, target, quirk);
            if (f > -MAX_EXP_F && f < MAX_EXP_F) fma_row4(work, __fmul_rn(1.f - lut[lut_index(f)], lr), target);
        }
        float f[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) f[k] = dot128_refblas(x, c[k], quirk);
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            if (tt[k] == word_index) continue;                       // pyx:234-235
            if (f[k] <= -MAX_EXP_F || f[k] >= MAX_EXP_F) continue;  // pyx:240-241
            fma_row4(work, __fmul_rn(0.f - lut[lut_index(f[k])], lr), c[k]);  // pyx:243-245
        }
        const Row4 nx = {x.v0 + work.v0, x.v1 + work.v1, x.v2 + work.v2, x.v3 + work.v3};  // pyx:247
        return nx;
    };
    auto sample = [&](int64_t q) -> uint32_t {  // the 2*NEG samples of edge q, one per lane
        const uint64_t rnd = seeds ? seeds[q] : (splitmix64(base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        return (lane < 2 * NEG) ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
    };
    if (n_edges <= 0) return;
    uint32_t e0 = edges[0], e1 = edges[1];
    uint32_t tq_cur = sample(0);
    uint32_t tq_nxt = n_edges > 1 ? sample(1) : 0u;
    Row4 r0 = ld_row4(node + (int64_t)e0 * D, lane);
    Row4 r1 = ld_row4(node + (int64_t)e1 * D, lane);
    uint32_t ta[NEG], tb[NEG];
    Row4 ca[NEG], cb[NEG];
#pragma unroll
    for (int k = 0; k < NEG; k++) {
        ta[k] = __shfl_sync(FULL, tq_cur, k);
        tb[k] = __shfl_sync(FULL, tq_cur, NEG + k);
        ca[k] = ld_row4(node + (int64_t)ta[k] * D, lane);
        cb[k] = ld_row4(node + (int64_t)tb[k] * D, lane);
    }
    for (int64_t q = 0; q < n_edges; q++) {  // node_embeddings.py:70-71
        const bool has_next = q + 1 < n_edges;
        uint32_t ne0 = 0, ne1 = 0, tq_new = 0;
        uint32_t tna[NEG], tnb[NEG];
        Row4 nr0 = r0, nr1 = r1;
        Row4 nca[NEG], ncb[NEG];
        bool stale = false;
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            tna[k] = __shfl_sync(FULL, tq_nxt, k);
            tnb[k] = __shfl_sync(FULL, tq_nxt, NEG + k);
            nca[k] = ca[k];
            ncb[k] = cb[k];
        }
        if (has_next) {
            ne0 = edges[2 * q + 2];
            ne1 = edges[2 * q + 3];
            if (q + 2 < n_edges) tq_new = sample(q + 2);
            nr0 = ld_row4(node + (int64_t)ne0 * D, lane);
            nr1 = ld_row4(node + (int64_t)ne1 * D, lane);
            stale = ne0 == e0 || ne0 == e1 || ne1 == e0 || ne1 == e1;
#pragma unroll
            for (int k = 0; k < NEG; k++) {
                nca[k] = ld_row4(node + (int64_t)tna[k] * D, lane);
                ncb[k] = ld_row4(node + (int64_t)tnb[k] * D, lane);
                stale = stale || tna[k] == e0 || tna[k] == e1 || tnb[k] == e0 || tnb[k] == e1;
            }
        }
        // ---- edge q ----
        r0 = directed(r0, r1, e1, ta, ca);  // pyx:444 (targets are read-only: a sample equal to e0 is the old row)
        st_row4(node + (int64_t)e0 * D, lane, r0);
        if (e1 == e0) r1 = r0;
#pragma unroll
        for (int k = 0; k < NEG; k++)
            if (tb[k] == e0) cb[k] = r0;  // pyx:447 sees the updated row e0, also as a sample
        r1 = directed(r1, r0, e0, tb, cb);
        st_row4(node + (int64_t)e1 * D, lane, r1);
        if (!has_next) break;
        if (stale) {  // re-read after this edge's stores
            nr0 = ld_row4(node + (int64_t)ne0 * D, lane);
            nr1 = ld_row4(node + (int64_t)ne1 * D, lane);
#pragma unroll
            for (int k = 0; k < NEG; k++) {
                nca[k] = ld_row4(node + (int64_t)tna[k] * D, lane);
                ncb[k] = ld_row4(node + (int64_t)tnb[k] * D, lane);
            }
        }
        e0 = ne0; e1 = ne1; r0 = nr0; r1 = nr1;
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            ta[k] = tna[k]; tb[k] = tnb[k]; ca[k] = nca[k]; cb[k] = ncb[k];
        }
        tq_nxt = tq_new;
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
int launch_o2_flow(float *, float *, int64_t, const uint32_t *, const int64_t *, int64_t, const uint64_t *, uint64_t,
                   const uint32_t *, uint64_t, int, int, float, float, bool, int64_t *, int, cudaStream_t);  // sgns_flow.cu
int launch_o2_ordered(float *node, float *ctx, int64_t n_rows, int size, const uint32_t *walks, const int64_t *walk_off,
                      int64_t n_walks, const uint64_t *seeds, uint64_t base_seed, const uint32_t *table,
                      uint64_t table_len, int window, int negative, float lr, float lambda, bool quirk, int64_t *n_tokens,
                      cudaStream_t st) {
    Sampler S{table, make_table_mod(table_len)};
    const bool disjoint = node + n_rows * size <= ctx || ctx + n_rows * size <= node;
    const int variant = comemb_opts().variant;
    if (size == 128 && disjoint && variant == COMEMB_VARIANT_ORDERED_FLOW) {
        // the same stream as a dataflow graph over many warps (sgns_flow.cu), same bits.  Opt-in: on the walk corpora
        // of BASELINE.json the stream's dependency chain leaves 1.1-1.5x of parallelism and a
        // pair costs one warp 2.9 us against 0.72 us on the one-CTA team below.
        const int rc = launch_o2_flow(node, ctx, n_rows, walks, walk_off, n_walks, seeds, base_seed, table, table_len, window,
                                      negative, lr, lambda, quirk, n_tokens, comemb_opts().max_warps, st);
        if (rc != COMEMB_E_UNSUPPORTED) return rc;
    }
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC && disjoint && !(comemb_opts().variant == COMEMB_VARIANT_ORDERED_PIPE || comemb_opts().variant == COMEMB_VARIANT_ORDERED_PLAIN) &&
        window <= TEAM_MAX_WINDOW) {  // warp per target row + scheduling warp, same bits
        switch (negative) {
#define COMEMB_CASE(N)                                                                                             \
    case N:                                                                                                        \
        o2_ordered_d128_team_kernel<N><<<1, 32 * (N + 2), 0, st>>>(node, ctx, walks, walk_off, n_walks, seeds,     \
                                                                   base_seed, S, window, lr, lambda, quirk,        \
                                                                   n_tokens, comemb_lut_device());                 \
        return (int)cudaGetLastError();
            COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            default: break;
        }
    }
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC && disjoint && comemb_opts().variant != COMEMB_VARIANT_ORDERED_PLAIN) {  // pipelined single warp
        switch (negative) {
#define COMEMB_CASE(N)                                                                                             \
    case N:                                                                                                        \
        CUDA_TRY(cudaFuncSetAttribute(o2_ordered_d128_pipe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      MAX_SENTENCE_LEN * 4));                                                      \
        o2_ordered_d128_pipe_kernel<N><<<1, 32, MAX_SENTENCE_LEN * 4, st>>>(node, ctx, walks, walk_off, n_walks, seeds, base_seed, S, \
                                                         window, lr, lambda, quirk, n_tokens, comemb_lut_device()); \
        return (int)cudaGetLastError();
            COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            default: break;
        }
    }
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC) {  // register-resident fast path, same bits
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
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC && comemb_opts().variant != COMEMB_VARIANT_ORDERED_PLAIN) {  // pipelined fast path, same bits
        switch (negative) {
#define COMEMB_CASE(N)                                                                                              \
    case N:                                                                                                         \
        o1_ordered_d128_pipe_kernel<N><<<1, 32, 0, st>>>(node, edges, n_edges, seeds, base_seed, S, lr, quirk,      \
                                                         comemb_lut_device());                                      \
        return (int)cudaGetLastError();
            COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            default: break;
        }
    }
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC) {  // register-resident fast path, same bits
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
