// fused_sgns.cuh -- the SGNS half of the fused pass for ONE centre of one walk (size 128), shared by the round-synchronous
// and the asynchronous fused kernels (fused_round.cu, fused_async.cu).  One warp; lane l owns the float4 column l of every row.
//
// For every row v of the centre's window, in position order (utils/training_sdg_inner.c:3364-3668):
//   o3 term of x_j: taken from the result buffer (tensor-core result for the row's value at the start of the centre), or --
//     for a node that already occurred at an earlier window position -- computed in-warp from the CURRENT value (fp32 FMAs
//     against the L2-resident inv_cov), so that a walk sees the reference's sequential semantics;
//   SGNS pair on the o2 size-128 code path (LCG jump constants, samples fetched one pair ahead, transposed 8-slot reduction,
//     lane-parallel sigma): g = (label - sigma) * lr (c:1813), context += g*lambda1*x_j unless is_node_embedding
//     (c:1822-1859);
//   combined write  x_j = fma(lambda1, work, x_j) + clip(-lambda2 * w * Y_j, +-0.1*lr)   (c:1870, c:3668).
#pragma once
#include "comemb_common.cuh"

namespace fused {

constexpr int D = 128;
constexpr int INFO_INWARP = 1 << 30;  // info[v]: community (or -1: no o3 term); this bit = o3 in-warp from the current value

struct SgnsArgs {
    float *node, *ctx;
    const uint32_t *table;
    TableMod mod;
    const float *mu, *inv_cov;
    const float *weight;   // top-1 form (per table row) ...
    const float *pi;       // ... or dense [n_rows, K] (then `dense`)
    float *ybuf;           // o3 results: w * Y per slot (dense: sum over communities, cleared by the consumer)
    int K;
    bool dense, o3_on, is_node;
    float lr, lambda1, nl2, clipv;
    bool y_evict_first = false;  // result slots are written once and read once: keep them from displacing table rows in L2
};

__device__ __forceinline__ float clipf(float v, float c) { return fminf(fmaxf(v, -c), c); }

// tokS/infS: this warp's window rows and their info words; lut: sigma table (shared).
// rnd/tnext: the walk's LCG state and the prefetched samples of the next pair (advanced here).
template <bool ATOMIC, int NEG>
__device__ __forceinline__ void sgns_centre(const SgnsArgs &P, const uint32_t wi, const int V, const uint32_t *tokS,
                                            const int32_t *infS, const float *lut, const int64_t slot0,
                                            uint64_t &rnd, uint32_t &tnext, const uint64_t myA, const uint64_t myC,
                                            const int lane) {
    constexpr LcgJump<NEG> J{};
    const bool dense = P.dense, o3_on = P.o3_on, is_node = P.is_node;
    const int K = P.K;
    const float lr = P.lr, lambda1 = P.lambda1, nl2 = P.nl2, clipv = P.clipv;
    float *const node_l = P.node + 4 * lane, *const ctx_l = P.ctx + 4 * lane;
    const int pi_slot = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi_slot == 0 ? 1.f : 0.f;
    float *pos_ptr = ctx_l + (int64_t)wi * D;
    float4 cpos = __ldcg(reinterpret_cast<const float4 *>(pos_ptr));
    float4 dpos = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = 0; v < V; v++) {
        const uint32_t wj = tokS[v];
        float *row1_ptr = node_l + (int64_t)wj * D;
        const float4 r1 = __ldcg(reinterpret_cast<const float4 *>(row1_ptr));
        // ---- o3 term of x_j -----------------------------------------------------------------------------------------
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        float ysc = 1.f;
        bool any = false;
        if (o3_on) {
            const int inf = infS[v];
            if (!(inf & INFO_INWARP) && (dense || inf >= 0)) {  // taken from the tensor-core result
                float *yp = P.ybuf + (slot0 + v) * D + 4 * lane;
                y = P.y_evict_first ? ldcg4_hint(yp, l2_policy_evict_first()) : __ldcg(reinterpret_cast<const float4 *>(yp));
                if (dense)
                    __stcg(reinterpret_cast<float4 *>(yp), make_float4(0.f, 0.f, 0.f, 0.f));
                else
                    ysc = __ldg(P.weight + wj);  // top-1 form: the slot holds Y, the responsibility is applied here
                any = true;
            } else if (inf >= 0 && (inf & INFO_INWARP)) {  // repeated node: from the current value, in-warp
                const int k0 = dense ? 0 : (inf & ~INFO_INWARP), k1 = dense ? K : k0 + 1;
                for (int k = k0; k < k1; k++) {
                    const float p = dense ? __ldg(P.pi + (int64_t)wj * K + k) : __ldg(P.weight + wj);
                    if (p == 0.f) continue;
                    const float4 mk = __ldg(reinterpret_cast<const float4 *>(P.mu + (int64_t)k * D + 4 * lane));
                    const float4 dv = make_float4(r1.x - mk.x, r1.y - mk.y, r1.z - mk.z, r1.w - mk.w);  // diff[4*lane .. +3]
                    const float4 *S = reinterpret_cast<const float4 *>(P.inv_cov + (int64_t)k * D * D) + lane;
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
                    for (int b4 = 0; b4 < 32; b4++) {  // column-major read: operand element (a,b) = S[b*128 + a]; b ascending
                        const float4 s0 = __ldg(S + (4 * b4 + 0) * 32), s1 = __ldg(S + (4 * b4 + 1) * 32);
                        const float4 s2 = __ldg(S + (4 * b4 + 2) * 32), s3 = __ldg(S + (4 * b4 + 3) * 32);
                        const float d0 = __shfl_sync(FULL, dv.x, b4), d1 = __shfl_sync(FULL, dv.y, b4);
                        const float d2 = __shfl_sync(FULL, dv.z, b4), d3 = __shfl_sync(FULL, dv.w, b4);
                        t.x = fmaf(s0.x, d0, t.x); t.y = fmaf(s0.y, d0, t.y); t.z = fmaf(s0.z, d0, t.z); t.w = fmaf(s0.w, d0, t.w);
                        t.x = fmaf(s1.x, d1, t.x); t.y = fmaf(s1.y, d1, t.y); t.z = fmaf(s1.z, d1, t.z); t.w = fmaf(s1.w, d1, t.w);
                        t.x = fmaf(s2.x, d2, t.x); t.y = fmaf(s2.y, d2, t.y); t.z = fmaf(s2.z, d2, t.z); t.w = fmaf(s2.w, d2, t.w);
                        t.x = fmaf(s3.x, d3, t.x); t.y = fmaf(s3.y, d3, t.y); t.z = fmaf(s3.z, d3, t.z); t.w = fmaf(s3.w, d3, t.w);
                    }
                    y.x = fmaf(p, t.x, y.x); y.y = fmaf(p, t.y, y.y);
                    y.z = fmaf(p, t.z, y.z); y.w = fmaf(p, t.w, y.w);
                    any = true;
                }
            }
        }
        // ---- SGNS pair (centre wi, row wj) ---------------------------------------------------------------------------
        if (is_node) cpos = __ldcg(reinterpret_cast<const float4 *>(pos_ptr));  // the "context" table may be the node table
        const uint32_t tmine = tnext;
        tnext = (lane < NEG) ? __ldg(P.table + table_slot((myA * rnd + myC) & LCG_MASK, P.mod)) : 0u;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
        uint32_t tt[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) tt[k] = __shfl_sync(FULL, tmine, k);
        bool anydup = false;
#pragma unroll
        for (int k = 1; k < NEG; k++)
#pragma unroll
            for (int a = 0; a < k; a++) anydup = anydup || (tt[a] == tt[k]);
        float4 work = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!anydup || is_node) {  // no context writes with is_node_embedding: equal samples cannot interact
            float4 c[NEG];
#pragma unroll
            for (int k = 0; k < NEG; k++) c[k] = __ldcg(reinterpret_cast<const float4 *>(ctx_l + (int64_t)tt[k] * D));
            float p[8];
            p[0] = fmaf(r1.w, cpos.w, fmaf(r1.z, cpos.z, fmaf(r1.y, cpos.y, fmaf(r1.x, cpos.x, 0.f))));
#pragma unroll
            for (int k = 0; k < 7; k++)
                p[k + 1] = k < NEG ? fmaf(r1.w, c[k < NEG ? k : 0].w,
                                          fmaf(r1.z, c[k < NEG ? k : 0].z,
                                               fmaf(r1.y, c[k < NEG ? k : 0].y,
                                                    fmaf(r1.x, c[k < NEG ? k : 0].x, 0.f))))
                                   : 0.f;
            const float fm = reduce8_transposed(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], lane);
            bool live = pi_slot == 0;
#pragma unroll
            for (int k = 0; k < NEG; k++) live = live || (pi_slot == k + 1 && tt[k] != wi);
            float gm = 0.f;
            if (live && fm > -MAX_EXP_F && fm < MAX_EXP_F) gm = __fmul_rn(my_label - lut[lut_index(fm)], lr);  // c:1813
            {
                const float gg = __shfl_sync(FULL, gm, lane_of_p(0));
                const float gl = __fmul_rn(gg, lambda1);  // c:1822
                work.x = fmaf(gg, cpos.x, work.x); work.y = fmaf(gg, cpos.y, work.y);
                work.z = fmaf(gg, cpos.z, work.z); work.w = fmaf(gg, cpos.w, work.w);
                if (!is_node) {  // c:1840-1859
                    if (ATOMIC) {
                        dpos.x = fmaf(gl, r1.x, dpos.x); dpos.y = fmaf(gl, r1.y, dpos.y);
                        dpos.z = fmaf(gl, r1.z, dpos.z); dpos.w = fmaf(gl, r1.w, dpos.w);
                    }
                    cpos.x = fmaf(gl, r1.x, cpos.x); cpos.y = fmaf(gl, r1.y, cpos.y);
                    cpos.z = fmaf(gl, r1.z, cpos.z); cpos.w = fmaf(gl, r1.w, cpos.w);
                }
            }
#pragma unroll
            for (int k = 0; k < NEG; k++) {
                const float gg = __shfl_sync(FULL, gm, lane_of_p(k + 1));
                const float gl = __fmul_rn(gg, lambda1);
                work.x = fmaf(gg, c[k].x, work.x); work.y = fmaf(gg, c[k].y, work.y);
                work.z = fmaf(gg, c[k].z, work.z); work.w = fmaf(gg, c[k].w, work.w);
                if (gg != 0.f && !is_node) {
                    float *cp = ctx_l + (int64_t)tt[k] * D;
                    if (ATOMIC)
                        red_add4(cp, make_float4(__fmul_rn(gl, r1.x), __fmul_rn(gl, r1.y), __fmul_rn(gl, r1.z),
                                                 __fmul_rn(gl, r1.w)));
                    else
                        st4(cp, make_float4(fmaf(gl, r1.x, c[k].x), fmaf(gl, r1.y, c[k].y), fmaf(gl, r1.z, c[k].z),
                                            fmaf(gl, r1.w, c[k].w)));
                }
            }
        } else {  // equal samples inside one pair: target by target, re-reading rows
            {
                const float f = warp_sum_xor(
                    fmaf(r1.w, cpos.w, fmaf(r1.z, cpos.z, fmaf(r1.y, cpos.y, fmaf(r1.x, cpos.x, 0.f)))));
                if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                    const float gg = __fmul_rn(1.f - lut[lut_index(f)], lr), gl = __fmul_rn(gg, lambda1);
                    work.x = fmaf(gg, cpos.x, work.x); work.y = fmaf(gg, cpos.y, work.y);
                    work.z = fmaf(gg, cpos.z, work.z); work.w = fmaf(gg, cpos.w, work.w);
                    if (ATOMIC) {
                        dpos.x = fmaf(gl, r1.x, dpos.x); dpos.y = fmaf(gl, r1.y, dpos.y);
                        dpos.z = fmaf(gl, r1.z, dpos.z); dpos.w = fmaf(gl, r1.w, dpos.w);
                    }
                    cpos.x = fmaf(gl, r1.x, cpos.x); cpos.y = fmaf(gl, r1.y, cpos.y);
                    cpos.z = fmaf(gl, r1.z, cpos.z); cpos.w = fmaf(gl, r1.w, cpos.w);
                }
            }
#pragma unroll 1
            for (int k = 0; k < NEG; k++) {
                const uint32_t tkk = __shfl_sync(FULL, tmine, k);
                if (tkk == wi) continue;
                float *cp = ctx_l + (int64_t)tkk * D;
                const float4 c = __ldcg(reinterpret_cast<const float4 *>(cp));
                const float f = warp_sum_xor(fmaf(r1.w, c.w, fmaf(r1.z, c.z, fmaf(r1.y, c.y, fmaf(r1.x, c.x, 0.f)))));
                if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                const float gg = __fmul_rn(0.f - lut[lut_index(f)], lr), gl = __fmul_rn(gg, lambda1);
                work.x = fmaf(gg, c.x, work.x); work.y = fmaf(gg, c.y, work.y);
                work.z = fmaf(gg, c.z, work.z); work.w = fmaf(gg, c.w, work.w);
                if (ATOMIC)
                    red_add4(cp, make_float4(__fmul_rn(gl, r1.x), __fmul_rn(gl, r1.y), __fmul_rn(gl, r1.z),
                                             __fmul_rn(gl, r1.w)));
                else
                    st4(cp, make_float4(fmaf(gl, r1.x, c.x), fmaf(gl, r1.y, c.y), fmaf(gl, r1.z, c.z),
                                        fmaf(gl, r1.w, c.w)));
            }
        }
        // combined write: x_j = fma(lambda1, work, x_j) + work_o3   (c:1870, c:3668); work_o3 = clip(-lambda2 * w * Y)
        float4 o3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (any)
            o3 = make_float4(clipf(__fmul_rn(nl2, __fmul_rn(ysc, y.x)), clipv), clipf(__fmul_rn(nl2, __fmul_rn(ysc, y.y)), clipv),
                             clipf(__fmul_rn(nl2, __fmul_rn(ysc, y.z)), clipv), clipf(__fmul_rn(nl2, __fmul_rn(ysc, y.w)), clipv));
        if (ATOMIC)
            red_add4(row1_ptr, make_float4(fmaf(lambda1, work.x, o3.x), fmaf(lambda1, work.y, o3.y),
                                           fmaf(lambda1, work.z, o3.z), fmaf(lambda1, work.w, o3.w)));
        else
            st4(row1_ptr, make_float4(fmaf(lambda1, work.x, r1.x) + o3.x, fmaf(lambda1, work.y, r1.y) + o3.y,
                                      fmaf(lambda1, work.z, r1.z) + o3.z, fmaf(lambda1, work.w, r1.w) + o3.w));
    }
    if (!is_node) {
        if (ATOMIC)
            red_add4(pos_ptr, dpos);
        else
            st4(pos_ptr, cpos);
    }
}

}  // namespace fused
