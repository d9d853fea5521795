// fused_sg.cu -- HOGWILD form of the legacy fused pass (stale train_sg: per window pair, the o3 community gradient of
// x_j from its current value, the SGNS pair update, and the combined write; utils/training_sdg_inner.c:1597-1905,
// 2520-2715, 2988-3740).  The ORDERED form lives in sgns_ordered.cu.
//
// One warp = one walk (sequential inside, walks race).  Per pair:
//   (1) work_o3[a] = clip(-lambda2 * sum_k sum_b (pi[j,k] * inv_cov[k][b][a]) * (x_j[b] - mu_k[b]), +-0.1*lr)
//       -- the reference's sgemm reads inv_cov column-major, i.e. uses its transpose; lanes own the output coordinates
//       a of their float4 slices, so the 32 lanes read 512 contiguous bytes of inv_cov for every b (coalesced, L2
//       resident: K*d*d*4 B); x_j is staged in shared memory; exactly-zero pi[j,k] are skipped;
//   (2) SGNS targets one after the other (every target re-reads its row: duplicates see each other's update),
//       g = (label - sigma)*lr, context += g*lambda1*x_j unless is_node_embedding;
//   (3) x_j = fma(lambda1, work, x_j) + work_o3.
// Per pair this costs 2*K_live*d^2 flop and K_live*d*d*4 bytes of L2 reads against 7168 B for the SGNS part: the
// fused pass is bound by the o3 term (FP64-FMA / L2), not by HBM.
#include "comemb_common.cuh"

int64_t hogwild_get_max_warps();

namespace {

constexpr int WARPS = 8;

struct SgParams {
    float *node, *negemb;
    int d;
    const uint32_t *walks;
    const int64_t *walk_off;
    int64_t n_walks;
    const int32_t *rw;
    const uint64_t *seeds;
    uint64_t base_seed;
    const uint32_t *table;
    TableMod mod;
    const float *mu, *inv_cov, *pi;
    int K, window, negative;
    float lr, lambda1, lambda2;
    int is_node_embedding;
    const float *glut;
};

template <int NCH, bool ATOMIC>
__global__ void __launch_bounds__(WARPS * 32) sg_fused_hogwild_kernel(const SgParams P) {
    extern __shared__ float smem[];
    float *lut = smem;                                                       // [1000]
    float *xs = smem + EXP_TABLE_SIZE + (size_t)(threadIdx.x >> 5) * P.d;    // this warp's copy of x_j
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int d = P.d, W = P.window, K = P.K;
    const float clipv = __double2float_rn(__dmul_rn((double)P.lr, 0.1));  // c:2556
    const float nl2 = -P.lambda2;                                         // c:3132
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS;

    for (int64_t w = warp0; w < P.n_walks; w += n_warps) {
        const uint32_t *path = P.walks + P.walk_off[w];
        const int32_t *rw = P.rw ? P.rw + P.walk_off[w] : nullptr;
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, P.walk_off[w + 1] - P.walk_off[w]);
        uint64_t rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        for (int i = 0; i < len; i++) {
            const uint32_t wi = __ldg(path + i);
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int r = rw ? rw[i] : 0;
            const int j1 = min(len, i + W + 1 - r);
            for (int j = max(0, i - W + r); j < j1; j++) {
                const uint32_t wj = __ldg(path + j);
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1_ptr = P.node + (int64_t)wj * d;
                float x[4 * NCH], work[4 * NCH], wo3[4 * NCH];
#pragma unroll
                for (int m = 0; m < NCH; m++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int e = 128 * m + 4 * lane + c;
                        x[4 * m + c] = e < d ? __ldcg(row1_ptr + e) : 0.f;
                        work[4 * m + c] = 0.f;
                        wo3[4 * m + c] = 0.f;
                    }
                // (1) o3 gradient
                if (nl2 != 0.f) {
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < NCH; m++)
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const int e = 128 * m + 4 * lane + c;
                            if (e < d) xs[e] = x[4 * m + c];
                        }
                    __syncwarp();
                    double acc[4 * NCH];
#pragma unroll
                    for (int q = 0; q < 4 * NCH; q++) acc[q] = 0.0;
                    for (int k = 0; k < K; k++) {
                        const float p = P.pi[(int64_t)wj * K + k];
                        if (p == 0.f) continue;  // contributes exactly 0
                        const float *S = P.inv_cov + (int64_t)k * d * d;
                        const float *mu = P.mu + (int64_t)k * d;
                        double t[4 * NCH];
#pragma unroll
                        for (int q = 0; q < 4 * NCH; q++) t[q] = 0.0;
                        for (int b = 0; b < d; b++) {
                            const double df = (double)(xs[b] - __ldg(mu + b));
#pragma unroll
                            for (int m = 0; m < NCH; m++)
#pragma unroll
                                for (int c = 0; c < 4; c++) {
                                    const int a = 128 * m + 4 * lane + c;
                                    if (a < d)  // column-major read: operand element (a,b) = S[b*d + a]
                                        t[4 * m + c] = __dadd_rn(
                                            t[4 * m + c], __dmul_rn((double)__fmul_rn(p, S[(int64_t)b * d + a]), df));
                                }
                        }
#pragma unroll
                        for (int q = 0; q < 4 * NCH; q++) acc[q] = __dadd_rn(acc[q], t[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 4 * NCH; q++) {
                        const float v = __fmul_rn(nl2, __double2float_rn(acc[q]));
                        wo3[q] = v < -clipv ? -clipv : (v > clipv ? clipv : v);
                    }
                }
                // (2) SGNS, target by target
                for (int dd = 0; dd < P.negative + 1; dd++) {
                    uint32_t target;
                    float label;
                    if (dd == 0) {
                        target = wi;
                        label = 1.f;
                    } else {
                        target = __ldg(P.table + table_slot(rnd, P.mod));
                        rnd = lcg_next(rnd);
                        if (target == wi) continue;
                        label = 0.f;
                    }
                    float *cp = P.negemb + (int64_t)target * d;
                    float c[4 * NCH];
                    float part = 0.f;
#pragma unroll
                    for (int m = 0; m < NCH; m++)
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int e = 128 * m + 4 * lane + q;
                            // when the "context" table is the node table, a target equal to x_j's row is x_j itself
                            c[4 * m + q] = e < d ? ((P.is_node_embedding && target == wj) ? x[4 * m + q] : __ldcg(cp + e)) : 0.f;
                            part = fmaf(x[4 * m + q], c[4 * m + q], part);
                        }
                    const float f = warp_sum_xor(part);
                    if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;
                    const float g = __fmul_rn(label - lut[lut_index(f)], P.lr);  // c:1813
                    const float gl = __fmul_rn(g, P.lambda1);                    // c:1822
#pragma unroll
                    for (int q = 0; q < 4 * NCH; q++) work[q] = fmaf(g, c[q], work[q]);
                    if (!P.is_node_embedding) {  // c:1840-1859
#pragma unroll
                        for (int m = 0; m < NCH; m++)
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int e = 128 * m + 4 * lane + q;
                                if (e < d) {
                                    if (ATOMIC)
                                        atomicAdd(cp + e, __fmul_rn(gl, x[4 * m + q]));
                                    else
                                        cp[e] = fmaf(gl, x[4 * m + q], c[4 * m + q]);
                                }
                            }
                    }
                }
                // (3) combined write: x_j = fma(lambda1, work, x_j) + work_o3   (c:1870, c:3668)
#pragma unroll
                for (int m = 0; m < NCH; m++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int e = 128 * m + 4 * lane + q;
                        if (e < d) {
                            const float nx = fmaf(P.lambda1, work[4 * m + q], x[4 * m + q]) + wo3[4 * m + q];
                            if (ATOMIC)
                                atomicAdd(row1_ptr + e, nx - x[4 * m + q]);
                            else
                                row1_ptr[e] = nx;
                        }
                    }
            }
        }
    }
}

template <int NCH>
int launch_t(const SgParams &P, bool atomic, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (P.n_walks + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)sms * 4;
    if (hogwild_get_max_warps() > 0 && (hogwild_get_max_warps() + WARPS - 1) / WARPS < cap)
        cap = (hogwild_get_max_warps() + WARPS - 1) / WARPS;
    const int grid = (int)(want < cap ? want : cap);
    const size_t smem = (EXP_TABLE_SIZE + (size_t)WARPS * P.d) * sizeof(float);
    if (atomic)
        sg_fused_hogwild_kernel<NCH, true><<<grid, WARPS * 32, smem, st>>>(P);
    else
        sg_fused_hogwild_kernel<NCH, false><<<grid, WARPS * 32, smem, st>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_sg_fused_hogwild(float *node, float *negemb, int size, const uint32_t *walks, const int64_t *walk_off,
                            int64_t n_walks, const int32_t *reduced_windows, const uint64_t *seeds, uint64_t base_seed,
                            const uint32_t *table, uint64_t table_len, const float *mu, const float *inv_cov,
                            const float *pi, int K, int window, int negative, float lr, float lambda1, float lambda2,
                            int is_node_embedding, bool atomic, cudaStream_t st) {
    if (size > 512) return COMEMB_E_UNSUPPORTED;
    if (n_walks == 0) return 0;
    SgParams P;
    P.node = node; P.negemb = negemb; P.d = size; P.walks = walks; P.walk_off = walk_off; P.n_walks = n_walks;
    P.rw = reduced_windows; P.seeds = seeds; P.base_seed = base_seed; P.table = table; P.mod = make_table_mod(table_len);
    P.mu = mu; P.inv_cov = inv_cov; P.pi = pi; P.K = K; P.window = window; P.negative = negative;
    P.lr = lr; P.lambda1 = lambda1; P.lambda2 = lambda2; P.is_node_embedding = is_node_embedding;
    P.glut = comemb_lut_device();
    if (size <= 128) return launch_t<1>(P, atomic, st);
    if (size <= 256) return launch_t<2>(P, atomic, st);
    return launch_t<4>(P, atomic, st);
}
