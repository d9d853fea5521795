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


int fused_pi_to_top1(const float *, int64_t, int, int32_t *, float *, int *, cudaStream_t);
int launch_sg_fused_async(float *, float *, const uint32_t *, const int64_t *, int64_t, const int32_t *, const uint64_t *,
                          uint64_t, const uint32_t *, uint64_t, const float *, const float *, int, int, int, float, float,
                          float, int, bool, const int32_t *, const float *, cudaStream_t);
int launch_sg_fused_round(float *, float *, const uint32_t *, const int64_t *, int64_t, const int32_t *, const uint64_t *,
                          uint64_t, const uint32_t *, uint64_t, const float *, const float *, const float *, int, int, int,
                          float, float, float, int, bool, int64_t, const int32_t *, const float *, cudaStream_t);

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

// ---- fast fused kernel: size 128, one-hot pi, context table != node table, 2*window <= 24 -------------------------------
// The generic kernel above does one 128x128 mat-vec per pair against an inv_cov block it re-reads from L2 every time
// (64 KB per pair: ~9x the bytes of the SGNS part).  Here the o3 term of a whole centre window is computed at once:
//   * x_j of all <= 24 window positions is staged in shared memory as diff_j = x_j - mu_{c(j)} (pi is one-hot:
//     every node has one community c(j) and weight w(j); rows with pi == 0 have no o3 term);
//   * positions are grouped by community; for each group Y = inv_cov_c^T . [diff_j]  (128 x 128 by 128 x m) runs on
//     the tensor cores with mma.sync m16n8k8 TF32 (fp32 accumulate), A fragments streamed from L2 once per 16
//     vectors, B fragments from the padded (conflict-free) staging rows; Y overwrites the staging rows;
//   * the SGNS part of each pair is the o2 d=128 code path (LCG jump constants, one-pair-ahead sample fetch,
//     transposed 8-slot reduction, lane-parallel sigma), with g = (label-sigma)*lr, context += g*lambda1*x_j, and the
//     final write  x_j = fma(lambda1, work, x_j) + clip(-lambda2 * w(j) * Y_j, +-0.1*lr).
// Differences to the sequential reference, both inside Hogwild tolerance: the o3 term of a window is computed from
// the x_j at the start of the centre (identical unless one node occupies two window positions), and inv_cov / diff
// enter the tensor cores as TF32 (relative error ~1e-3 of a term that is clipped to 0.1*lr).
constexpr int FW = 6;         // warps per block: 2 blocks/SM = 12 warps (13.5 KB staging per warp, <= 168 registers)
constexpr int VMAX = 24;      // window positions held per centre
constexpr int DSTRIDE = 132;  // floats per staging row: 128 + 4 pad -> B-fragment loads hit 32 distinct banks

__device__ __forceinline__ uint32_t to_tf32(float x) {  // round-to-nearest TF32 (feeding raw fp32 bits would truncate)
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// m16n8k8 TF32, fp32 accumulate.  A = {(row g,k t),(row g+8,k t),(row g,k t+4),(row g+8,k t+4)}, B = {(k t,col g),(k t+4,col g)}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4 &a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
          "r"(b0), "r"(b1));
}

// inv_cov [K][b][a] -> TF32-rounded, fragment-major copy: out[c][kt][mt][lane][e], lane = 4g+t, with
//   e0 = S[8kt+t  ][16g+2mt]   e1 = S[8kt+t  ][16g+2mt+1]     (MMA rows g and g+8 of tile mt <-> outputs a, a+1)
//   e2 = S[8kt+t+4][16g+2mt]   e3 = S[8kt+t+4][16g+2mt+1]
// i.e. tile mt's row r stands for output coordinate a = 16*(r%8) + 2*mt + r/8, so that a lane ends up owning 16
// consecutive outputs per vector (vectorised write-back).
__global__ void tile_inv_cov_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 3), lane = (int)((i >> 2) & 31), mt = (int)((i >> 7) & 7), kt = (int)((i >> 10) & 15);
        const int64_t c = i >> 14;
        const int g = lane >> 2, t = lane & 3;
        const int bq = 8 * kt + t + ((e >> 1) ? 4 : 0), aq = 16 * g + 2 * mt + (e & 1);
        out[i] = __uint_as_float(to_tf32(in[c * 16384 + (int64_t)bq * 128 + aq]));
    }
}

// per row: the single community with pi != 0 (or -1) and its weight; *not_onehot |= 1 if a row has more than one
__global__ void pi_dominant_kernel(const float *__restrict__ pi, int64_t n, int K, int32_t *__restrict__ comm,
                                   float *__restrict__ weight, int *not_onehot) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int cnt = 0, idx = -1;
    float w = 0.f;
    for (int k = 0; k < K; k++) {
        const float p = pi[r * K + k];
        if (p != 0.f) {
            cnt++;
            idx = k;
            w = p;
        }
    }
    comm[r] = cnt == 1 ? idx : -1;
    weight[r] = cnt == 1 ? w : 0.f;
    if (cnt > 1) atomicOr(not_onehot, 1);
}

struct SgFastParams {
    float *node, *ctx;
    const uint32_t *walks;
    const int64_t *walk_off;
    int64_t n_walks;
    const int32_t *rw;
    const uint64_t *seeds;
    uint64_t base_seed;
    const uint32_t *table;
    TableMod mod;
    const float *mu, *inv_cov;
    const int32_t *comm;
    const float *weight;
    int window;
    float lr, lambda1, lambda2;
    const float *glut;
};

template <bool ATOMIC, int NEG>
__global__ void __launch_bounds__(FW * 32, 2) sg_fused_d128_kernel(const SgFastParams P) {
    constexpr int D = 128;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    extern __shared__ float dyn[];  // per warp: stage[(VMAX+1)*DSTRIDE] | tok[VMAX] | comm[VMAX] | wgt[VMAX]
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    const int lane = threadIdx.x & 31;
    constexpr int PER_WARP = (VMAX + 1) * DSTRIDE + 3 * VMAX;
    float *stage = dyn + (size_t)(threadIdx.x >> 5) * PER_WARP;
    uint32_t *tokS = reinterpret_cast<uint32_t *>(stage + (VMAX + 1) * DSTRIDE);
    int32_t *commS = reinterpret_cast<int32_t *>(tokS + VMAX);
    float *wgtS = reinterpret_cast<float *>(commS + VMAX);
    for (int e = lane; e < DSTRIDE; e += 32) stage[VMAX * DSTRIDE + e] = 0.f;  // the all-zero padding row
    __syncthreads();
    const int W = P.window;
    const float lr = P.lr, lambda1 = P.lambda1;
    const float clipv = __double2float_rn(__dmul_rn((double)P.lr, 0.1));  // c:2556
    const float nl2 = -P.lambda2;                                         // c:3132
    float *const node_l = P.node + 4 * lane, *const ctx_l = P.ctx + 4 * lane;
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int pi_slot = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi_slot == 0 ? 1.f : 0.f;
    const int g = lane >> 2, t = lane & 3;  // mma fragment coordinates
    const int64_t warp0 = (int64_t)blockIdx.x * FW + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * FW;

    for (int64_t w = warp0; w < P.n_walks; w += n_warps) {
        const uint32_t *path = P.walks + P.walk_off[w];
        const int32_t *rw = P.rw ? P.rw + P.walk_off[w] : nullptr;
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, P.walk_off[w + 1] - P.walk_off[w]);
        uint64_t rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        uint32_t tnext = (lane < NEG) ? __ldg(P.table + table_slot((myA * rnd + myC) & LCG_MASK, P.mod)) : 0u;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;

        for (int i = 0; i < len; i++) {
            const uint32_t wi = __ldg(path + i);
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int r = rw ? rw[i] : 0;
            const int ja = max(0, i - W + r), jb = min(len, i + W + 1 - r);
            // window slots: lane l looks at position ja + l (the window spans <= 2W+1 <= 25 positions)
            const int jl = ja + lane;
            uint32_t tk = COMEMB_TOKEN_NONE;
            if (jl < jb && jl != i) tk = __ldg(path + jl);
            const bool valid = tk != COMEMB_TOKEN_NONE;
            const unsigned vm = __ballot_sync(FULL, valid);
            const int V = __popc(vm);
            if (V == 0) continue;
            const int slot = __popc(vm & ((1u << lane) - 1u));
            const int my_comm = (valid && nl2 != 0.f) ? __ldg(P.comm + tk) : -1;
            __syncwarp();
            if (valid) {
                tokS[slot] = tk;
                commS[slot] = my_comm;
                wgtS[slot] = my_comm >= 0 ? __ldg(P.weight + tk) : 0.f;
            }
            __syncwarp();
            // ---- (1) o3 term of the whole window -----------------------------------------------------------------------
            if (nl2 != 0.f) {
                for (int v = 0; v < V; v++) {  // stage diff_v = x_v - mu_{c(v)}
                    const int cv = commS[v];
                    if (cv < 0) continue;
                    const float4 x = __ldcg(reinterpret_cast<const float4 *>(node_l + (int64_t)tokS[v] * D));
                    const float4 m = __ldg(reinterpret_cast<const float4 *>(P.mu + (int64_t)cv * D + 4 * lane));
                    *reinterpret_cast<float4 *>(stage + v * DSTRIDE + 4 * lane) =
                        make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
                }
                __syncwarp();
                unsigned remaining = __ballot_sync(FULL, my_comm >= 0);
                while (remaining) {  // one pass per community present in the window
                    const int leader = __ffs(remaining) - 1;
                    const int c = __shfl_sync(FULL, my_comm, leader);
                    const bool member = my_comm == c;
                    remaining &= ~__ballot_sync(FULL, member);
                    const unsigned ms = __reduce_or_sync(FULL, member ? (1u << slot) : 0u);  // member SLOTS
                    const int m = __popc(ms);
                    const float *Sc = P.inv_cov + (int64_t)c * D * D;
                    for (int base = 0; base < m; base += 16) {  // two n-tiles (16 vectors) per sweep over inv_cov_c
                        const int i0 = base + g, i1 = base + 8 + g;
                        const float *d0 = stage + (i0 < m ? (int)__fns(ms, 0, i0 + 1) : VMAX) * DSTRIDE;
                        const float *d1 = stage + (i1 < m ? (int)__fns(ms, 0, i1 + 1) : VMAX) * DSTRIDE;
                        const bool two = (m - base) > 8;
                        float acc0[8][4], acc1[8][4];
#pragma unroll
                        for (int mt = 0; mt < 8; mt++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc0[mt][q] = acc1[mt][q] = 0.f;
                        // operand A = inv_cov_c^T (A[a][b] = S[b][a], the reference's column-major read) comes from the
                        // FRAGMENT-MAJOR, TF32-rounded copy built by tile_inv_cov_kernel: for every (kt, mt) the 32 lanes'
                        // operand quads are 512 consecutive bytes -> one fully coalesced 128-bit load per MMA, landing
                        // directly in its operand register quad.  The 8 quads of slice kt+1 are requested before the MMAs
                        // of slice kt run (two register sets that swap by name).
                        const float4 *af = reinterpret_cast<const float4 *>(Sc) + lane;
                        float4 aA[8], aB[8];
                        auto request = [&](float4 (&a)[8], int kt) {
#pragma unroll
                            for (int mt = 0; mt < 8; mt++) a[mt] = __ldg(af + (size_t)kt * 256 + mt * 32);
                        };
                        auto sweep = [&](const float4 (&a)[8], int kt) {
                            const int b0 = 8 * kt;
                            const uint32_t bb0 = to_tf32(d0[b0 + t]), bb1 = to_tf32(d0[b0 + t + 4]);
                            const uint32_t cc0 = to_tf32(d1[b0 + t]), cc1 = to_tf32(d1[b0 + t + 4]);
#pragma unroll
                            for (int mt = 0; mt < 8; mt++) {
                                mma_tf32(acc0[mt], a[mt], bb0, bb1);
                                if (two) mma_tf32(acc1[mt], a[mt], cc0, cc1);
                            }
                        };
                        request(aA, 0);
#pragma unroll 1
                        for (int kt = 0; kt < 16; kt += 2) {
                            request(aB, kt + 1);
                            sweep(aA, kt);
                            if (kt + 2 < 16) request(aA, kt + 2);
                            sweep(aB, kt + 1);
                        }
                        __syncwarp();  // every lane is done reading these members' diff rows
                        // C fragment of tile mt: c0:(a=16g+2mt, col 2t) c1:(same a, col 2t+1) c2:(a+1, col 2t) c3:(a+1, col 2t+1)
                        // -> per column this lane owns the 16 consecutive outputs a = 16g .. 16g+15
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if (h == 1 && !two) break;
                            const int e0 = base + 8 * h + 2 * t, e1 = e0 + 1;
                            float *y0 = e0 < m ? stage + (int)__fns(ms, 0, e0 + 1) * DSTRIDE + 16 * g : nullptr;
                            float *y1 = e1 < m ? stage + (int)__fns(ms, 0, e1 + 1) * DSTRIDE + 16 * g : nullptr;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const float *f0 = h == 0 ? acc0[2 * q] : acc1[2 * q];
                                const float *f1 = h == 0 ? acc0[2 * q + 1] : acc1[2 * q + 1];
                                if (y0) *reinterpret_cast<float4 *>(y0 + 4 * q) = make_float4(f0[0], f0[2], f1[0], f1[2]);
                                if (y1) *reinterpret_cast<float4 *>(y1 + 4 * q) = make_float4(f0[1], f0[3], f1[1], f1[3]);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            // ---- (2) SGNS pairs of the window, in position order -------------------------------------------------------
            float *pos_ptr = ctx_l + (int64_t)wi * D;
            float4 cpos = __ldcg(reinterpret_cast<const float4 *>(pos_ptr));
            float4 dpos = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int v = 0; v < V; v++) {
                const uint32_t wj = tokS[v];
                float *row1_ptr = node_l + (int64_t)wj * D;
                const float4 r1 = __ldcg(reinterpret_cast<const float4 *>(row1_ptr));
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
                if (!anydup) {
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
                        if (ATOMIC) {
                            dpos.x = fmaf(gl, r1.x, dpos.x); dpos.y = fmaf(gl, r1.y, dpos.y);
                            dpos.z = fmaf(gl, r1.z, dpos.z); dpos.w = fmaf(gl, r1.w, dpos.w);
                        }
                        cpos.x = fmaf(gl, r1.x, cpos.x); cpos.y = fmaf(gl, r1.y, cpos.y);
                        cpos.z = fmaf(gl, r1.z, cpos.z); cpos.w = fmaf(gl, r1.w, cpos.w);
                    }
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        const float gg = __shfl_sync(FULL, gm, lane_of_p(k + 1));
                        const float gl = __fmul_rn(gg, lambda1);
                        work.x = fmaf(gg, c[k].x, work.x); work.y = fmaf(gg, c[k].y, work.y);
                        work.z = fmaf(gg, c[k].z, work.z); work.w = fmaf(gg, c[k].w, work.w);
                        float *cp = ctx_l + (int64_t)tt[k] * D;
                        if (gg != 0.f) {
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
                // (3) combined write: x_j = fma(lambda1, work, x_j) + clip(-lambda2 * w * Y_j)   (c:1870, c:3668)
                float4 o3 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nl2 != 0.f && commS[v] >= 0) {
                    const float4 y = *reinterpret_cast<const float4 *>(stage + v * DSTRIDE + 4 * lane);
                    const float sc = wgtS[v];
                    o3.x = fminf(fmaxf(__fmul_rn(nl2, __fmul_rn(sc, y.x)), -clipv), clipv);
                    o3.y = fminf(fmaxf(__fmul_rn(nl2, __fmul_rn(sc, y.y)), -clipv), clipv);
                    o3.z = fminf(fmaxf(__fmul_rn(nl2, __fmul_rn(sc, y.z)), -clipv), clipv);
                    o3.w = fminf(fmaxf(__fmul_rn(nl2, __fmul_rn(sc, y.w)), -clipv), clipv);
                }
                if (ATOMIC)
                    red_add4(row1_ptr, make_float4(fmaf(lambda1, work.x, o3.x), fmaf(lambda1, work.y, o3.y),
                                                   fmaf(lambda1, work.z, o3.z), fmaf(lambda1, work.w, o3.w)));
                else
                    st4(row1_ptr, make_float4(fmaf(lambda1, work.x, r1.x) + o3.x, fmaf(lambda1, work.y, r1.y) + o3.y,
                                              fmaf(lambda1, work.z, r1.z) + o3.z, fmaf(lambda1, work.w, r1.w) + o3.w));
            }
            if (ATOMIC)
                red_add4(pos_ptr, dpos);
            else
                st4(pos_ptr, cpos);
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
    if (comemb_opts().max_warps > 0 && (comemb_opts().max_warps + WARPS - 1) / WARPS < cap)
        cap = (comemb_opts().max_warps + WARPS - 1) / WARPS;
    const int grid = (int)(want < cap ? want : cap);
    const size_t smem = (EXP_TABLE_SIZE + (size_t)WARPS * P.d) * sizeof(float);
    if (atomic)
        sg_fused_hogwild_kernel<NCH, true><<<grid, WARPS * 32, smem, st>>>(P);
    else
        sg_fused_hogwild_kernel<NCH, false><<<grid, WARPS * 32, smem, st>>>(P);
    return (int)cudaGetLastError();
}

template <int NEG>
int launch_fast_t(const SgFastParams &F, bool atomic, cudaStream_t st) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)FW * ((VMAX + 1) * DSTRIDE + 3 * VMAX) * sizeof(float);
    auto launch = [&](auto kernel) -> int {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FW * 32, smem);
        if (per_sm < 1) per_sm = 1;
        int64_t cap = (int64_t)sms * per_sm;
        if (comemb_opts().max_warps > 0 && (comemb_opts().max_warps + FW - 1) / FW < cap)
            cap = (comemb_opts().max_warps + FW - 1) / FW;
        const int64_t want = (F.n_walks + FW - 1) / FW;
        kernel<<<(int)(want < cap ? want : cap), FW * 32, smem, st>>>(F);
        return (int)cudaGetLastError();
    };
    return atomic ? launch(sg_fused_d128_kernel<true, NEG>) : launch(sg_fused_d128_kernel<false, NEG>);
}

int launch_fast(const SgFastParams &F, int negative, bool atomic, cudaStream_t st) {
    switch (negative) {
        case 3: return launch_fast_t<3>(F, atomic, st);
        case 4: return launch_fast_t<4>(F, atomic, st);
        default: return launch_fast_t<5>(F, atomic, st);
    }
}

}  // namespace

int launch_sg_fused_hogwild(float *node, float *negemb, int size, const uint32_t *walks, const int64_t *walk_off,
                            int64_t n_walks, const int32_t *reduced_windows, const uint64_t *seeds, uint64_t base_seed,
                            const uint32_t *table, uint64_t table_len, const float *mu, const float *inv_cov,
                            const float *pi, int K, int window, int negative, float lr, float lambda1, float lambda2,
                            int is_node_embedding, bool atomic, int64_t n_rows, const int32_t *top1_comm,
                            const float *top1_weight, cudaStream_t st) {
    // n_rows: rows of pi (the one-hot scan).  top1_comm / top1_weight: pi in top-1 form supplied by the caller
    // (comemb_sg_fused_top1; no dense pi then).
    if (size > 512) return COMEMB_E_UNSUPPORTED;
    if (n_walks == 0) return 0;
    SgParams P;
    P.node = node; P.negemb = negemb; P.d = size; P.walks = walks; P.walk_off = walk_off; P.n_walks = n_walks;
    P.rw = reduced_windows; P.seeds = seeds; P.base_seed = base_seed; P.table = table; P.mod = make_table_mod(table_len);
    P.mu = mu; P.inv_cov = inv_cov; P.pi = pi; P.K = K; P.window = window; P.negative = negative;
    P.lr = lr; P.lambda1 = lambda1; P.lambda2 = lambda2; P.is_node_embedding = is_node_embedding;
    P.glut = comemb_lut_device();
    const bool top1 = top1_comm != nullptr;  // the caller already holds pi in top-1 form (and no dense pi)
    const int variant = comemb_opts().variant;
    if (size == 128 && (variant == COMEMB_VARIANT_DEFAULT || variant == COMEMB_VARIANT_TENSOR ||
                        variant == COMEMB_VARIANT_ROUNDSYNC)) {
        // tcgen05 kernels.  pi in top-1 form (given, or a dense pi that turns out one-hot): the asynchronous kernel
        // (fused_async.cu); dense pi, lambda2 == 0 or COMEMB_VARIANT_ROUNDSYNC: the round-synchronous one (fused_round.cu).
        if (lambda2 != 0.f && variant != COMEMB_VARIANT_ROUNDSYNC && (top1 || (pi && n_rows > 0))) {
            const int32_t *c1 = top1_comm;
            const float *w1 = top1_weight;
            char *tmp = nullptr;
            if (!top1) {
                int h_flag = 1;
                CUDA_TRY(cudaMallocAsync(&tmp, 16 + (size_t)n_rows * 8, st));
                int32_t *cc = reinterpret_cast<int32_t *>(tmp + 16);
                float *ww = reinterpret_cast<float *>(tmp + 16 + (size_t)n_rows * 4);
                cudaError_t e = cudaMemsetAsync(tmp, 0, 16, st);
                if (e == cudaSuccess) e = (cudaError_t)fused_pi_to_top1(pi, n_rows, K, cc, ww, reinterpret_cast<int *>(tmp), st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, tmp, 4, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) {
                    cudaFreeAsync(tmp, st);
                    return (int)e;
                }
                if (h_flag == 0) {
                    c1 = cc;
                    w1 = ww;
                }
            }
            int r = COMEMB_E_UNSUPPORTED;
            if (c1)
                r = launch_sg_fused_async(node, negemb, walks, walk_off, n_walks, reduced_windows, seeds, base_seed, table,
                                          table_len, mu, inv_cov, K, window, negative, lr, lambda1, lambda2,
                                          is_node_embedding, atomic, c1, w1, st);
            if (tmp) cudaFreeAsync(tmp, st);
            if (r != COMEMB_E_UNSUPPORTED) return r;
        }
        const int r = launch_sg_fused_round(node, negemb, walks, walk_off, n_walks, reduced_windows, seeds, base_seed, table,
                                            table_len, mu, inv_cov, pi, K, window, negative, lr, lambda1, lambda2,
                                            is_node_embedding, atomic, n_rows, top1_comm, top1_weight, st);
        if (r != COMEMB_E_UNSUPPORTED) return r;
    }
    // round-1 fast path: size 128, NEG in {3,4,5}, separate context table, window <= 12, pi one-hot (or lambda2 == 0)
    if (size == 128 && !is_node_embedding && negemb != node && 2 * window <= VMAX && negative >= 3 && negative <= 5 &&
        (comemb_opts().variant != COMEMB_VARIANT_GENERIC || top1)) {
        int32_t *comm = nullptr;
        float *weight = nullptr;
        bool ok = true;
        if (lambda2 != 0.f && !top1) {
            int *flag = nullptr, h_flag = 0;
            if (n_rows <= 0) ok = false;
            if (ok) {
                CUDA_TRY(cudaMallocAsync(&comm, (size_t)n_rows * sizeof(int32_t), st));
                CUDA_TRY(cudaMallocAsync(&weight, (size_t)n_rows * sizeof(float), st));
                CUDA_TRY(cudaMallocAsync(&flag, sizeof(int), st));
                CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
                pi_dominant_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(pi, n_rows, K, comm, weight, flag);
                CUDA_TRY(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                CUDA_TRY(cudaFreeAsync(flag, st));
                ok = h_flag == 0;  // dense pi -> generic kernel
            }
        }
        float *inv_r = nullptr;
        if (ok && lambda2 != 0.f) {  // inv_cov re-tiled + rounded to TF32 once per call (3.3 MB at K=50)
            const int64_t ne = (int64_t)K * size * size;
            CUDA_TRY(cudaMallocAsync(&inv_r, (size_t)ne * sizeof(float), st));
            tile_inv_cov_kernel<<<148 * 4, 256, 0, st>>>(inv_cov, inv_r, ne);
        }
        if (ok) {
            SgFastParams F;
            F.node = node; F.ctx = negemb; F.walks = walks; F.walk_off = walk_off; F.n_walks = n_walks;
            F.rw = reduced_windows; F.seeds = seeds; F.base_seed = base_seed; F.table = table; F.mod = P.mod;
            F.mu = mu; F.inv_cov = inv_r ? inv_r : inv_cov;
            F.comm = top1 ? top1_comm : comm; F.weight = top1 ? top1_weight : weight; F.window = window;
            F.lr = lr; F.lambda1 = lambda1; F.lambda2 = lambda2; F.glut = P.glut;
            const int e = launch_fast(F, negative, atomic, st);
            if (inv_r) CUDA_TRY(cudaFreeAsync(inv_r, st));
            if (comm) CUDA_TRY(cudaFreeAsync(comm, st));
            if (weight) CUDA_TRY(cudaFreeAsync(weight, st));
            return e;
        }
        if (comm) CUDA_TRY(cudaFreeAsync(comm, st));
        if (weight) CUDA_TRY(cudaFreeAsync(weight, st));
    }
    if (top1) return COMEMB_E_UNSUPPORTED;  // the generic kernel needs the dense pi
    if (size <= 128) return launch_t<1>(P, atomic, st);
    if (size <= 256) return launch_t<2>(P, atomic, st);
    return launch_t<4>(P, atomic, st);
}
