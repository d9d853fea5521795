// gmm_mstep.cu -- the covariance half of the GMM M-step on the tensor cores (the producer of o3's inputs;
// /root/reference/ADSCModel/community_embeddings.py:16-37 -> sklearn _estimate_gaussian_covariances_full):
//      S_k[a][b] = sum_i r_ik (x_i[a] - mu_k[a]) (x_i[b] - mu_k[b])              k < K,  a, b < 128
// (the caller divides by n_k and adds reg_covar).  2*N*K*d^2 flop -- as much as the E-step -- which the library path
// runs as [kc, d, N] x [kc, N, d] batched GEMMs over two materialised [K, N, d] temporaries (5 GB at N=1e5, K=50).
//
// Here one CTA owns 4 components (their four 128 x 128 fp32 accumulators fill the SM's 512 TMEM columns) and a strided
// share of the 32-point tiles.  A tile's points are read once into registers (thread = one point x 8 features); for each
// of the 4 components the CTA writes the transposed, centred tile  B[b][i] = x_i[b] - mu_k[b]  and its weighted copy
// A[a][i] = r_ik B[a][i]  as 3xTF32 hi/lo images in the K-major SWIZZLE_128B layout (K = the point index: 32 points =
// one 128-byte swizzle row, lanes = points -> conflict-free stores) and one thread issues
//      D_k += A_lo.B_hi^T + A_hi.B_lo^T + A_hi.B_hi^T          12 tcgen05.mma (M=128, N=128, K=8) per component and tile
// The images are double-buffered: the tensor cores work on stage s while all warps build stage s^1.  At the end every
// CTA adds its four accumulators to S with red.global.add.v4.f32 (S is zeroed by the launcher).
#include <algorithm>

#include "comemb_common.cuh"
#include "umma.cuh"

namespace {

constexpr int D = 128;
constexpr int TP = 32;  // points per tile = one K-atom
constexpr int CG = 4;   // components per CTA: 4 x 128 TMEM columns
constexpr int WARPS = 16;
constexpr int FPT = D / WARPS;  // features per thread
constexpr int NV = FPT / 4;     // float4 loads per thread and tile
constexpr int IMG = D * TP * 4;  // one [128 features x 32 points] image: 16 KB
constexpr int STAGE = 4 * IMG;   // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_MU = 2 * STAGE;
constexpr int SMEM_BAR = SMEM_MU + CG * D * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 64;

struct MstepParams {
    const float *x;      // [n][128]
    int64_t n;
    const float *resp;   // [n][K]
    const float *means;  // [K][128]
    int K;
    float *scatter;      // [K][128][128], zeroed
    int ranges;          // CTAs per component group
};

__global__ void __launch_bounds__(WARPS * 32, 1) gmm_mstep_kernel(const MstepParams P) {
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *mu_s = reinterpret_cast<float *>(smem + SMEM_MU);
    uint64_t *bar_mma = reinterpret_cast<uint64_t *>(smem + SMEM_BAR);  // [2]: the MMAs that read stage s are complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = (int)(blockIdx.x / P.ranges) * CG, range = (int)(blockIdx.x % P.ranges);
    const int nc = min(CG, P.K - k0);
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        umma::mbar_init(&bar_mma[0], 1);
        umma::mbar_init(&bar_mma[1], 1);
        umma::fence_mbar_init();
    }
    for (int e = threadIdx.x; e < CG * D; e += WARPS * 32)
        mu_s[e] = (e / D) < nc ? __ldg(P.means + (int64_t)(k0 + e / D) * D + (e % D)) : 0.f;
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t taddr = *tmem_slot;
    const int64_t tiles = (P.n + TP - 1) / TP;
    const uint32_t idesc = umma::idesc_tf32_m128(D);

    float4 xv[NV], xn[NV];
    float wv[CG], wn[CG];
    auto load = [&](int64_t t, float4(&xx)[NV], float(&ww)[CG]) {  // thread = point `lane` of the tile, features FPT*warp..
        const int64_t i = t * TP + lane;
        if (i < P.n) {
#pragma unroll
            for (int q = 0; q < NV; q++) xx[q] = __ldg(reinterpret_cast<const float4 *>(P.x + i * D + FPT * warp + 4 * q));
#pragma unroll
            for (int c = 0; c < CG; c++) ww[c] = c < nc ? __ldg(P.resp + i * P.K + k0 + c) : 0.f;
        } else {
#pragma unroll
            for (int q = 0; q < NV; q++) xx[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CG; c++) ww[c] = 0.f;
        }
    };
    int64_t t = range;
    if (t < tiles) load(t, xv, wv);
    uint32_t batch = 0;    // (tile, component) products issued so far; stage = batch & 1
    uint32_t started = 0;  // components whose accumulator has been written
    for (; t < tiles; t += P.ranges) {
        if (t + P.ranges < tiles) load(t + P.ranges, xn, wn);
#pragma unroll
        for (int c = 0; c < CG; c++) {
            if (c >= nc) break;
            // responsibilities that are exactly 0 (exp() underflow once the clusters separate) contribute exactly 0:
            // a tile without any weight for this component is skipped (every warp holds the same 32 weights)
            if (__all_sync(FULL, wv[c] == 0.f)) continue;
            const uint32_t s = batch & 1u;
            if (batch >= 2) umma::mbar_wait(&bar_mma[s], ((batch >> 1) - 1u) & 1u);  // the previous use of this stage
            char *stg = smem + s * STAGE;
            const float w = wv[c];
            const float *mu = mu_s + c * D + FPT * warp;
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const float xs[4] = {xv[q].x, xv[q].y, xv[q].z, xv[q].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int a = FPT * warp + 4 * q + e;
                    const float dv = xs[e] - mu[4 * q + e];
                    const float av = w * dv;  // a tail point (beyond n) has w = 0: its A column is 0, B's does not matter
                    const float dh = umma::tf32_round(dv), ah = umma::tf32_round(av);
                    const uint32_t off = (uint32_t)(a * 128 + ((((lane >> 2) ^ (a & 7))) << 4) + ((lane & 3) << 2));
                    *reinterpret_cast<float *>(stg + off) = ah;
                    *reinterpret_cast<float *>(stg + IMG + off) = umma::tf32_round(av - ah);
                    *reinterpret_cast<float *>(stg + 2 * IMG + off) = dh;
                    *reinterpret_cast<float *>(stg + 3 * IMG + off) = umma::tf32_round(dv - dh);
                }
            }
            umma::fence_proxy_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                umma::tc_fence_after();
                const uint32_t base = umma::smem_u32(stg);
                const uint32_t d = taddr + (uint32_t)(c * D);
                bool acc = (started >> c) & 1u;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {  // small terms first
                    const uint32_t a0 = base + (pass == 0 ? IMG : 0);
                    const uint32_t b0 = base + 2 * IMG + (pass == 1 ? IMG : 0);
#pragma unroll
                    for (int ks = 0; ks < TP / umma::KSTEP; ks++) {
                        umma::mma_tf32(d, umma::smem_desc_sw128(a0 + 32 * ks, 1024), umma::smem_desc_sw128(b0 + 32 * ks, 1024),
                                       idesc, acc);
                        acc = true;
                    }
                }
                umma::mma_commit(&bar_mma[s]);
            }
            started |= 1u << c;
            batch++;
        }
#pragma unroll
        for (int q = 0; q < NV; q++) xv[q] = xn[q];
#pragma unroll
        for (int c = 0; c < CG; c++) wv[c] = wn[c];
    }
    if (batch >= 2) umma::mbar_wait(&bar_mma[(batch - 2) & 1u], ((batch - 2) >> 1) & 1u);
    if (batch >= 1) umma::mbar_wait(&bar_mma[(batch - 1) & 1u], ((batch - 1) >> 1) & 1u);
    umma::tc_fence_after();
    {
        // accumulator row a = TMEM lane; the WARPS/4 warps of a lane quarter split the 128 columns
        constexpr int COLS = D / (WARPS / 4);
        const int a = 32 * (warp & 3) + lane;
        for (int c = 0; c < nc; c++) {
            if (!((started >> c) & 1u)) continue;
            float *dst = P.scatter + ((int64_t)(k0 + c) * D + a) * D + COLS * (warp >> 2);
#pragma unroll 1
            for (int ch = 0; ch < COLS / 16; ch++) {
                float v[16];
                umma::tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(c * D + COLS * (warp >> 2) + 16 * ch), v);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    red_add4(dst + 16 * ch + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, 512);
}

}  // namespace

// d_x [n][128], d_resp [n][K], d_means [K][128] -> d_scatter [K][128][128] = sum_i r_ik (x_i - mu_k)(x_i - mu_k)^T
int launch_gmm_mstep(const float *d_x, int64_t n, const float *d_resp, const float *d_means, int K, float *d_scatter,
                     cudaStream_t st) {
    if (K <= 0) return 0;
    CUDA_TRY(cudaMemsetAsync(d_scatter, 0, (size_t)K * D * D * sizeof(float), st));
    if (n <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int groups = (K + CG - 1) / CG;
    const int64_t tiles = (n + TP - 1) / TP;
    MstepParams P;
    P.x = d_x; P.n = n; P.resp = d_resp; P.means = d_means; P.K = K; P.scatter = d_scatter;
    P.ranges = (int)std::max<int64_t>(1, std::min<int64_t>(sms / groups, tiles));
    const int smem = SMEM_TOTAL + 1024;
    CUDA_TRY(cudaFuncSetAttribute(gmm_mstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gmm_mstep_kernel<<<groups * P.ranges, WARPS * 32, smem, st>>>(P);
    return (int)cudaGetLastError();
}
