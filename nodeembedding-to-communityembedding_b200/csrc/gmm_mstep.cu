// gmm_mstep.cu -- the covariance half of the GMM M-step on the tensor cores (the producer of o3's inputs;
// /root/reference/ADSCModel/community_embeddings.py:16-37 -> sklearn _estimate_gaussian_covariances_full):
//      S_k[a][b] = sum_i r_ik (x_i[a] - mu_k[a]) (x_i[b] - mu_k[b])              k < K,  a, b < 128
// (the caller divides by n_k and adds reg_covar).  2*N*K*d^2 flop -- as much as the E-step -- which the library path
// runs as [kc, d, N] x [kc, N, d] batched GEMMs over two materialised [K, N, d] temporaries (5 GB at N=1e5, K=50).
//
// Here one CTA owns 4 components (their four 128 x 128 fp32 accumulators fill the SM's 512 TMEM columns) and a strided
// share of the 32-point tiles.  A tile's points are read once into registers (thread = 4 points x 2 features); for each
// of the 4 components the CTA writes the transposed, centred tile  B[b][i] = x_i[b] - mu_k[b]  and its weighted copy
// A[a][i] = r_ik B[a][i]  as 3xTF32 hi/lo images in the K-major SWIZZLE_128B layout (K = the point index: 32 points =
// one 128-byte swizzle row, a thread's 4 points = one 16-byte chunk -> 128-bit conflict-free stores) and one thread issues
//      D_k += A_lo.B_hi^T + A_hi.B_lo^T + A_hi.B_hi^T          12 tcgen05.mma (M=128, N=128, K=8) per component and tile
// by a 17th warp that does nothing else (issuing blocks a thread for about the MMAs' duration).  The images are
// double-buffered and handed over through mbarriers -- no CTA-wide barrier per tile: the tensor cores work on stage s
// while the 16 staging warps build stage s^1.  At the end every
// CTA adds its four accumulators to S with red.global.add.v4.f32 (S is zeroed by the launcher).
#include <algorithm>

#include "comemb_common.cuh"
#include "umma.cuh"

namespace {

constexpr int D = 128;
constexpr int TP = 32;  // points per tile = one K-atom
constexpr int CG = 4;   // components per CTA: 4 x 128 TMEM columns
constexpr int WARPS = 16;
constexpr int IMG = D * TP * 4;  // one [128 features x 32 points] image: 16 KB
constexpr int STAGE = 4 * IMG;   // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_MU = 2 * STAGE;
constexpr int SMEM_BAR = SMEM_MU + CG * D * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 128;

struct MstepParams {
    const float *x;      // [n][128]
    int64_t n;
    const float *resp;   // [n][K]
    const float *means;  // [K][128]
    int K;
    float *scatter;      // [K][128][128], zeroed
    int ranges;          // CTAs per component group
};

__global__ void __launch_bounds__((WARPS + 1) * 32, 1) gmm_mstep_kernel(const MstepParams P) {
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *mu_s = reinterpret_cast<float *>(smem + SMEM_MU);
    uint64_t *bar_mma = reinterpret_cast<uint64_t *>(smem + SMEM_BAR);  // [2]: the MMAs that read stage s are complete
    uint64_t *bar_stage = bar_mma + 2;                                  // [2]: the 16 staging warps have written stage s
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 4);
    int *desc_s = reinterpret_cast<int *>(bar_mma + 5);                 // [2]: what stage s holds: component | accumulate << 8, -1 = end
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = (int)(blockIdx.x / P.ranges) * CG, range = (int)(blockIdx.x % P.ranges);
    const int nc = min(CG, P.K - k0);
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        umma::mbar_init(&bar_mma[0], 1);
        umma::mbar_init(&bar_mma[1], 1);
        umma::mbar_init(&bar_stage[0], WARPS);
        umma::mbar_init(&bar_stage[1], WARPS);
        umma::fence_mbar_init();
    }
    for (int e = threadIdx.x; e < CG * D; e += (WARPS + 1) * 32)
        mu_s[e] = (e / D) < nc ? __ldg(P.means + (int64_t)(k0 + e / D) * D + (e % D)) : 0.f;
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t taddr = *tmem_slot;
    const int64_t tiles = (P.n + TP - 1) / TP;
    const uint32_t idesc = umma::idesc_tf32_m128(D);

    if (warp == WARPS) {
        // ---- the issuing warp: stage s is complete -> 12 tcgen05.mma into the component's accumulator -> commit.  Issuing
        // blocks the thread for about the duration of the MMAs; on a warp of its own that costs the staging warps nothing.
        for (uint32_t batch = 0;; batch++) {
            const uint32_t st = batch & 1u;
            umma::mbar_wait(&bar_stage[st], (batch >> 1) & 1u);
            const int dsc = desc_s[st];
            if (dsc < 0) break;
            umma::tc_fence_after();
            if (lane == 0) {
                const uint32_t base = umma::smem_u32(smem + st * STAGE);
                const uint32_t d = taddr + (uint32_t)((dsc & 0xFF) * D);
                bool acc = (dsc >> 8) != 0;
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
                umma::mma_commit(&bar_mma[st]);
            }
            __syncwarp();
        }
        umma::tc_fence_before();
        __syncthreads();  // matches the staging warps' final barrier
        return;
    }

    // thread = 4 consecutive points (4*pg .. 4*pg+3 of the tile) x 2 features (2*fp, 2*fp+1): the four points of a feature
    // are one 16-byte chunk of the K-major image, so every (feature, image) is ONE 128-bit shared store; a warp covers
    // all 32 points for 8 features (4 rows of 128 bytes per store instruction: bank-conflict-free)
    const int pg = threadIdx.x & 7, fp = threadIdx.x >> 3;
    static_assert(WARPS * 32 == 8 * (D / 2), "one thread per (point group, feature pair)");
    float2 xv[4], xn[4];
    float wv[4][CG], wn[4][CG];
    auto load = [&](int64_t t, float2(&xx)[4], float(&ww)[4][CG]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int64_t i = t * TP + 4 * pg + q;
            if (i < P.n) {
                xx[q] = __ldg(reinterpret_cast<const float2 *>(P.x + i * D + 2 * fp));
#pragma unroll
                for (int c = 0; c < CG; c++) ww[q][c] = c < nc ? __ldg(P.resp + i * P.K + k0 + c) : 0.f;
            } else {
                xx[q] = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < CG; c++) ww[q][c] = 0.f;
            }
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
            // a tile without any weight for this component is skipped (every warp covers all 32 points of the tile)
            if (__all_sync(FULL, wv[0][c] == 0.f && wv[1][c] == 0.f && wv[2][c] == 0.f && wv[3][c] == 0.f)) continue;
            const uint32_t s = batch & 1u;
            if (batch >= 2) umma::mbar_wait(&bar_mma[s], ((batch >> 1) - 1u) & 1u);  // the previous use of this stage
            char *stg = smem + s * STAGE;
            const float2 mu = *reinterpret_cast<const float2 *>(mu_s + c * D + 2 * fp);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int a = 2 * fp + e;
                const float m = e ? mu.y : mu.x;
                float dv[4], av[4], dh[4], ah[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    dv[q] = (e ? xv[q].y : xv[q].x) - m;
                    av[q] = wv[q][c] * dv[q];  // a tail point (beyond n) has w = 0: its A column is 0, B's does not matter
                    dh[q] = umma::tf32_round(dv[q]);
                    ah[q] = umma::tf32_round(av[q]);
                }
                const uint32_t off = (uint32_t)(a * 128 + ((pg ^ (a & 7)) << 4));
                *reinterpret_cast<float4 *>(stg + off) = make_float4(ah[0], ah[1], ah[2], ah[3]);
                *reinterpret_cast<float4 *>(stg + IMG + off) =
                    make_float4(umma::tf32_round(av[0] - ah[0]), umma::tf32_round(av[1] - ah[1]),
                                umma::tf32_round(av[2] - ah[2]), umma::tf32_round(av[3] - ah[3]));
                *reinterpret_cast<float4 *>(stg + 2 * IMG + off) = make_float4(dh[0], dh[1], dh[2], dh[3]);
                *reinterpret_cast<float4 *>(stg + 3 * IMG + off) =
                    make_float4(umma::tf32_round(dv[0] - dh[0]), umma::tf32_round(dv[1] - dh[1]),
                                umma::tf32_round(dv[2] - dh[2]), umma::tf32_round(dv[3] - dh[3]));
            }
            if (threadIdx.x == 0) desc_s[s] = c | (int)(((started >> c) & 1u) << 8);
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {  // release: this warp's part of stage s (and warp 0's descriptor) is written
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&bar_stage[s])) : "memory");
            }
            started |= 1u << c;
            batch++;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            xv[q] = xn[q];
#pragma unroll
            for (int c = 0; c < CG; c++) wv[q][c] = wn[q][c];
        }
    }
    {  // end marker for the issuing warp, through the same hand-off
        const uint32_t s = batch & 1u;
        if (batch >= 2) umma::mbar_wait(&bar_mma[s], ((batch >> 1) - 1u) & 1u);
        if (threadIdx.x == 0) desc_s[s] = -1;
        __syncwarp();
        if (lane == 0)
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&bar_stage[s])) : "memory");
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
    gmm_mstep_kernel<<<groups * P.ranges, (WARPS + 1) * 32, smem, st>>>(P);
    return (int)cudaGetLastError();
}
