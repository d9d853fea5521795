// o3_gemm.cu -- the community gradient step (Community2Vec.train, ADSCModel/community_embeddings.py:61-77) at size 128
// with pi in top-1 form, as a grouped GEMM on the 5th-generation tensor cores:
//
//   rows are bucketed by community (histogram + scan + scatter, three small kernels); a tile job is (community c,
//   up to TN rows of it).  A persistent CTA per SM walks a contiguous range of jobs; for each it
//     * keeps inv_cov_c resident in shared memory as the tcgen05 A operand (hi/lo TF32 images, 128 KB, fetched by
//       the TMA engine with cp.async.bulk only when the community changes),
//     * gathers the rows (prefetched into registers while the previous tile is in the tensor cores, then kept in shared
//       memory across the `iters` iterations), forms diff = x - mu_c, splits it into hi/lo TF32 and stores the swizzled B
//       operand,
//     * one elected thread issues 48 tcgen05.mma (3xTF32, see umma.cuh) with the accumulator in TMEM,
//     * all warps read the accumulator back (tcgen05.ld) and apply  x -= clip(w * G * (beta/K), +-5) * lr.
//
// Accuracy: 3xTF32 products with fp32 accumulation -- the same fp32-level result as the reference's numpy matmul up to
// summation order (tests: <= 1e-5 against the reference's golden output; the previous CUDA-core kernel mirrored the
// ORACLE's double-accumulated dot bit for bit, which the reference itself does not do).
#include <vector>

#include "comemb_common.cuh"
#include "umma.cuh"

namespace {

constexpr int D = 128;
constexpr int A_IMG_BYTES = 2 * D * D * 4;  // hi + lo images of one community: 128 KB

// ---- operand preparation: P [K][128][128] fp32 -> per community {hi image, lo image} with A[m][k] = P[c][k][m] --------------
// (o3 HEAD passes inv_cov with its blocks transposed, and G = inv_cov . diff needs A[m][k] = inv_cov[m][k] = P[k][m];
//  the fused pass passes inv_cov itself and needs Y = inv_cov^T . diff, again A[m][k] = P[k][m].)
__global__ void umma_prep_a_kernel(const float *__restrict__ P, char *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i >> 14;
        const int k = (int)((i >> 7) & 127), m = (int)(i & 127);  // consecutive threads read consecutive m: coalesced
        const float v = P[i];
        const float hi = umma::tf32_round(v);
        const float lo = umma::tf32_round(v - hi);
        char *img = out + c * A_IMG_BYTES;
        const uint32_t off = umma::sw128_offset(D, m, k);
        *reinterpret_cast<float *>(img + off) = hi;
        *reinterpret_cast<float *>(img + D * D * 4 + off) = lo;
    }
}

// ---- bucketing rows by community ---------------------------------------------------------------------------------------------
__global__ void o3g_count_kernel(const uint32_t *rows, int64_t n_sel, const int32_t *comm, const float *weight, int K,
                                 int *count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sel; s += stride) {
        const uint32_t r = rows ? rows[s] : (uint32_t)s;
        const int c = comm[r];
        if (c >= 0 && c < K && weight[r] != 0.f) atomicAdd(count + c, 1);
    }
}

// one block: exclusive scan of the K counts -> row cursor per community, and the tile job list
template <int TN>
__global__ void o3g_scan_kernel(const int *count, int K, int *cursor, int4 *jobs, int *n_jobs) {
    __shared__ int s_row[1024], s_tile[1024];
    __shared__ int carry_row, carry_tile;
    if (threadIdx.x == 0) carry_row = carry_tile = 0;
    __syncthreads();
    for (int base = 0; base < K; base += 1024) {
        const int c = base + threadIdx.x;
        const int n = c < K ? count[c] : 0, t = (n + TN - 1) / TN;
        s_row[threadIdx.x] = n;
        s_tile[threadIdx.x] = t;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
            const int a = threadIdx.x >= o ? s_row[threadIdx.x - o] : 0, b = threadIdx.x >= o ? s_tile[threadIdx.x - o] : 0;
            __syncthreads();
            s_row[threadIdx.x] += a;
            s_tile[threadIdx.x] += b;
            __syncthreads();
        }
        const int row0 = carry_row + s_row[threadIdx.x] - n, tile0 = carry_tile + s_tile[threadIdx.x] - t;
        if (c < K) {
            cursor[c] = row0;
            for (int q = 0; q < t; q++) jobs[tile0 + q] = make_int4(c, row0 + q * TN, min(TN, n - q * TN), 0);
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_row += s_row[1023];
            carry_tile += s_tile[1023];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_jobs = carry_tile;
}

__global__ void o3g_scatter_kernel(const uint32_t *rows, int64_t n_sel, const int32_t *comm, const float *weight, int K,
                                   int *cursor, uint32_t *srows) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sel; s += stride) {
        const uint32_t r = rows ? rows[s] : (uint32_t)s;
        const int c = comm[r];
        if (c >= 0 && c < K && weight[r] != 0.f) srows[atomicAdd(cursor + c, 1)] = r;
    }
}

// ---- the grouped GEMM ------------------------------------------------------------------------------------------------------------
struct O3GemmParams {
    float *node;
    float *gbuf;            // ACCUM mode: [n_sel][128] gradient accumulators (red.add), indexed by selection slot
    const uint32_t *sslot;  // ACCUM mode: selection slot of every bucketed entry
    const float *sw;        // ACCUM mode: responsibility of every bucketed entry
    const uint32_t *srows;
    const int4 *jobs;
    const int *n_jobs;
    const float *mu;
    const char *a_img;
    const float *weight;
    float scale, lr;
    int iters;
};

template <int TN>
struct O3GemmSmem {
    static constexpr int A_HI = 0, A_LO = D * D * 4, B_HI = 2 * D * D * 4, B_LO = B_HI + TN * D * 4;
    static constexpr int MU = B_LO + TN * D * 4, ROW = MU + D * 4, WGT = ROW + TN * 4, BAR = WGT + TN * 4;
    static constexpr int TOTAL = BAR + 32;
};

// ACCUM = false: top-1 pi, the epilogue applies the update.  ACCUM = true: sparse pi with several non-zero responsibilities
// per row -- one bucketed entry per (row, community), the epilogue adds w * G into the row's accumulator (red.add; a
// second kernel applies the clipped update), one pass per launch.
template <int TN, int WARPS, bool ACCUM>
__global__ void __launch_bounds__(WARPS * 32, 1) o3_gemm_kernel(const O3GemmParams P) {
    using L = O3GemmSmem<TN>;
    static_assert(TN % 16 == 0 && TN <= 256 && WARPS % 4 == 0, "tile shape");
    constexpr uint32_t TMEM_COLS = TN <= 32 ? 32 : TN <= 64 ? 64 : TN <= 128 ? 128 : 256;
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *mu_s = reinterpret_cast<float *>(smem + L::MU);
    uint32_t *row_s = reinterpret_cast<uint32_t *>(smem + L::ROW);
    float *wgt_s = reinterpret_cast<float *>(smem + L::WGT);
    uint64_t *bar_a = reinterpret_cast<uint64_t *>(smem + L::BAR), *bar_mma = bar_a + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_a + 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (warp == 0) umma::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 0) {
        umma::mbar_init(bar_a, 1);
        umma::mbar_init(bar_mma, 1);
        umma::fence_mbar_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t taddr = *tmem_slot;
    const uint32_t a_hi = umma::smem_u32(smem + L::A_HI), a_lo = umma::smem_u32(smem + L::A_LO);
    const uint32_t b_hi = umma::smem_u32(smem + L::B_HI), b_lo = umma::smem_u32(smem + L::B_LO);

    const int n_jobs = *P.n_jobs;
    const int j0 = (int)((int64_t)n_jobs * blockIdx.x / gridDim.x), j1 = (int)((int64_t)n_jobs * (blockIdx.x + 1) / gridDim.x);
    float *x_s = reinterpret_cast<float *>(smem + ((L::TOTAL + 15) & ~15));  // [TN][128] the tile's rows (fp32) for the epilogue
    int cur_c = -1;
    uint32_t par_a = 0, par_m = 0;
    bool a_pending = false;
    constexpr int RPW = (TN + WARPS - 1) / WARPS;  // rows per warp
    uint32_t rowv[RPW];
    float4 xv[RPW];
    float wv[RPW];
    // rows of job j into registers: all gathers in flight at once, issued while the previous tile is still in the tensor cores
    auto prefetch = [&](int j) {
        const int4 job = __ldg(P.jobs + j);
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = warp + q * WARPS;
            rowv[q] = r < job.z ? __ldg(P.srows + job.y + r) : 0u;
        }
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = warp + q * WARPS;
            if (r < job.z) {
                xv[q] = __ldcg(reinterpret_cast<const float4 *>(P.node + (int64_t)rowv[q] * D + 4 * lane));
                wv[q] = ACCUM ? __ldg(P.sw + job.y + r) : __ldg(P.weight + rowv[q]);
                if (ACCUM) rowv[q] = __ldg(P.sslot + job.y + r);  // from here on the "row" is the accumulator slot
            }
        }
    };
    if (j0 < j1) prefetch(j0);
    for (int j = j0; j < j1; j++) {
        const int4 job = __ldg(P.jobs + j);
        const int c = job.x, cnt = job.z;
        const int n16 = (cnt + 15) & ~15;
        if (c != cur_c) {  // every MMA that read the resident A has completed (bar_mma was waited on below)
            if (threadIdx.x == 0) {
                umma::mbar_expect_tx(bar_a, A_IMG_BYTES);
                const char *src = P.a_img + (int64_t)c * A_IMG_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++) umma::bulk_g2s(smem + L::A_HI + q * 16384, src + q * 16384, 16384, bar_a);
            }
            if (warp == 1) *reinterpret_cast<float4 *>(mu_s + 4 * lane) =
                __ldg(reinterpret_cast<const float4 *>(P.mu + (int64_t)c * D + 4 * lane));
            a_pending = true;
            cur_c = c;
            __syncthreads();  // mu_s visible
        }
        const float4 m = *reinterpret_cast<const float4 *>(mu_s + 4 * lane);
        // the tile's rows: registers -> shared memory (kept there across the iterations: nothing is re-read from global)
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = warp + q * WARPS;
            if (r < cnt) {
                *reinterpret_cast<float4 *>(x_s + r * D + 4 * lane) = xv[q];
                if (lane == 0) {
                    row_s[r] = rowv[q];
                    wgt_s[r] = wv[q];
                }
            }
        }
        const int iters = ACCUM ? 1 : P.iters;
        for (int it = 0; it < iters; it++) {
            __syncthreads();  // x_s of this iteration is complete
            // ---- B operand: diff rows, hi/lo split, swizzled K-major ----------------------------------------------------
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int r = warp + q * WARPS;
                if (r < cnt) {
                    const float4 x = *reinterpret_cast<const float4 *>(x_s + r * D + 4 * lane);
                    const float4 df = make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
                    const float4 hi = make_float4(umma::tf32_round(df.x), umma::tf32_round(df.y), umma::tf32_round(df.z),
                                                  umma::tf32_round(df.w));
                    const float4 lo = make_float4(umma::tf32_round(df.x - hi.x), umma::tf32_round(df.y - hi.y),
                                                  umma::tf32_round(df.z - hi.z), umma::tf32_round(df.w - hi.w));
                    const uint32_t off = umma::sw128_offset(TN, r, 4 * lane);
                    *reinterpret_cast<float4 *>(smem + L::B_HI + off) = hi;
                    *reinterpret_cast<float4 *>(smem + L::B_LO + off) = lo;
                }
            }
            umma::fence_proxy_async_smem();
            __syncthreads();
            if (warp == 0) {
                if (a_pending) {
                    umma::mbar_wait(bar_a, par_a);
                    par_a ^= 1;
                }
                umma::tc_fence_after();
                if (lane == 0) {
                    umma::issue_3xtf32(taddr, a_hi, a_lo, b_hi, b_lo, TN, n16);
                    umma::mma_commit(bar_mma);
                }
                __syncwarp();
            }
            a_pending = false;
            if (it == iters - 1 && j + 1 < j1) prefetch(j + 1);  // the next tile's rows travel while the tensor cores work
            umma::mbar_wait(bar_mma, par_m);
            par_m ^= 1;
            umma::tc_fence_after();
            // ---- epilogue: thread = output coordinate a of TMEM lane quarter (warp % 4), 16 rows at a time ----------------
            const int a = 32 * (warp & 3) + lane;
            for (int ch = warp >> 2; ch * 16 < n16; ch += WARPS / 4) {
                float v[16];
                umma::tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ch * 16), v);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int n = ch * 16 + q;
                    if (n < cnt) {
                        if (ACCUM) {
                            atomicAdd(P.gbuf + (int64_t)row_s[n] * D + a, __fmul_rn(wgt_s[n], v[q]));  // :71-72 summed over k
                        } else {
                            float g = __fmul_rn(__fmul_rn(wgt_s[n], v[q]), P.scale);  // :71, :76
                            g = fminf(fmaxf(g, -5.f), 5.f);                           // :77
                            const float nx = x_s[n * D + a] - __fmul_rn(g, P.lr);
                            x_s[n * D + a] = nx;
                            if (it == iters - 1) P.node[(int64_t)row_s[n] * D + a] = nx;
                        }
                    }
                }
            }
            umma::tc_fence_before();
        }
        __syncthreads();  // accumulator, operand images and x_s are free for the next tile
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, TMEM_COLS);
}

}  // namespace

// P: [K][128][128] fp32; out: K * 128 KB operand images (see umma_prep_a_kernel)
int launch_umma_prep_a(const float *P, char *out, int K, cudaStream_t st) {
    const int64_t n = (int64_t)K * D * D;
    if (n <= 0) return 0;
    umma_prep_a_kernel<<<(unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8), 256, 0, st>>>(P, out, n);
    return (int)cudaGetLastError();
}

// top-1 form at size 128; returns COMEMB_E_UNSUPPORTED when the shape does not fit (the caller falls back)
int launch_o3_gemm(float *node, const uint32_t *rows, int64_t n_sel, const float *mu, const float *inv_cov_t,
                   const int32_t *comm, const float *weight, int K, float scale, float lr, int iters, cudaStream_t st) {
    constexpr int TN = 64, WARPS = 16;
    using L = O3GemmSmem<TN>;
    if (n_sel <= 0 || iters <= 0) return 0;
    if (n_sel >= (1LL << 31) || K > (1 << 20)) return COMEMB_E_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t max_jobs = (n_sel + TN - 1) / TN + K;
    char *scratch = nullptr;
    const size_t off_count = 0, off_cursor = off_count + (size_t)K * 4, off_njobs = off_cursor + (size_t)K * 4;
    const size_t off_jobs = (off_njobs + 4 + 15) & ~(size_t)15, off_srows = off_jobs + (size_t)max_jobs * 16;
    const size_t off_img = (off_srows + (size_t)n_sel * 4 + 1023) & ~(size_t)1023;
    const size_t total = off_img + (size_t)K * A_IMG_BYTES;
    CUDA_TRY(cudaMallocAsync(&scratch, total, st));
    auto fail = [&](cudaError_t e) {
        cudaFreeAsync(scratch, st);
        return (int)e;
    };
    int *count = reinterpret_cast<int *>(scratch + off_count), *cursor = reinterpret_cast<int *>(scratch + off_cursor);
    int *n_jobs = reinterpret_cast<int *>(scratch + off_njobs);
    int4 *jobs = reinterpret_cast<int4 *>(scratch + off_jobs);
    uint32_t *srows = reinterpret_cast<uint32_t *>(scratch + off_srows);
    char *img = scratch + off_img;
    cudaError_t e = cudaMemsetAsync(count, 0, (size_t)K * 4, st);
    if (e != cudaSuccess) return fail(e);
    const int g1 = (int)((n_sel + 255) / 256 < sms * 8 ? (n_sel + 255) / 256 : sms * 8);
    o3g_count_kernel<<<g1, 256, 0, st>>>(rows, n_sel, comm, weight, K, count);
    o3g_scan_kernel<TN><<<1, 1024, 0, st>>>(count, K, cursor, jobs, n_jobs);
    o3g_scatter_kernel<<<g1, 256, 0, st>>>(rows, n_sel, comm, weight, K, cursor, srows);
    int r = launch_umma_prep_a(inv_cov_t, img, K, st);
    if (r) return fail((cudaError_t)r);
    O3GemmParams P;
    P.node = node; P.srows = srows; P.jobs = jobs; P.n_jobs = n_jobs; P.mu = mu; P.a_img = img; P.weight = weight;
    P.scale = scale; P.lr = lr; P.iters = iters;
    const int smem = L::TOTAL + 1024 + 16 + TN * D * 4;  // + the tile's rows in fp32
    P.gbuf = nullptr; P.sslot = nullptr; P.sw = nullptr;
    e = cudaFuncSetAttribute(o3_gemm_kernel<TN, WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(e);
    const int grid = (int)(max_jobs < sms ? max_jobs : sms);
    o3_gemm_kernel<TN, WARPS, false><<<grid, WARPS * 32, smem, st>>>(P);
    e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    return (int)e;
}

// ---- sparse pi (several non-zero responsibilities per row): entries (row, community, weight) bucketed by community ---------------
namespace {
__global__ void o3s_count_kernel(const uint32_t *rows, int64_t n_sel, const float *pi, int K, int *count) {
    // one warp per selected row: lanes scan the row's K responsibilities (coalesced)
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = w0; s < n_sel; s += nw) {
        const int64_t r = rows ? rows[s] : s;
        for (int k = lane; k < K; k += 32)
            if (pi[r * K + k] != 0.f) atomicAdd(count + k, 1);
    }
}
__global__ void o3s_scatter_kernel(const uint32_t *rows, int64_t n_sel, const float *pi, int K, int *cursor, uint32_t *srows,
                                   uint32_t *sslot, float *sw) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = w0; s < n_sel; s += nw) {
        const int64_t r = rows ? rows[s] : s;
        for (int k = lane; k < K; k += 32) {
            const float p = pi[r * K + k];
            if (p != 0.f) {
                const int at = atomicAdd(cursor + k, 1);
                srows[at] = (uint32_t)r;
                sslot[at] = (uint32_t)s;
                sw[at] = p;
            }
        }
    }
}
// x_r -= clip(G_s * scale, +-5) * lr  (community_embeddings.py:76-77), and the accumulator is cleared for the next iteration
__global__ void o3s_apply_kernel(float *node, const uint32_t *rows, int64_t n_sel, float *gbuf, float scale, float lr) {
    const int64_t n4 = n_sel * (D / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / (D / 4);
        const int q = (int)(i % (D / 4));
        const int64_t r = rows ? rows[s] : s;
        float4 g = *reinterpret_cast<float4 *>(gbuf + s * D + 4 * q);
        *reinterpret_cast<float4 *>(gbuf + s * D + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *xp = reinterpret_cast<float4 *>(node + r * D + 4 * q);
        float4 x = *xp;
        x.x -= __fmul_rn(fminf(fmaxf(__fmul_rn(g.x, scale), -5.f), 5.f), lr);
        x.y -= __fmul_rn(fminf(fmaxf(__fmul_rn(g.y, scale), -5.f), 5.f), lr);
        x.z -= __fmul_rn(fminf(fmaxf(__fmul_rn(g.z, scale), -5.f), 5.f), lr);
        x.w -= __fmul_rn(fminf(fmaxf(__fmul_rn(g.w, scale), -5.f), 5.f), lr);
        *xp = x;
    }
}
}  // namespace

// dense pi pointer whose rows are SPARSE (at most max_nnz_per_row non-zeros on average): size 128.  Returns
// COMEMB_E_UNSUPPORTED when pi is too dense to pay off (the caller keeps the one-row-per-warp kernel).
int launch_o3_gemm_sparse(float *node, const uint32_t *rows, int64_t n_sel, const float *mu, const float *inv_cov_t,
                          const float *pi, int K, float scale, float lr, int iters, cudaStream_t st) {
    constexpr int TN = 64, WARPS = 16;
    using L = O3GemmSmem<TN>;
    if (n_sel <= 0 || iters <= 0) return 0;
    if (n_sel >= (1LL << 28) || K > (1 << 20)) return COMEMB_E_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // pass 1: entries per community (the host needs the total to size the entry arrays: one small synchronous copy)
    int *count = nullptr;
    CUDA_TRY(cudaMallocAsync(&count, (size_t)(2 * K + 4) * 4, st));
    auto fail0 = [&](cudaError_t e) {
        cudaFreeAsync(count, st);
        return (int)e;
    };
    cudaError_t e = cudaMemsetAsync(count, 0, (size_t)K * 4, st);
    if (e != cudaSuccess) return fail0(e);
    const int gw = (int)((n_sel + 7) / 8 < sms * 8 ? (n_sel + 7) / 8 : sms * 8);
    o3s_count_kernel<<<gw, 256, 0, st>>>(rows, n_sel, pi, K, count);
    std::vector<int> h_count(K);
    if ((e = cudaMemcpyAsync(h_count.data(), count, (size_t)K * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail0(e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail0(e);
    int64_t nnz = 0;
    for (int k = 0; k < K; k++) nnz += h_count[k];
    if (nnz > 8 * n_sel || nnz >= (1LL << 31)) return fail0((cudaError_t)0), COMEMB_E_UNSUPPORTED;
    if (nnz == 0) return fail0((cudaError_t)0), 0;
    const int64_t max_jobs = (nnz + TN - 1) / TN + K;
    char *scratch = nullptr;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        const size_t at = off;
        off = (off + bytes + 1023) & ~(size_t)1023;
        return at;
    };
    const size_t o_jobs = carve((size_t)max_jobs * 16), o_rows = carve((size_t)nnz * 4), o_slot = carve((size_t)nnz * 4);
    const size_t o_w = carve((size_t)nnz * 4), o_g = carve((size_t)n_sel * D * 4), o_img = carve((size_t)K * A_IMG_BYTES);
    if ((e = cudaMallocAsync(&scratch, off, st)) != cudaSuccess) return fail0(e);
    auto fail = [&](cudaError_t err) {
        cudaFreeAsync(scratch, st);
        cudaFreeAsync(count, st);
        return (int)err;
    };
    int *cursor = count + K, *n_jobs = count + 2 * K;
    int4 *jobs = reinterpret_cast<int4 *>(scratch + o_jobs);
    uint32_t *srows = reinterpret_cast<uint32_t *>(scratch + o_rows), *sslot = reinterpret_cast<uint32_t *>(scratch + o_slot);
    float *sw = reinterpret_cast<float *>(scratch + o_w), *gbuf = reinterpret_cast<float *>(scratch + o_g);
    o3g_scan_kernel<TN><<<1, 1024, 0, st>>>(count, K, cursor, jobs, n_jobs);
    o3s_scatter_kernel<<<gw, 256, 0, st>>>(rows, n_sel, pi, K, cursor, srows, sslot, sw);
    if ((e = cudaMemsetAsync(gbuf, 0, (size_t)n_sel * D * 4, st)) != cudaSuccess) return fail(e);
    int r = launch_umma_prep_a(inv_cov_t, scratch + o_img, K, st);
    if (r) return fail((cudaError_t)r);
    O3GemmParams P;
    P.node = node; P.gbuf = gbuf; P.sslot = sslot; P.sw = sw; P.srows = srows; P.jobs = jobs; P.n_jobs = n_jobs; P.mu = mu;
    P.a_img = scratch + o_img; P.weight = nullptr; P.scale = scale; P.lr = lr; P.iters = 1;
    const int smem = L::TOTAL + 1024 + 16 + TN * D * 4;
    if ((e = cudaFuncSetAttribute(o3_gemm_kernel<TN, WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return fail(e);
    const int grid = (int)(max_jobs < sms ? max_jobs : sms);
    const int64_t n4 = n_sel * (D / 4);
    for (int it = 0; it < iters; it++) {  // the gradient is frozen per iteration (:64-73), applied once (:76-77)
        o3_gemm_kernel<TN, WARPS, true><<<grid, WARPS * 32, smem, st>>>(P);
        o3s_apply_kernel<<<(int)((n4 + 255) / 256 < sms * 8 ? (n4 + 255) / 256 : sms * 8), 256, 0, st>>>(node, rows, n_sel, gbuf,
                                                                                                         scale, lr);
    }
    e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    cudaFreeAsync(count, st);
    return (int)e;
}

// ==== GMM E-step (sklearn _estimate_log_gaussian_prob, 'full'):  sq[n][k] = || x_n . P_k - mu_k . P_k ||^2 =========
// The same contraction with A = P_k^T (precision Cholesky factor, resident hi/lo TF32 images) and B = 64 consecutive points:
// D[j][n] = sum_b P_k[b][j] x_n[b] in TMEM; the epilogue subtracts bias[k][j] = (mu_k . P_k)[j], squares, and reduces over
// the 128 output coordinates j (= TMEM lanes = threads) with a transposed shuffle reduction, so only the [N, K] matrix of
// squared norms ever reaches memory (the library formulation materialises the [N, K*d] product).  Jobs (k, tile) are
// walked by persistent CTAs in k-major order: P_k is fetched (TMA bulk copy) once per CTA and component.
namespace {

// 16 values per lane, summed over the 32 lanes: 8+4+2+1+1 shuffles.  On return the lanes whose bit 0 is clear hold the total of
// column ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1).
__device__ __forceinline__ float reduce16_transposed(const float (&v)[16], int lane) {
    float r8[8], r4[4], r2[2];
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; i++) r8[i] = (b16 ? v[8 + i] : v[i]) + __shfl_xor_sync(FULL, b16 ? v[i] : v[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; i++) r4[i] = (b8 ? r8[4 + i] : r8[i]) + __shfl_xor_sync(FULL, b8 ? r8[i] : r8[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; i++) r2[i] = (b4 ? r4[2 + i] : r4[i]) + __shfl_xor_sync(FULL, b4 ? r4[i] : r4[2 + i], 4);
    float t = (b2 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b2 ? r2[0] : r2[1], 2);
    t += __shfl_xor_sync(FULL, t, 1);
    return t;
}

struct EstepParams {
    const float *x;
    int64_t n;
    int K;
    const char *a_img;
    const float *bias;  // [K][128] = mu_k . P_k
    float *sq;          // [n][K]
};

template <int TN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) gmm_estep_kernel(const EstepParams P) {
    using L = O3GemmSmem<TN>;
    static_assert(TN == 64 && WARPS % 4 == 0, "tile shape");
    constexpr uint32_t TMEM_COLS = 64;
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *bias_s = reinterpret_cast<float *>(smem + L::MU);     // [128]
    float *part_s = reinterpret_cast<float *>(smem + ((L::TOTAL + 15) & ~15));  // [TN][4] partial sums per TMEM lane quarter
    uint64_t *bar_a = reinterpret_cast<uint64_t *>(smem + L::BAR), *bar_mma = bar_a + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_a + 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) umma::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 0) {
        umma::mbar_init(bar_a, 1);
        umma::mbar_init(bar_mma, 1);
        umma::fence_mbar_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t taddr = *tmem_slot;
    const uint32_t a_hi = umma::smem_u32(smem + L::A_HI), a_lo = umma::smem_u32(smem + L::A_LO);
    const uint32_t b_hi = umma::smem_u32(smem + L::B_HI), b_lo = umma::smem_u32(smem + L::B_LO);
    const int64_t tiles = (P.n + TN - 1) / TN, n_jobs = tiles * P.K;
    const int64_t j0 = n_jobs * blockIdx.x / gridDim.x, j1 = n_jobs * (blockIdx.x + 1) / gridDim.x;
    int cur_k = -1;
    uint32_t par_a = 0, par_m = 0;
    bool a_pending = false;
    constexpr int RPW = (TN + WARPS - 1) / WARPS;
    float4 xv[RPW];
    auto prefetch = [&](int64_t job) {  // the tile's rows are consecutive: coalesced, all in flight at once
        const int64_t r0 = (job % tiles) * TN;
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int64_t r = r0 + warp + q * WARPS;
            xv[q] = (warp + q * WARPS < TN && r < P.n) ? __ldg(reinterpret_cast<const float4 *>(P.x + r * D + 4 * lane))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (j0 < j1) prefetch(j0);
    for (int64_t j = j0; j < j1; j++) {
        const int k = (int)(j / tiles);
        const int64_t r0 = (j % tiles) * TN;
        const int cnt = (int)min((int64_t)TN, P.n - r0);
        const int n16 = (cnt + 15) & ~15;
        if (k != cur_k) {
            if (threadIdx.x == 0) {
                umma::mbar_expect_tx(bar_a, A_IMG_BYTES);
                const char *src = P.a_img + (int64_t)k * A_IMG_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++) umma::bulk_g2s(smem + L::A_HI + q * 16384, src + q * 16384, 16384, bar_a);
            }
            if (warp == 1) *reinterpret_cast<float4 *>(bias_s + 4 * lane) =
                __ldg(reinterpret_cast<const float4 *>(P.bias + (int64_t)k * D + 4 * lane));
            a_pending = true;
            cur_k = k;
        }
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = warp + q * WARPS;
            if (r < TN) {
                const float4 x = xv[q];
                const float4 hi = make_float4(umma::tf32_round(x.x), umma::tf32_round(x.y), umma::tf32_round(x.z),
                                              umma::tf32_round(x.w));
                const float4 lo = make_float4(umma::tf32_round(x.x - hi.x), umma::tf32_round(x.y - hi.y),
                                              umma::tf32_round(x.z - hi.z), umma::tf32_round(x.w - hi.w));
                const uint32_t off = umma::sw128_offset(TN, r, 4 * lane);
                *reinterpret_cast<float4 *>(smem + L::B_HI + off) = hi;
                *reinterpret_cast<float4 *>(smem + L::B_LO + off) = lo;
            }
        }
        umma::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (a_pending) {
                umma::mbar_wait(bar_a, par_a);
                par_a ^= 1;
            }
            umma::tc_fence_after();
            if (lane == 0) {
                umma::issue_3xtf32(taddr, a_hi, a_lo, b_hi, b_lo, TN, n16);
                umma::mma_commit(bar_mma);
            }
            __syncwarp();
        }
        a_pending = false;
        if (j + 1 < j1) prefetch(j + 1);  // the next tile's rows travel while the tensor cores work
        umma::mbar_wait(bar_mma, par_m);
        par_m ^= 1;
        umma::tc_fence_after();
        // epilogue: thread = output coordinate jj of TMEM lane quarter (warp % 4); warps with the same quarter share the chunks
        const int jj = 32 * (warp & 3) + lane;
        const float bj = bias_s[jj];
        for (int ch = warp >> 2; ch * 16 < n16; ch += WARPS / 4) {
            float v[16];
            umma::tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ch * 16), v);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float y = v[q] - bj;
                v[q] = y * y;
            }
            const float tot = reduce16_transposed(v, lane);
            if (!(lane & 1)) {
                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                part_s[(ch * 16 + col) * 4 + (warp & 3)] = tot;
            }
        }
        umma::tc_fence_before();
        __syncthreads();
        if (threadIdx.x < cnt) {
            const float4 p = *reinterpret_cast<const float4 *>(part_s + 4 * threadIdx.x);
            P.sq[(r0 + threadIdx.x) * P.K + k] = (p.x + p.y) + (p.z + p.w);
        }
        __syncthreads();  // part_s, the B images and the accumulator are free again
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, TMEM_COLS);
}

}  // namespace

// d_x [n][128], d_prec_chol [K][128][128] (P_k as sklearn stores it), d_bias [K][128] = mu_k . P_k, d_sq [n][K] (out)
int launch_gmm_estep(const float *d_x, int64_t n, const float *d_prec_chol, const float *d_bias, int K, float *d_sq,
                     cudaStream_t st) {
    constexpr int TN = 64, WARPS = 16;
    using L = O3GemmSmem<TN>;
    if (n <= 0 || K <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    char *img = nullptr;
    CUDA_TRY(cudaMallocAsync(&img, (size_t)K * A_IMG_BYTES, st));
    int r = launch_umma_prep_a(d_prec_chol, img, K, st);
    if (r) {
        cudaFreeAsync(img, st);
        return r;
    }
    EstepParams P;
    P.x = d_x; P.n = n; P.K = K; P.a_img = img; P.bias = d_bias; P.sq = d_sq;
    const int smem = L::TOTAL + 1024 + 1024;
    cudaError_t e = cudaFuncSetAttribute(gmm_estep_kernel<TN, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) {
        const int64_t jobs = ((n + TN - 1) / TN) * K;
        gmm_estep_kernel<TN, WARPS><<<(int)(jobs < sms ? jobs : sms), WARPS * 32, smem, st>>>(P);
        e = cudaGetLastError();
    }
    cudaFreeAsync(img, st);
    return (int)e;
}
