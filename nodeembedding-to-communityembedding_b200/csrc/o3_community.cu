// o3_community.cu -- the community (GMM) gradient step, HEAD form: Community2Vec.train
// (ADSCModel/community_embeddings.py:61-77).
//
//   G_r = sum_k pi[r,k] * inv_cov[k] @ (x_r - mu_k)          (gradient frozen per iteration, :64-73)
//   x_r -= clip(G_r * (float)(beta/K), -5, 5) * lr            (:76-77)
//
// G_r depends on row r only, so the reference's "accumulate a full [N,d] gradient, then apply" is exactly a per-row
// update; there is no gradient buffer here.  One warp owns one row; lane a-slices own output coordinates
// a = lane, lane+32, ...; inv_cov is read through its per-block TRANSPOSE so that the 32 lanes read 128 contiguous
// bytes for every b.  Exactly-zero pi[r,k] (sklearn's predict_proba is one-hot to fp32 on separated data) are skipped:
// they contribute exactly 0.  Arithmetic mirrors the oracle (oracle_o3_batch): (float)(pi*S) products, double dot
// over b in index order, float accumulation over communities.
//
// Not HBM-bound: 2*K_live*d^2 flop per row against L2-resident inv_cov (K*d*d*4 B: 3.3 MB at K=50, d=128).
#include "comemb_common.cuh"

namespace {

constexpr int O3_WARPS = 8;
constexpr int O3_MAX_SLICES = 16;  // size <= 512

__global__ void __launch_bounds__(O3_WARPS * 32)
    o3_batch_kernel(float *node, int64_t n_rows, int size, const uint32_t *rows, int64_t n_sel, const float *mu,
                    const float *inv_cov_t, const float *pi, const int32_t *comm, const float *weight, int K,
                    float scale, float lr, int iters) {
    extern __shared__ float smem[];
    float *diff = smem + (size_t)(threadIdx.x >> 5) * size;  // this warp's (x - mu_k)
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * O3_WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * O3_WARPS;
    const int n_slices = (size + 31) / 32;
    for (int64_t s = warp0; s < n_sel; s += n_warps) {
        const int64_t r = rows ? (int64_t)rows[s] : s;
        float *x = node + r * size;
        // dense pi row, or the top-1 form (one community + weight per row; comm < 0: no community)
        const float *p_row = pi ? pi + r * K : nullptr;
        const int only = pi ? -1 : comm[r];
        const float only_w = pi ? 0.f : weight[r];
        for (int it = 0; it < iters; it++) {
            float grad[O3_MAX_SLICES];
#pragma unroll
            for (int m = 0; m < O3_MAX_SLICES; m++) grad[m] = 0.f;
            for (int k = (pi ? 0 : max(only, 0)); k < (pi ? K : only + 1); k++) {
                const float p = pi ? p_row[k] : only_w;
                if (p == 0.f) continue;
                __syncwarp();
                for (int b = lane; b < size; b += 32) diff[b] = x[b] - mu[(int64_t)k * size + b];  // :68
                __syncwarp();
                const float *St = inv_cov_t + (int64_t)k * size * size;
#pragma unroll
                for (int m = 0; m < O3_MAX_SLICES; m++) {
                    if (m >= n_slices) break;
                    const int a = lane + 32 * m;
                    if (a < size) {
                        double t = 0.0;
                        for (int b = 0; b < size; b++)  // :69-71
                            t = __dadd_rn(t, __dmul_rn((double)__fmul_rn(p, St[(int64_t)b * size + a]), (double)diff[b]));
                        grad[m] = grad[m] + __double2float_rn(t);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < O3_MAX_SLICES; m++) {
                if (m >= n_slices) break;
                const int a = lane + 32 * m;
                if (a < size) {
                    float g = __fmul_rn(grad[m], scale);                 // :76
                    g = fminf(fmaxf(g, -5.f), 5.f);                      // :77 clip
                    x[a] = x[a] - __fmul_rn(g, lr);
                }
            }
            __syncwarp();
        }
    }
}

__global__ void transpose_blocks_kernel(const float *in, float *out, int size) {
    __shared__ float tile[32][33];
    const float *src = in + (int64_t)blockIdx.z * size * size;
    float *dst = out + (int64_t)blockIdx.z * size * size;
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (x < size && y0 + j < size) tile[j][threadIdx.x] = src[(int64_t)(y0 + j) * size + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        if (x < size && y0 + j < size) dst[(int64_t)(y0 + j) * size + x] = tile[threadIdx.x][j];
}

__global__ void scale_kernel(float *x, int64_t n, float s) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    float4 *x4 = reinterpret_cast<float4 *>(x);
    for (int64_t q = i; q < n4; q += stride) {
        float4 v = x4[q];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        x4[q] = v;
    }
    for (int64_t q = n4 * 4 + i; q < n; q += stride) x[q] *= s;
}

// sum over window pairs of -log(sigmoid(x_j . c_i)), exact sigmoid, double accumulation; one warp per walk
__global__ void __launch_bounds__(256) o2_pos_loss_kernel(const float *node, const float *ctx, int size,
                                                          const uint32_t *walks, const int64_t *walk_off,
                                                          int64_t n_walks, int window, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    double loss = 0.0, pairs = 0.0;
    for (int64_t w = warp0; w < n_walks; w += (int64_t)gridDim.x * 8) {
        const uint32_t *path = walks + walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
        for (int i = 0; i < len; i++) {
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int j1 = min(len, i + window + 1);
            for (int j = max(0, i - window); j < j1; j++) {
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                double acc = 0.0;
                for (int e = lane; e < size; e += 32)
                    acc += (double)node[(int64_t)wj * size + e] * (double)ctx[(int64_t)wi * size + e];
                for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                // -log(sigmoid(z)) = log1p(exp(-z)) computed stably
                loss += acc > 0 ? log1p(exp(-acc)) : (-acc + log1p(exp(acc)));
                pairs += 1.0;
            }
        }
    }
    if (lane == 0 && pairs > 0) {
        atomicAdd(out, loss);
        atomicAdd(out + 1, pairs);
    }
}

}  // namespace

int launch_o3_batch(float *node, int64_t n_rows, int size, const uint32_t *rows, int64_t n_sel, const float *mu,
                    const float *inv_cov_t, const float *pi, const int32_t *comm, const float *weight, int K,
                    double beta, float lr, int iters, cudaStream_t st) {
    if (size > 32 * O3_MAX_SLICES) return COMEMB_E_UNSUPPORTED;
    if (n_sel <= 0 || iters <= 0) return 0;
    const float scale = (float)(beta / (double)K);  // numpy: float32 array *= python float (beta/k)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (n_sel + O3_WARPS - 1) / O3_WARPS;
    int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    size_t smem = (size_t)O3_WARPS * size * sizeof(float);
    o3_batch_kernel<<<grid, O3_WARPS * 32, smem, st>>>(node, n_rows, size, rows, n_sel, mu, inv_cov_t, pi, comm, weight, K, scale, lr,
                                                       iters);
    return (int)cudaGetLastError();
}

int launch_transpose_blocks(const float *in, float *out, int K, int size, cudaStream_t st) {
    if (K <= 0 || size <= 0) return 0;
    dim3 grid((size + 31) / 32, (size + 31) / 32, K), block(32, 8);
    transpose_blocks_kernel<<<grid, block, 0, st>>>(in, out, size);
    return (int)cudaGetLastError();
}

int launch_scale(float *x, int64_t n, float s, cudaStream_t st) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return COMEMB_E_ARG;
    scale_kernel<<<148 * 8, 256, 0, st>>>(x, n, s);
    return (int)cudaGetLastError();
}

int launch_o2_pos_loss(const float *node, const float *ctx, int size, const uint32_t *walks, const int64_t *walk_off,
                       int64_t n_walks, int window, double *out, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(out, 0, 2 * sizeof(double), st));
    if (n_walks <= 0) return 0;
    int64_t want = (n_walks + 7) / 8;
    int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    o2_pos_loss_kernel<<<grid, 256, 0, st>>>(node, ctx, size, walks, walk_off, n_walks, window, out);
    return (int)cudaGetLastError();
}
