// sgns_flow.cu -- ORDERED mode of o2 executed as a dataflow graph: the reference's sequential update stream, bit for
// bit, on thousands of warps.
//
// The reference's single worker (context_embeddings.py:83-84 -> pyx:454-509) applies the pair updates of walk 0, then
// walk 1, ... ; pair p+1 sees every write of pair p.  Which rows a pair touches does not depend on any floating-point
// value: the node row of walk[j], the context row of walk[i] and `negative` rows drawn from a per-walk LCG that
// advances exactly `negative` steps per pair (pyx:128-136).  Two pair updates that touch different rows commute
// exactly, so the sequential result is fixed by the ORDER OF THE TOUCHES OF EACH ROW alone.  This file
//   1. enumerates every (row, pair) touch of a chunk of walks in stream order              flow_enumerate_kernel
//   2. stable-sorts the touches by row (cub radix sort) and gives each touch its ticket =
//      the number of earlier touches of the same row                                        flow_mark/assign_kernel
//   3. replays the walks on one warp each; a pair starts when served[row] has reached the
//      ticket of each of its touches and adds one to served[row] when it is done with it    o2_flow_d128_kernel
// Walks are claimed in increasing order from a counter, and a walk only ever waits for touches of EARLIER walks, so
// the earliest unfinished walk can always run: no deadlock for any number of resident warps.  The arithmetic of a
// pair is the code of o2_ordered_d128_kernel (sgns_ordered.cu), so tables are bit-identical to the one-warp replay,
// to oracle/comemb_oracle.c and to the reference's golden vectors (tests/test_gpu_parity.py::test_o2_flow_*).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "comemb_common.cuh"
#include "ordered_d128.cuh"

namespace {
using namespace ordered;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// rows are written by other SMs while this kernel runs: read them at L2, never from this SM's L1
__device__ __forceinline__ Row4 ldcg_row4(const float *row, int lane) {
    Row4 r;
    r.v0 = __ldcg(row + lane); r.v1 = __ldcg(row + lane + 32); r.v2 = __ldcg(row + lane + 64); r.v3 = __ldcg(row + lane + 96);
    return r;
}

// ---- 1. pairs per walk (pyx:494-507) and non-None tokens (pyx:490) ---------------------------------------------------
__global__ void flow_count_kernel(const uint32_t *walks, const int64_t *walk_off, int64_t n_walks, int window,
                                  int64_t *pairs, int64_t *n_tokens) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_walks) return;
    if (w == n_walks) {
        pairs[w] = 0;
        return;
    }
    const uint32_t *path = walks + walk_off[w];
    const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
    int64_t cnt = 0, tok = 0;
    for (int i = 0; i < len; i++) {
        if (path[i] == COMEMB_TOKEN_NONE) continue;
        tok++;
        const int j1 = min(len, i + window + 1);
        for (int j = max(0, i - window); j < j1; j++) cnt += (j != i && path[j] != COMEMB_TOKEN_NONE);
    }
    pairs[w] = cnt;
    if (n_tokens && tok) atomicAdd((unsigned long long *)n_tokens, (unsigned long long)tok);
}

// ---- the touches of a chunk of walks in stream order: touch (pair p, slot s) has index (pair_off[w]-pair_off[w0]+p)*T+s;
// slot 0 = node row of walk[j], slot 1 = context row of walk[i], slot 2+k = context row of sample k (none if it equals
// walk[i], pyx:135-136).  Keys: node row r -> r, context row r -> n_rows + r, none -> 2*n_rows.
template <int NEG>
__global__ void __launch_bounds__(256)
    flow_enumerate_kernel(const uint32_t *walks, const int64_t *walk_off, int64_t w0, int64_t w1, const uint64_t *seeds,
                          uint64_t base_seed, Sampler S, int window, uint32_t n_rows, const int64_t *pair_off,
                          uint32_t *keys, uint32_t *vals) {
    constexpr int T = NEG + 2;
    constexpr LcgJump<NEG> J{};
    const int lane = threadIdx.x & 31;
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k + 2) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t p0 = pair_off[w0];
    for (int64_t w = w0 + gwarp; w < w1; w += nwarps) {
        const uint32_t *path = walks + walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
        uint64_t rnd = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        int64_t t = (pair_off[w] - p0) * T + lane;
        for (int i = 0; i < len; i++) {
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int j1 = min(len, i + window + 1);
            for (int j = max(0, i - window); j < j1; j++) {
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                uint32_t key = wj;
                if (lane == 1) key = n_rows + wi;
                if (lane >= 2 && lane < T) {
                    const uint32_t smp = S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)];
                    key = smp == wi ? 2u * n_rows : n_rows + smp;
                }
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                if (lane < T) {
                    keys[t] = key;
                    vals[t] = (uint32_t)t;
                }
                t += T;
            }
        }
    }
}

// ---- 2. tickets: position of a touch among the (stream-ordered) touches of its row -------------------------------------
__global__ void flow_mark_kernel(const uint32_t *keys_sorted, int64_t n, uint32_t n_keys, uint32_t *start) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = keys_sorted[s];
    if (k < n_keys && (s == 0 || keys_sorted[s - 1] != k)) start[k] = (uint32_t)s;
}
__global__ void flow_assign_kernel(const uint32_t *keys_sorted, const uint32_t *vals_sorted, int64_t n, uint32_t n_keys,
                                   const uint32_t *start, uint32_t *ticket) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = keys_sorted[s];
    ticket[vals_sorted[s]] = k < n_keys ? (uint32_t)s - start[k] : 0u;
}

// ---- 3. the replay: one warp per walk, rows in registers, pair arithmetic of o2_ordered_d128_kernel ----------------------
template <int NEG>
__global__ void __launch_bounds__(256)
    o2_flow_d128_kernel(float *node, float *ctx, const uint32_t *walks, const int64_t *walk_off, int64_t w0, int64_t w1,
                        const uint64_t *seeds, uint64_t base_seed, Sampler S, int window, float lr, float lambda,
                        bool quirk, uint32_t n_rows, const int64_t *pair_off, const uint32_t *ticket, uint32_t *served,
                        unsigned long long *cursor, const float *g_exp_table) {
    constexpr int D = 128;
    constexpr int T = NEG + 2;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = g_exp_table[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr uint32_t smp_lanes = ((1u << NEG) - 1u) << 2;  // lanes 2 .. NEG+1 hold the samples
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k + 2) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int64_t p0 = pair_off[w0];
    for (;;) {
        unsigned long long wq = 0;
        if (lane == 0) wq = atomicAdd(cursor, 1ULL);  // walks are claimed in stream order
        const int64_t w = w0 + (int64_t)__shfl_sync(FULL, wq, 0);
        if (w >= w1) break;
        const uint32_t *path = walks + walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, walk_off[w + 1] - walk_off[w]);
        uint64_t rnd = seeds ? seeds[w] : (splitmix64(base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        const uint32_t *tk_ptr = ticket + (pair_off[w] - p0) * T + lane;
        const bool smp_lane = lane >= 2 && lane < T;
        uint32_t tnext = smp_lane ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
        for (int i = 0; i < len; i++) {  // pyx:494
            const uint32_t wi = path[i];
            if (wi == COMEMB_TOKEN_NONE) continue;
            float *pos_ptr = ctx + (int64_t)wi * D;
            Row4 cpos = {0.f, 0.f, 0.f, 0.f};
            bool held = false;        // the positive context row is held (and kept in registers) across the window
            uint32_t pos_ticket = 0;  // lane 1: ticket of the window's latest touch of it
            const int j1 = min(len, i + window + 1);
            for (int j = max(0, i - window); j < j1; j++) {  // pyx:503
                const uint32_t wj = path[j];
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                const uint32_t tk = lane < T ? __ldg(tk_ptr) : 0u;
                tk_ptr += T;
                const uint32_t tmine = tnext;
                tnext = smp_lane ? S.table[table_slot((myA * rnd + myC) & LCG_MASK, S.mod)] : 0u;
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                // this lane's touch: counter, whether it exists, and how many earlier lanes of this pair touch the same row
                const bool touches = lane == 0 || (lane == 1) || (smp_lane && tmine != wi);
                const uint32_t key = lane == 0 ? wj : n_rows + (lane == 1 ? wi : tmine);
                const uint32_t same = __match_any_sync(FULL, touches ? key : 0xFFFFFFFFu - lane) & smp_lanes;
                const bool dup_lane = smp_lane && touches;
                const uint32_t need = tk - (dup_lane ? __popc(same & lt_mask) : 0u);
                const bool last_of_row = !dup_lane || (same >> (lane + 1)) == 0u;
                if (lane == 1) pos_ticket = tk;
                uint32_t *cnt = served + key;
                bool ok = !touches || (lane == 1 && held) || ld_relaxed_u32(cnt) >= need;
                while (!__all_sync(FULL, ok)) {
                    if (!ok) {
                        __nanosleep(64);
                        ok = ld_relaxed_u32(cnt) >= need;
                    }
                }
                fence_acq_rel();
                __syncwarp();
                float *row1_ptr = node + (int64_t)wj * D;
                const Row4 x = ldcg_row4(row1_ptr, lane);
                if (!held) {
                    cpos = ldcg_row4(pos_ptr, lane);
                    held = true;
                }
                uint32_t t[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) t[k] = __shfl_sync(FULL, tmine, k + 2);
                bool anydup = false;
#pragma unroll
                for (int k = 1; k < NEG; k++)
#pragma unroll
                    for (int a = 0; a < k; a++) anydup = anydup || (t[a] == t[k]);
                Row4 work = {0.f, 0.f, 0.f, 0.f};  // pyx:126
                {                                  // positive target, pyx:129-131
                    const float f = dot128_refblas(x, cpos, quirk);
                    if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                        const float g = __fmul_rn(__fmul_rn(1.f - lut[lut_index(f)], lr), lambda);
                        fma_row4(work, g, cpos);  // pyx:146
                        fma_row4(cpos, g, x);     // pyx:147 (registers; stored when the centre ends)
                    }
                }
                if (!anydup) {
                    Row4 c[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) c[k] = ldcg_row4(ctx + (int64_t)t[k] * D, lane);
                    float f[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) f[k] = dot128_refblas(x, c[k], quirk);
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        if (t[k] == wi) continue;                                // pyx:135-136 (row not held: value unused)
                        if (f[k] <= -MAX_EXP_F || f[k] >= MAX_EXP_F) continue;  // pyx:141-142
                        const float g = __fmul_rn(__fmul_rn(0.f - lut[lut_index(f[k])], lr), lambda);
                        fma_row4(work, g, c[k]);
                        fma_row4(c[k], g, x);
                        st_row4(ctx + (int64_t)t[k] * D, lane, c[k]);
                    }
                } else {  // equal samples inside one pair: one after the other, re-reading the row
#pragma unroll 1
                    for (int k = 0; k < NEG; k++) {
                        const uint32_t tkk = __shfl_sync(FULL, tmine, k + 2);
                        if (tkk == wi) continue;
                        float *cp = ctx + (int64_t)tkk * D;
                        Row4 c = ldcg_row4(cp, lane);
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
                __syncwarp();
                if (touches && lane != 1 && last_of_row) st_release_u32(cnt, tk + 1u);
            }
            if (held) {
                st_row4(pos_ptr, lane, cpos);
                __syncwarp();
                if (lane == 1) st_release_u32(served + n_rows + wi, pos_ticket + 1u);
            }
        }
    }
}

}  // namespace

// ---- launcher ------------------------------------------------------------------------------------------------------------
// Returns COMEMB_E_UNSUPPORTED when the shape is not handled here (the caller falls back to the one-stream kernels).
int launch_o2_flow(float *node, float *ctx, int64_t n_rows, const uint32_t *walks, const int64_t *walk_off,
                   int64_t n_walks, const uint64_t *seeds, uint64_t base_seed, const uint32_t *table, uint64_t table_len,
                   int window, int negative, float lr, float lambda, bool quirk, int64_t *n_tokens, int max_warps,
                   cudaStream_t st) {
    if (negative < 1 || negative > 7 || n_rows <= 0 || 2 * n_rows + 1 > 0x7FFFFFFFLL) return COMEMB_E_UNSUPPORTED;
    if (n_walks <= 0) return 0;
    Sampler S{table, make_table_mod(table_len)};
    const int T = negative + 2;
    const uint32_t n_keys = (uint32_t)(2 * n_rows);
    int end_bit = 1;
    while ((1ULL << end_bit) <= n_keys) end_bit++;

    int64_t *pair_off = nullptr;
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0;
    CUDA_TRY(cudaMallocAsync(&pair_off, (size_t)(n_walks + 1) * sizeof(int64_t), st));
    flow_count_kernel<<<(unsigned)((n_walks + 1 + 255) / 256), 256, 0, st>>>(walks, walk_off, n_walks, window, pair_off,
                                                                              n_tokens);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, pair_off, pair_off, n_walks + 1, st);
    std::vector<int64_t> off_h((size_t)n_walks + 1);
    cudaError_t e = cudaMallocAsync(&scan_tmp, scan_bytes ? scan_bytes : 1, st);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, pair_off, pair_off, n_walks + 1, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(off_h.data(), pair_off, off_h.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (scan_tmp) cudaFreeAsync(scan_tmp, st);
    if (e != cudaSuccess) {
        cudaFreeAsync(pair_off, st);
        return (int)e;
    }

    // chunks of whole walks with at most `cap` touches (a single longer walk forms its own chunk)
    int64_t cap = 1LL << 27;
    if (const char *e = getenv("COMEMB_FLOW_CHUNK_TOUCHES")) cap = std::max<int64_t>(1, atoll(e));  // tests: force several chunks
    int64_t biggest = 0;
    for (int64_t a = 0; a < n_walks;) {
        int64_t b = a + 1;
        while (b < n_walks && (off_h[b + 1] - off_h[a]) * T <= cap) b++;
        biggest = std::max(biggest, (off_h[b] - off_h[a]) * T);
        a = b;
    }
    if (biggest >= (1LL << 31)) {
        cudaFreeAsync(pair_off, st);
        return COMEMB_E_UNSUPPORTED;
    }
    int rc = 0;
    uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr, *start = nullptr,
             *served = nullptr;
    unsigned long long *cursor = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_bytes = 0;
    if (biggest > 0) {
        const size_t nb = (size_t)biggest * sizeof(uint32_t);
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)biggest, 0, end_bit,
                                        st);
        if (cudaMallocAsync(&keys_in, nb, st) || cudaMallocAsync(&keys_out, nb, st) || cudaMallocAsync(&vals_in, nb, st) ||
            cudaMallocAsync(&vals_out, nb, st) || cudaMallocAsync(&start, (size_t)n_keys * 4, st) ||
            cudaMallocAsync(&served, (size_t)n_keys * 4, st) || cudaMallocAsync(&cursor, 8, st) ||
            cudaMallocAsync(&sort_tmp, sort_bytes ? sort_bytes : 1, st))
            rc = (int)cudaGetLastError();
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int64_t a = 0; rc == 0 && biggest > 0 && a < n_walks;) {
        int64_t b = a + 1;
        while (b < n_walks && (off_h[b + 1] - off_h[a]) * T <= cap) b++;
        const int64_t n_touch = (off_h[b] - off_h[a]) * T;
        if (n_touch > 0) {
            int64_t warps = std::min<int64_t>(b - a, max_warps > 0 ? max_warps : (int64_t)sms * 32);
            const unsigned blocks = (unsigned)((warps + 7) / 8);
            cudaMemsetAsync(served, 0, (size_t)n_keys * 4, st);
            cudaMemsetAsync(cursor, 0, 8, st);
            const unsigned eb = (unsigned)std::min<int64_t>((b - a + 7) / 8, (int64_t)sms * 8);
            const unsigned tb = (unsigned)((n_touch + 255) / 256);
            switch (negative) {
#define COMEMB_CASE(N)                                                                                                  \
    case N:                                                                                                             \
        flow_enumerate_kernel<N><<<eb, 256, 0, st>>>(walks, walk_off, a, b, seeds, base_seed, S, window,               \
                                                     (uint32_t)n_rows, pair_off, keys_in, vals_in);                     \
        cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)n_touch, 0,   \
                                        end_bit, st);                                                                   \
        flow_mark_kernel<<<tb, 256, 0, st>>>(keys_out, n_touch, n_keys, start);                                         \
        flow_assign_kernel<<<tb, 256, 0, st>>>(keys_out, vals_out, n_touch, n_keys, start, vals_in);                   \
        o2_flow_d128_kernel<N><<<blocks, 256, 0, st>>>(node, ctx, walks, walk_off, a, b, seeds, base_seed, S, window,  \
                                                       lr, lambda, quirk, (uint32_t)n_rows, pair_off, vals_in, served,  \
                                                       cursor, comemb_lut_device());                                    \
        break;
                COMEMB_CASE(1) COMEMB_CASE(2) COMEMB_CASE(3) COMEMB_CASE(4) COMEMB_CASE(5) COMEMB_CASE(6) COMEMB_CASE(7)
#undef COMEMB_CASE
            }
            rc = (int)cudaGetLastError();
        }
        a = b;
    }
    void *bufs[] = {pair_off, keys_in, keys_out, vals_in, vals_out, start, served, cursor, sort_tmp};
    for (void *p : bufs)
        if (p) cudaFreeAsync(p, st);
    return rc;
}
