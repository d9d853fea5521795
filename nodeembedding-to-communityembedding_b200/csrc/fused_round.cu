// fused_round.cu -- HOGWILD form of the legacy fused pass (stale train_sg: per window pair the o3 community gradient of
// x_j, the SGNS pair update and the combined write; utils/training_sdg_inner.c:1597-1905, 2520-2715, 2988-3740) at size
// 128 as ONE persistent, CTA-cooperative kernel whose o3 half runs on the 5th-generation tensor cores.
//
// Why a new structure: the o3 term is a 128x128 mat-vec per pair against the inverse covariance of x_j's community.
// A window holds ~4.5 different communities (SBM, 50 blocks), so a warp that owns a walk can batch only ~4 rows per
// community and has to stream 64 KB of inv_cov from L2 for each of them (round 1: 15 KB of L2 traffic per pair, twice
// the SGNS part, through registers into mma.sync).  Here the batching is done ACROSS the ~3500 walks in flight:
//
//   one CTA per SM (cooperative launch), every warp owns a walk and advances through it one centre per ROUND:
//   [stage]  the warp lists the rows of its next window; every row whose o3 term can be taken from the row's value at
//            the start of the centre (= every row except a node that occupies two window positions) is appended to
//            the request list of its community: (row, slot in the result buffer [, pi weight]);
//   -- grid barrier --
//   [gemm]   the CTAs split the request lists into tiles of <= 64 rows of one community.  Per tile: inv_cov_c is the
//            resident tcgen05 A operand in shared memory (hi/lo TF32 images, 128 KB, fetched by the TMA engine only
//            when the CTA moves to another community); all warps gather the rows, form x - mu_c, split hi/lo and
//            store the swizzled B operand; one thread issues 48 tcgen05.mma (3xTF32 = fp32-level accuracy, accumulator
//            in TMEM); the warps read the accumulator back and write w * Y to the rows' slots (L2-resident buffer);
//   -- grid barrier --
//   [sgns]   every warp runs the SGNS pairs of its centre in position order exactly like the o2 kernel (size-128
//            specialisation: LCG jump constants, samples fetched one pair ahead, transposed 8-slot reduction,
//            lane-parallel sigma), with g = (label - sigma) * lr, context += g*lambda1*x_j (unless is_node_embedding)
//            and the write  x_j = fma(lambda1, work, x_j) + clip(-lambda2 * w * Y_j, +-0.1*lr);  a row that repeats
//            a node of the same window gets its o3 term in-warp from the CURRENT value (fp32 FMAs against the L2-resident
//            inv_cov), so that a walk sees exactly the reference's sequential semantics.
//
// Per pair this moves 512 B (row gather) + 512 B (result write) + 512 B (result read) through L2 on top of the SGNS
// traffic, instead of ~15 KB, and the inv_cov traffic drops to 128 KB per (CTA, community switch).
// Any pi: top-1 form (community + weight per row: plain result stores) or dense rows (one request per non-zero
// responsibility, results accumulated with red.add and cleared by the consumer); negative 1..7; window span <= 64;
// is_node_embedding 0/1; window shrinking; atomic or plain scatter.
#include <cstdio>
#include <cstdlib>

#include "comemb_common.cuh"
#include "fused_sgns.cuh"
#include "umma.cuh"

int launch_umma_prep_a(const float *P, char *out, int K, cudaStream_t st);

namespace {

constexpr int D = 128;
constexpr int TN = 64;                       // rows per GEMM tile (tcgen05 N)
constexpr int VMAX = 64;                     // window rows per centre (2*window <= 64)
constexpr int A_IMG_BYTES = 2 * D * D * 4;   // hi + lo operand images of one community
constexpr int KMAX = 1024;                   // communities (job scan lives in shared memory)
constexpr long long BARRIER_TIMEOUT = 6000000000LL;  // ~3 s of SM clocks: a protocol bug must end the kernel, not hang the GPU

struct RoundParams {
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
    const char *a_img;
    const int32_t *comm;   // top-1 form (per table row) ...
    const float *weight;
    const float *pi;       // ... or dense [n_rows, K] (comm == nullptr)
    int K, window, is_node;
    float lr, lambda1, lambda2;
    const float *glut;
    // scratch
    float *ybuf;           // [total warps * vslots][128]
    uint32_t *brow, *bslot;
    float *bw;             // dense pi only
    int64_t cap;           // entries per community list
    int *count;            // [2][K + 1]: per-community request counts of the round; [K] = number of staged warps
    unsigned *bar;         // grid barrier: {arrivals, generation}
    unsigned long long *walk_cursor;
    int *err;
    long long *stats;      // optional (debug): block 0 accumulates clock64 cycles per phase {bar1, gemm, bar2, sgns, stage, rounds}
    int vslots;            // result slots per warp (>= 2*window)
    int64_t active_warps;  // Hogwild concurrency cap (warps that take walks)
};

template <int NW>
struct RoundSmem {
    static constexpr int A_HI = 0, A_LO = D * D * 4, B_HI = 2 * D * D * 4, B_LO = B_HI + TN * D * 4;
    static constexpr int MU = B_LO + TN * D * 4;          // float[128]
    static constexpr int ROW = MU + D * 4;                // uint32[TN] (tile: result slots)
    static constexpr int WGT = ROW + TN * 4;              // float[TN]
    static constexpr int LUT = WGT + TN * 4;              // float[1000]
    static constexpr int SCAN = LUT + 4096;               // int[KMAX + 1] exclusive tile prefix
    static constexpr int WTOK = SCAN + (KMAX + 1) * 4 + 12;  // per warp: uint32 tok[VMAX]
    static constexpr int WINF = WTOK + NW * VMAX * 4;     // per warp: int32 info[VMAX] (community or -1, bit 30 = in-warp o3)
    static constexpr int BAR = (WINF + NW * VMAX * 4 + 15) & ~15;
    static constexpr int TOTAL = BAR + 64;
};

using fused::INFO_INWARP;

// ---- grid-wide barrier (all CTAs are co-resident: cooperative launch) ------------------------------------------------------------
// Returns false when the barrier timed out (another CTA died): the caller leaves its loops and tears down.
__device__ __forceinline__ bool grid_sync(unsigned *bar, unsigned n_ctas, int *err, int *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        int ok = 1;
        if (atomicAdd(bar, 1u) == n_ctas - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (*gen == g) {
                __nanosleep(40);
                if (clock64() - t0 > BARRIER_TIMEOUT) {
                    *err = 1;
                    ok = 0;
                    break;
                }
            }
        }
        __threadfence();
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}

template <bool ATOMIC, int NEG, int NW>
__global__ void __launch_bounds__(NW * 32, 1) sg_round_kernel(const RoundParams P) {
    using L = RoundSmem<NW>;
    constexpr LcgJump<NEG> J{};
    constexpr uint32_t TMEM_COLS = 64;
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *mu_s = reinterpret_cast<float *>(smem + L::MU);
    uint32_t *slot_s = reinterpret_cast<uint32_t *>(smem + L::ROW);
    float *wgt_s = reinterpret_cast<float *>(smem + L::WGT);
    float *lut = reinterpret_cast<float *>(smem + L::LUT);
    int *scan_s = reinterpret_cast<int *>(smem + L::SCAN);
    uint64_t *bar_a = reinterpret_cast<uint64_t *>(smem + L::BAR), *bar_mma = bar_a + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_a + 2);
    int *s_flag = reinterpret_cast<int *>(bar_a + 3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *tokS = reinterpret_cast<uint32_t *>(smem + L::WTOK) + warp * VMAX;
    int32_t *infS = reinterpret_cast<int32_t *>(smem + L::WINF) + warp * VMAX;

    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
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

    const int W = P.window, K = P.K;
    const bool dense = P.comm == nullptr;
    const bool o3_on = P.lambda2 != 0.f;
    const float lr = P.lr, lambda1 = P.lambda1;
    const float clipv = __double2float_rn(__dmul_rn((double)P.lr, 0.1));  // c:2556
    const float nl2 = -P.lambda2;                                         // c:3132
    const bool is_node = P.is_node != 0;
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
    const int64_t gwarp = (int64_t)blockIdx.x * NW + warp;
    const bool walker = gwarp < P.active_warps;
    const int64_t slot0 = gwarp * P.vslots;
    fused::SgnsArgs SA;
    SA.node = P.node; SA.ctx = P.ctx; SA.table = P.table; SA.mod = P.mod; SA.mu = P.mu; SA.inv_cov = P.inv_cov;
    SA.weight = P.weight; SA.pi = P.pi; SA.ybuf = P.ybuf; SA.K = K; SA.dense = dense; SA.o3_on = o3_on;
    SA.is_node = is_node; SA.lr = lr; SA.lambda1 = lambda1; SA.nl2 = nl2; SA.clipv = clipv;

    // ---- per-warp walk state (registers; the kernel is persistent) ---------------------------------------------------------
    const uint32_t *path = nullptr;
    const int32_t *rwp = nullptr;
    int len = 0, ci = -1, V = 0;   // ci: current centre; V: rows of its window
    uint32_t wi = 0;               // centre token
    uint64_t rnd = 0;
    uint32_t tnext = 0;
    bool have = false;             // a staged centre is waiting for its sgns phase
    bool exhausted = !walker;

    // advance to the next centre with a non-empty window and stage it (requests into count/lists of parity `par`)
    auto stage_next = [&](int par) {
        have = false;
        while (!exhausted) {
            ci++;
            if (ci >= len) {  // next walk
                unsigned long long w = 0;
                if (lane == 0) w = atomicAdd(P.walk_cursor, 1ULL);
                w = __shfl_sync(FULL, w, 0);
                if ((int64_t)w >= P.n_walks) {
                    exhausted = true;
                    break;
                }
                const int64_t o0 = __ldg(P.walk_off + w), o1 = __ldg(P.walk_off + w + 1);
                path = P.walks + o0;
                rwp = P.rw ? P.rw + o0 : nullptr;
                len = (int)min((int64_t)MAX_SENTENCE_LEN, o1 - o0);
                rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
                tnext = (lane < NEG) ? __ldg(P.table + table_slot((myA * rnd + myC) & LCG_MASK, P.mod)) : 0u;
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                ci = -1;
                continue;
            }
            wi = __ldg(path + ci);
            if (wi == COMEMB_TOKEN_NONE) continue;
            const int r = rwp ? rwp[ci] : 0;
            const int ja = max(0, ci - W + r), jb = min(len, ci + W + 1 - r);
            // window rows in position order (span <= 2W+1 <= 65 positions: three lane passes cover it)
            int v = 0;
            __syncwarp();
            for (int base = ja; base < jb; base += 32) {
                const int jl = base + lane;
                uint32_t tk = COMEMB_TOKEN_NONE;
                if (jl < jb && jl != ci) tk = __ldg(path + jl);
                const bool valid = tk != COMEMB_TOKEN_NONE;
                const unsigned vm = __ballot_sync(FULL, valid);
                if (valid) tokS[v + __popc(vm & ((1u << lane) - 1u))] = tk;
                v += __popc(vm);
            }
            V = v;
            if (V == 0) continue;
            __syncwarp();
            if (o3_on) {
                for (int base = 0; base < V; base += 32) {
                    const int vv = base + lane;
                    const bool mine = vv < V;
                    const uint32_t tk = mine ? tokS[vv] : 0u;
                    bool dup = false;  // an earlier window position holds the same node: o3 from the current value, in-warp
                    for (int u = 0; u < vv && mine; u++) dup = dup || (tokS[u] == tk);
                    if (!dense) {
                        int c = -1;
                        float wgt = 0.f;
                        if (mine) {
                            c = __ldg(P.comm + tk);
                            wgt = c >= 0 ? __ldg(P.weight + tk) : 0.f;
                            if (c >= K || wgt == 0.f) c = -1;
                        }
                        const int key = (mine && c >= 0 && !dup) ? c : -1 - lane;  // unique negative keys for non-requests
                        const unsigned peers = __match_any_sync(FULL, key);
                        if (key >= 0) {
                            const int leader = __ffs(peers) - 1;
                            int basep = 0;
                            if (lane == leader) basep = atomicAdd(P.count + par * (K + 1) + c, __popc(peers));
                            basep = __shfl_sync(peers, basep, leader);
                            const int64_t at = (int64_t)c * P.cap + basep + __popc(peers & ((1u << lane) - 1u));
                            P.brow[at] = tk;
                            P.bslot[at] = (uint32_t)(slot0 + vv);
                        }
                        if (mine) infS[vv] = c < 0 ? -1 : (dup ? (c | INFO_INWARP) : c);
                    } else if (mine) {
                        infS[vv] = dup ? INFO_INWARP : 0;
                    }
                }
                if (dense) {  // one request per non-zero responsibility of every non-repeated row
                    __syncwarp();
                    for (int vv = 0; vv < V; vv++) {
                        if (infS[vv] & INFO_INWARP) continue;
                        const uint32_t tk = tokS[vv];
                        for (int k = lane; k < K; k += 32) {
                            const float p = __ldg(P.pi + (int64_t)tk * K + k);
                            if (p != 0.f) {
                                const int64_t at = (int64_t)k * P.cap + atomicAdd(P.count + par * (K + 1) + k, 1);
                                P.brow[at] = tk;
                                P.bslot[at] = (uint32_t)(slot0 + vv);
                                P.bw[at] = p;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            have = true;
            break;
        }
        if (have && lane == 0) atomicAdd(P.count + par * (K + 1) + K, 1);
    };

    int par = 0;
    stage_next(par);
    int cur_c = -1;
    uint32_t par_a = 0, par_m = 0;
    bool a_pending = false;
    bool ok = true;

    long long t_prev = clock64(), acc[6] = {0, 0, 0, 0, 0, 0};
    auto lap = [&](int slot) {
        const long long now = clock64();
        acc[slot] += now - t_prev;
        t_prev = now;
    };
    while (true) {
        if (!(ok = grid_sync(P.bar, gridDim.x, P.err, s_flag))) break;
        lap(0);
        const int *cnt_g = P.count + par * (K + 1);
        if (__ldcg(cnt_g + K) == 0) break;  // no warp staged a centre: every walk is finished
        // ---- [gemm] -----------------------------------------------------------------------------------------------------------
        if (blockIdx.x == 0)  // clear the other parity's counters for the staging that follows this round's sgns phase
            for (int e = threadIdx.x; e <= K; e += blockDim.x) P.count[(par ^ 1) * (K + 1) + e] = 0;
        if (o3_on) {
            // exclusive prefix of tiles per community (warp 0, 32 communities per step)
            if (warp == 0) {
                int carry = 0;
                for (int base = 0; base < K; base += 32) {
                    const int c = base + lane;
                    const int t = c < K ? (__ldcg(cnt_g + c) + TN - 1) / TN : 0;
                    int incl = t;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int up = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += up;
                    }
                    if (c < K) scan_s[c] = carry + incl - t;
                    carry += __shfl_sync(FULL, incl, 31);
                }
                if (lane == 0) scan_s[K] = carry;
            }
            __syncthreads();
            const int n_jobs = scan_s[K];
            const int j0 = (int)((int64_t)n_jobs * blockIdx.x / gridDim.x);
            const int j1 = (int)((int64_t)n_jobs * (blockIdx.x + 1) / gridDim.x);
            int c = 0;
            constexpr int RPW = (TN + NW - 1) / NW;
            for (int j = j0; j < j1; j++) {
                // community of job j: last c with scan_s[c] <= j (counts can be zero: skip empty communities)
                {
                    int lo = c, hi = K - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (scan_s[mid] <= j) lo = mid; else hi = mid - 1;
                    }
                    c = lo;
                }
                const int start = (j - scan_s[c]) * TN;
                const int cnt = min(TN, __ldcg(cnt_g + c) - start);
                const int n16 = (cnt + 15) & ~15;
                const int64_t lbase = (int64_t)c * P.cap + start;
                // rows of the tile: issue the gathers before anything that waits
                uint32_t rowv[RPW], slotv[RPW];
                float4 xv[RPW];
#pragma unroll
                for (int q = 0; q < RPW; q++) {
                    const int r = warp + q * NW;
                    rowv[q] = r < cnt ? __ldcg(P.brow + lbase + r) : 0u;
                    slotv[q] = r < cnt ? __ldcg(P.bslot + lbase + r) : 0u;
                }
#pragma unroll
                for (int q = 0; q < RPW; q++) {
                    const int r = warp + q * NW;
                    if (r < cnt) xv[q] = __ldcg(reinterpret_cast<const float4 *>(node_l + (int64_t)rowv[q] * D));
                }
                if (c != cur_c) {  // every MMA that read the resident A has completed (bar_mma is waited on per tile)
                    if (threadIdx.x == 0) {
                        umma::mbar_expect_tx(bar_a, A_IMG_BYTES);
                        const char *src = P.a_img + (int64_t)c * A_IMG_BYTES;
#pragma unroll
                        for (int q = 0; q < 8; q++) umma::bulk_g2s(smem + L::A_HI + q * 16384, src + q * 16384, 16384, bar_a);
                    }
                    if (warp == 1)
                        *reinterpret_cast<float4 *>(mu_s + 4 * lane) =
                            __ldg(reinterpret_cast<const float4 *>(P.mu + (int64_t)c * D + 4 * lane));
                    a_pending = true;
                    cur_c = c;
                    __syncthreads();
                }
                const float4 m = *reinterpret_cast<const float4 *>(mu_s + 4 * lane);
#pragma unroll
                for (int q = 0; q < RPW; q++) {
                    const int r = warp + q * NW;
                    if (r < cnt) {
                        const float4 x = xv[q];
                        const float4 df = make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
                        const float4 hi = make_float4(umma::tf32_round(df.x), umma::tf32_round(df.y),
                                                      umma::tf32_round(df.z), umma::tf32_round(df.w));
                        const float4 lo = make_float4(umma::tf32_round(df.x - hi.x), umma::tf32_round(df.y - hi.y),
                                                      umma::tf32_round(df.z - hi.z), umma::tf32_round(df.w - hi.w));
                        const uint32_t off = umma::sw128_offset(TN, r, 4 * lane);
                        *reinterpret_cast<float4 *>(smem + L::B_HI + off) = hi;
                        *reinterpret_cast<float4 *>(smem + L::B_LO + off) = lo;
                        if (lane == 0) {
                            slot_s[r] = slotv[q];
                            if (dense) wgt_s[r] = __ldcg(P.bw + lbase + r);
                        }
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
                umma::mbar_wait(bar_mma, par_m);
                par_m ^= 1;
                umma::tc_fence_after();
                // epilogue: thread = output coordinate a (TMEM lane), 16 rows of the tile at a time
                const int a = 32 * (warp & 3) + lane;
                for (int ch = warp >> 2; ch * 16 < n16; ch += NW / 4) {
                    float v[16];
                    umma::tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ch * 16), v);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const int n = ch * 16 + q;
                        if (n < cnt) {
                            float *yp = P.ybuf + (int64_t)slot_s[n] * D + a;
                            if (dense)
                                atomicAdd(yp, __fmul_rn(wgt_s[n], v[q]));  // one slot receives the terms of several communities
                            else
                                *yp = v[q];  // top-1 form: the consumer applies the responsibility
                        }
                    }
                }
                umma::tc_fence_before();
                __syncthreads();  // accumulator, B images and the tile's slot list are free again
            }
        }
        __syncthreads();
        lap(1);
        if (!(ok = grid_sync(P.bar, gridDim.x, P.err, s_flag))) break;
        lap(2);
        // ---- [sgns] of the staged centre, then stage the next one -------------------------------------------------------------
        if (have)
            fused::sgns_centre<ATOMIC, NEG>(SA, wi, V, tokS, infS, lut, slot0, rnd, tnext, myA, myC, lane);
        par ^= 1;
        lap(3);
        stage_next(par);
        lap(4);
        acc[5]++;
    }
    if (P.stats && blockIdx.x == 0 && (threadIdx.x & 31) == 0)  // per warp of block 0: bar1 gemm bar2 sgns stage rounds
        for (int q = 0; q < 6; q++) P.stats[warp * 6 + q] = acc[q];
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, TMEM_COLS);
}

// pi one-hot? -> top-1 form (community, weight); *flag |= 1 when a row has several non-zero responsibilities
__global__ void pi_top1_kernel(const float *__restrict__ pi, int64_t n, int K, int32_t *__restrict__ comm,
                               float *__restrict__ weight, int *flag) {
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
    if (cnt > 1) atomicOr(flag, 1);
}

template <int NEG>
cudaError_t launch_round_t(const RoundParams &P, bool atomic, int grid, cudaStream_t st) {
    constexpr int NW = 24;
    const int smem = RoundSmem<NW>::TOTAL + 1024;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        void *args[] = {const_cast<RoundParams *>(&P)};
        return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kernel), dim3(grid), dim3(NW * 32), args, (size_t)smem, st);
    };
    return atomic ? go(sg_round_kernel<true, NEG, NW>) : go(sg_round_kernel<false, NEG, NW>);
}

}  // namespace

// dense pi -> top-1 form; *d_flag (zeroed by the caller) receives 1 when a row has several non-zero responsibilities
int fused_pi_to_top1(const float *pi, int64_t n_rows, int K, int32_t *comm, float *weight, int *d_flag, cudaStream_t st) {
    if (n_rows <= 0) return COMEMB_E_ARG;
    pi_top1_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(pi, n_rows, K, comm, weight, d_flag);
    return (int)cudaGetLastError();
}

// Returns COMEMB_E_UNSUPPORTED when the shape does not fit this kernel (the caller falls back to the generic one).
int launch_sg_fused_round(float *node, float *negemb, const uint32_t *walks, const int64_t *walk_off, int64_t n_walks,
                          const int32_t *reduced_windows, const uint64_t *seeds, uint64_t base_seed, const uint32_t *table,
                          uint64_t table_len, const float *mu, const float *inv_cov, const float *pi, int K, int window,
                          int negative, float lr, float lambda1, float lambda2, int is_node_embedding, bool atomic,
                          int64_t n_rows, const int32_t *top1_comm, const float *top1_weight, cudaStream_t st) {
    constexpr int NW = 24;
    if (negative < 1 || negative > 7 || window < 1 || 2 * window > VMAX) return COMEMB_E_UNSUPPORTED;
    if (lambda2 != 0.f && (K < 1 || K > KMAX || !mu || !inv_cov || (!pi && !top1_comm))) return COMEMB_E_UNSUPPORTED;
    if (!is_node_embedding && negemb == node) return COMEMB_E_UNSUPPORTED;  // context updates would alias the cached rows
    if (n_walks <= 0) return 0;
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return COMEMB_E_UNSUPPORTED;
    const bool o3_on = lambda2 != 0.f;
    if (!o3_on) K = 1;
    int64_t warps = (int64_t)sms * NW;
    if (comemb_opts().max_warps > 0 && comemb_opts().max_warps < warps) warps = comemb_opts().max_warps;
    if (n_walks < warps) warps = n_walks;
    const int grid = (int)((warps + NW - 1) / NW);
    const int64_t total_warps = (int64_t)grid * NW;
    const int vslots = 2 * window;
    const int64_t cap = total_warps * vslots;  // every request of a round could fall into one community
    const bool dense = o3_on && !top1_comm;
    // pi dense but one-hot -> top-1 form on the fly (plain stores instead of red.add accumulation)
    char *scratch = nullptr;
    const size_t sz_y = (size_t)total_warps * vslots * D * 4;
    const size_t sz_list = o3_on ? (size_t)K * cap * 4 : 0;
    const size_t sz_img = o3_on ? (size_t)K * A_IMG_BYTES : 0;
    const size_t sz_top1 = dense ? (size_t)n_rows * 8 + 16 : 0;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        const size_t at = off;
        off = (off + bytes + 1023) & ~(size_t)1023;
        return at;
    };
    const size_t o_ctl = carve(256 + (size_t)2 * (K + 1) * 4), o_y = carve(sz_y), o_row = carve(sz_list), o_slot = carve(sz_list);
    const size_t o_w = carve(dense ? sz_list : 0), o_img = carve(sz_img), o_t1 = carve(sz_top1);
    if (off > ((size_t)24 << 30)) return COMEMB_E_UNSUPPORTED;
    CUDA_TRY(cudaMallocAsync(&scratch, off, st));
    auto fail = [&](cudaError_t e) {
        cudaFreeAsync(scratch, st);
        return (int)e;
    };
    cudaError_t e = cudaMemsetAsync(scratch + o_ctl, 0, 256 + (size_t)2 * (K + 1) * 4, st);
    if (e != cudaSuccess) return fail(e);
    const int32_t *comm = top1_comm;
    const float *weight = top1_weight;
    const float *pi_dense = nullptr;
    if (dense) {
        if (n_rows <= 0) return fail(cudaErrorInvalidValue);
        int32_t *c = reinterpret_cast<int32_t *>(scratch + o_t1 + 16);
        float *w = reinterpret_cast<float *>(scratch + o_t1 + 16 + (size_t)n_rows * 4);
        int *flag = reinterpret_cast<int *>(scratch + o_t1), h_flag = 0;
        if ((e = cudaMemsetAsync(flag, 0, 4, st)) != cudaSuccess) return fail(e);
        pi_top1_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(pi, n_rows, K, c, w, flag);
        if ((e = cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
        if (h_flag == 0) {
            comm = c;
            weight = w;
        } else {
            pi_dense = pi;
            if ((e = cudaMemsetAsync(scratch + o_y, 0, sz_y, st)) != cudaSuccess) return fail(e);  // red.add accumulators
        }
    }
    if (o3_on) {
        const int r = launch_umma_prep_a(inv_cov, scratch + o_img, K, st);
        if (r) return fail((cudaError_t)r);
    }
    RoundParams P;
    P.node = node; P.ctx = negemb; P.walks = walks; P.walk_off = walk_off; P.n_walks = n_walks; P.rw = reduced_windows;
    P.seeds = seeds; P.base_seed = base_seed; P.table = table; P.mod = make_table_mod(table_len);
    P.mu = mu; P.inv_cov = inv_cov; P.a_img = scratch + o_img; P.comm = comm; P.weight = weight; P.pi = pi_dense;
    P.K = K; P.window = window; P.is_node = is_node_embedding; P.lr = lr; P.lambda1 = lambda1; P.lambda2 = lambda2;
    P.glut = comemb_lut_device();
    P.ybuf = reinterpret_cast<float *>(scratch + o_y);
    P.brow = reinterpret_cast<uint32_t *>(scratch + o_row); P.bslot = reinterpret_cast<uint32_t *>(scratch + o_slot);
    P.bw = reinterpret_cast<float *>(scratch + o_w); P.cap = cap;
    P.bar = reinterpret_cast<unsigned *>(scratch + o_ctl);
    P.walk_cursor = reinterpret_cast<unsigned long long *>(scratch + o_ctl + 16);
    P.err = reinterpret_cast<int *>(scratch + o_ctl + 32);
    P.count = reinterpret_cast<int *>(scratch + o_ctl + 256);
    P.vslots = vslots; P.active_warps = warps;
    static const bool want_stats = getenv("COMEMB_ROUND_STATS") != nullptr;
    long long *d_stats = nullptr;
    if (want_stats) {
        cudaMalloc(&d_stats, NW * 6 * sizeof(long long));
        cudaMemset(d_stats, 0, NW * 6 * sizeof(long long));
    }
    P.stats = d_stats;
    switch (negative) {
        case 1: e = launch_round_t<1>(P, atomic, grid, st); break;
        case 2: e = launch_round_t<2>(P, atomic, grid, st); break;
        case 3: e = launch_round_t<3>(P, atomic, grid, st); break;
        case 4: e = launch_round_t<4>(P, atomic, grid, st); break;
        case 5: e = launch_round_t<5>(P, atomic, grid, st); break;
        case 6: e = launch_round_t<6>(P, atomic, grid, st); break;
        default: e = launch_round_t<7>(P, atomic, grid, st); break;
    }
    if (want_stats) {
        long long h[NW * 6];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(d_stats);
        for (int w = 0; w < NW; w += 5)
            fprintf(stderr, "[round stats] block 0 warp %2d: rounds %lld  cycles/round: bar1 %lld gemm %lld bar2 %lld sgns %lld stage %lld\n",
                    w, h[w * 6 + 5], h[w * 6 + 0] / (h[w * 6 + 5] + 1), h[w * 6 + 1] / (h[w * 6 + 5] + 1),
                    h[w * 6 + 2] / (h[w * 6 + 5] + 1), h[w * 6 + 3] / (h[w * 6 + 5] + 1), h[w * 6 + 4] / (h[w * 6 + 5] + 1));
    }
    cudaFreeAsync(scratch, st);
    return (int)e;
}
