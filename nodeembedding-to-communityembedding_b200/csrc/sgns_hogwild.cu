// sgns_hogwild.cu -- HOGWILD mode of o1 / o2 on sm_100a: lock-free SGD, one warp per reference "worker job".
//
// Mapping (Context2Vec.train's worker threads, context_embeddings.py:72-98 -> warps):
//   * one warp owns one walk (or one centre-chunk of a walk) and replays it sequentially exactly like train_o2
//     (pyx:494-508): same pair order, same LCG stream (chunks start at the right stream position via O(log n) LCG
//     skip-ahead, so the negative-sample indices are those the reference would draw from the same seed);
//   * walks run concurrently and race on shared rows without locks, as the reference's threads do.
// Data movement per pair (the roofline's 7168 B at size=128, neg=5): the node row and neg+1 context rows are
// gathered with one 128-bit load per lane per row (a 512 B row = one fully coalesced warp request), all negative
// rows of a pair are in flight together, dots are reduced with xor-shuffles, sigma comes from the 1000-entry table
// staged in shared memory, rows are scattered back with 128-bit stores (or red.global.add.v4.f32 with
// COMEMB_F_ATOMIC).  The positive context row of a centre stays in registers across its whole window.
// Table rows are loaded with ld.global.cg (L2-coherent): other warps' updates are seen as soon as they reach L2.
#include "comemb_common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int NEGB = 5;  // negatives gathered per batch (neg=5 -> one batch)


template <int NCH>
struct Row {
    float v[4 * NCH];
};

template <int NCH, bool VEC>
__device__ __forceinline__ Row<NCH> row_load(const float *p, int d, int lane) {
    Row<NCH> r;
#pragma unroll
    for (int m = 0; m < NCH; m++) {
        const int base = 128 * m + 4 * lane;
        if (VEC) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base < d) q = ldcg4(p + base);
            r.v[4 * m + 0] = q.x; r.v[4 * m + 1] = q.y; r.v[4 * m + 2] = q.z; r.v[4 * m + 3] = q.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) r.v[4 * m + c] = (base + c < d) ? __ldcg(p + base + c) : 0.f;
        }
    }
    return r;
}

template <int NCH, bool VEC>
__device__ __forceinline__ void row_store(float *p, const Row<NCH> &r, int d, int lane) {
#pragma unroll
    for (int m = 0; m < NCH; m++) {
        const int base = 128 * m + 4 * lane;
        if (VEC) {
            if (base < d) st4(p + base, make_float4(r.v[4 * m], r.v[4 * m + 1], r.v[4 * m + 2], r.v[4 * m + 3]));
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (base + c < d) p[base + c] = r.v[4 * m + c];
        }
    }
}

template <int NCH, bool VEC>
__device__ __forceinline__ void row_red(float *p, const Row<NCH> &r, int d, int lane) {
#pragma unroll
    for (int m = 0; m < NCH; m++) {
        const int base = 128 * m + 4 * lane;
        if (VEC) {
            if (base < d) red_add4(p + base, make_float4(r.v[4 * m], r.v[4 * m + 1], r.v[4 * m + 2], r.v[4 * m + 3]));
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (base + c < d) atomicAdd(p + base + c, r.v[4 * m + c]);
        }
    }
}

// lane-local part of the canonical dot: fma chain over the lane's elements in increasing order, from +0
template <int NCH>
__device__ __forceinline__ float dot_part(const Row<NCH> &a, const Row<NCH> &b) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 4 * NCH; e++) acc = fmaf(a.v[e], b.v[e], acc);
    return acc;
}

template <int NCH>
__device__ __forceinline__ void row_fma(Row<NCH> &y, float a, const Row<NCH> &x) {  // y += a*x
#pragma unroll
    for (int e = 0; e < 4 * NCH; e++) y.v[e] = fmaf(a, x.v[e], y.v[e]);
}

template <int NCH>
__device__ __forceinline__ Row<NCH> row_scaled(float a, const Row<NCH> &x) {
    Row<NCH> r;
#pragma unroll
    for (int e = 0; e < 4 * NCH; e++) r.v[e] = __fmul_rn(a, x.v[e]);
    return r;
}

template <int NCH>
__device__ __forceinline__ Row<NCH> row_zero() {
    Row<NCH> r;
#pragma unroll
    for (int e = 0; e < 4 * NCH; e++) r.v[e] = 0.f;
    return r;
}

struct Draw {
    const uint32_t *table;
    TableMod mod;
    const uint32_t *alias;  // {threshold, alias} pairs or nullptr
    uint32_t n_alias;
};

// One negative drawn from LCG state r (pyx:133): the table slot, or with the alias sampler the (bucket, coin) pair
// taken from the same 32 random bits.
__device__ __forceinline__ uint32_t draw_fetch(const Draw &D, uint64_t r) {
    if (D.alias) {
        const uint64_t a = (r >> 16) * (uint64_t)D.n_alias;  // 32-bit uniform x n -> bucket in the high word
        const uint32_t bucket = (uint32_t)(a >> 32), coin = (uint32_t)a;
        const uint2 e = __ldg(reinterpret_cast<const uint2 *>(D.alias) + bucket);
        return coin < e.x ? bucket : e.y;
    }
    return __ldg(D.table + table_slot(r, D.mod));
}

__device__ __forceinline__ float sgns_g(float f, float label, float lr, float lambda, const float *lut) {
    return __fmul_rn(__fmul_rn(label - lut[lut_index(f)], lr), lambda);  // pyx:143-144
}

// ---- o2 -------------------------------------------------------------------------------------------------------------------
struct O2Params {
    float *node, *ctx;
    int d;
    const uint32_t *walks;
    const int64_t *walk_off;
    int64_t n_walks;
    const uint64_t *seeds;
    uint64_t base_seed;
    Draw draw;
    int window, negative;
    float lr, lambda;
    int centres_per_unit, units_per_walk;
    int64_t *n_tokens;
    const float *glut;
    // row-partitioned tables (n_shards > 0): rows [s*rows_per_shard, (s+1)*rows_per_shard) live in shard s, which may be
    // memory of a peer GPU mapped over NVLink (CUDA IPC); n_shards == 0: one flat table at node/ctx
    float *node_shard[8], *ctx_shard[8];
    uint32_t rows_per_shard;
    int n_shards;
};

template <int NCH, bool VEC, bool ATOMIC, int DFIX, int MINB = 1>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) o2_hogwild_kernel(const O2Params P) {
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int d = DFIX ? DFIX : P.d;
    const int W = P.window, negative = P.negative;
    const float lr = P.lr, lambda = P.lambda;
    float *const node = P.node, *const ctx = P.ctx;
    const int64_t n_units = P.n_walks * P.units_per_walk;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    for (int64_t u = warp0; u < n_units; u += n_warps) {
        const int64_t w = u / P.units_per_walk;
        const int q = (int)(u - w * P.units_per_walk);
        const uint32_t *path = P.walks + P.walk_off[w];
        int len = (int)min((int64_t)MAX_SENTENCE_LEN, P.walk_off[w + 1] - P.walk_off[w]);  // pyx:480
        const int c0 = P.centres_per_unit ? q * P.centres_per_unit : 0;
        const int c1 = P.centres_per_unit ? min(len, c0 + P.centres_per_unit) : len;
        if (c0 >= len) continue;
        uint64_t rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        if (q == 0 && P.n_tokens) {  // train_o2's return value (pyx:490)
            int cnt = 0;
            for (int i = lane; i < len; i += 32) cnt += (__ldg(path + i) != COMEMB_TOKEN_NONE);
            cnt = __reduce_add_sync(FULL, cnt);
            if (lane == 0 && cnt) atomicAdd(reinterpret_cast<unsigned long long *>(P.n_tokens), (unsigned long long)cnt);
        }
        if (c0 > 0) {  // position of this chunk in the walk's LCG stream: `negative` draws per preceding pair
            int pairs = 0;
            for (int i = lane; i < c0; i += 32) {
                if (__ldg(path + i) == COMEMB_TOKEN_NONE) continue;
                const int j1 = min(len, i + W + 1);
                for (int j = max(0, i - W); j < j1; j++) pairs += (j != i && __ldg(path + j) != COMEMB_TOKEN_NONE);
            }
            pairs = __reduce_add_sync(FULL, pairs);
            rnd = lcg_skip(rnd, (uint64_t)pairs * (uint64_t)negative);
        }

        for (int i = c0; i < c1; i++) {  // pyx:494
            const uint32_t wi = __ldg(path + i);
            if (wi == COMEMB_TOKEN_NONE) continue;
            float *pos_ptr = ctx + (int64_t)wi * d;
            Row<NCH> cpos = row_load<NCH, VEC>(pos_ptr, d, lane);
            Row<NCH> dpos = row_zero<NCH>();  // ATOMIC: accumulated delta of the positive row
            const int j1 = min(len, i + W + 1);
            for (int j = max(0, i - W); j < j1; j++) {  // pyx:503
                const uint32_t wj = __ldg(path + j);
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1_ptr = node + (int64_t)wj * d;
                const Row<NCH> row1 = row_load<NCH, VEC>(row1_ptr, d, lane);
                Row<NCH> work = row_zero<NCH>();
                bool pos_done = false;
                for (int kb = 0; kb < negative || !pos_done; kb += NEGB) {
                    const int nb = max(0, min(NEGB, negative - kb));
                    // draw nb negatives: every lane steps the LCG, lane k fetches sample k (pyx:133-134)
                    uint64_t mine = 0;
#pragma unroll
                    for (int k = 0; k < NEGB; k++)
                        if (k < nb) {
                            if (lane == k) mine = rnd;
                            rnd = lcg_next(rnd);
                        }
                    const uint32_t tmine = (lane < nb) ? draw_fetch(P.draw, mine) : wi;
                    uint32_t t[NEGB];
                    Row<NCH> c[NEGB];
                    unsigned actmask = pos_done ? 0u : 1u;  // bit 0 = the positive target (first batch only)
#pragma unroll
                    for (int k = 0; k < NEGB; k++) {
                        t[k] = __shfl_sync(FULL, tmine, k);
                        const bool a = (k < nb) && (t[k] != wi);  // pyx:135-136
                        if (a) {
                            actmask |= 2u << k;
                            c[k] = row_load<NCH, VEC>(ctx + (int64_t)t[k] * d, d, lane);
                        } else {
                            c[k] = row_zero<NCH>();
                        }
                    }
                    // an earlier sample of this batch hitting the same row must be seen by the later one
                    unsigned dupmask = 0;
#pragma unroll
                    for (int k = 1; k < NEGB; k++)
#pragma unroll
                        for (int a = 0; a < k; a++)
                            if (t[a] == t[k] && ((actmask >> (k + 1)) & (actmask >> (a + 1)) & 1u)) dupmask |= 1u << k;
                    // all dots of the batch at once: positive in slot 0, negative k in p_{k+1}
                    const float p0 = pos_done ? 0.f : dot_part<NCH>(row1, cpos);
                    const float fm = reduce8_transposed(p0, dot_part<NCH>(row1, c[0]), dot_part<NCH>(row1, c[1]),
                                                        dot_part<NCH>(row1, c[2]), dot_part<NCH>(row1, c[3]),
                                                        dot_part<NCH>(row1, c[4]), 0.f, 0.f, lane);
                    // this lane's slot holds p_i with lane_of_p(i) == (lane & 28): i = (b4<<2)|(b3<<1)|b2
                    const int pi = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
                    float gm = 0.f;
                    if (((actmask >> pi) & 1u) && fm > -MAX_EXP_F && fm < MAX_EXP_F)
                        gm = sgns_g(fm, pi == 0 ? 1.f : 0.f, lr, lambda, lut);  // pyx:141-144
                    if (!pos_done) {  // positive target (label 1), pyx:129-131
                        const float g = __shfl_sync(FULL, gm, lane_of_p(0));
                        if (g != 0.f) {
                            row_fma<NCH>(work, g, cpos);  // pyx:146
                            if (ATOMIC) row_fma<NCH>(dpos, g, row1);
                            row_fma<NCH>(cpos, g, row1);  // pyx:147 (kept in registers)
                        }
                        pos_done = true;
                    }
#pragma unroll
                    for (int k = 0; k < NEGB; k++) {
                        float g = __shfl_sync(FULL, gm, lane_of_p(k + 1));
                        if (dupmask & (1u << k)) {  // rare: redo this target against the refreshed row
#pragma unroll
                            for (int a = 0; a < k; a++)
                                if (t[a] == t[k] && ((actmask >> (a + 1)) & 1u)) c[k] = c[a];
                            const float fk = warp_sum_xor(dot_part<NCH>(row1, c[k]));
                            g = (fk > -MAX_EXP_F && fk < MAX_EXP_F) ? sgns_g(fk, 0.f, lr, lambda, lut) : 0.f;
                        }
                        if (g == 0.f) continue;  // inactive, skipped (pyx:135-136, 141-142) -- sigma is never 0 or 1
                        row_fma<NCH>(work, g, c[k]);  // pyx:146
                        float *cp = ctx + (int64_t)t[k] * d;
                        if (ATOMIC) {
                            row_red<NCH, VEC>(cp, row_scaled<NCH>(g, row1), d, lane);
                            if (dupmask) row_fma<NCH>(c[k], g, row1);
                        } else {
                            row_fma<NCH>(c[k], g, row1);  // pyx:147
                            row_store<NCH, VEC>(cp, c[k], d, lane);
                        }
                    }
                }
                // pyx:149
                if (ATOMIC) {
                    row_red<NCH, VEC>(row1_ptr, work, d, lane);
                } else {
                    Row<NCH> r = row1;
#pragma unroll
                    for (int e = 0; e < 4 * NCH; e++) r.v[e] = r.v[e] + work.v[e];
                    row_store<NCH, VEC>(row1_ptr, r, d, lane);
                }
            }
            if (ATOMIC)
                row_red<NCH, VEC>(pos_ptr, dpos, d, lane);
            else
                row_store<NCH, VEC>(pos_ptr, cpos, d, lane);
        }
    }
}

// ---- o2, headline shape: size == 128 (one float4 per lane), NEG negatives known at compile time ----------------------------
// Same semantics as o2_hogwild_kernel, leaner instruction stream:
//   * lane k < NEG jumps straight to its own LCG state (affine skip constants A_k, C_k), every lane then advances the
//     walk's state by NEG steps with one multiply-add;  the NEXT pair's table lookups are issued one pair ahead;
//   * all NEG+1 rows are loaded unconditionally (a dropped sample just gets g = 0), no per-row predicate moves;
//   * one transposed reduction + one lane-parallel sigma evaluation per pair;
//   * duplicate samples inside a pair (which must see each other's update, pyx:146-147) are detected with plain
//     warp-uniform compares (match.any kept the ADU pipe 53 % busy) and handled by a sequential path that re-reads
//     rows from memory;
//   * SHARDED: row r lives in shard r / rows_per_shard (possibly a peer GPU's memory, see sharded.py).
template <bool ATOMIC, int NEG, int MINB, bool HINT, bool SHARDED>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) o2_hogwild_d128_kernel(const O2Params P) {
    static_assert(NEG >= 1 && NEG <= 7, "positive + negatives must fit the 8 reduction slots");
    constexpr int D = 128;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int W = P.window;
    const float lr = P.lr, lambda = P.lambda;
    float *const node_l = P.node + 4 * lane, *const ctx_l = P.ctx + 4 * lane;  // this lane's float4 column
    // row r of a table -> this lane's 16 bytes of it.  SHARDED: the owning shard's base (possibly peer memory)
    auto NODE_ROW = [&](uint32_t r) -> float * {
        if (SHARDED) {
            const uint32_t sh = r / P.rows_per_shard;
            return P.node_shard[sh] + (int64_t)(r - sh * P.rows_per_shard) * D + 4 * lane;
        }
        return node_l + (int64_t)r * D;
    };
    auto CTX_ROW = [&](uint32_t r) -> float * {
        if (SHARDED) {
            const uint32_t sh = r / P.rows_per_shard;
            return P.ctx_shard[sh] + (int64_t)(r - sh * P.rows_per_shard) * D + 4 * lane;
        }
        return ctx_l + (int64_t)r * D;
    };
    const Draw draw = P.draw;
    const uint64_t pol_keep = HINT ? l2_policy_evict_last() : 0, pol_stream = HINT ? l2_policy_evict_first() : 0;
    auto LDROW = [&](const float *p) { return HINT ? ldcg4_hint(p, pol_keep) : __ldcg(reinterpret_cast<const float4 *>(p)); };
    auto STROW = [&](float *p, float4 v) { if (HINT) st4_hint(p, v, pol_keep); else st4(p, v); };
    auto LDTOK = [&](const uint32_t *p) { return HINT ? ldg_u32_hint(p, pol_stream) : __ldg(p); };
    auto FETCH = [&](uint64_t r) -> uint32_t {
        if (HINT && !draw.alias) return ldg_u32_hint(draw.table + table_slot(r, draw.mod), pol_stream);
        return draw_fetch(draw, r);
    };
    // this lane's jump constants (lanes >= NEG never use theirs)
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    // slot i of reduce8_transposed ends on lanes with (b4,b3,b2) == i; label 1 only for slot 0 (the positive)
    const int pi = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi == 0 ? 1.f : 0.f;
    const int64_t n_units = P.n_walks * P.units_per_walk;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    for (int64_t u = warp0; u < n_units; u += n_warps) {
        const int64_t w = u / P.units_per_walk;
        const int q = (int)(u - w * P.units_per_walk);
        const uint32_t *path = P.walks + P.walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, P.walk_off[w + 1] - P.walk_off[w]);  // pyx:480
        const int c0 = P.centres_per_unit ? q * P.centres_per_unit : 0;
        const int c1 = P.centres_per_unit ? min(len, c0 + P.centres_per_unit) : len;
        if (c0 >= len) continue;
        uint64_t rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        if (q == 0 && P.n_tokens) {  // train_o2's return value (pyx:490)
            int cnt = 0;
            for (int i = lane; i < len; i += 32) cnt += (__ldg(path + i) != COMEMB_TOKEN_NONE);
            cnt = __reduce_add_sync(FULL, cnt);
            if (lane == 0 && cnt) atomicAdd(reinterpret_cast<unsigned long long *>(P.n_tokens), (unsigned long long)cnt);
        }
        if (c0 > 0) {
            int pairs = 0;
            for (int i = lane; i < c0; i += 32) {
                if (__ldg(path + i) == COMEMB_TOKEN_NONE) continue;
                const int j1 = min(len, i + W + 1);
                for (int j = max(0, i - W); j < j1; j++) pairs += (j != i && __ldg(path + j) != COMEMB_TOKEN_NONE);
            }
            pairs = __reduce_add_sync(FULL, pairs);
            rnd = lcg_skip(rnd, (uint64_t)pairs * (uint64_t)NEG);
        }
        // samples of the NEXT pair to run, fetched one pair ahead (the draw depends only on the LCG stream, which
        // advances by exactly NEG per pair whatever the pair turns out to be)
        uint32_t tnext = (lane < NEG) ? FETCH((myA * rnd + myC) & LCG_MASK) : 0xFFFFFF00u + lane;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;

        for (int i = c0; i < c1; i++) {  // pyx:494
            const uint32_t wi = LDTOK(path + i);
            if (wi == COMEMB_TOKEN_NONE) continue;
            float *pos_ptr = CTX_ROW(wi);
            float4 cpos = LDROW(pos_ptr);
            float4 dpos = make_float4(0.f, 0.f, 0.f, 0.f);  // ATOMIC: accumulated delta of the positive row
            const int j1 = min(len, i + W + 1);
            for (int j = max(0, i - W); j < j1; j++) {  // pyx:503
                const uint32_t wj = LDTOK(path + j);
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1_ptr = NODE_ROW(wj);
                const float4 r1 = LDROW(row1_ptr);
                const uint32_t tmine = tnext;  // this pair's samples (lane k holds sample k)
                tnext = (lane < NEG) ? FETCH((myA * rnd + myC) & LCG_MASK) : 0xFFFFFF00u + lane;
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                uint32_t t[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) t[k] = __shfl_sync(FULL, tmine, k);
                // two equal samples in one pair?  (warp-uniform: every lane holds all NEG indices; plain compares --
                // match.any costs ~100 ADU-pipe cycles per pair here)
                bool anydup = false;
#pragma unroll
                for (int k = 1; k < NEG; k++)
#pragma unroll
                    for (int a = 0; a < k; a++) anydup = anydup || (t[a] == t[k]);
                float4 work = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!anydup) {
                    float4 c[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) c[k] = LDROW(CTX_ROW(t[k]));
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
                    // is my slot a live target?  slot 0 always; slot k+1 iff sample k != centre (pyx:135-136)
                    bool live = pi == 0;
#pragma unroll
                    for (int k = 0; k < NEG; k++) live = live || (pi == k + 1 && t[k] != wi);
                    float gm = 0.f;
                    if (live && fm > -MAX_EXP_F && fm < MAX_EXP_F)  // pyx:141-144
                        gm = __fmul_rn(__fmul_rn(my_label - lut[lut_index(fm)], lr), lambda);
                    {
                        const float g = __shfl_sync(FULL, gm, lane_of_p(0));
                        work.x = fmaf(g, cpos.x, work.x); work.y = fmaf(g, cpos.y, work.y);  // pyx:146
                        work.z = fmaf(g, cpos.z, work.z); work.w = fmaf(g, cpos.w, work.w);
                        if (ATOMIC) {
                            dpos.x = fmaf(g, r1.x, dpos.x); dpos.y = fmaf(g, r1.y, dpos.y);
                            dpos.z = fmaf(g, r1.z, dpos.z); dpos.w = fmaf(g, r1.w, dpos.w);
                        }
                        cpos.x = fmaf(g, r1.x, cpos.x); cpos.y = fmaf(g, r1.y, cpos.y);  // pyx:147 (registers)
                        cpos.z = fmaf(g, r1.z, cpos.z); cpos.w = fmaf(g, r1.w, cpos.w);
                    }
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        const float g = __shfl_sync(FULL, gm, lane_of_p(k + 1));
                        work.x = fmaf(g, c[k].x, work.x); work.y = fmaf(g, c[k].y, work.y);  // pyx:146
                        work.z = fmaf(g, c[k].z, work.z); work.w = fmaf(g, c[k].w, work.w);
                        float *cp = CTX_ROW(t[k]);
                        if (g != 0.f) {  // g == 0 <=> dropped or saturated target (sigma is never exactly 0 or 1)
                            if (ATOMIC) {
                                red_add4(cp, make_float4(__fmul_rn(g, r1.x), __fmul_rn(g, r1.y), __fmul_rn(g, r1.z),
                                                         __fmul_rn(g, r1.w)));
                            } else {  // pyx:147
                                STROW(cp, make_float4(fmaf(g, r1.x, c[k].x), fmaf(g, r1.y, c[k].y), fmaf(g, r1.z, c[k].z),
                                                      fmaf(g, r1.w, c[k].w)));
                            }
                        }
                    }
                } else {
                    // sequential path: every target re-reads its row after the previous target's write
                    {
                        const float f = warp_sum_xor(
                            fmaf(r1.w, cpos.w, fmaf(r1.z, cpos.z, fmaf(r1.y, cpos.y, fmaf(r1.x, cpos.x, 0.f)))));
                        if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                            const float g = sgns_g(f, 1.f, lr, lambda, lut);
                            work.x = fmaf(g, cpos.x, work.x); work.y = fmaf(g, cpos.y, work.y);
                            work.z = fmaf(g, cpos.z, work.z); work.w = fmaf(g, cpos.w, work.w);
                            if (ATOMIC) {
                                dpos.x = fmaf(g, r1.x, dpos.x); dpos.y = fmaf(g, r1.y, dpos.y);
                                dpos.z = fmaf(g, r1.z, dpos.z); dpos.w = fmaf(g, r1.w, dpos.w);
                            }
                            cpos.x = fmaf(g, r1.x, cpos.x); cpos.y = fmaf(g, r1.y, cpos.y);
                            cpos.z = fmaf(g, r1.z, cpos.z); cpos.w = fmaf(g, r1.w, cpos.w);
                        }
                    }
#pragma unroll 1
                    for (int k = 0; k < NEG; k++) {
                        const uint32_t tk = __shfl_sync(FULL, tmine, k);
                        if (tk == wi) continue;  // pyx:135-136
                        float *cp = CTX_ROW(tk);
                        const float4 c = LDROW(cp);
                        const float f = warp_sum_xor(fmaf(r1.w, c.w, fmaf(r1.z, c.z, fmaf(r1.y, c.y, fmaf(r1.x, c.x, 0.f)))));
                        if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;  // pyx:141-142
                        const float g = sgns_g(f, 0.f, lr, lambda, lut);
                        work.x = fmaf(g, c.x, work.x); work.y = fmaf(g, c.y, work.y);
                        work.z = fmaf(g, c.z, work.z); work.w = fmaf(g, c.w, work.w);
                        if (ATOMIC)
                            red_add4(cp, make_float4(__fmul_rn(g, r1.x), __fmul_rn(g, r1.y), __fmul_rn(g, r1.z),
                                                     __fmul_rn(g, r1.w)));
                        else
                            STROW(cp, make_float4(fmaf(g, r1.x, c.x), fmaf(g, r1.y, c.y), fmaf(g, r1.z, c.z),
                                                  fmaf(g, r1.w, c.w)));
                    }
                }
                // pyx:149
                if (ATOMIC)
                    red_add4(row1_ptr, work);
                else
                    STROW(row1_ptr, make_float4(r1.x + work.x, r1.y + work.y, r1.z + work.z, r1.w + work.w));
            }
            if (ATOMIC)
                red_add4(pos_ptr, dpos);
            else
                STROW(pos_ptr, cpos);
        }
    }
}

// ---- o2, sizes 64 and 256: the size-128 instruction stream with another row width ---------------------------------------------
// NV float4 per lane: size 256 = lane l owns elements 4l..4l+3 and 128+4l..128+4l+3 (the layout of the any-size kernel and
// of the oracle's warp-order model, so a single warp stays bit-exact); HALF: size 64 = lane l owns elements 2l, 2l+1 (one
// 64-bit access per lane, 256 contiguous bytes per row; the oracle models this order as DOT_WARP2).  Everything else as
// o2_hogwild_d128_kernel: compile-time NEG, LCG jump constants, samples one pair ahead, transposed reduction,
// lane-parallel sigma, duplicate samples on the sequential path.  (The any-size kernel spends ~650 warp-instructions per
// pair on run-time loops; this one ~330 x NV.)
template <int NV>
struct RowV {
    float4 v[NV];
};

// HALF: the lane's two elements travel in .x/.y, .z/.w stay +0 (fma(0, 0, acc) == acc: the sums are untouched)
template <int NV, bool HALF>
struct RowOps {
    static constexpr int D = HALF ? 64 : 128 * NV;
    static constexpr int LANE_STRIDE = HALF ? 2 : 4;  // elements between consecutive lanes' pointers
    static __device__ __forceinline__ RowV<NV> ld(const float *p) {
        RowV<NV> r;
        if (HALF) {
            const float2 t = __ldcg(reinterpret_cast<const float2 *>(p));
            r.v[0] = make_float4(t.x, t.y, 0.f, 0.f);
        } else {
#pragma unroll
            for (int m = 0; m < NV; m++) r.v[m] = ldcg4(p + 128 * m);
        }
        return r;
    }
    static __device__ __forceinline__ void st(float *p, const RowV<NV> &r) {
        if (HALF) {
            *reinterpret_cast<float2 *>(p) = make_float2(r.v[0].x, r.v[0].y);
        } else {
#pragma unroll
            for (int m = 0; m < NV; m++) st4(p + 128 * m, r.v[m]);
        }
    }
    static __device__ __forceinline__ void red(float *p, const RowV<NV> &r) {
        if (HALF) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(r.v[0].x), "f"(r.v[0].y) : "memory");
        } else {
#pragma unroll
            for (int m = 0; m < NV; m++) red_add4(p + 128 * m, r.v[m]);
        }
    }
    static __device__ __forceinline__ float dot(const RowV<NV> &a, const RowV<NV> &b) {  // element index ascending
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < NV; m++) {
            acc = fmaf(a.v[m].x, b.v[m].x, acc);
            acc = fmaf(a.v[m].y, b.v[m].y, acc);
            acc = fmaf(a.v[m].z, b.v[m].z, acc);
            acc = fmaf(a.v[m].w, b.v[m].w, acc);
        }
        return acc;
    }
    static __device__ __forceinline__ void fma(RowV<NV> &y, float g, const RowV<NV> &x) {  // y += g * x
#pragma unroll
        for (int m = 0; m < NV; m++) {
            y.v[m].x = fmaf(g, x.v[m].x, y.v[m].x); y.v[m].y = fmaf(g, x.v[m].y, y.v[m].y);
            y.v[m].z = fmaf(g, x.v[m].z, y.v[m].z); y.v[m].w = fmaf(g, x.v[m].w, y.v[m].w);
        }
    }
    static __device__ __forceinline__ RowV<NV> mul(float g, const RowV<NV> &x) {
        RowV<NV> r;
#pragma unroll
        for (int m = 0; m < NV; m++)
            r.v[m] = make_float4(__fmul_rn(g, x.v[m].x), __fmul_rn(g, x.v[m].y), __fmul_rn(g, x.v[m].z), __fmul_rn(g, x.v[m].w));
        return r;
    }
    static __device__ __forceinline__ RowV<NV> zero() {
        RowV<NV> r;
#pragma unroll
        for (int m = 0; m < NV; m++) r.v[m] = make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    static __device__ __forceinline__ RowV<NV> add(const RowV<NV> &a, const RowV<NV> &b) {
        RowV<NV> r;
#pragma unroll
        for (int m = 0; m < NV; m++)
            r.v[m] = make_float4(a.v[m].x + b.v[m].x, a.v[m].y + b.v[m].y, a.v[m].z + b.v[m].z, a.v[m].w + b.v[m].w);
        return r;
    }
};

template <bool ATOMIC, int NEG, int NV, bool HALF>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NV == 1 ? 3 : 2) o2_hogwild_dx_kernel(const O2Params P) {
    static_assert(NEG >= 1 && NEG <= 7 && (NV == 1 || NV == 2) && !(HALF && NV != 1), "shape");
    using R = RowOps<NV, HALF>;
    constexpr int D = R::D;
    constexpr LcgJump<NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int W = P.window;
    const float lr = P.lr, lambda = P.lambda;
    float *const node_l = P.node + (HALF ? 2 : 4) * lane, *const ctx_l = P.ctx + (HALF ? 2 : 4) * lane;
    const Draw draw = P.draw;
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int pi = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi == 0 ? 1.f : 0.f;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    for (int64_t w = warp0; w < P.n_walks; w += n_warps) {
        const uint32_t *path = P.walks + P.walk_off[w];
        const int len = (int)min((int64_t)MAX_SENTENCE_LEN, P.walk_off[w + 1] - P.walk_off[w]);  // pyx:480
        uint64_t rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
        if (P.n_tokens) {  // train_o2's return value (pyx:490)
            int cnt = 0;
            for (int i = lane; i < len; i += 32) cnt += (__ldg(path + i) != COMEMB_TOKEN_NONE);
            cnt = __reduce_add_sync(FULL, cnt);
            if (lane == 0 && cnt) atomicAdd(reinterpret_cast<unsigned long long *>(P.n_tokens), (unsigned long long)cnt);
        }
        uint32_t tnext = (lane < NEG) ? draw_fetch(draw, (myA * rnd + myC) & LCG_MASK) : 0xFFFFFF00u + lane;
        rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
        for (int i = 0; i < len; i++) {  // pyx:494
            const uint32_t wi = __ldg(path + i);
            if (wi == COMEMB_TOKEN_NONE) continue;
            float *pos_ptr = ctx_l + (int64_t)wi * D;
            RowV<NV> cpos = R::ld(pos_ptr), dpos = R::zero();
            const int j1 = min(len, i + W + 1);
            for (int j = max(0, i - W); j < j1; j++) {  // pyx:503
                const uint32_t wj = __ldg(path + j);
                if (j == i || wj == COMEMB_TOKEN_NONE) continue;
                float *row1_ptr = node_l + (int64_t)wj * D;
                const RowV<NV> r1 = R::ld(row1_ptr);
                const uint32_t tmine = tnext;
                tnext = (lane < NEG) ? draw_fetch(draw, (myA * rnd + myC) & LCG_MASK) : 0xFFFFFF00u + lane;
                rnd = (J.A[NEG] * rnd + J.C[NEG]) & LCG_MASK;
                uint32_t t[NEG];
#pragma unroll
                for (int k = 0; k < NEG; k++) t[k] = __shfl_sync(FULL, tmine, k);
                bool anydup = false;
#pragma unroll
                for (int k = 1; k < NEG; k++)
#pragma unroll
                    for (int a = 0; a < k; a++) anydup = anydup || (t[a] == t[k]);
                RowV<NV> work = R::zero();
                if (!anydup) {
                    RowV<NV> c[NEG];
#pragma unroll
                    for (int k = 0; k < NEG; k++) c[k] = R::ld(ctx_l + (int64_t)t[k] * D);
                    float p[8];
                    p[0] = R::dot(r1, cpos);
#pragma unroll
                    for (int k = 0; k < 7; k++) p[k + 1] = k < NEG ? R::dot(r1, c[k < NEG ? k : 0]) : 0.f;
                    const float fm = reduce8_transposed(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], lane);
                    bool live = pi == 0;
#pragma unroll
                    for (int k = 0; k < NEG; k++) live = live || (pi == k + 1 && t[k] != wi);
                    float gm = 0.f;
                    if (live && fm > -MAX_EXP_F && fm < MAX_EXP_F)  // pyx:141-144
                        gm = __fmul_rn(__fmul_rn(my_label - lut[lut_index(fm)], lr), lambda);
                    {
                        const float g = __shfl_sync(FULL, gm, lane_of_p(0));
                        R::fma(work, g, cpos);  // pyx:146
                        if (ATOMIC) R::fma(dpos, g, r1);
                        R::fma(cpos, g, r1);  // pyx:147 (registers)
                    }
#pragma unroll
                    for (int k = 0; k < NEG; k++) {
                        const float g = __shfl_sync(FULL, gm, lane_of_p(k + 1));
                        R::fma(work, g, c[k]);  // pyx:146
                        if (g != 0.f) {
                            float *cp = ctx_l + (int64_t)t[k] * D;
                            if (ATOMIC) {
                                R::red(cp, R::mul(g, r1));
                            } else {  // pyx:147
                                RowV<NV> nc = c[k];
                                R::fma(nc, g, r1);
                                R::st(cp, nc);
                            }
                        }
                    }
                } else {  // equal samples inside one pair: target by target, re-reading rows
                    {
                        const float f = warp_sum_xor(R::dot(r1, cpos));
                        if (f > -MAX_EXP_F && f < MAX_EXP_F) {
                            const float g = sgns_g(f, 1.f, lr, lambda, lut);
                            R::fma(work, g, cpos);
                            if (ATOMIC) R::fma(dpos, g, r1);
                            R::fma(cpos, g, r1);
                        }
                    }
#pragma unroll 1
                    for (int k = 0; k < NEG; k++) {
                        const uint32_t tk = __shfl_sync(FULL, tmine, k);
                        if (tk == wi) continue;  // pyx:135-136
                        float *cp = ctx_l + (int64_t)tk * D;
                        const RowV<NV> c = R::ld(cp);
                        const float f = warp_sum_xor(R::dot(r1, c));
                        if (f <= -MAX_EXP_F || f >= MAX_EXP_F) continue;  // pyx:141-142
                        const float g = sgns_g(f, 0.f, lr, lambda, lut);
                        R::fma(work, g, c);
                        if (ATOMIC) {
                            R::red(cp, R::mul(g, r1));
                        } else {
                            RowV<NV> nc = c;
                            R::fma(nc, g, r1);
                            R::st(cp, nc);
                        }
                    }
                }
                if (ATOMIC) {  // pyx:149
                    R::red(row1_ptr, work);
                } else {
                    RowV<NV> nr;
#pragma unroll
                    for (int m = 0; m < NV; m++)
                        nr.v[m] = make_float4(r1.v[m].x + work.v[m].x, r1.v[m].y + work.v[m].y, r1.v[m].z + work.v[m].z,
                                              r1.v[m].w + work.v[m].w);
                    R::st(row1_ptr, nr);
                }
            }
            if (ATOMIC)
                R::red(pos_ptr, dpos);
            else
                R::st(pos_ptr, cpos);
        }
    }
}

// ---- o1 -------------------------------------------------------------------------------------------------------------------
struct O1Params {
    float *node;
    int d;
    const uint32_t *edges;
    int64_t n_edges;
    const uint64_t *seeds;
    uint64_t base_seed;
    Draw draw;
    int negative;
    float lr;
    int64_t stride;
    const float *glut;
};

// One directed update (fast_o1, pyx:205-249) with row1 already in registers.  `other_idx/other` = a row this warp
// holds a newer copy of than global memory might (the row it updated a moment ago); samples equal to it use that copy.
template <int NCH, bool VEC>
__device__ __forceinline__ Row<NCH> o1_directed(const O1Params &P, const float *lut, int lane, uint32_t word_index,
                                                const Row<NCH> &target_row, const Row<NCH> &row1, uint32_t other_idx,
                                                const Row<NCH> &other, bool has_other, uint64_t &rnd) {
    const int d = P.d;
    Row<NCH> work = row_zero<NCH>();
    {
        const float f = warp_sum_xor(dot_part<NCH>(row1, target_row));
        if (f > -MAX_EXP_F && f < MAX_EXP_F) {
            const float g = __fmul_rn(1.f - lut[lut_index(f)], P.lr);  // pyx:243
            row_fma<NCH>(work, g, target_row);                         // pyx:245
        }
    }
    for (int kb = 0; kb < P.negative; kb += NEGB) {
        const int nb = min(NEGB, P.negative - kb);
        uint64_t mine = 0;
#pragma unroll
        for (int k = 0; k < NEGB; k++)
            if (k < nb) {
                if (lane == k) mine = rnd;
                rnd = lcg_next(rnd);
            }
        const uint32_t tmine = (lane < nb) ? draw_fetch(P.draw, mine) : 0u;
        Row<NCH> c[NEGB];
        bool act[NEGB];
#pragma unroll
        for (int k = 0; k < NEGB; k++) {
            const uint32_t t = __shfl_sync(FULL, tmine, k);
            act[k] = (k < nb) && (t != word_index);
            if (act[k]) {
                if (has_other && t == other_idx)
                    c[k] = other;
                else
                    c[k] = row_load<NCH, VEC>(P.node + (int64_t)t * d, d, lane);
            }
        }
        float f[NEGB];
#pragma unroll
        for (int k = 0; k < NEGB; k++) f[k] = act[k] ? dot_part<NCH>(row1, c[k]) : 0.f;
#pragma unroll
        for (int k = 0; k < NEGB; k++) f[k] = warp_sum_xor(f[k]);
#pragma unroll
        for (int k = 0; k < NEGB; k++) {
            if (!act[k] || f[k] <= -MAX_EXP_F || f[k] >= MAX_EXP_F) continue;
            const float g = __fmul_rn(0.f - lut[lut_index(f[k])], P.lr);
            row_fma<NCH>(work, g, c[k]);
        }
    }
    return work;
}

template <int NCH, bool VEC, bool ATOMIC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) o1_hogwild_kernel(const O1Params P) {
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int d = P.d;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    for (int64_t u = warp0; u < P.n_edges; u += n_warps) {
        const int64_t q = P.stride > 1 ? (int64_t)(((uint64_t)u * (uint64_t)P.stride) % (uint64_t)P.n_edges) : u;  // both < 2^32
        const uint32_t e0 = __ldg(P.edges + 2 * q), e1 = __ldg(P.edges + 2 * q + 1);
        uint64_t rnd = P.seeds ? P.seeds[q] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        float *p0 = P.node + (int64_t)e0 * d, *p1 = P.node + (int64_t)e1 * d;
        Row<NCH> r0 = row_load<NCH, VEC>(p0, d, lane);
        Row<NCH> r1 = (e1 == e0) ? r0 : row_load<NCH, VEC>(p1, d, lane);
        // pyx:444: row e0 against target e1
        Row<NCH> work = o1_directed<NCH, VEC>(P, lut, lane, e1, r1, r0, 0u, r0, false, rnd);
        if (ATOMIC) row_red<NCH, VEC>(p0, work, d, lane);
#pragma unroll
        for (int e = 0; e < 4 * NCH; e++) r0.v[e] = r0.v[e] + work.v[e];
        if (!ATOMIC) row_store<NCH, VEC>(p0, r0, d, lane);
        if (e1 == e0) r1 = r0;
        // pyx:447: row e1 against the UPDATED row e0 (held in registers)
        work = o1_directed<NCH, VEC>(P, lut, lane, e0, r0, r1, e0, r0, true, rnd);
        if (ATOMIC) {
            row_red<NCH, VEC>(p1, work, d, lane);
        } else {
#pragma unroll
            for (int e = 0; e < 4 * NCH; e++) r1.v[e] = r1.v[e] + work.v[e];
            row_store<NCH, VEC>(p1, r1, d, lane);
        }
    }
}

// ---- o1, headline shape: size == 128, NEG known at compile time -------------------------------------------------------------
// Same result as o1_hogwild_kernel with the o2 d=128 instruction diet: the 2*NEG samples of an edge are fetched with one
// load (lane k jumps straight to LCG state k), each directed update reduces its NEG+1 dots with one transposed 8-slot
// reduction and evaluates sigma once per lane; targets are read-only in o1 (pyx:245), so there is no duplicate path.
template <bool ATOMIC, int NEG>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 3) o1_hogwild_d128_kernel(const O1Params P) {
    static_assert(NEG >= 1 && NEG <= 7, "positive + negatives must fit the 8 reduction slots");
    constexpr int D = 128;
    constexpr LcgJump<2 * NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float lr = P.lr;
    float *const node_l = P.node + 4 * lane;
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < 2 * NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int pi = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi == 0 ? 1.f : 0.f;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    // one directed update (fast_o1): row x against `target` (label 1) and samples tt[0..NEG) (label 0); returns work
    auto directed = [&](const float4 &x, const float4 &target, uint32_t word_index, const uint32_t (&tt)[NEG],
                        const float4 (&c)[NEG]) -> float4 {
        float p[8];
        p[0] = fmaf(x.w, target.w, fmaf(x.z, target.z, fmaf(x.y, target.y, fmaf(x.x, target.x, 0.f))));
#pragma unroll
        for (int k = 0; k < 7; k++)
            p[k + 1] = k < NEG ? fmaf(x.w, c[k < NEG ? k : 0].w,
                                      fmaf(x.z, c[k < NEG ? k : 0].z,
                                           fmaf(x.y, c[k < NEG ? k : 0].y, fmaf(x.x, c[k < NEG ? k : 0].x, 0.f))))
                               : 0.f;
        const float fm = reduce8_transposed(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], lane);
        bool live = pi == 0;
#pragma unroll
        for (int k = 0; k < NEG; k++) live = live || (pi == k + 1 && tt[k] != word_index);  // pyx:234-235
        float gm = 0.f;
        if (live && fm > -MAX_EXP_F && fm < MAX_EXP_F) gm = __fmul_rn(my_label - lut[lut_index(fm)], lr);  // pyx:243
        float4 work = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const float g = __shfl_sync(FULL, gm, lane_of_p(0));
            work.x = fmaf(g, target.x, work.x); work.y = fmaf(g, target.y, work.y);
            work.z = fmaf(g, target.z, work.z); work.w = fmaf(g, target.w, work.w);
        }
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            const float g = __shfl_sync(FULL, gm, lane_of_p(k + 1));
            work.x = fmaf(g, c[k].x, work.x); work.y = fmaf(g, c[k].y, work.y);  // pyx:245
            work.z = fmaf(g, c[k].z, work.z); work.w = fmaf(g, c[k].w, work.w);
        }
        return work;
    };

    for (int64_t u = warp0; u < P.n_edges; u += n_warps) {
        const int64_t q = P.stride > 1 ? (int64_t)(((uint64_t)u * (uint64_t)P.stride) % (uint64_t)P.n_edges) : u;
        const uint32_t e0 = __ldg(P.edges + 2 * q), e1 = __ldg(P.edges + 2 * q + 1);
        const uint64_t rnd = P.seeds ? P.seeds[q] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        // all 2*NEG samples of the edge with one load: lane k draws sample k (pyx:232-233, one stream per edge)
        const uint32_t tmine = (lane < 2 * NEG) ? draw_fetch(P.draw, (myA * rnd + myC) & LCG_MASK) : 0u;
        float *p0 = node_l + (int64_t)e0 * D, *p1 = node_l + (int64_t)e1 * D;
        float4 r0 = __ldcg(reinterpret_cast<const float4 *>(p0));
        float4 r1 = (e1 == e0) ? r0 : __ldcg(reinterpret_cast<const float4 *>(p1));
        uint32_t ta[NEG], tb[NEG];
        float4 ca[NEG], cb[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            ta[k] = __shfl_sync(FULL, tmine, k);
            tb[k] = __shfl_sync(FULL, tmine, NEG + k);
            ca[k] = __ldcg(reinterpret_cast<const float4 *>(node_l + (int64_t)ta[k] * D));
            cb[k] = __ldcg(reinterpret_cast<const float4 *>(node_l + (int64_t)tb[k] * D));
        }
        // pyx:444: row e0 against target e1
        float4 work = directed(r0, r1, e1, ta, ca);
        if (ATOMIC) red_add4(p0, work);
        r0 = make_float4(r0.x + work.x, r0.y + work.y, r0.z + work.z, r0.w + work.w);
        if (!ATOMIC) st4(p0, r0);
        if (e1 == e0) r1 = r0;
        // pyx:447: row e1 against the UPDATED row e0; samples equal to e0 must see that update too
#pragma unroll
        for (int k = 0; k < NEG; k++)
            if (tb[k] == e0) cb[k] = r0;
        work = directed(r1, r0, e0, tb, cb);
        if (ATOMIC)
            red_add4(p1, work);
        else
            st4(p1, make_float4(r1.x + work.x, r1.y + work.y, r1.z + work.z, r1.w + work.w));
    }
}

// ---- o1, sizes 64 and 256: o1_hogwild_d128_kernel with another row width (RowOps, as o2_hogwild_dx_kernel) -------------------
template <bool ATOMIC, int NEG, int NV, bool HALF>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NV == 1 ? 3 : 2) o1_hogwild_dx_kernel(const O1Params P) {
    static_assert(NEG >= 1 && NEG <= 7 && (NV == 1 || NV == 2) && !(HALF && NV != 1), "shape");
    using R = RowOps<NV, HALF>;
    constexpr int D = R::D;
    constexpr LcgJump<2 * NEG> J{};
    __shared__ float lut[EXP_TABLE_SIZE];
    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float lr = P.lr;
    float *const node_l = P.node + R::LANE_STRIDE * lane;
    uint64_t myA = 1, myC = 0;
#pragma unroll
    for (int k = 0; k < 2 * NEG; k++)
        if (lane == k) {
            myA = J.A[k];
            myC = J.C[k];
        }
    const int pi = ((lane >> 4) & 1) << 2 | ((lane >> 3) & 1) << 1 | ((lane >> 2) & 1);
    const float my_label = pi == 0 ? 1.f : 0.f;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    // one directed update (fast_o1): row x against `target` (label 1) and samples tt[0..NEG) (label 0); returns work
    auto directed = [&](const RowV<NV> &x, const RowV<NV> &target, uint32_t word_index, const uint32_t (&tt)[NEG],
                        const RowV<NV> (&c)[NEG]) -> RowV<NV> {
        float p[8];
        p[0] = R::dot(x, target);
#pragma unroll
        for (int k = 0; k < 7; k++) p[k + 1] = k < NEG ? R::dot(x, c[k < NEG ? k : 0]) : 0.f;
        const float fm = reduce8_transposed(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], lane);
        bool live = pi == 0;
#pragma unroll
        for (int k = 0; k < NEG; k++) live = live || (pi == k + 1 && tt[k] != word_index);  // pyx:234-235
        float gm = 0.f;
        if (live && fm > -MAX_EXP_F && fm < MAX_EXP_F) gm = __fmul_rn(my_label - lut[lut_index(fm)], lr);  // pyx:243
        RowV<NV> work = R::zero();
        R::fma(work, __shfl_sync(FULL, gm, lane_of_p(0)), target);
#pragma unroll
        for (int k = 0; k < NEG; k++) R::fma(work, __shfl_sync(FULL, gm, lane_of_p(k + 1)), c[k]);  // pyx:245
        return work;
    };

    for (int64_t u = warp0; u < P.n_edges; u += n_warps) {
        const int64_t q = P.stride > 1 ? (int64_t)(((uint64_t)u * (uint64_t)P.stride) % (uint64_t)P.n_edges) : u;
        const uint32_t e0 = __ldg(P.edges + 2 * q), e1 = __ldg(P.edges + 2 * q + 1);
        const uint64_t rnd = P.seeds ? P.seeds[q] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)q)) & LCG_MASK);
        const uint32_t tmine = (lane < 2 * NEG) ? draw_fetch(P.draw, (myA * rnd + myC) & LCG_MASK) : 0u;
        float *p0 = node_l + (int64_t)e0 * D, *p1 = node_l + (int64_t)e1 * D;
        RowV<NV> r0 = R::ld(p0);
        RowV<NV> r1 = (e1 == e0) ? r0 : R::ld(p1);
        uint32_t ta[NEG], tb[NEG];
        RowV<NV> ca[NEG], cb[NEG];
#pragma unroll
        for (int k = 0; k < NEG; k++) {
            ta[k] = __shfl_sync(FULL, tmine, k);
            tb[k] = __shfl_sync(FULL, tmine, NEG + k);
            ca[k] = R::ld(node_l + (int64_t)ta[k] * D);
        }
        // pyx:444: row e0 against target e1
        RowV<NV> work = directed(r0, r1, e1, ta, ca);
        if (ATOMIC) R::red(p0, work);
        r0 = R::add(r0, work);
        if (!ATOMIC) R::st(p0, r0);
        if (e1 == e0) r1 = r0;
        // pyx:447: row e1 against the UPDATED row e0; samples equal to e0 must see that update too.  The second
        // direction's sample rows are loaded here (after the first direction) to halve the registers held at once.
#pragma unroll
        for (int k = 0; k < NEG; k++) cb[k] = tb[k] == e0 ? r0 : R::ld(node_l + (int64_t)tb[k] * D);
        work = directed(r1, r0, e0, tb, cb);
        if (ATOMIC)
            R::red(p1, work);
        else
            R::st(p1, R::add(r1, work));
    }
}

template <typename K>
int grid_for(K kernel, int64_t n_units, size_t dyn_smem = 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS_PER_BLOCK * 32, dyn_smem);
    if (comemb_opts().blocks_per_sm > 0) per_sm = comemb_opts().blocks_per_sm;
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_units + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int64_t cap = (int64_t)sms * per_sm;  // a multiple of the SM count: every SM holds its full complement of warps
    if (comemb_opts().max_warps > 0) {
        const int64_t lim = (comemb_opts().max_warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
        if (lim < cap) cap = lim;
    }
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <int NCH, bool VEC, int DFIX = 0, int MINB = 1>
int launch_o2_t(const O2Params &P, bool atomic, cudaStream_t st) {
    const int64_t n_units = P.n_walks * P.units_per_walk;
    if (atomic) {
        auto k = o2_hogwild_kernel<NCH, VEC, true, DFIX, MINB>;
        k<<<grid_for(k, n_units), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    } else {
        auto k = o2_hogwild_kernel<NCH, VEC, false, DFIX, MINB>;
        k<<<grid_for(k, n_units), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    }
    return (int)cudaGetLastError();
}

template <int NEG>
int launch_o2_d128(const O2Params &P, bool atomic, cudaStream_t st) {
    const int64_t n_units = P.n_walks * P.units_per_walk;
    const bool hint = comemb_opts().variant == COMEMB_VARIANT_L2_HINTS;  // experiment: L2 eviction-priority hints
#define COMEMB_LAUNCH(ATOM, HINT)                                            \
    do {                                                                     \
        auto k = o2_hogwild_d128_kernel<ATOM, NEG, 3, HINT, false>;          \
        k<<<grid_for(k, n_units), WARPS_PER_BLOCK * 32, 0, st>>>(P);         \
    } while (0)
    if (P.n_shards > 0) {  // row-partitioned tables: always red.add (remote rows are updated over NVLink)
        auto k = o2_hogwild_d128_kernel<true, NEG, 3, false, true>;
        k<<<grid_for(k, n_units), WARPS_PER_BLOCK * 32, 0, st>>>(P);
        return (int)cudaGetLastError();
    }
    if (atomic) {
        if (hint) COMEMB_LAUNCH(true, true); else COMEMB_LAUNCH(true, false);
    } else {
        if (hint) COMEMB_LAUNCH(false, true); else COMEMB_LAUNCH(false, false);
    }
#undef COMEMB_LAUNCH
    return (int)cudaGetLastError();
}

template <int NEG, int NV, bool HALF>
int launch_o2_dx(const O2Params &P, bool atomic, cudaStream_t st) {
    if (atomic) {
        auto k = o2_hogwild_dx_kernel<true, NEG, NV, HALF>;
        k<<<grid_for(k, P.n_walks), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    } else {
        auto k = o2_hogwild_dx_kernel<false, NEG, NV, HALF>;
        k<<<grid_for(k, P.n_walks), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    }
    return (int)cudaGetLastError();
}

template <int NV, bool HALF>
int launch_o2_dx_neg(const O2Params &P, int negative, bool atomic, cudaStream_t st) {
    switch (negative) {
        case 1: return launch_o2_dx<1, NV, HALF>(P, atomic, st);
        case 2: return launch_o2_dx<2, NV, HALF>(P, atomic, st);
        case 3: return launch_o2_dx<3, NV, HALF>(P, atomic, st);
        case 4: return launch_o2_dx<4, NV, HALF>(P, atomic, st);
        case 5: return launch_o2_dx<5, NV, HALF>(P, atomic, st);
        case 6: return launch_o2_dx<6, NV, HALF>(P, atomic, st);
        default: return launch_o2_dx<7, NV, HALF>(P, atomic, st);
    }
}

template <int NCH, bool VEC>
int launch_o1_t(const O1Params &P, bool atomic, cudaStream_t st) {
    if (atomic) {
        auto k = o1_hogwild_kernel<NCH, VEC, true>;
        k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    } else {
        auto k = o1_hogwild_kernel<NCH, VEC, false>;
        k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);
    }
    return (int)cudaGetLastError();
}

}  // namespace

int launch_o2_hogwild(float *node, float *ctx, int size, const uint32_t *walks, const int64_t *walk_off,
                      int64_t n_walks, const uint64_t *seeds, uint64_t base_seed, const uint32_t *table,
                      uint64_t table_len, const uint32_t *alias, uint32_t n_alias, int window, int negative, float lr,
                      float lambda, bool atomic, int64_t *n_tokens, cudaStream_t st) {
    if (size > 512) return COMEMB_E_UNSUPPORTED;
    if (n_walks == 0) return 0;
    O2Params P;
    P.node = node; P.ctx = ctx; P.d = size; P.walks = walks; P.walk_off = walk_off; P.n_walks = n_walks;
    P.seeds = seeds; P.base_seed = base_seed;
    P.draw = Draw{table, make_table_mod(table_len), alias, n_alias};
    P.window = window; P.negative = negative; P.lr = lr; P.lambda = lambda;
    P.centres_per_unit = 0; P.units_per_walk = 1;
    if (comemb_opts().centres_per_unit > 0 && comemb_opts().max_walk_len > 0) {
        P.centres_per_unit = comemb_opts().centres_per_unit;
        P.units_per_walk = (comemb_opts().max_walk_len + P.centres_per_unit - 1) / P.centres_per_unit;
    }
    P.n_tokens = n_tokens;
    P.glut = comemb_lut_device();
    P.n_shards = 0;
    P.rows_per_shard = 1;
    const bool vec = (size % 4) == 0;
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC) {  // the headline shape (variant 9 forces the generic kernel: tests)
        switch (negative) {
            case 1: return launch_o2_d128<1>(P, atomic, st);
            case 2: return launch_o2_d128<2>(P, atomic, st);
            case 3: return launch_o2_d128<3>(P, atomic, st);
            case 4: return launch_o2_d128<4>(P, atomic, st);
            case 5: return launch_o2_d128<5>(P, atomic, st);
            case 6: return launch_o2_d128<6>(P, atomic, st);
            case 7: return launch_o2_d128<7>(P, atomic, st);
            default: break;
        }
    }
    // sizes 64 and 256: the specialised instruction stream (no alias sampler / centre-chunked units there)
    if ((size == 64 || size == 256) && negative >= 1 && negative <= 7 && comemb_opts().variant != COMEMB_VARIANT_GENERIC &&
        P.units_per_walk == 1 && (reinterpret_cast<uintptr_t>(node) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0)
        return size == 64 ? launch_o2_dx_neg<1, true>(P, negative, atomic, st) : launch_o2_dx_neg<2, false>(P, negative, atomic, st);
    if (size <= 128) return vec ? launch_o2_t<1, true>(P, atomic, st) : launch_o2_t<1, false>(P, atomic, st);
    if (size <= 256) return vec ? launch_o2_t<2, true>(P, atomic, st) : launch_o2_t<2, false>(P, atomic, st);
    return vec ? launch_o2_t<4, true>(P, atomic, st) : launch_o2_t<4, false>(P, atomic, st);
}

int launch_o1_hogwild(float *node, int size, const uint32_t *edges, int64_t n_edges, const uint64_t *seeds,
                      uint64_t base_seed, const uint32_t *table, uint64_t table_len, const uint32_t *alias,
                      uint32_t n_alias, int negative, float lr, bool atomic, int64_t stride, cudaStream_t st) {
    if (size > 512) return COMEMB_E_UNSUPPORTED;
    if (n_edges == 0) return 0;
    O1Params P;
    P.node = node; P.d = size; P.edges = edges; P.n_edges = n_edges; P.seeds = seeds; P.base_seed = base_seed;
    P.draw = Draw{table, make_table_mod(table_len), alias, n_alias};
    if (stride > 1 && n_edges > 0xFFFFFFFFLL) return COMEMB_E_UNSUPPORTED;
    P.negative = negative; P.lr = lr; P.stride = stride > 1 ? stride % n_edges : 0;
    P.glut = comemb_lut_device();
    const bool vec = (size % 4) == 0;
    if (size == 128 && comemb_opts().variant != COMEMB_VARIANT_GENERIC) {
#define COMEMB_O1(N)                                                                  \
    case N:                                                                           \
        if (atomic) {                                                                 \
            auto k = o1_hogwild_d128_kernel<true, N>;                                 \
            k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);            \
        } else {                                                                      \
            auto k = o1_hogwild_d128_kernel<false, N>;                                \
            k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);            \
        }                                                                             \
        return (int)cudaGetLastError();
        switch (negative) {
            COMEMB_O1(1) COMEMB_O1(2) COMEMB_O1(3) COMEMB_O1(4) COMEMB_O1(5) COMEMB_O1(6) COMEMB_O1(7)
            default: break;
        }
#undef COMEMB_O1
    }
    if ((size == 64 || size == 256) && negative >= 1 && negative <= 7 && comemb_opts().variant != COMEMB_VARIANT_GENERIC &&
        (reinterpret_cast<uintptr_t>(node) & 15) == 0) {  // the specialised instruction stream at another row width
#define COMEMB_O1X(N)                                                                 \
    case N:                                                                           \
        if (size == 64) {                                                             \
            if (atomic) {                                                             \
                auto k = o1_hogwild_dx_kernel<true, N, 1, true>;                      \
                k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);        \
            } else {                                                                  \
                auto k = o1_hogwild_dx_kernel<false, N, 1, true>;                     \
                k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);        \
            }                                                                         \
        } else {                                                                      \
            if (atomic) {                                                             \
                auto k = o1_hogwild_dx_kernel<true, N, 2, false>;                     \
                k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);        \
            } else {                                                                  \
                auto k = o1_hogwild_dx_kernel<false, N, 2, false>;                    \
                k<<<grid_for(k, P.n_edges), WARPS_PER_BLOCK * 32, 0, st>>>(P);        \
            }                                                                         \
        }                                                                             \
        return (int)cudaGetLastError();
        switch (negative) {
            COMEMB_O1X(1) COMEMB_O1X(2) COMEMB_O1X(3) COMEMB_O1X(4) COMEMB_O1X(5) COMEMB_O1X(6) COMEMB_O1X(7)
            default: break;
        }
#undef COMEMB_O1X
    }
    if (size <= 128) return vec ? launch_o1_t<1, true>(P, atomic, st) : launch_o1_t<1, false>(P, atomic, st);
    if (size <= 256) return vec ? launch_o1_t<2, true>(P, atomic, st) : launch_o1_t<2, false>(P, atomic, st);
    return vec ? launch_o1_t<4, true>(P, atomic, st) : launch_o1_t<4, false>(P, atomic, st);
}

// Row-partitioned o2 (SURVEY 8e partition B): tables split into n_shards contiguous row blocks, each possibly on a
// peer GPU (pointers mapped through CUDA IPC); the kernel gathers rows and scatters red.add updates straight over
// NVLink -- the exchange is fused into the SGD kernel, there is no separate collective.
int launch_o2_hogwild_sharded(float *const *node_shards, float *const *ctx_shards, int n_shards, int64_t rows_per_shard,
                              int size, const uint32_t *walks, const int64_t *walk_off, int64_t n_walks,
                              const uint64_t *seeds, uint64_t base_seed, const uint32_t *table, uint64_t table_len,
                              int window, int negative, float lr, float lambda, int64_t *n_tokens, cudaStream_t st) {
    if (size != 128 || n_shards < 1 || n_shards > 8 || rows_per_shard <= 0 || rows_per_shard > 0xFFFFFFFFLL)
        return COMEMB_E_UNSUPPORTED;
    if (n_walks == 0) return 0;
    O2Params P;
    P.node = node_shards[0]; P.ctx = ctx_shards[0]; P.d = size; P.walks = walks; P.walk_off = walk_off;
    P.n_walks = n_walks; P.seeds = seeds; P.base_seed = base_seed;
    P.draw = Draw{table, make_table_mod(table_len), nullptr, 0};
    P.window = window; P.negative = negative; P.lr = lr; P.lambda = lambda;
    P.centres_per_unit = 0; P.units_per_walk = 1; P.n_tokens = n_tokens; P.glut = comemb_lut_device();
    P.n_shards = n_shards; P.rows_per_shard = (uint32_t)rows_per_shard;
    for (int s = 0; s < 8; s++) {
        P.node_shard[s] = s < n_shards ? node_shards[s] : nullptr;
        P.ctx_shard[s] = s < n_shards ? ctx_shards[s] : nullptr;
    }
    switch (negative) {
        case 3: return launch_o2_d128<3>(P, true, st);
        case 4: return launch_o2_d128<4>(P, true, st);
        case 5: return launch_o2_d128<5>(P, true, st);
        default: return COMEMB_E_UNSUPPORTED;
    }
}
