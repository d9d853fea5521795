/*
 * comemb_oracle.c -- CPU restatement of the ComEmb SGD hot path.   *** TEST INFRASTRUCTURE, NOT PRODUCT ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product package never imports, links or executes anything under oracle/.
 *
 * Every function cites the reference lines it restates ("pyx:N" = /root/reference/utils/training_sdg_inner.pyx:N).
 * Parity of this restatement is PINNED: tests/test_oracle_golden.py checks it against vectors produced by the
 * reference's own compiled Cython module (oracle/_ref, built by oracle/build_ref.py; generator:
 * tests/golden/make_golden.py).  With dot model ORACLE_DOT_REFBLAS_QUIRK the restatement is bit-exact to the
 * reference as built in the authoring container (Cython 3.3.0, scipy 1.18.1 / OpenBLAS 0.3.31.dev "SkylakeX" kernels,
 * FAST_VERSION == 0), see oracle_dot() below.
 *
 * Build:  gcc -O2 -mfma -ffp-contract=off -fPIC -shared comemb_oracle.c -o libcomemb_oracle.so -lm   (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXP_TABLE_SIZE 1000 /* pyx:89 */
#define MAX_EXP 6           /* pyx:90 */
#define MAX_SENTENCE_LEN 10000 /* pyx:18 */
#define LCG_MUL 25214903917ULL /* pyx:134 */
#define LCG_MASK 281474976710655ULL /* pyx:121 (2^48-1) */
#define TOKEN_NONE 0xFFFFFFFFu /* our encoding of a `None` path entry (pyx:485-486) */

enum {
    ORACLE_DOT_REFBLAS_QUIRK = 0, /* FAST_VERSION 0 on the golden machine: SkylakeX sdot read back as a double */
    ORACLE_DOT_REFBLAS = 1,       /* FAST_VERSION 1: same sdot, plain float return (pyx:190) */
    ORACLE_DOT_WARP = 2,          /* summation order of the CUDA Hogwild kernels (lane-strided fma + xor butterfly) */
    ORACLE_DOT_WARP2 = 3          /* same, size-64 specialisation: lane l owns elements 2l, 2l+1 */
};

static float EXP_TABLE[EXP_TABLE_SIZE];
static int lut_ready = 0;

/* pyx:531-533 with the generated-C arithmetic types: ((float)i/(float)1000)*2.0-1.0)*6.0 in double, exp in double,
 * stored as float; then (float)((double)T/((double)T+1.0)). */
void oracle_init_lut(float *out) {
    for (int i = 0; i < EXP_TABLE_SIZE; i++) {
        float e = (float)exp(((((double)((float)i / (float)EXP_TABLE_SIZE)) * 2.0) - 1.0) * 6.0);
        EXP_TABLE[i] = (float)((double)e / ((double)e + 1.0));
    }
    lut_ready = 1;
    if (out) memcpy(out, EXP_TABLE, sizeof(EXP_TABLE));
}

/* ---- dot product models -------------------------------------------------------------------------------------------
 * The reference calls BLAS sdot through a function pointer (pyx:80-81, 140, 190).  Summation order is a property of
 * the BLAS build; the one observed for the golden vectors is OpenBLAS' SkylakeX sdot kernel:
 *   blocks of 64: four 16-lane fp32 FMA accumulators; fold 16->8 lanes; one optional block of 32 on four 8-lane
 *   accumulators; ((a0+a1)+a2)+a3; fold 8->4; two horizontal adds; tail (<32) summed in double from fp32 products;
 *   result = (float)(tail + (double)vec_sum).
 * FAST_VERSION 0 (pyx:536-541) then reads that float return register as a double: the SSE register still holds the
 * upper 32 bits of the double intermediate, the lower 32 bits are the float's own bit pattern; pyx:140 casts that
 * double back to float.  Verified bit-for-bit against scipy's sdot (see tests/golden/make_golden.py). */
static float sdot_skx(const float *x, const float *y, int n, int quirk) {
    float a5[4][16];
    float a[4][8];
    int i = 0, k, l;
    memset(a5, 0, sizeof(a5));
    int n64 = n & ~63;
    for (; i < n64; i += 64)
        for (k = 0; k < 4; k++)
            for (l = 0; l < 16; l++) a5[k][l] = fmaf(x[i + 16 * k + l], y[i + 16 * k + l], a5[k][l]);
    for (k = 0; k < 4; k++)
        for (l = 0; l < 8; l++) a[k][l] = a5[k][l] + a5[k][l + 8];
    int n32 = n & ~31;
    for (; i < n32; i += 32)
        for (k = 0; k < 4; k++)
            for (l = 0; l < 8; l++) a[k][l] = fmaf(x[i + 8 * k + l], y[i + 8 * k + l], a[k][l]);
    float v[8], h[4];
    for (l = 0; l < 8; l++) v[l] = ((a[0][l] + a[1][l]) + a[2][l]) + a[3][l];
    for (l = 0; l < 4; l++) h[l] = v[l] + v[l + 4];
    float my = (h[0] + h[1]) + (h[2] + h[3]);
    double dot = 0.0;
    for (; i < n; i++) {
        float p = y[i] * x[i];
        dot += (double)p;
    }
    dot += (double)my;
    float fl = (float)dot;
    if (!quirk) return fl;
    uint64_t db;
    uint32_t fb;
    memcpy(&db, &dot, 8);
    memcpy(&fb, &fl, 4);
    db = (db & 0xFFFFFFFF00000000ULL) | (uint64_t)fb;
    double seen;
    memcpy(&seen, &db, 8);
    return (float)seen;
}

/* Summation order of the CUDA Hogwild kernels: lane l (0..31) owns elements 128m+4l..128m+4l+3 (per_lane = 4; the size-64
 * specialisation: 2l, 2l+1, per_lane = 2), fma-accumulates them in increasing order starting from +0, then an
 * xor-butterfly (16,8,4,2,1) adds across lanes. */
static float sdot_warp(const float *x, const float *y, int n, int per_lane) {
    float v[32], w[32];
    for (int l = 0; l < 32; l++) {
        float acc = 0.0f;
        for (int base = 0; base < n; base += 32 * per_lane)
            for (int c = 0; c < per_lane; c++) {
                int e = base + per_lane * l + c;
                if (e < n) acc = fmaf(x[e], y[e], acc);
            }
        v[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; l++) w[l] = v[l] + v[l ^ off];
        memcpy(v, w, sizeof(v));
    }
    return v[0];
}

float oracle_dot(const float *x, const float *y, int n, int model) {
    if (model == ORACLE_DOT_WARP) return sdot_warp(x, y, n, 4);
    if (model == ORACLE_DOT_WARP2) return sdot_warp(x, y, n, 2);
    return sdot_skx(x, y, n, model == ORACLE_DOT_REFBLAS_QUIRK);
}

/* y += a*x as OpenBLAS SkylakeX saxpy does it (one FMA per element; verified bit-for-bit). pyx:146-149 */
static inline void saxpy_fma(int n, float a, const float *x, float *y) {
    for (int i = 0; i < n; i++) y[i] = fmaf(a, x[i], y[i]);
}

/* ---- A2: fast0_o2 / fast1_o2, pyx:105-151 / 155-201 --------------------------------------------------------------- */
uint64_t oracle_fast_o2(int negative, const uint32_t *table, uint64_t table_len, float *node, float *ctx, int size,
                        uint32_t word_index, uint32_t word2_index, float lr, float lambda, float *work,
                        uint64_t next_random, int dot_model) {
    int64_t row1 = (int64_t)word2_index * size, row2; /* pyx:120 (64-bit here; the reference wraps at 2^32 elements) */
    float f, g, label;
    uint32_t target;
    memset(work, 0, (size_t)size * sizeof(float)); /* pyx:126 */
    for (int d = 0; d < negative + 1; d++) {       /* pyx:128 */
        if (d == 0) {
            target = word_index;
            label = 1.0f;
        } else {
            target = table[(next_random >> 16) % table_len];             /* pyx:133 */
            next_random = (next_random * LCG_MUL + 11ULL) & LCG_MASK;    /* pyx:134 */
            if (target == word_index) continue;                          /* pyx:135-136 */
            label = 0.0f;
        }
        row2 = (int64_t)target * size;
        f = oracle_dot(&node[row1], &ctx[row2], size, dot_model); /* pyx:140 */
        if ((double)f <= -6.0 || (double)f >= 6.0) continue;      /* pyx:141-142 */
        f = EXP_TABLE[(int)(((double)f + 6.0) * (double)((EXP_TABLE_SIZE / MAX_EXP) / 2))]; /* pyx:143, int 83 */
        g = ((label - f) * lr) * lambda;                                               /* pyx:144 */
        saxpy_fma(size, g, &ctx[row2], work);        /* pyx:146 work += g*ctx[t] (old value) */
        saxpy_fma(size, g, &node[row1], &ctx[row2]); /* pyx:147 ctx[t] += g*node[row1] */
    }
    saxpy_fma(size, 1.0f, work, &node[row1]); /* pyx:149 */
    return next_random;
}

/* ---- A4: fast0_o1 / fast1_o1, pyx:205-249 / 252-296 (both tables are the node table, targets read-only) ----------- */
uint64_t oracle_fast_o1(int negative, const uint32_t *table, uint64_t table_len, float *node, int size,
                        uint32_t word_index, uint32_t word2_index, float lr, float *work, uint64_t next_random,
                        int dot_model) {
    int64_t row1 = (int64_t)word2_index * size, row2;
    float f, g, label;
    uint32_t target;
    memset(work, 0, (size_t)size * sizeof(float)); /* pyx:225 */
    for (int d = 0; d < negative + 1; d++) {
        if (d == 0) {
            target = word_index;
            label = 1.0f;
        } else {
            target = table[(next_random >> 16) % table_len];          /* pyx:232 */
            next_random = (next_random * LCG_MUL + 11ULL) & LCG_MASK; /* pyx:233 */
            if (target == word_index) continue;
            label = 0.0f;
        }
        row2 = (int64_t)target * size;
        f = oracle_dot(&node[row1], &node[row2], size, dot_model); /* pyx:239 */
        if ((double)f <= -6.0 || (double)f >= 6.0) continue;
        f = EXP_TABLE[(int)(((double)f + 6.0) * (double)((EXP_TABLE_SIZE / MAX_EXP) / 2))];
        g = (label - f) * lr;                 /* pyx:243 */
        saxpy_fma(size, g, &node[row2], work); /* pyx:245 */
    }
    saxpy_fma(size, 1.0f, work, &node[row1]); /* pyx:247 */
    return next_random;
}

/* ---- A3: train_o2, pyx:454-509.  `path` holds row indices (Vocab.index), TOKEN_NONE for None; `seed` is the value
 * pyx:477 builds from two np.random.randint draws (the caller draws them).  Returns #non-None tokens (pyx:490). ----- */
int64_t oracle_train_o2(float *node, float *ctx, int size, const uint32_t *path, int64_t len, float lr, int negative,
                        int window, const uint32_t *table, uint64_t table_len, float lambda, uint64_t seed,
                        int dot_model, float *work) {
    if (!lut_ready) oracle_init_lut(NULL);
    int64_t path_len = len < MAX_SENTENCE_LEN ? len : MAX_SENTENCE_LEN; /* pyx:480 */
    int64_t result = 0;
    uint64_t next_random = seed;
    for (int64_t i = 0; i < path_len; i++)
        if (path[i] != TOKEN_NONE) result++;
    for (int64_t i = 0; i < path_len; i++) { /* pyx:494 */
        if (path[i] == TOKEN_NONE) continue;
        int64_t j = i - window;
        if (j < 0) j = 0;
        int64_t k = i + window + 1;
        if (k > path_len) k = path_len;
        for (; j < k; j++) { /* pyx:503 */
            if (j == i || path[j] == TOKEN_NONE) continue;
            next_random = oracle_fast_o2(negative, table, table_len, node, ctx, size, path[i], path[j], lr, lambda,
                                         work, next_random, dot_model); /* pyx:507 */
        }
    }
    return result;
}

/* ---- A5: train_o1, pyx:407-450: two directed updates sharing one LCG stream. -------------------------------------- */
int64_t oracle_train_o1(float *node, int size, uint32_t e0, uint32_t e1, float lr, int negative,
                        const uint32_t *table, uint64_t table_len, uint64_t seed, int dot_model, float *work) {
    if (!lut_ready) oracle_init_lut(NULL);
    uint64_t next_random = seed;
    next_random = oracle_fast_o1(negative, table, table_len, node, size, e1, e0, lr, work, next_random, dot_model); /* pyx:444 */
    next_random = oracle_fast_o1(negative, table, table_len, node, size, e0, e1, lr, work, next_random, dot_model); /* pyx:447 */
    return 2;
}

/* Whole-corpus drivers = what the learners' single worker does (context_embeddings.py:83-84, node_embeddings.py:70-71)
 * with one seed per call, in order. */
int64_t oracle_o2_walks(float *node, float *ctx, int size, const uint32_t *walks, const int64_t *walk_off,
                        int64_t n_walks, const uint64_t *seeds, float lr, int negative, int window,
                        const uint32_t *table, uint64_t table_len, float lambda, int dot_model) {
    float *work = (float *)malloc((size_t)size * sizeof(float));
    int64_t tot = 0;
    for (int64_t w = 0; w < n_walks; w++)
        tot += oracle_train_o2(node, ctx, size, walks + walk_off[w], walk_off[w + 1] - walk_off[w], lr, negative, window,
                               table, table_len, lambda, seeds[w], dot_model, work);
    free(work);
    return tot;
}

int64_t oracle_o1_edges(float *node, int size, const uint32_t *edges, int64_t n_edges, const uint64_t *seeds, float lr,
                        int negative, const uint32_t *table, uint64_t table_len, int dot_model) {
    float *work = (float *)malloc((size_t)size * sizeof(float));
    int64_t tot = 0;
    for (int64_t e = 0; e < n_edges; e++)
        tot += oracle_train_o1(node, size, edges[2 * e], edges[2 * e + 1], lr, negative, table, table_len, seeds[e],
                               dot_model, work);
    free(work);
    return tot;
}

/* ---- A9: Community2Vec.train, ADSCModel/community_embeddings.py:61-77 (HEAD o3, full-batch).
 * grad_i = sum_k pi[i,k] * inv_cov[k] @ (x_i - mu_k)  from x frozen at the start of each iter;
 * x -= clip(grad * beta/K, -5, 5) * lr.  `rows` = model.vocab[node].index for the nodes passed in (duplicates add up,
 * as `grad_input[node_index] += ...` does per chunk -- within one chunk numpy's fancy += keeps only the last duplicate;
 * the callers pass each node once, and so do we).  Each [d]x[d,d] product is summed in double then rounded to float32,
 * communities accumulate in float32 like batch_grad_input: the reference's own order inside np.matmul
 * ([n,d,d]x[n,d,1]) is BLAS-dependent, tests compare at 1e-5 relative. */
void oracle_o3_batch(float *node, int64_t n_rows, int size, const uint32_t *rows, int64_t n_sel, const float *mu,
                     const float *inv_cov, const float *pi, int K, double beta, float lr, int iters) {
    float *grad = (float *)malloc((size_t)n_rows * size * sizeof(float));
    float *acc = (float *)malloc((size_t)size * sizeof(float));
    float *diff = (float *)malloc((size_t)size * sizeof(float));
    for (int it = 0; it < iters; it++) {
        memset(grad, 0, (size_t)n_rows * size * sizeof(float));
        for (int64_t s = 0; s < n_sel; s++) {
            int64_t r = rows[s];
            const float *x = node + r * size;
            for (int a = 0; a < size; a++) acc[a] = 0.0f; /* :66 batch_grad_input (float32) */
            for (int k = 0; k < K; k++) {
                float p = pi[r * K + k];
                for (int b = 0; b < size; b++) diff[b] = x[b] - mu[(int64_t)k * size + b]; /* :68 */
                const float *S = inv_cov + (int64_t)k * size * size;
                for (int a = 0; a < size; a++) { /* :69-71  m = pi*inv_cov (float32) ; m @ diff */
                    double t = 0.0;
                    for (int b = 0; b < size; b++) t += (double)(float)(p * S[(int64_t)a * size + b]) * (double)diff[b];
                    acc[a] += (float)t; /* :72 float32 += */
                }
            }
            for (int a = 0; a < size; a++) grad[r * size + a] += acc[a]; /* :73 */
        }
        float scale = (float)(beta / (double)K); /* :76 float32 array *= python float (beta/k) */
        for (int64_t e = 0; e < n_rows * size; e++) {
            float g = grad[e] * scale;
            g = g < -5.0f ? -5.0f : (g > 5.0f ? 5.0f : g); /* :77 */
            node[e] -= g * lr;
        }
    }
    free(grad);
    free(acc);
    free(diff);
}

/* ---- A12: Model.make_table, ADSCModel/model.py:97-122.  counts[i] = degree of the node with rank i (sorted ids);
 * min_id = min(vocab keys) (asserted == 1, model.py:66).  Table VALUES are node ids used later as row indices. -------- */
void oracle_make_table(const double *counts, int64_t vocab_size, int64_t min_id, double power, uint32_t *table,
                       int64_t table_size) {
    /* :110 float(sum([...])): CPython >= 3.12 sums floats with Neumaier compensation (Python/bltinmodule.c cs_add) */
    double z = 0.0, comp = 0.0;
    for (int64_t i = 0; i < vocab_size; i++) {
        double x = pow(counts[i], power), t = z + x;
        if (fabs(z) >= fabs(x)) comp += (z - t) + x; else comp += (x - t) + z;
        z = t;
    }
    if (comp != 0.0 && isfinite(comp)) z += comp;
    int64_t widx = min_id;                                               /* :112 */
    double d1 = pow(counts[widx - min_id], power) / z;                   /* :114 self.vocab[widx] -> id widx */
    for (int64_t t = 0; t < table_size; t++) {                           /* :115 */
        table[t] = (uint32_t)widx;
        if (1.0 * (double)t / (double)table_size > d1) { /* :117 */
            widx += 1;
            /* :119 self.vocab[widx]: a KeyError in the reference if widx > max id; guard keeps the last mass */
            if (widx - min_id < vocab_size) d1 += pow(counts[widx - min_id], power) / z;
        }
        if (widx >= vocab_size) widx = vocab_size - 1; /* :120-121 */
    }
}

/* ---- W: CPython `random.Random` (MT19937) consumption model + DeepWalk walks, utils/graph_utils.py:20-46, 191-197 - */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt_t;

static void mt_init_genrand(mt_t *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static void mt_init_by_array(mt_t *s, const uint32_t *key, int klen) {
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0, k = (624 > klen ? 624 : klen);
    for (; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000u;
    s->idx = 624;
}

static uint32_t mt_u32(mt_t *s) {
    if (s->idx >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (s->mt[k] & 0x80000000u) | (s->mt[(k + 1) % 624] & 0x7fffffffu);
            s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* random.Random(seed) for a non-negative int seed: init_by_array over its 32-bit little-endian words. */
static void py_seed(mt_t *s, uint64_t seed) {
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    mt_init_by_array(s, key, key[1] ? 2 : 1);
}

static int bit_length(uint64_t n) {
    int k = 0;
    while (n) { k++; n >>= 1; }
    return k;
}

/* Random._randbelow_with_getrandbits (n < 2^32 here; getrandbits(k<=32) = u32 >> (32-k)) */
static uint32_t py_randbelow(mt_t *s, uint64_t n) {
    int k = bit_length(n);
    if (k <= 32) {
        uint32_t r = mt_u32(s) >> (32 - k);
        while ((uint64_t)r >= n) r = mt_u32(s) >> (32 - k);
        return r;
    }
    /* k == 33 only arises for randint(0, 2**32..): not used by the path */
    for (;;) {
        uint64_t lo = mt_u32(s);
        uint64_t hi = mt_u32(s) >> (64 - k);
        uint64_t r = (hi << 32) | lo;
        if (r < n) return (uint32_t)r;
    }
}

static double py_random(mt_t *s) {
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

/* rand.randint(0, 2**31) drawn from random.Random(parent_seed): the per-file seed of graph_utils.py:150. */
uint64_t oracle_walk_file_seed(uint64_t parent_seed) {
    mt_t s;
    py_seed(&s, parent_seed);
    return (uint64_t)py_randbelow(&s, (1ULL << 31) + 1ULL);
}

/* build_deepwalk_corpus_iter (graph_utils.py:191-197) over a CSR whose row order is list(G.nodes()) and whose column
 * order is list(G.neighbors(v)).  Walk tokens are ROW numbers of that CSR; out is [num_paths*n, path_length] padded
 * with TOKEN_NONE, out_len the true lengths.  rand = random.Random(seed). */
void oracle_walks(const int64_t *rowptr, const uint32_t *col, int64_t n, int num_paths, int path_length, double alpha,
                  uint64_t seed, uint32_t *out, int32_t *out_len) {
    mt_t s;
    py_seed(&s, seed);
    uint32_t *nodes = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    for (int64_t i = 0; i < n; i++) nodes[i] = (uint32_t)i;
    int64_t w = 0;
    for (int cnt = 0; cnt < num_paths; cnt++) {
        for (int64_t i = n - 1; i >= 1; i--) { /* rand.shuffle(nodes), graph_utils.py:194 */
            uint32_t j = py_randbelow(&s, (uint64_t)i + 1);
            uint32_t t = nodes[i]; nodes[i] = nodes[j]; nodes[j] = t;
        }
        for (int64_t q = 0; q < n; q++, w++) { /* __random_walk__, graph_utils.py:20-46 */
            uint32_t *path = out + w * path_length;
            int len = 1;
            path[0] = nodes[q];
            while (len < path_length) {
                uint32_t cur = path[len - 1];
                int64_t deg = rowptr[cur + 1] - rowptr[cur];
                if (deg <= 0) break;                       /* :39, :45 */
                if (py_random(&s) >= alpha)                /* :40 */
                    path[len++] = col[rowptr[cur] + py_randbelow(&s, (uint64_t)deg)]; /* :41 */
                else
                    path[len++] = path[0];                 /* :43 */
            }
            out_len[w] = len;
            for (int p = len; p < path_length; p++) path[p] = TOKEN_NONE;
        }
    }
    free(nodes);
}

/* ---- A7: the legacy fused pass (stale train_sg; C bodies training_sdg_inner.c:1597-1905,
 * 2520-2715, 2988-3740).  Per pair (i,j): (1) o3 gradient of x_j from the CURRENT x_j:
 *     work1 = sum_k sgemm_colmajor(pi[j,k]*inv_cov[k]) (x_j - mu_k)   i.e. uses inv_cov[k] TRANSPOSED
 *     work_o3 = clip(-lambda2 * work1, +-0.1*lr)
 * (2) SGNS pair with g=(label-f)*lr, gl=g*lambda1: work += g*neg[t]; if !is_node_embedding: neg[t] += gl*x_j;
 *     x_j += lambda1*work;   (3) x_j += work_o3.
 * reduced_windows[i] (np.random.randint(window) per token when window>1, drawn by the caller) shrink the window.
 * When is_node_embedding != 0 the "context" table is the node table itself (o1-style). */
int64_t oracle_train_sg(float *node, float *negemb, int size, const uint32_t *path, int64_t len,
                        const int32_t *reduced_windows, float lr, int negative, int window, const uint32_t *table,
                        uint64_t table_len, const float *mu, const float *inv_cov, const float *pi, int K,
                        float lambda1, float lambda2, int is_node_embedding, uint64_t seed, int dot_model) {
    if (!lut_ready) oracle_init_lut(NULL);
    int64_t path_len = len < MAX_SENTENCE_LEN ? len : MAX_SENTENCE_LEN;
    float *work = (float *)malloc((size_t)size * sizeof(float));
    float *work_o3 = (float *)calloc((size_t)size, sizeof(float));
    double *acc = (double *)malloc((size_t)size * sizeof(double));
    float *diff = (float *)malloc((size_t)size * sizeof(float));
    uint64_t next_random = seed;
    int64_t result = 0;
    for (int64_t i = 0; i < path_len; i++)
        if (path[i] != TOKEN_NONE) result++;
    for (int64_t i = 0; i < path_len; i++) {
        if (path[i] == TOKEN_NONE) continue;
        int rw = reduced_windows ? reduced_windows[i] : 0;
        int64_t j = i - window + rw;
        if (j < 0) j = 0;
        int64_t k = i + window + 1 - rw;
        if (k > path_len) k = path_len;
        for (; j < k; j++) {
            if (j == i || path[j] == TOKEN_NONE) continue;
            uint32_t word_index = path[i], word2_index = path[j];
            int64_t row1 = (int64_t)word2_index * size;
            /* (1) o3 */
            float clipv = (float)((double)lr * 0.1); /* c:2556 `cdef REAL_t _alpha = alpha * 0.1` (double product) */
            float nl2 = -lambda2;    /* c:3132 */
            memset(work_o3, 0, (size_t)size * sizeof(float));
            if (nl2 != 0.0f) {
                for (int a = 0; a < size; a++) acc[a] = 0.0;
                for (int c = 0; c < K; c++) {
                    float p = pi[(int64_t)word2_index * K + c];
                    const float *S = inv_cov + (int64_t)c * size * size;
                    for (int b = 0; b < size; b++) diff[b] = node[row1 + b] - mu[(int64_t)c * size + b];
                    for (int a = 0; a < size; a++) { /* column-major read: element (a,b) of the operand is S[b*size+a] */
                        double t = 0.0;
                        for (int b = 0; b < size; b++) t += (double)(float)(p * S[(int64_t)b * size + a]) * (double)diff[b];
                        acc[a] += t;
                    }
                }
                for (int a = 0; a < size; a++) {
                    float v = nl2 * (float)acc[a];
                    work_o3[a] = v < -clipv ? -clipv : (v > clipv ? clipv : v);
                }
            }
            /* (2) SGNS */
            memset(work, 0, (size_t)size * sizeof(float));
            for (int d = 0; d < negative + 1; d++) {
                uint32_t target;
                float label;
                if (d == 0) {
                    target = word_index;
                    label = 1.0f;
                } else {
                    target = table[(next_random >> 16) % table_len];
                    next_random = (next_random * LCG_MUL + 11ULL) & LCG_MASK;
                    if (target == word_index) continue;
                    label = 0.0f;
                }
                int64_t row2 = (int64_t)target * size;
                float f = oracle_dot(&node[row1], &negemb[row2], size, dot_model);
                if ((double)f <= -6.0 || (double)f >= 6.0) continue;
                f = EXP_TABLE[(int)(((double)f + 6.0) * 83.0)];
                float g = (label - f) * lr; /* c:1813 */
                float gl = g * lambda1;     /* c:1822 */
                saxpy_fma(size, g, &negemb[row2], work);
                if (!is_node_embedding) saxpy_fma(size, gl, &node[row1], &negemb[row2]); /* c:1840-1859 */
            }
            saxpy_fma(size, lambda1, work, &node[row1]); /* c:1870 */
            /* (3) */
            saxpy_fma(size, 1.0f, work_o3, &node[row1]); /* c:3668 */
        }
    }
    free(work); free(work_o3); free(acc); free(diff);
    return result;
}

/* LCG helper exposed for tests of the device skip-ahead. */
uint64_t oracle_lcg_advance(uint64_t x, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) x = (x * LCG_MUL + 11ULL) & LCG_MASK;
    return x;
}
