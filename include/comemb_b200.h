/*
 * comemb_b200.h -- C ABI of libcomemb_b200.so: the ComEmb SGD hot path (o1 / o2 / o3, walks, sampler tables) as
 * hand-written sm_100a CUDA.  Plain pointers and sizes only; every pointer named d_* is a DEVICE pointer on the
 * current CUDA device, `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * Each entry point names the reference interface it replaces ("pyx" = utils/training_sdg_inner.pyx of
 * andompesta/nodeembedding-to-communityembedding).  The reference binds its hot path as a CPython extension module;
 * the equivalent binding of this library is a ctypes stub, used by
 * nodeembedding-to-communityembedding_b200/utils/training_sdg_inner.py.
 *
 * Return value: 0 = ok; <0 = COMEMB_E_* argument error; >0 = cudaError_t of the failed CUDA call.
 * Calls are asynchronous on `stream` unless stated otherwise.  There is no CPU fallback anywhere in this library.
 */
#ifndef COMEMB_B200_H
#define COMEMB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMEMB_ABI_VERSION 2

#define COMEMB_TOKEN_NONE 0xFFFFFFFFu /* a `None` entry of a path (pyx:485-486) / padding of a short walk */

/* error codes */
#define COMEMB_E_ARG (-1)         /* null pointer / negative size / unsupported value */
#define COMEMB_E_UNSUPPORTED (-2) /* e.g. Hogwild kernels need size <= 512 */
#define COMEMB_E_NOINIT (-3)      /* comemb_init() not called on this device */

/* execution modes */
#define COMEMB_MODE_ORDERED 0 /* one logical stream: reproduces the reference's sequential updates exactly */
#define COMEMB_MODE_HOGWILD 1 /* lock-free parallel SGD: one warp per walk / edge (the reference's worker threads) */

/* flags (bit-or) */
#define COMEMB_F_DOT_FLOAT 1u   /* ORDERED: model FAST_VERSION 1 (plain float sdot) instead of FAST_VERSION 0 (pyx:536-547) */
#define COMEMB_F_ATOMIC 2u      /* HOGWILD: scatter with red.global.add.v4.f32 instead of plain stores */
#define COMEMB_F_ALIAS 4u       /* HOGWILD: negatives from the alias table instead of the unigram table */
#define COMEMB_F_SEED_HASH 8u   /* seeds==NULL: per-unit LCG seeds derived on device from base_seed */

/* ---- module init: replaces `init()` + EXP_TABLE (pyx:92-95, 512-549) ------------------------------------------------
 * Builds the 1000-entry sigmoid table on the host exactly as pyx:531-533 (double exp, float store) and uploads it
 * to the current device.  Returns FAST_VERSION-compatible 0 on success (this library always models the float-table
 * arithmetic; which sdot return convention ORDERED mode mimics is the COMEMB_F_DOT_FLOAT flag).  Synchronous. */
int comemb_init(void);
/* copy of the sigmoid table as uploaded (host buffer of 1000 floats); for tests */
int comemb_get_lut(float *h_lut1000);
int comemb_abi_version(void);
/* ---- per-thread launch options ---------------------------------------------------------------------------------------------
 * The reference entry points are called concurrently from worker threads (pyx:443, 493).  Everything that steers a launch
 * besides its arguments lives in this struct; comemb_set_opts() stores a copy in THREAD-LOCAL storage and it applies to the
 * launches the calling thread makes afterwards (NULL restores the defaults), so two threads -- or one thread per GPU -- never
 * see each other's settings.  No entry point keeps process-global mutable state. */
#define COMEMB_VARIANT_DEFAULT 0       /* the fastest kernel for the shape */
#define COMEMB_VARIANT_ROUNDSYNC 3     /* fused pass: the round-synchronous tcgen05 kernel even where the asynchronous one applies */
#define COMEMB_VARIANT_TENSOR 4        /* tensor-core kernels even below the size where they become the default (tests) */
#define COMEMB_VARIANT_L2_HINTS 5      /* size-128 Hogwild o2: L2 eviction-priority hints (experiment) */
#define COMEMB_VARIANT_ROUND1 6        /* round-1 kernels: fp64-pipe top-1 o3, per-warp mma.sync fused pass */
#define COMEMB_VARIANT_ORDERED_PIPE 7  /* ORDERED size 128 on one software-pipelined warp */
#define COMEMB_VARIANT_ORDERED_PLAIN 8 /* ORDERED size 128 on one plain warp */
#define COMEMB_VARIANT_GENERIC 9       /* the any-size kernels even where a size-128 specialisation exists (tests) */
#define COMEMB_VARIANT_ORDERED_FLOW 10 /* ORDERED o2/o1 size 128 as a dataflow graph on many warps, whatever the table size */
#define COMEMB_VARIANT_ORDERED_TEAM 11 /* ORDERED o2 size 128 on one CTA (warp per target row), whatever the table size */
typedef struct comemb_opts {
    int32_t centres_per_unit; /* Hogwild o2: 0 = one warp per walk (the reference's per-thread granularity); > 0 = one warp
                                 per chunk of that many centres (needs max_walk_len; the chunk starts at the right position
                                 of the walk's LCG stream) */
    int32_t max_walk_len;
    int32_t blocks_per_sm;    /* resident CTAs per SM (0 = occupancy query) */
    int32_t variant;          /* COMEMB_VARIANT_* */
    int64_t max_warps;        /* HOGWILD concurrency cap: at most this many walks/edges in flight (0 = fill the GPU).  The
                                 reference's `workers` plays this role; the learners pass max(workers, n_rows/28) so that
                                 the expected number of concurrent updates per row stays below ~1/4 on small graphs */
} comemb_opts_t;
int comemb_set_opts(const comemb_opts_t *opts);
int comemb_get_opts(comemb_opts_t *out);
/* Shorthands that edit the calling thread's options: blocks_per_sm % 100 = resident CTAs per SM, blocks_per_sm / 100 =
 * COMEMB_VARIANT_*. */
int comemb_set_tuning(int centres_per_unit, int max_walk_len, int blocks_per_sm);
int comemb_set_max_warps(int64_t max_warps);
const char *comemb_error_string(int code);

/* ---- o2: replaces train_o2 (pyx:454-509) applied to a batch of paths, i.e. the worker loop of
 * Context2Vec.train (ADSCModel/context_embeddings.py:83-84) -------------------------------------------------------------
 * d_node, d_ctx : float32 [n_rows, size] row-major, updated in place (pyx:457-458)
 * d_walks       : uint32 tokens = ROW indices (Vocab.index), COMEMB_TOKEN_NONE = None; walk w is
 *                 d_walks[d_walk_off[w] .. d_walk_off[w+1])   (int64 offsets, n_walks+1 of them);
 *                 only the first 10000 tokens of a walk are used (MAX_SENTENCE_LEN, pyx:18, 480)
 * d_seeds       : uint64 per walk = (2^24)*randint(0,2^24)+randint(0,2^24) as pyx:477 draws it (host draws them in
 *                 walk order); NULL with COMEMB_F_SEED_HASH -> derived from base_seed
 * d_table       : uint32 [table_len] unigram^0.75 table whose VALUES are used as row indices (model.py:97-122)
 * d_alias       : (COMEMB_F_ALIAS) uint32 [2*n_alias] = {threshold, alias} pairs from comemb_build_alias; else NULL
 * lr, lambda    : py_lr, py_alpha of pyx:454 (g = (label - sigma)*lr*lambda, pyx:144)
 * d_n_tokens    : optional int64[1]; receives sum of train_o2 return values (#non-None tokens, pyx:490)
 * ORDERED: one warp replays all walks in order (bit-exact to the reference on its golden build).
 * HOGWILD: one warp per walk (sequential inside the walk like a reference worker thread, walks concurrent). */
int comemb_o2_walks(float *d_node, float *d_ctx, int64_t n_rows, int size, const uint32_t *d_walks,
                    const int64_t *d_walk_off, int64_t n_walks, const uint64_t *d_seeds, uint64_t base_seed,
                    const uint32_t *d_table, uint64_t table_len, const uint32_t *d_alias, uint32_t n_alias,
                    int window, int negative, float lr, float lambda, int mode, uint32_t flags, int64_t *d_n_tokens,
                    void *stream);

/* ---- o2 over ROW-PARTITIONED tables (HOGWILD, size 128, negative in 3..5, red.add scatter) -----------------------------
 * Table row r lives in shard r / rows_per_shard at local row r % rows_per_shard.  h_node_shards / h_ctx_shards are HOST
 * arrays of n_shards (<= 8) DEVICE pointers; a shard may be memory of a peer GPU of the same node mapped into this
 * process (CUDA IPC) with peer access enabled (comemb_enable_peer_access): rows are then gathered and updated straight
 * over NVLink from inside the SGD kernel -- no separate collective.  Other arguments as comemb_o2_walks. */
int comemb_o2_walks_sharded(float *const *h_node_shards, float *const *h_ctx_shards, int n_shards,
                            int64_t rows_per_shard, int size, const uint32_t *d_walks, const int64_t *d_walk_off,
                            int64_t n_walks, const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table,
                            uint64_t table_len, int window, int negative, float lr, float lambda, uint32_t flags,
                            int64_t *d_n_tokens, void *stream);
/* cudaIpcOpenMemHandle of a 64-byte handle exported by a peer process (torch: storage._share_cuda_()[1]) into the
 * CURRENT device's context with lazy peer access; *h_out_ptr = mapped base + offset_bytes.  Synchronous. */
int comemb_ipc_open(const void *h_handle64, int64_t offset_bytes, void **h_out_ptr);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device (idempotent). */
int comemb_enable_peer_access(int peer_device);

/* ---- o1: replaces train_o1 (pyx:407-450) applied to a batch of edges, i.e. the worker loop of
 * Node2Vec.train (ADSCModel/node_embeddings.py:70-71) -------------------------------------------------------------------
 * d_edges : uint32 [n_edges, 2] row indices; per edge: update row e0 against target e1, then row e1 against the
 *           updated e0, one LCG stream (pyx:444-448).  `edge_stride` (HOGWILD only, 0/1 = identity): edge number
 *           u is taken as (u*edge_stride) mod n_edges so that concurrently running warps do not share a hub row. */
int comemb_o1_edges(float *d_node, int64_t n_rows, int size, const uint32_t *d_edges, int64_t n_edges,
                    const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                    const uint32_t *d_alias, uint32_t n_alias, int negative, float lr, int mode, uint32_t flags,
                    int64_t edge_stride, void *stream);

/* ---- o3 (HEAD): replaces Community2Vec.train (ADSCModel/community_embeddings.py:61-77) ----------------------------------
 * For every selected row r (d_rows, or all n_rows when NULL), `iters` times:
 *    G = sum_k pi[r,k] * inv_cov[k] @ (x_r - mu_k);  x_r -= clip(G * (float)(beta/K), -5, 5) * lr
 * d_inv_cov_t is inv_cov with each [size,size] block TRANSPOSED (comemb_transpose_blocks) so that lanes read it
 * coalesced.  Exactly-zero pi entries are skipped (they contribute exactly 0). */
int comemb_o3_batch(float *d_node, int64_t n_rows, int size, const uint32_t *d_rows, int64_t n_sel,
                    const float *d_mu, const float *d_inv_cov_t, const float *d_pi, int K, double beta, float lr,
                    int iters, void *stream);
/* The same update with pi in TOP-1 form (what fp32 predict_proba gives on separated data, and the only form that fits
 * at 50M nodes x 1000 communities): d_comm[r] = the row's community or -1, d_weight[r] = its responsibility.
 * At size 128 the selected rows are sorted by community on the device (16 bytes of stream-ordered scratch per row,
 * cudaMallocAsync on `stream`) and rows with weight exactly 1 are processed 8 per warp sharing the inv_cov reads; the
 * result is the same bit for bit.  A row must not appear twice in d_rows. */
int comemb_o3_batch_top1(float *d_node, int64_t n_rows, int size, const uint32_t *d_rows, int64_t n_sel,
                         const float *d_mu, const float *d_inv_cov_t, const int32_t *d_comm, const float *d_weight, int K,
                         double beta, float lr, int iters, void *stream);
/* ---- GMM E-step: the distance part of Community2Vec.fit's GaussianMixture (ADSCModel/community_embeddings.py:16-37;
 * sklearn _estimate_log_gaussian_prob, covariance_type 'full') -----------------------------------------------------------------
 *    d_sq[n][k] = || x_n . P_k - mu_k . P_k ||^2      P_k = d_prec_chol[k] (precision Cholesky factor, [size][size]),
 *                                                     d_bias[k] = mu_k . P_k  ([K][size])
 * size 128: tcgen05 3xTF32 tiles with P_k resident in shared memory, the squared norm reduced in the epilogue -- only the
 * [n, K] result reaches memory.  Other sizes: COMEMB_E_UNSUPPORTED (the caller uses a library GEMM). */
int comemb_gmm_estep(const float *d_x, int64_t n, int size, const float *d_prec_chol, const float *d_bias, int K,
                     float *d_sq, void *stream);
/* ---- GMM M-step, covariance part (sklearn _estimate_gaussian_covariances_full before the division by n_k) ------------------
 *    d_scatter[k][a][b] = sum_n d_resp[n][k] (x_n[a] - mu_k[a]) (x_n[b] - mu_k[b])        d_means = mu, [K][size]
 * size 128: tcgen05 3xTF32, four components per CTA accumulate in TMEM, the centred / weighted tiles are built in shared
 * memory -- no [K, n, size] temporary.  Tiles whose 32 responsibilities for a component are all exactly 0 are skipped.
 * Other sizes: COMEMB_E_UNSUPPORTED (the caller uses library GEMMs). */
int comemb_gmm_mstep(const float *d_x, int64_t n, int size, const float *d_resp, const float *d_means, int K,
                     float *d_scatter, void *stream);
/* out[k][b][a] = in[k][a][b] for K blocks of size x size */
int comemb_transpose_blocks(const float *d_in, float *d_out, int K, int size, void *stream);

/* ---- legacy fused pass: replaces the stale train_sg (utils/training_sdg_inner.c:2736, 2988-3740; Python twin
 * utils/embedding.py:15-72): per pair o3 gradient of x_j + SGNS pair + combined write ------------------------------------
 * d_reduced_windows: int32 per token (np.random.randint(window) per token, drawn by the host) or NULL.
 * inv_cov is read column-major as the reference's sgemm does, i.e. pass d_inv_cov UNtransposed. */
int comemb_sg_fused(float *d_node, float *d_negemb, int64_t n_rows, int size, const uint32_t *d_walks,
                    const int64_t *d_walk_off, int64_t n_walks, const int32_t *d_reduced_windows,
                    const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                    const float *d_mu, const float *d_inv_cov, const float *d_pi, int K, int window, int negative,
                    float lr, float lambda1, float lambda2, int is_node_embedding, int mode, uint32_t flags,
                    void *stream);

/* The fused pass (HOGWILD, tensor-core kernel: size 128, negative 3..5, window <= 12, context table != node table) with
 * pi in top-1 form (see comemb_o3_batch_top1); no dense pi is needed. */
int comemb_sg_fused_top1(float *d_node, float *d_negemb, int64_t n_rows, int size, const uint32_t *d_walks,
                         const int64_t *d_walk_off, int64_t n_walks, const int32_t *d_reduced_windows,
                         const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                         const float *d_mu, const float *d_inv_cov, const int32_t *d_comm, const float *d_weight, int K,
                         int window, int negative, float lr, float lambda1, float lambda2, uint32_t flags, void *stream);

/* ---- Python-twin semantics of the fused pass: replaces the fallback train_sg / gradient_update / community_sdg of
 * utils/embedding.py:15-98 (exact sigmoid, negatives redrawn until != both nodes, vectorised update where duplicate
 * targets keep the last write, float64 intermediates, inv_cov NOT transposed).  The host enumerates the window pairs and
 * draws the targets with the twin's own np.random loop; d_pair_row[p] = row of node2, d_targets[p*(negative+1)+k] =
 * target rows (positive first).  One warp, sequential.  negative <= 7. */
int comemb_sg_twin(float *d_node, float *d_ctx, int64_t n_rows, int size, const uint32_t *d_pair_row,
                   const uint32_t *d_targets, int64_t n_pairs, int negative, double alpha, double lambda1,
                   double lambda2, const float *d_mu, const float *d_inv_cov, const float *d_pi, int K,
                   int is_node_embedding, void *stream);

/* ---- walks: replaces __random_walk__ / build_deepwalk_corpus_iter (utils/graph_utils.py:20-46, 191-197) -----------------
 * CSR over row numbers: d_rowptr int64 [n+1], d_col uint32.  Output d_walks uint32 [num_paths*n, path_length] padded
 * with COMEMB_TOKEN_NONE, d_lens int32 [num_paths*n].
 * ORDERED: one thread replays CPython's random.Random(seed) (MT19937) stream: per pass a Fisher-Yates shuffle of the
 *          start nodes, per step one random() and one choice() -- bit-exact walks.
 * HOGWILD: one thread per walk, counter-based generator keyed by (seed, walk); start nodes are a per-pass
 *          permutation; same walk distribution, different stream.
 * first_walk/n_out select a contiguous shard of the num_paths*n walks (multi-GPU sharding; HOGWILD only). */
int comemb_walks_csr(const int64_t *d_rowptr, const uint32_t *d_col, int64_t n, int num_paths, int path_length,
                     double alpha, uint64_t seed, int mode, int64_t first_walk, int64_t n_out, uint32_t *d_walks,
                     int32_t *d_lens, void *stream);

/* Frequent-node down-sampling of device-resident walks: replaces the filter of prepare_sentences
 * (utils/embedding.py:126-136) with Model.precalc_sampling's probabilities (ADSCModel/model.py:69-81).  In place: token t
 * survives with probability d_keep_prob[t]; dropped tokens are removed (the walk shrinks), the tail becomes
 * COMEMB_TOKEN_NONE, d_lens (optional) is updated.  Counter-based generator keyed by (seed, walk, position). */
int comemb_downsample_walks(uint32_t *d_walks, int32_t *d_lens, int64_t n_walks, int path_length,
                             const float *d_keep_prob, uint64_t seed, void *stream);

/* ---- sampler tables: replaces Model.make_table (ADSCModel/model.py:97-122) ----------------------------------------------
 * h_counts: HOST double [vocab_size] (node degree by row; the O(vocab) run-boundary recurrence runs on the host with
 * the same libm pow() as CPython), d_table: DEVICE uint32 [table_size] filled by a kernel, with the reference's
 * id-as-row quirk.  Bit-exact to the reference for contiguous ids 1..vocab_size.  Synchronous. */
int comemb_make_table(const double *h_counts, int64_t vocab_size, double power, uint32_t *d_table, int64_t table_size,
                      void *stream);
/* alias table equivalent to drawing table[u % table_len] with u uniform: weights = run lengths of the table.
 * d_alias uint32 [2*n_rows].  Synchronous (host-side Vose construction on the run lengths). */
int comemb_build_alias(const uint32_t *d_table, int64_t table_len, int64_t n_rows, uint32_t *d_alias, void *stream);

/* ---- multi-GPU replica averaging epilogue: x = x * scale (after an NCCL sum all-reduce) ---------------------------------- */
int comemb_scale(float *d_x, int64_t n, float scale, void *stream);

/* Measurement utility (bench.py's L2 roofline denominator): the access mix of the SGNS kernels at streaming rate -- per warp
 * one coalesced 512-byte row gather (ld.global.cg) plus one red.global.add.v4.f32 of zeros into another scattered row,
 * 8 rows in flight -- over d_buf = [n_rows][128] floats (n_rows a power of two), `passes` times.  d_buf is left unchanged; bytes moved per pass =
 * 2 * n_rows * 512.  With a buffer that fits L2 this is the L2 gather/scatter peak. */
int comemb_row_probe(float *d_buf, int64_t n_rows, int passes, float *d_sink, void *stream);

/* SGNS objective of a batch of walks (evaluation only; no reference counterpart for o2, mirrors Node2Vec.loss,
 * node_embeddings.py:26-31, for window pairs): sum over pairs of -log sigmoid(x_j . c_i), exact sigmoid, double
 * accumulation.  d_out: double[2] = {sum, n_pairs}. */
int comemb_o2_pos_loss(const float *d_node, const float *d_ctx, int size, const uint32_t *d_walks,
                       const int64_t *d_walk_off, int64_t n_walks, int window, double *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COMEMB_B200_H */
