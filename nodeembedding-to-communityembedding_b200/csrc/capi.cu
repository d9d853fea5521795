// capi.cu -- extern "C" entry points of libcomemb_b200.so (see include/comemb_b200.h for the contract).
#include <math.h>
#include <string.h>

#include "comemb_common.cuh"


// launchers implemented in the other translation units
int launch_o2_ordered(float *, float *, int64_t, int, const uint32_t *, const int64_t *, int64_t, const uint64_t *, uint64_t,
                      const uint32_t *, uint64_t, int, int, float, float, bool, int64_t *, cudaStream_t);
int launch_o1_ordered(float *, int, const uint32_t *, int64_t, const uint64_t *, uint64_t, const uint32_t *, uint64_t,
                      int, float, bool, cudaStream_t);
int launch_sg_fused_ordered(float *, float *, int, const uint32_t *, const int64_t *, int64_t, const int32_t *,
                            const uint64_t *, uint64_t, const uint32_t *, uint64_t, const float *, const float *,
                            const float *, int, int, int, float, float, float, int, bool, cudaStream_t);
int launch_o2_hogwild(float *, float *, int, const uint32_t *, const int64_t *, int64_t, const uint64_t *, uint64_t,
                      const uint32_t *, uint64_t, const uint32_t *, uint32_t, int, int, float, float, bool, int64_t *,
                      cudaStream_t);
int launch_o1_hogwild(float *, int, const uint32_t *, int64_t, const uint64_t *, uint64_t, const uint32_t *, uint64_t,
                      const uint32_t *, uint32_t, int, float, bool, int64_t, cudaStream_t);
int launch_sg_fused_hogwild(float *, float *, int, const uint32_t *, const int64_t *, int64_t, const int32_t *,
                            const uint64_t *, uint64_t, const uint32_t *, uint64_t, const float *, const float *,
                            const float *, int, int, int, float, float, float, int, bool, int64_t, const int32_t *,
                            const float *, cudaStream_t);
int launch_o2_hogwild_sharded(float *const *, float *const *, int, int64_t, int, const uint32_t *, const int64_t *, int64_t,
                              const uint64_t *, uint64_t, const uint32_t *, uint64_t, int, int, float, float, int64_t *,
                              cudaStream_t);
int launch_sg_twin(float *, float *, int, const uint32_t *, const uint32_t *, int64_t, int, double, double, double,
                   const float *, const float *, const float *, int, int, cudaStream_t);
int launch_o3_batch(float *, int64_t, int, const uint32_t *, int64_t, const float *, const float *, const float *,
                    const int32_t *, const float *, int, double, float, int, cudaStream_t);
int launch_transpose_blocks(const float *, float *, int, int, cudaStream_t);
int launch_gmm_estep(const float *, int64_t, const float *, const float *, int, float *, cudaStream_t);
int launch_gmm_mstep(const float *, int64_t, const float *, const float *, int, float *, cudaStream_t);  // gmm_mstep.cu
int launch_scale(float *, int64_t, float, cudaStream_t);
int launch_row_probe(float *, int64_t, int, float *, cudaStream_t);
int launch_o2_pos_loss(const float *, const float *, int, const uint32_t *, const int64_t *, int64_t, int, double *,
                       cudaStream_t);
int launch_walks(const int64_t *, const uint32_t *, int64_t, int, int, double, uint64_t, int, int64_t, int64_t,
                 uint32_t *, int32_t *, cudaStream_t);
int launch_downsample_walks(uint32_t *, int32_t *, int64_t, int, const float *, uint64_t, cudaStream_t);
int host_make_table(const double *, int64_t, double, uint32_t *, int64_t, cudaStream_t);
int host_build_alias(const uint32_t *, int64_t, int64_t, uint32_t *, cudaStream_t);

namespace {
constexpr int MAX_DEVICES = 64;
float *g_lut_dev[MAX_DEVICES] = {nullptr};
float g_host_lut[EXP_TABLE_SIZE];
thread_local comemb_opts_t t_opts = {0, 0, 0, 0, 0};
}  // namespace

const comemb_opts_t &comemb_opts() { return t_opts; }

int comemb_check_init() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return COMEMB_E_NOINIT;
    return g_lut_dev[dev] ? 0 : COMEMB_E_NOINIT;
}

const float *comemb_lut_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    return g_lut_dev[dev];
}

#define REQUIRE_INIT()                   \
    do {                                 \
        int _r = comemb_check_init();    \
        if (_r) return _r;               \
    } while (0)

extern "C" {

int comemb_abi_version(void) { return COMEMB_ABI_VERSION; }

const char *comemb_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case COMEMB_E_ARG: return "comemb: invalid argument";
        case COMEMB_E_UNSUPPORTED: return "comemb: unsupported configuration (Hogwild/o3 kernels need size <= 512)";
        case COMEMB_E_NOINIT: return "comemb: comemb_init() has not been called on this device";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "comemb: unknown error";
    }
}

int comemb_init(void) {
    // pyx:531-533 with the arithmetic types of the generated C: the argument is built from (float)i/(float)1000 in
    // double, exp() in double, stored as float; then (float)((double)T/((double)T + 1.0)).
    for (int i = 0; i < EXP_TABLE_SIZE; i++) {
        const float e = (float)exp(((((double)((float)i / (float)EXP_TABLE_SIZE)) * 2.0) - 1.0) * 6.0);
        g_host_lut[i] = (float)((double)e / ((double)e + 1.0));
    }
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return COMEMB_E_ARG;
    if (!g_lut_dev[dev]) CUDA_TRY(cudaMalloc(&g_lut_dev[dev], sizeof(g_host_lut)));
    CUDA_TRY(cudaMemcpy(g_lut_dev[dev], g_host_lut, sizeof(g_host_lut), cudaMemcpyHostToDevice));
    {   // the launchers take stream-ordered scratch (cudaMallocAsync): keep the pool's memory across calls instead of
        // returning it to the driver at every synchronisation (the default release threshold is 0)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = 2ull << 30, cur = 0;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur);
            if (cur < keep) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

int comemb_get_lut(float *h_lut1000) {
    REQUIRE_INIT();
    if (!h_lut1000) return COMEMB_E_ARG;
    CUDA_TRY(cudaMemcpy(h_lut1000, comemb_lut_device(), sizeof(float) * EXP_TABLE_SIZE, cudaMemcpyDeviceToHost));
    return 0;
}

int comemb_set_opts(const comemb_opts_t *opts) {
    if (!opts) {
        t_opts = comemb_opts_t{0, 0, 0, 0, 0};
        return 0;
    }
    if (opts->centres_per_unit < 0 || opts->max_walk_len < 0 || opts->blocks_per_sm < 0 || opts->variant < 0 ||
        opts->max_warps < 0)
        return COMEMB_E_ARG;
    t_opts = *opts;
    return 0;
}

int comemb_get_opts(comemb_opts_t *out) {
    if (!out) return COMEMB_E_ARG;
    *out = t_opts;
    return 0;
}

int comemb_set_tuning(int centres_per_unit, int max_walk_len, int blocks_per_sm) {
    if (centres_per_unit < 0 || max_walk_len < 0 || blocks_per_sm < 0) return COMEMB_E_ARG;
    t_opts.centres_per_unit = centres_per_unit;
    t_opts.max_walk_len = max_walk_len;
    t_opts.blocks_per_sm = blocks_per_sm % 100;
    t_opts.variant = blocks_per_sm / 100;
    return 0;
}

int comemb_set_max_warps(int64_t max_warps) {
    if (max_warps < 0) return COMEMB_E_ARG;
    t_opts.max_warps = max_warps;
    return 0;
}

int comemb_o2_walks(float *d_node, float *d_ctx, int64_t n_rows, int size, const uint32_t *d_walks,
                    const int64_t *d_walk_off, int64_t n_walks, const uint64_t *d_seeds, uint64_t base_seed,
                    const uint32_t *d_table, uint64_t table_len, const uint32_t *d_alias, uint32_t n_alias, int window,
                    int negative, float lr, float lambda, int mode, uint32_t flags, int64_t *d_n_tokens, void *stream) {
    REQUIRE_INIT();
    if (!d_node || !d_ctx || n_rows <= 0 || size <= 0 || n_walks < 0 || window < 0 || negative < 0) return COMEMB_E_ARG;
    if (n_walks > 0 && (!d_walks || !d_walk_off)) return COMEMB_E_ARG;
    if (negative > 0 && (!d_table || table_len == 0)) return COMEMB_E_ARG;
    if (!d_seeds && !(flags & COMEMB_F_SEED_HASH)) return COMEMB_E_ARG;
    if ((flags & COMEMB_F_ALIAS) && (!d_alias || n_alias == 0 || mode != COMEMB_MODE_HOGWILD)) return COMEMB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_walks == 0) return 0;
    if (mode == COMEMB_MODE_ORDERED)
        return launch_o2_ordered(d_node, d_ctx, n_rows, size, d_walks, d_walk_off, n_walks, d_seeds, base_seed, d_table,
                                 table_len, window, negative, lr, lambda, !(flags & COMEMB_F_DOT_FLOAT), d_n_tokens, st);
    if (mode == COMEMB_MODE_HOGWILD)
        return launch_o2_hogwild(d_node, d_ctx, size, d_walks, d_walk_off, n_walks, d_seeds, base_seed, d_table,
                                 table_len, (flags & COMEMB_F_ALIAS) ? d_alias : nullptr, n_alias, window, negative, lr,
                                 lambda, (flags & COMEMB_F_ATOMIC) != 0, d_n_tokens, st);
    return COMEMB_E_ARG;
}

int comemb_o2_walks_sharded(float *const *h_node_shards, float *const *h_ctx_shards, int n_shards,
                            int64_t rows_per_shard, int size, const uint32_t *d_walks, const int64_t *d_walk_off,
                            int64_t n_walks, const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table,
                            uint64_t table_len, int window, int negative, float lr, float lambda, uint32_t flags,
                            int64_t *d_n_tokens, void *stream) {
    REQUIRE_INIT();
    if (!h_node_shards || !h_ctx_shards || n_shards < 1 || n_walks < 0 || window < 0 || negative < 0) return COMEMB_E_ARG;
    for (int s = 0; s < n_shards && s < 8; s++)
        if (!h_node_shards[s] || !h_ctx_shards[s]) return COMEMB_E_ARG;
    if (n_walks > 0 && (!d_walks || !d_walk_off)) return COMEMB_E_ARG;
    if (negative > 0 && (!d_table || table_len == 0)) return COMEMB_E_ARG;
    if (!d_seeds && !(flags & COMEMB_F_SEED_HASH)) return COMEMB_E_ARG;
    return launch_o2_hogwild_sharded(h_node_shards, h_ctx_shards, n_shards, rows_per_shard, size, d_walks, d_walk_off,
                                     n_walks, d_seeds, base_seed, d_table, table_len, window, negative, lr, lambda,
                                     d_n_tokens, (cudaStream_t)stream);
}

int comemb_ipc_open(const void *h_handle64, int64_t offset_bytes, void **h_out_ptr) {
    // Map a peer process' allocation INTO THE CURRENT DEVICE's context (cudaIpcMemLazyEnablePeerAccess turns on peer
    // access between the current device and the exporting device), so kernels launched here can dereference it.
    if (!h_handle64 || !h_out_ptr || offset_bytes < 0) return COMEMB_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, sizeof(h));
    void *base = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *h_out_ptr = static_cast<char *>(base) + offset_bytes;
    return 0;
}

int comemb_enable_peer_access(int peer_device) {
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (peer_device == dev) return 0;
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return COMEMB_E_UNSUPPORTED;
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    return (int)e;
}

int comemb_o1_edges(float *d_node, int64_t n_rows, int size, const uint32_t *d_edges, int64_t n_edges,
                    const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                    const uint32_t *d_alias, uint32_t n_alias, int negative, float lr, int mode, uint32_t flags,
                    int64_t edge_stride, void *stream) {
    REQUIRE_INIT();
    if (!d_node || n_rows <= 0 || size <= 0 || n_edges < 0 || negative < 0 || edge_stride < 0) return COMEMB_E_ARG;
    if (n_edges > 0 && !d_edges) return COMEMB_E_ARG;
    if (negative > 0 && (!d_table || table_len == 0)) return COMEMB_E_ARG;
    if (!d_seeds && !(flags & COMEMB_F_SEED_HASH)) return COMEMB_E_ARG;
    if ((flags & COMEMB_F_ALIAS) && (!d_alias || n_alias == 0 || mode != COMEMB_MODE_HOGWILD)) return COMEMB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_edges == 0) return 0;
    if (mode == COMEMB_MODE_ORDERED)
        return launch_o1_ordered(d_node, size, d_edges, n_edges, d_seeds, base_seed, d_table, table_len, negative, lr,
                                 !(flags & COMEMB_F_DOT_FLOAT), st);
    if (mode == COMEMB_MODE_HOGWILD)
        return launch_o1_hogwild(d_node, size, d_edges, n_edges, d_seeds, base_seed, d_table, table_len,
                                 (flags & COMEMB_F_ALIAS) ? d_alias : nullptr, n_alias, negative, lr,
                                 (flags & COMEMB_F_ATOMIC) != 0, edge_stride, st);
    return COMEMB_E_ARG;
}

int comemb_o3_batch(float *d_node, int64_t n_rows, int size, const uint32_t *d_rows, int64_t n_sel, const float *d_mu,
                    const float *d_inv_cov_t, const float *d_pi, int K, double beta, float lr, int iters, void *stream) {
    if (!d_node || n_rows <= 0 || size <= 0 || K <= 0 || !d_mu || !d_inv_cov_t || !d_pi || iters < 0) return COMEMB_E_ARG;
    if (!d_rows) n_sel = n_rows;
    if (n_sel < 0) return COMEMB_E_ARG;
    return launch_o3_batch(d_node, n_rows, size, d_rows, n_sel, d_mu, d_inv_cov_t, d_pi, nullptr, nullptr, K, beta, lr,
                           iters, (cudaStream_t)stream);
}

int comemb_o3_batch_top1(float *d_node, int64_t n_rows, int size, const uint32_t *d_rows, int64_t n_sel,
                         const float *d_mu, const float *d_inv_cov_t, const int32_t *d_comm, const float *d_weight, int K,
                         double beta, float lr, int iters, void *stream) {
    if (!d_node || n_rows <= 0 || size <= 0 || K <= 0 || !d_mu || !d_inv_cov_t || !d_comm || !d_weight || iters < 0)
        return COMEMB_E_ARG;
    if (!d_rows) n_sel = n_rows;
    if (n_sel < 0) return COMEMB_E_ARG;
    return launch_o3_batch(d_node, n_rows, size, d_rows, n_sel, d_mu, d_inv_cov_t, nullptr, d_comm, d_weight, K, beta, lr,
                           iters, (cudaStream_t)stream);
}

int comemb_gmm_estep(const float *d_x, int64_t n, int size, const float *d_prec_chol, const float *d_bias, int K,
                     float *d_sq, void *stream) {
    if (!d_x || !d_prec_chol || !d_bias || !d_sq || n < 0 || K <= 0) return COMEMB_E_ARG;
    if (size != 128) return COMEMB_E_UNSUPPORTED;
    return launch_gmm_estep(d_x, n, d_prec_chol, d_bias, K, d_sq, (cudaStream_t)stream);
}

int comemb_gmm_mstep(const float *d_x, int64_t n, int size, const float *d_resp, const float *d_means, int K,
                     float *d_scatter, void *stream) {
    if (!d_x || !d_resp || !d_means || !d_scatter || n < 0 || K <= 0) return COMEMB_E_ARG;
    if (size != 128) return COMEMB_E_UNSUPPORTED;
    return launch_gmm_mstep(d_x, n, d_resp, d_means, K, d_scatter, (cudaStream_t)stream);
}

int comemb_transpose_blocks(const float *d_in, float *d_out, int K, int size, void *stream) {
    if (!d_in || !d_out || K < 0 || size < 0 || d_in == d_out) return COMEMB_E_ARG;
    return launch_transpose_blocks(d_in, d_out, K, size, (cudaStream_t)stream);
}

int comemb_sg_fused(float *d_node, float *d_negemb, int64_t n_rows, int size, const uint32_t *d_walks,
                    const int64_t *d_walk_off, int64_t n_walks, const int32_t *d_reduced_windows,
                    const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                    const float *d_mu, const float *d_inv_cov, const float *d_pi, int K, int window, int negative,
                    float lr, float lambda1, float lambda2, int is_node_embedding, int mode, uint32_t flags,
                    void *stream) {
    REQUIRE_INIT();
    if (!d_node || !d_negemb || n_rows <= 0 || size <= 0 || n_walks < 0 || window < 0 || negative < 0) return COMEMB_E_ARG;
    if (n_walks > 0 && (!d_walks || !d_walk_off)) return COMEMB_E_ARG;
    if (negative > 0 && (!d_table || table_len == 0)) return COMEMB_E_ARG;
    if (lambda2 != 0.f && (K <= 0 || !d_mu || !d_inv_cov || !d_pi)) return COMEMB_E_ARG;
    if (!d_seeds && !(flags & COMEMB_F_SEED_HASH)) return COMEMB_E_ARG;
    if (n_walks == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == COMEMB_MODE_ORDERED)
        return launch_sg_fused_ordered(d_node, d_negemb, size, d_walks, d_walk_off, n_walks, d_reduced_windows, d_seeds,
                                       base_seed, d_table, table_len, d_mu, d_inv_cov, d_pi, K, window, negative, lr,
                                       lambda1, lambda2, is_node_embedding, !(flags & COMEMB_F_DOT_FLOAT), st);
    if (mode == COMEMB_MODE_HOGWILD)
        return launch_sg_fused_hogwild(d_node, d_negemb, size, d_walks, d_walk_off, n_walks, d_reduced_windows, d_seeds,
                                       base_seed, d_table, table_len, d_mu, d_inv_cov, d_pi, K, window, negative, lr,
                                       lambda1, lambda2, is_node_embedding, (flags & COMEMB_F_ATOMIC) != 0, n_rows,
                                       nullptr, nullptr, st);
    return COMEMB_E_ARG;
}

int comemb_sg_twin(float *d_node, float *d_ctx, int64_t n_rows, int size, const uint32_t *d_pair_row,
                   const uint32_t *d_targets, int64_t n_pairs, int negative, double alpha, double lambda1,
                   double lambda2, const float *d_mu, const float *d_inv_cov, const float *d_pi, int K,
                   int is_node_embedding, void *stream) {
    if (!d_node || !d_ctx || n_rows <= 0 || size <= 0 || n_pairs < 0 || negative < 0) return COMEMB_E_ARG;
    if (n_pairs > 0 && (!d_pair_row || !d_targets)) return COMEMB_E_ARG;
    if (lambda2 > 0.0 && (K <= 0 || !d_mu || !d_inv_cov || !d_pi)) return COMEMB_E_ARG;
    return launch_sg_twin(d_node, d_ctx, size, d_pair_row, d_targets, n_pairs, negative, alpha, lambda1, lambda2, d_mu,
                          d_inv_cov, d_pi, K, is_node_embedding, (cudaStream_t)stream);
}

int comemb_sg_fused_top1(float *d_node, float *d_negemb, int64_t n_rows, int size, const uint32_t *d_walks,
                         const int64_t *d_walk_off, int64_t n_walks, const int32_t *d_reduced_windows,
                         const uint64_t *d_seeds, uint64_t base_seed, const uint32_t *d_table, uint64_t table_len,
                         const float *d_mu, const float *d_inv_cov, const int32_t *d_comm, const float *d_weight, int K,
                         int window, int negative, float lr, float lambda1, float lambda2, uint32_t flags, void *stream) {
    REQUIRE_INIT();
    if (!d_node || !d_negemb || d_node == d_negemb || n_rows <= 0 || size != 128 || n_walks < 0 || window < 1 ||
        2 * window > 64 || negative < 1 || negative > 7 || !d_comm || !d_weight || !d_mu || !d_inv_cov || K <= 0)
        return COMEMB_E_UNSUPPORTED;  // this entry exists for the tensor-core kernels only
    if (n_walks > 0 && (!d_walks || !d_walk_off)) return COMEMB_E_ARG;
    if (!d_table || table_len == 0) return COMEMB_E_ARG;
    if (!d_seeds && !(flags & COMEMB_F_SEED_HASH)) return COMEMB_E_ARG;
    if (n_walks == 0) return 0;
    return launch_sg_fused_hogwild(d_node, d_negemb, size, d_walks, d_walk_off, n_walks, d_reduced_windows, d_seeds,
                                   base_seed, d_table, table_len, d_mu, d_inv_cov, nullptr, K, window, negative, lr,
                                   lambda1, lambda2, 0, (flags & COMEMB_F_ATOMIC) != 0, n_rows, d_comm, d_weight,
                                   (cudaStream_t)stream);
}

int comemb_walks_csr(const int64_t *d_rowptr, const uint32_t *d_col, int64_t n, int num_paths, int path_length,
                     double alpha, uint64_t seed, int mode, int64_t first_walk, int64_t n_out, uint32_t *d_walks,
                     int32_t *d_lens, void *stream) {
    if (!d_rowptr || !d_col || !d_walks || n < 0 || num_paths < 0 || path_length < 0) return COMEMB_E_ARG;
    if (mode != COMEMB_MODE_ORDERED && mode != COMEMB_MODE_HOGWILD) return COMEMB_E_ARG;
    return launch_walks(d_rowptr, d_col, n, num_paths, path_length, alpha, seed, mode, first_walk, n_out, d_walks,
                        d_lens, (cudaStream_t)stream);
}

int comemb_downsample_walks(uint32_t *d_walks, int32_t *d_lens, int64_t n_walks, int path_length,
                             const float *d_keep_prob, uint64_t seed, void *stream) {
    if (!d_walks || !d_keep_prob || n_walks < 0 || path_length < 0) return COMEMB_E_ARG;
    return launch_downsample_walks(d_walks, d_lens, n_walks, path_length, d_keep_prob, seed, (cudaStream_t)stream);
}

int comemb_make_table(const double *h_counts, int64_t vocab_size, double power, uint32_t *d_table, int64_t table_size,
                      void *stream) {
    if (!h_counts || !d_table) return COMEMB_E_ARG;
    return host_make_table(h_counts, vocab_size, power, d_table, table_size, (cudaStream_t)stream);
}

int comemb_build_alias(const uint32_t *d_table, int64_t table_len, int64_t n_rows, uint32_t *d_alias, void *stream) {
    if (!d_table || !d_alias) return COMEMB_E_ARG;
    return host_build_alias(d_table, table_len, n_rows, d_alias, (cudaStream_t)stream);
}

int comemb_scale(float *d_x, int64_t n, float scale, void *stream) {
    if (!d_x || n < 0) return COMEMB_E_ARG;
    return launch_scale(d_x, n, scale, (cudaStream_t)stream);
}

int comemb_row_probe(float *d_buf, int64_t n_rows, int passes, float *d_sink, void *stream) {
    if (!d_buf || !d_sink || n_rows < 0 || passes < 0) return COMEMB_E_ARG;
    return launch_row_probe(d_buf, n_rows, passes, d_sink, (cudaStream_t)stream);
}

int comemb_o2_pos_loss(const float *d_node, const float *d_ctx, int size, const uint32_t *d_walks,
                       const int64_t *d_walk_off, int64_t n_walks, int window, double *d_out, void *stream) {
    if (!d_node || !d_ctx || !d_out || size <= 0 || n_walks < 0 || window < 0) return COMEMB_E_ARG;
    if (n_walks > 0 && (!d_walks || !d_walk_off)) return COMEMB_E_ARG;
    return launch_o2_pos_loss(d_node, d_ctx, size, d_walks, d_walk_off, n_walks, window, d_out, (cudaStream_t)stream);
}

}  // extern "C"
