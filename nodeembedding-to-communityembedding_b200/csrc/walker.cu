// walker.cu -- truncated random walks with restarts on a CSR graph: replaces __random_walk__ /
// build_deepwalk_corpus_iter (utils/graph_utils.py:20-46, 191-197) so the walk -> pair stream never touches the host.
//
// ORDERED: the reference consumes ONE CPython random.Random (MT19937) stream sequentially: per pass a
//   Fisher-Yates shuffle of the start nodes (Random.shuffle -> _randbelow), then per step `random()` (two 32-bit
//   outputs) and `choice(neighbors)` (_randbelow(deg): k = deg.bit_length() bits per try, rejection).  One thread
//   replays that stream with the generator state in shared memory: bit-exact walks, serial by construction.
// HOGWILD: one thread per walk, counter-based generator (splitmix64 of (seed, walk, step)); start nodes are a
//   per-pass pseudo-random permutation (cycle-walking Feistel network), so every node starts exactly one walk per
//   pass as in the reference.  Same walk distribution, different random stream.
#include "comemb_common.cuh"

namespace {

struct MT {
    uint32_t *mt;  // [624] in shared memory
    int idx;
};

__device__ void mt_init_by_array(MT &s, const uint32_t *key, int klen) {
    s.mt[0] = 19650218u;
    for (int i = 1; i < 624; i++) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0, k = (624 > klen ? 624 : klen);
    for (; k; k--) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
    }
    s.mt[0] = 0x80000000u;
    s.idx = 624;
}

__device__ uint32_t mt_u32(MT &s) {
    if (s.idx >= 624) {
        for (int k = 0; k < 624; k++) {
            const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[(k + 1) % 624] & 0x7fffffffu);
            s.mt[k] = s.mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s.idx = 0;
    }
    uint32_t y = s.mt[s.idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// Random._randbelow_with_getrandbits for n < 2^32: k = n.bit_length(); r = getrandbits(k) until r < n
__device__ uint32_t py_randbelow(MT &s, uint32_t n) {
    const int k = 32 - __clz(n);
    uint32_t r = mt_u32(s) >> (32 - k);
    while (r >= n) r = mt_u32(s) >> (32 - k);
    return r;
}

__device__ double py_random(MT &s) {
    const uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(32) walks_ordered_kernel(const int64_t *rowptr, const uint32_t *col, int64_t n,
                                                           int num_paths, int L, double alpha, uint64_t seed,
                                                           uint32_t *nodes, uint32_t *walks, int32_t *lens) {
    __shared__ uint32_t state[624];
    if (threadIdx.x != 0) return;
    MT s{state, 624};
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    mt_init_by_array(s, key, key[1] ? 2 : 1);  // random.Random(seed) for a non-negative int
    for (int64_t i = 0; i < n; i++) nodes[i] = (uint32_t)i;
    int64_t w = 0;
    for (int cnt = 0; cnt < num_paths; cnt++) {
        for (int64_t i = n - 1; i >= 1; i--) {  // rand.shuffle(nodes), graph_utils.py:194
            const uint32_t j = py_randbelow(s, (uint32_t)(i + 1));
            const uint32_t t = nodes[i];
            nodes[i] = nodes[j];
            nodes[j] = t;
        }
        for (int64_t q = 0; q < n; q++, w++) {  // __random_walk__, graph_utils.py:20-46
            uint32_t *path = walks + w * L;
            int len = 1;
            const uint32_t start = nodes[q];
            uint32_t cur = start;
            path[0] = start;
            while (len < L) {
                const int64_t r0 = rowptr[cur], deg = rowptr[cur + 1] - r0;
                if (deg <= 0) break;           // :39, :45
                if (py_random(s) >= alpha)     // :40
                    cur = col[r0 + py_randbelow(s, (uint32_t)deg)];  // :41
                else
                    cur = start;               // :43
                path[len++] = cur;
            }
            lens[w] = len;
            for (int p = len; p < L; p++) path[p] = COMEMB_TOKEN_NONE;
        }
    }
}

// pseudo-random permutation of [0, n): 4-round Feistel on 2*hb bits + cycle walking
__device__ __forceinline__ uint64_t permute_index(uint64_t x, uint64_t n, int hb, uint64_t key) {
    const uint64_t mask = (1ULL << hb) - 1ULL;
    do {
        uint64_t l = x >> hb, r = x & mask;
#pragma unroll
        for (int round = 0; round < 4; round++) {
            const uint64_t f = splitmix64(r ^ (key + 0x632BE59BD9B4E019ULL * (uint64_t)(round + 1))) & mask;
            const uint64_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << hb) | r;
    } while (x >= n);
    return x;
}

__global__ void __launch_bounds__(256) walks_hogwild_kernel(const int64_t *__restrict__ rowptr,
                                                            const uint32_t *__restrict__ col, int64_t n, int L,
                                                            float alpha, uint64_t seed, int hb, int64_t first_walk,
                                                            int64_t n_out, uint32_t *__restrict__ walks,
                                                            int32_t *__restrict__ lens) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const int64_t g = first_walk + o;
    const int64_t pass = g / n, slot = g - pass * n;
    const uint32_t start = (uint32_t)permute_index((uint64_t)slot, (uint64_t)n, hb, splitmix64(seed ^ (uint64_t)pass));
    const uint64_t key = splitmix64(seed ^ splitmix64((uint64_t)g + 0x5851F42D4C957F2DULL));
    uint32_t *path = walks + o * L;
    uint32_t cur = start;
    int len = 1;
    path[0] = start;
    while (len < L) {
        const int64_t r0 = rowptr[cur], deg = rowptr[cur + 1] - r0;
        if (deg <= 0) break;
        const uint64_t x = splitmix64(key + 0x9E3779B97F4A7C15ULL * (uint64_t)len);
        const float u = (float)(x >> 40) * (1.0f / 16777216.0f);  // 24 random bits in [0,1)
        if (u >= alpha)
            cur = col[r0 + (int64_t)(((x & 0xffffffffULL) * (uint64_t)deg) >> 32)];
        else
            cur = start;
        path[len++] = cur;
    }
    if (lens) lens[o] = len;
    for (int p = len; p < L; p++) path[p] = COMEMB_TOKEN_NONE;
}

// prepare_sentences' frequent-node down-sampling (utils/embedding.py:126-136) for walks that never leave the device:
// token t is kept with probability keep_prob[t] (Model.precalc_sampling, model.py:69-81); dropped tokens are REMOVED
// (the path gets shorter, as in the reference), the tail is padded with TOKEN_NONE.  One warp per walk, in place.
__global__ void __launch_bounds__(256) downsample_walks_kernel(uint32_t *walks, int32_t *lens, int64_t n_walks, int L,
                                                               const float *__restrict__ keep_prob, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= n_walks) return;
    uint32_t *path = walks + w * L;
    const int len = lens ? lens[w] : L;
    const uint64_t key = splitmix64(seed ^ splitmix64((uint64_t)w + 0x2545F4914F6CDD1DULL));
    int out = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        uint32_t t = COMEMB_TOKEN_NONE;
        bool keep = false;
        if (i < len) {
            t = path[i];
            if (t != COMEMB_TOKEN_NONE) {
                const float p = keep_prob[t];
                const float u = (float)(splitmix64(key + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1)) >> 40) * (1.0f / 16777216.0f);
                keep = p >= 1.0f || p >= u;  // embedding.py:135
            }
        }
        const unsigned m = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) path[out + __popc(m & ((1u << lane) - 1u))] = t;  // out + rank <= i: never overtakes unread tokens
        out += __popc(m);
        __syncwarp();
    }
    for (int i = out + lane; i < L; i += 32) path[i] = COMEMB_TOKEN_NONE;
    if (lens && lane == 0) lens[w] = out;
}

}  // namespace

int launch_downsample_walks(uint32_t *walks, int32_t *lens, int64_t n_walks, int L, const float *keep_prob, uint64_t seed,
                            cudaStream_t st) {
    if (n_walks <= 0 || L <= 0) return 0;
    downsample_walks_kernel<<<(unsigned)((n_walks + 7) / 8), 256, 0, st>>>(walks, lens, n_walks, L, keep_prob, seed);
    return (int)cudaGetLastError();
}

int launch_walks(const int64_t *rowptr, const uint32_t *col, int64_t n, int num_paths, int L, double alpha,
                 uint64_t seed, int mode, int64_t first_walk, int64_t n_out, uint32_t *walks, int32_t *lens,
                 cudaStream_t st) {
    if (n <= 0 || num_paths <= 0 || L <= 0) return 0;
    if (n > 0xFFFFFFFELL) return COMEMB_E_UNSUPPORTED;
    if (mode == COMEMB_MODE_ORDERED) {
        if (first_walk != 0 || n_out != (int64_t)num_paths * n || !lens) return COMEMB_E_ARG;
        uint32_t *nodes = nullptr;
        CUDA_TRY(cudaMallocAsync(&nodes, (size_t)n * sizeof(uint32_t), st));
        walks_ordered_kernel<<<1, 32, 0, st>>>(rowptr, col, n, num_paths, L, alpha, seed, nodes, walks, lens);
        int e = (int)cudaGetLastError();
        CUDA_TRY(cudaFreeAsync(nodes, st));
        return e;
    }
    if (first_walk < 0 || n_out < 0 || first_walk + n_out > (int64_t)num_paths * n) return COMEMB_E_ARG;
    if (n_out == 0) return 0;
    int bits = 1;
    while ((1LL << bits) < n) bits++;
    const int hb = (bits + 1) / 2;  // Feistel domain 2^(2*hb) >= n, < 4n
    walks_hogwild_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(rowptr, col, n, L, (float)alpha, seed, hb,
                                                                          first_walk, n_out, walks, lens);
    return (int)cudaGetLastError();
}
