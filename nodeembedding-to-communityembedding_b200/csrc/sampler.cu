// sampler.cu -- negative-sampling tables: Model.make_table (ADSCModel/model.py:97-122) and an alias table with the
// same distribution for the Hogwild kernels.
//
// make_table in the reference is an O(table_size) interpreted loop (1e8 iterations by default).  Its recurrence
//     table[t] = widx;  if t/table_size > d1: widx += 1; d1 += count[widx]**power / Z;  clamp widx
// advances widx by at most one per slot, so the first slot of value w obeys  first(w+1) = max(first(w), T(w)) + 1
// with T(w) the first t whose t/table_size exceeds the cumulative mass through w.  The host walks that O(vocab)
// recurrence in double precision with the same libm pow() and the same summation order as CPython (so the run
// boundaries are bit-exact), and a kernel fills the table at HBM speed.
#include <math.h>

#include <vector>

#include "comemb_common.cuh"

namespace {

__global__ void fill_table_kernel(const int64_t *__restrict__ first, int64_t n_vals, uint32_t min_val,
                                  uint32_t *__restrict__ table, int64_t table_size) {
    // one block-stride pass over slots; value found by binary search over the run starts
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < table_size;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n_vals - 1;  // last v with first[v] <= t
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (first[mid] <= t) lo = mid; else hi = mid - 1;
        }
        table[t] = min_val + (uint32_t)lo;
    }
}

__global__ void run_starts_kernel(const uint32_t *__restrict__ table, int64_t table_len, int64_t n_rows,
                                  int64_t *__restrict__ first) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < table_len;
         t += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = table[t];
        if ((t == 0 || table[t - 1] != v) && v < n_rows) first[v] = t;
    }
}

}  // namespace

int host_make_table(const double *h_counts, int64_t vocab_size, double power, uint32_t *d_table, int64_t table_size,
                    cudaStream_t st) {
    if (vocab_size < 2 || table_size <= 0) return COMEMB_E_ARG;
    // Z: float(sum([...])) -- CPython >= 3.12 sums floats with Neumaier compensation (model.py:110)
    std::vector<double> p((size_t)vocab_size);
    double z = 0.0, comp = 0.0;
    for (int64_t i = 0; i < vocab_size; i++) {
        const double x = pow(h_counts[i], power), t = z + x;
        p[(size_t)i] = x;
        if (fabs(z) >= fabs(x)) comp += (z - t) + x; else comp += (x - t) + z;
        z = t;
    }
    if (comp != 0.0 && std::isfinite(comp)) z += comp;
    // values are node ids 1..vocab_size-1 (ids start at 1, clamp at vocab_size-1: model.py:112, 120-121)
    const int64_t n_vals = vocab_size - 1;
    std::vector<int64_t> first((size_t)n_vals);
    const double size_d = (double)table_size;
    double d1 = p[0] / z;  // :114, id 1 = rank 0
    first[0] = 0;
    for (int64_t v = 0; v + 1 < n_vals; v++) {
        // T = first t with t/size > d1
        int64_t t = (int64_t)floor(d1 * size_d);
        if (t < 0) t = 0;
        while (t > 0 && (double)(t - 1) / size_d > d1) t--;
        while (!((double)t / size_d > d1)) t++;
        const int64_t at = t > first[(size_t)v] ? t : first[(size_t)v];  // the slot at which widx is bumped
        first[(size_t)(v + 1)] = at + 1;
        d1 += p[(size_t)(v + 1)] / z;  // :119
    }
    int64_t *d_first = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_first, (size_t)n_vals * sizeof(int64_t), st));
    CUDA_TRY(cudaMemcpyAsync(d_first, first.data(), (size_t)n_vals * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    fill_table_kernel<<<148 * 8, 256, 0, st>>>(d_first, n_vals, 1u, d_table, table_size);
    int e = (int)cudaGetLastError();
    CUDA_TRY(cudaStreamSynchronize(st));  // `first` is a host temporary
    CUDA_TRY(cudaFreeAsync(d_first, st));
    return e;
}

// Vose alias construction over run lengths of the table; entries {threshold (coin < threshold -> bucket), alias}.
int host_build_alias(const uint32_t *d_table, int64_t table_len, int64_t n_rows, uint32_t *d_alias, cudaStream_t st) {
    if (table_len <= 0 || n_rows <= 0 || n_rows > 0xFFFFFFFFLL) return COMEMB_E_ARG;
    int64_t *d_first = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_first, (size_t)n_rows * sizeof(int64_t), st));
    CUDA_TRY(cudaMemsetAsync(d_first, 0xFF, (size_t)n_rows * sizeof(int64_t), st));  // -1
    run_starts_kernel<<<148 * 8, 256, 0, st>>>(d_table, table_len, n_rows, d_first);
    std::vector<int64_t> first((size_t)n_rows);
    CUDA_TRY(cudaMemcpyAsync(first.data(), d_first, (size_t)n_rows * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaFreeAsync(d_first, st));
    // run length of value v = next present start - first[v] (the table is monotone, model.py:115-121)
    std::vector<double> w((size_t)n_rows, 0.0);
    int64_t next = table_len;
    for (int64_t v = n_rows - 1; v >= 0; v--)
        if (first[(size_t)v] >= 0) {
            w[(size_t)v] = (double)(next - first[(size_t)v]);
            next = first[(size_t)v];
        }
    std::vector<uint32_t> out((size_t)(2 * n_rows));
    std::vector<double> q((size_t)n_rows);
    std::vector<int64_t> small, large;
    for (int64_t v = 0; v < n_rows; v++) {
        q[(size_t)v] = w[(size_t)v] * (double)n_rows / (double)table_len;
        (q[(size_t)v] < 1.0 ? small : large).push_back(v);
        out[(size_t)(2 * v + 1)] = (uint32_t)v;
    }
    auto thresh = [](double pr) -> uint32_t {
        if (pr >= 1.0) return 0xFFFFFFFFu;
        if (pr <= 0.0) return 0u;
        return (uint32_t)(pr * 4294967296.0);
    };
    while (!small.empty() && !large.empty()) {
        const int64_t s = small.back(), l = large.back();
        small.pop_back();
        out[(size_t)(2 * s)] = thresh(q[(size_t)s]);
        out[(size_t)(2 * s + 1)] = (uint32_t)l;
        q[(size_t)l] = (q[(size_t)l] + q[(size_t)s]) - 1.0;
        if (q[(size_t)l] < 1.0) {
            large.pop_back();
            small.push_back(l);
        }
    }
    for (int64_t v : large) out[(size_t)(2 * v)] = 0xFFFFFFFFu;
    for (int64_t v : small) out[(size_t)(2 * v)] = 0xFFFFFFFFu;  // numerical leftovers
    CUDA_TRY(cudaMemcpyAsync(d_alias, out.data(), out.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}
