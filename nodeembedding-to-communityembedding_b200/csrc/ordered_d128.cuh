// ordered_d128.cuh -- register-resident 128-element rows and the reference-order dot product shared by the ORDERED
// kernels (sgns_ordered.cu: one sequential stream; sgns_flow.cu: the same stream executed as a dataflow graph).
#pragma once
#include "comemb_common.cuh"

namespace ordered {

struct Sampler {
    const uint32_t *table;
    TableMod mod;
};

struct Row4 {
    float v0, v1, v2, v3;  // elements t, t+32, t+64, t+96
};
__device__ __forceinline__ Row4 ld_row4(const float *row, int lane) {
    Row4 r;
    r.v0 = row[lane]; r.v1 = row[lane + 32]; r.v2 = row[lane + 64]; r.v3 = row[lane + 96];
    return r;
}
__device__ __forceinline__ void st_row4(float *row, int lane, const Row4 &r) {
    row[lane] = r.v0; row[lane + 32] = r.v1; row[lane + 64] = r.v2; row[lane + 96] = r.v3;
}
// dot of two 128-element rows in the reference's order (see dot_refblas): value on every lane
__device__ __forceinline__ float dot128_refblas(const Row4 &x, const Row4 &y, bool quirk) {
    float a0 = fmaf(x.v2, y.v2, fmaf(x.v0, y.v0, 0.f));  // accumulator q = t      : elements t, t+64
    float a1 = fmaf(x.v3, y.v3, fmaf(x.v1, y.v1, 0.f));  // accumulator q = t + 32 : elements t+32, t+96
    a0 = a0 + __shfl_down_sync(FULL, a0, 8);
    a1 = a1 + __shfl_down_sync(FULL, a1, 8);
    float v = a0 + __shfl_down_sync(FULL, a0, 16);
    v = v + a1;
    v = v + __shfl_down_sync(FULL, a1, 16);
    const float h = v + __shfl_down_sync(FULL, v, 4);
    const float p = h + __shfl_down_sync(FULL, h, 1);
    float my = p + __shfl_down_sync(FULL, p, 2);
    my = __shfl_sync(FULL, my, 0);
    if (!quirk) return my;  // (float)(0.0 + (double)my) == my
    const double dot = (double)my;
    return __double2float_rn(__hiloint2double(__double2hiint(dot), __float_as_int(my)));
}
__device__ __forceinline__ void fma_row4(Row4 &y, float a, const Row4 &x) {
    y.v0 = fmaf(a, x.v0, y.v0); y.v1 = fmaf(a, x.v1, y.v1); y.v2 = fmaf(a, x.v2, y.v2); y.v3 = fmaf(a, x.v3, y.v3);
}

}  // namespace ordered
