// fused_sg.cu -- HOGWILD form of the legacy fused pass (stale train_sg: o3 gradient of x_j + SGNS pair + combined
// write per pair; utils/training_sdg_inner.c:2988-3740).  The ORDERED form lives in sgns_ordered.cu.
#include "comemb_common.cuh"

int launch_sg_fused_hogwild(float *, float *, int, const uint32_t *, const int64_t *, int64_t, const int32_t *,
                            const uint64_t *, uint64_t, const uint32_t *, uint64_t, const float *, const float *,
                            const float *, int, int, int, float, float, float, int, bool, cudaStream_t) {
    return COMEMB_E_UNSUPPORTED;  // not built yet: callers get a loud error, never a silent fallback
}
