// azb_mlp_tc.cuh — tensor-core (tcgen05) path of the prior model.  Placeholder until the UMMA kernels land:
// creating a handle with AZB_MLP_TC fails loudly instead of falling back.
#pragma once
#include "azb_common.cuh"

struct AzbMlpTc {
    int unused;
};
static inline const char *azb_mlp_tc_create(AzbMlpTc &, uint32_t, const uint32_t *, uint64_t *) {
    return "AZB_MLP_TC is not built in this revision";
}
static inline void azb_mlp_tc_destroy(AzbMlpTc &) {}
static inline const char *azb_mlp_tc_load(AzbMlpTc &, const float *, cudaStream_t, uint64_t *) { return "not built"; }
static inline const char *azb_mlp_tc_forward(AzbMlpTc &, const float *, uint32_t, float *, uint32_t, uint32_t,
                                             cudaStream_t, uint64_t *) {
    return "not built";
}
