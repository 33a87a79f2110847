// Tensor-core (tcgen05 / TMEM) implementation of the final-layer kernels.  Placeholder until the bf16 path lands:
// requesting NCDE_PREC_BF16 fails loudly instead of silently running the fp32 kernels.
#pragma once
#include "solve_kernels.cuh"

namespace ncde {

static inline int tc_prepare(const ncde_problem_t*, int, int, int, int, int) {
    set_error("precision=bf16 (tcgen05 path) is not built into this library yet");
    return NCDE_ERR_UNSUPPORTED;
}
static inline int tc_field_fwd(const ncde_problem_t*, const FieldArgs&, cudaStream_t, int64_t*) { return NCDE_ERR_UNSUPPORTED; }
static inline int tc_field_bwd(const ncde_problem_t*, const FieldArgs&, cudaStream_t, int64_t*) { return NCDE_ERR_UNSUPPORTED; }

}  // namespace ncde
