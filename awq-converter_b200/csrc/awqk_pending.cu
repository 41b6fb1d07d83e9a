// Entry points declared in include/awqk.h whose kernels are not built yet in this revision.
// They fail loudly (AWQK_E_UNSUPPORTED) -- there is no CPU fallback behind any of them.
#include "awqk_common.cuh"

extern "C" int awqk_abs_colsum(const void*, int, int64_t, int64_t, double*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_alpha_grid(const double*, int64_t, int64_t, int, float*, float*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_fakequant_delta(const void*, int, int64_t, int64_t, int, int, int, const float*, int, void*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_sqerr_gemm(const void*, const void*, int64_t, int64_t, int64_t, int, double*, void*) { return AWQK_E_UNSUPPORTED; }
