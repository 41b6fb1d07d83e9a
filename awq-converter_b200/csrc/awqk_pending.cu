// Entry points declared in include/awqk.h whose kernels are not built yet in this revision.
// They fail loudly (AWQK_E_UNSUPPORTED) -- there is no CPU fallback behind any of them.
#include "awqk_common.cuh"

extern "C" int awqk_abs_colsum(const void*, int, int64_t, int64_t, double*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_alpha_grid(const double*, int64_t, int64_t, int, float*, float*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_fakequant_delta(const void*, int, int64_t, int64_t, int, int, int, const float*, int, void*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_sqerr_gemm(const void*, const void*, int64_t, int64_t, int64_t, int, double*, void*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_pipe_create(int, size_t, awqk_pipe**) { return AWQK_E_UNSUPPORTED; }
extern "C" void awqk_pipe_destroy(awqk_pipe*) {}
extern "C" int awqk_pipe_quant_host(awqk_pipe*, const void*, int, int64_t, int64_t, int, int, int, int, int32_t*, uint32_t*, void*, int32_t*, uint32_t*) { return AWQK_E_UNSUPPORTED; }
extern "C" int awqk_pipe_sync(awqk_pipe*) { return AWQK_E_UNSUPPORTED; }
