// K2: activation-aware alpha search.  No reference counterpart: the definition
// is oracle/awq_oracle.py::search_scales, which composes the reference's own group quantizer
// (awq.py:173-250, fp32 arithmetic).
//
//   awqk_abs_colsum      sum_t |X[t,k]| in fp64 (exact for bf16 inputs -> order independent)
//   awqk_alpha_grid      s_i[k] = clamp(m[k]^(i/n), 1e-4) / sqrt(max_k * min_k),  m = colsum / T
//   awqk_fakequant_delta dW_i = bf16( W - dequant(group_quant(W * s_i)) / s_i )   (stand-alone producer)
//   awqk_sqerr_gemm      err_i = sum_{t,c} ( X . dW_i^T )^2   (awqk_search_gemm2.cu: tcgen05 cta_group::2)
//   awqk_scale_search    the whole search for one tensor in ONE call: column statistic -> grid -> scores
//                        (fused producer + tcgen05 GEMM, awqk_search_fused.cu) -> device-side argmin and
//                        winning scale vector -> final column-scaled K1 pass.  No host synchronisation.
//
// err uses the delta form ||X (W - W^)^T||^2 (identical in exact arithmetic to ||X W^T - X W^^T||^2):
// half the flops, and rounding dW (not W^) to bf16 keeps the relative error of err ~1e-6.
#include <algorithm>

#include "awqk_search.cuh"

namespace awqk {

// ------------------------------------------------------------------------------------------
// column statistic
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float abs_as_float(T v);
template <>
__device__ __forceinline__ float abs_as_float<__nv_bfloat16>(__nv_bfloat16 v) { return fabsf(__bfloat162float(v)); }
template <>
__device__ __forceinline__ float abs_as_float<__half>(__half v) { return fabsf(__half2float(v)); }
template <>
__device__ __forceinline__ float abs_as_float<float>(float v) { return fabsf(v); }

template <typename T>
__global__ void __launch_bounds__(256)
abs_colsum_kernel(const T* __restrict__ x, int64_t Tn, int64_t K, int rows_per_cta, double* __restrict__ colsum) {
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t t1 = (t0 + rows_per_cta < Tn) ? t0 + rows_per_cta : Tn;
  double acc = 0.0;
  for (int64_t t = t0; t < t1; ++t) acc += (double)abs_as_float(x[t * K + k]);
  atomicAdd(colsum + k, acc);
}

// ------------------------------------------------------------------------------------------
// alpha grid: pass 1 raw powers + per-alpha min/max, pass 2 normalisation
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
alpha_grid_raw(const double* __restrict__ colsum, int64_t Tn, int64_t K, int n_grid, float* __restrict__ s_grid,
               unsigned int* __restrict__ mnmx /* [2*n_grid] float bits: min then max */) {
  const int i = blockIdx.y;
  const float alpha = (float)((double)i / (double)n_grid);
  float lmin = __int_as_float(0x7F800000), lmax = 0.0f;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < K; k += (int64_t)gridDim.x * 256) {
    const float m = (float)(colsum[k] / (double)Tn);
    float p;
    if (i == 0) p = 1.0f;                         // m^0
    else if (2 * i == n_grid) p = sqrtf(m);       // torch special-cases exponent 0.5
    else p = powf(m, alpha);
    p = (p != p) ? p : fmaxf(p, 1e-4f);           // torch.clamp keeps NaN
    s_grid[(int64_t)i * K + k] = p;
    lmin = fminf(lmin, p);
    lmax = fmaxf(lmax, p);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    lmin = fminf(lmin, __shfl_xor_sync(0xFFFFFFFFu, lmin, o));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) {                  // all values are positive: uint order == float order
    atomicMin(mnmx + i, __float_as_uint(lmin));
    atomicMax(mnmx + n_grid + i, __float_as_uint(lmax));
  }
}

__global__ void __launch_bounds__(256)
alpha_grid_norm(int64_t K, int n_grid, float* __restrict__ s_grid, const unsigned int* __restrict__ mnmx) {
  const int i = blockIdx.y;
  const float norm = sqrtf(__fmul_rn(__uint_as_float(mnmx[n_grid + i]), __uint_as_float(mnmx[i])));
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < K; k += (int64_t)gridDim.x * 256)
    s_grid[(int64_t)i * K + k] = __fdiv_rn(s_grid[(int64_t)i * K + k], norm);
}

__global__ void alpha_grid_init(int n_grid, unsigned int* mnmx) {
  const int i = threadIdx.x;
  if (i < n_grid) {
    mnmx[i] = 0x7F800000u;
    mnmx[n_grid + i] = 0u;
  }
}


// ------------------------------------------------------------------------------------------
// fake-quant delta, stand-alone: one pass over W for all n_s scale vectors.
// CTA = 4 rows x 1024 columns (warps 2r, 2r+1 = the two halves of row r), 16 elements per thread (a group of
// 128 = 8 lanes), <= 80 registers -> 3 CTAs (24 warps) per SM.  The 1024-column slab of s_a and 1/s_a is
// fetched ONCE per CTA and alpha with coalesced 16-byte loads into shared memory; every thread then reads its
// 16 floats with conflict-free LDS.128 (chunk c of a lane's 64-byte row sits at position (c + lane/2) % 4).
// ------------------------------------------------------------------------------------------
template <typename T, int G, int BITS>
__global__ void __launch_bounds__(256, 3)
fakequant_delta_kernel(const T* __restrict__ w, int64_t C, int64_t K, bool sym, const float* __restrict__ s_grid,
                       int n_s_total, __nv_bfloat16* __restrict__ dw) {
  const int per_z = (n_s_total + (int)gridDim.z - 1) / (int)gridDim.z;
  const int a_begin = (int)blockIdx.z * per_z;
  const int n_s = min(per_z, n_s_total - a_begin);
  if (n_s <= 0) return;
  s_grid += (int64_t)a_begin * K;
  dw += (int64_t)a_begin * C * K;
  __shared__ __align__(16) float sm_s[2][1024];
  __shared__ __align__(16) float sm_r[2][1024];   // refined 1/s, computed once per CTA (4 rows share it)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 4 + (warp >> 1);
  const int64_t col0 = (int64_t)blockIdx.y * 1024;
  const int64_t col = col0 + (warp & 1) * 512 + lane * 16;
  const bool valid = row < C && col < K;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  float2 wv[8];
  {
    Raw16<T> raw;
    if (valid) raw.load(w + row * K + col); else raw.zero();
    raw.unpack(wv);
  }
  // staging: thread t fetches floats [4t, 4t+4): 64-byte row t/4 (= half * 32 + lane), chunk t % 4,
  // stored at position (chunk + (lane >> 1)) & 3 -> quarter-warps read 8 distinct bank groups
  const int t = threadIdx.x;
  const int64_t gcol = col0 + 4 * t;
  const bool gvalid = gcol < K;
  const int st_off = (t >> 2) * 16 + ((((t & 3) + (((t >> 2) & 31) >> 1)) & 3) << 2);
  int ld_off[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) ld_off[c] = ((warp & 1) * 32 + lane) * 16 + (((c + (lane >> 1)) & 3) << 2);

  __nv_bfloat16* out = dw + row * K + col;
  const int64_t plane = C * K;
  float4 ns = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gvalid) ns = __ldg(reinterpret_cast<const float4*>(s_grid + gcol));
#pragma unroll 1
  for (int a = 0; a < n_s; ++a) {
    const int buf = a & 1;
    *reinterpret_cast<float4*>(&sm_s[buf][st_off]) = ns;
    *reinterpret_cast<float4*>(&sm_r[buf][st_off]) =
        make_float4(refined_rcp(ns.x), refined_rcp(ns.y), refined_rcp(ns.z), refined_rcp(ns.w));
    __syncthreads();                                  // slab a visible; buffer buf^1 is free again after this point
    if (a + 1 < n_s && gvalid) ns = __ldg(reinterpret_cast<const float4*>(s_grid + (int64_t)(a + 1) * K + gcol));
    float2 sv[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(&sm_s[buf][ld_off[c]]);
      sv[2 * c] = make_float2(v.x, v.y);
      sv[2 * c + 1] = make_float2(v.z, v.w);
    }
    uint32_t o[8];
    const float* rbuf = sm_r[buf];
    delta16<G, BITS>(wv, sv, [&](int i) {                      // reciprocals fetched on use (keeps 80 registers)
      const float4 rr = *reinterpret_cast<const float4*>(rbuf + ld_off[i >> 1]);
      return (i & 1) ? make_float2(rr.z, rr.w) : make_float2(rr.x, rr.y);
    }, sym, qmin, qmax, o);
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)a * plane);
      st_stream16(dst, make_uint4(o[0], o[1], o[2], o[3]));
      st_stream16(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
  }
}

template <typename T>
static int launch_delta_t(const T* w, int64_t C, int64_t K, int g, int bits, bool sym, const float* s, int n_s,
                          __nv_bfloat16* dw, cudaStream_t st) {
  // alpha slices: enough CTAs for ~4 waves of 3 x 148, but no more (each slice re-reads its W rows)
  const int64_t base_ctas = ceil_div(K, 1024) * ceil_div(C, 4);
  const int z = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_s, 4), ceil_div(1800, base_ctas)));
  if (ceil_div(K, 1024) > 65535 || ceil_div(C, 4) > 0x7FFFFFFF) return AWQK_E_BADARG;
  dim3 grid((unsigned)ceil_div(C, 4), (unsigned)ceil_div(K, 1024), (unsigned)z);
#define AWQK_DELTA(GG, BB) fakequant_delta_kernel<T, GG, BB><<<grid, 256, 0, st>>>(w, C, K, sym, s, n_s, dw)
  if (bits == 4) {
    if (g == 32) AWQK_DELTA(32, 4); else if (g == 64) AWQK_DELTA(64, 4); else AWQK_DELTA(128, 4);
  } else {
    if (g == 32) AWQK_DELTA(32, 8); else if (g == 64) AWQK_DELTA(64, 8); else AWQK_DELTA(128, 8);
  }
#undef AWQK_DELTA
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

int launch_fakequant_delta(const void* w, int dtype, int64_t C, int64_t K, int g, int bits, bool sym, const float* s,
                           int n_s, __nv_bfloat16* dw, cudaStream_t st) {
  if (dtype == AWQK_BF16) return launch_delta_t(reinterpret_cast<const __nv_bfloat16*>(w), C, K, g, bits, sym, s, n_s, dw, st);
  if (dtype == AWQK_FP16) return launch_delta_t(reinterpret_cast<const __half*>(w), C, K, g, bits, sym, s, n_s, dw, st);
  if (dtype == AWQK_FP32) return launch_delta_t(reinterpret_cast<const float*>(w), C, K, g, bits, sym, s, n_s, dw, st);
  return AWQK_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------
// device-side selection: err_mean = err_sum / (T*C), best = first strict minimum (ties -> smallest alpha; a NaN
// score never wins), best_s = s_grid[best].  Every CTA recomputes the argmin over the <= 256 scores.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
search_select_kernel(const double* __restrict__ err_sum, int n_grid, double denom, const float* __restrict__ s_grid,
                     int64_t K, double* __restrict__ err_mean, int32_t* __restrict__ best_idx,
                     float* __restrict__ best_s) {
  __shared__ int s_best;
  if (threadIdx.x == 0) {
    int b = 0;
    double bv = err_sum[0];
    for (int i = 1; i < n_grid; ++i) {
      const double v = err_sum[i];
      if (v < bv || (bv != bv && v == v)) { b = i; bv = v; }
    }
    s_best = b;
    if (blockIdx.x == 0 && best_idx != nullptr) *best_idx = b;
  }
  if (blockIdx.x == 0 && err_mean != nullptr)
    for (int i = threadIdx.x; i < n_grid; i += 256) err_mean[i] = err_sum[i] / denom;
  __syncthreads();
  const float* src = s_grid + (int64_t)s_best * K;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < K; k += (int64_t)gridDim.x * 256) best_s[k] = src[k];
}

// ------------------------------------------------------------------------------------------
// workspace plan of awqk_scale_search (all offsets 256-byte aligned)
// ------------------------------------------------------------------------------------------
struct SearchPlan {
  size_t off_colsum, off_mnmx, off_grid, off_err, off_sync, off_ring, total_min, total_pref;
  FusedPlan fused;
  int rc;
};
static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

static SearchPlan plan_search(int64_t C, int64_t K, int64_t T, int n_grid, bool own_grid) {
  SearchPlan p{};
  p.rc = fused_plan(C, K, T, n_grid, &p.fused);
  if (p.rc != AWQK_OK) return p;
  size_t o = 0;
  p.off_colsum = o; o += own_grid ? al256((size_t)K * 8) : 0;
  p.off_mnmx = o;   o += own_grid ? al256((size_t)n_grid * 2 * 4) : 0;
  p.off_grid = o;   o += own_grid ? al256((size_t)n_grid * K * 4) : 0;
  p.off_err = o;    o += al256((size_t)n_grid * 8);
  p.off_sync = o;   o += al256(p.fused.sync_bytes);
  p.off_ring = o;
  const size_t per_depth = (size_t)p.fused.g.smax * p.fused.entry_bytes;
  p.total_min = o + (size_t)p.fused.depth_min * per_depth;
  p.total_pref = o + (size_t)p.fused.depth_pref * per_depth;
  return p;
}

}  // namespace awqk

using namespace awqk;

extern "C" int awqk_abs_colsum(const void* x, int dtype, int64_t T, int64_t K, double* colsum, void* stream) {
  if (!x || !colsum || T <= 0 || K <= 0) return AWQK_E_BADARG;
  DeviceGuard guard(x);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows_per_cta = 64;
  dim3 grid((unsigned)ceil_div(K, 256), (unsigned)ceil_div(T, rows_per_cta));
  if (dtype == AWQK_BF16)
    abs_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, K, rows_per_cta, colsum);
  else if (dtype == AWQK_FP16)
    abs_colsum_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), T, K, rows_per_cta, colsum);
  else if (dtype == AWQK_FP32)
    abs_colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), T, K, rows_per_cta, colsum);
  else
    return AWQK_E_UNSUPPORTED;
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

extern "C" int awqk_alpha_grid(const double* colsum, int64_t T, int64_t K, int n_grid, float* s_grid,
                               float* workspace_2n, void* stream) {
  if (!colsum || !s_grid || !workspace_2n || T <= 0 || K <= 0 || n_grid <= 0 || n_grid > 256) return AWQK_E_BADARG;
  DeviceGuard guard(colsum);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned int* mnmx = reinterpret_cast<unsigned int*>(workspace_2n);
  alpha_grid_init<<<1, 256, 0, st>>>(n_grid, mnmx);
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(K, 256), 64), (unsigned)n_grid);
  alpha_grid_raw<<<grid, 256, 0, st>>>(colsum, T, K, n_grid, s_grid, mnmx);
  alpha_grid_norm<<<grid, 256, 0, st>>>(K, n_grid, s_grid, mnmx);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

extern "C" int awqk_fakequant_delta(const void* w, int dtype, int64_t C, int64_t K, int group_size, int bits,
                                    int symmetric, const float* s, int n_s, void* dw_bf16, void* stream) {
  if (!w || !s || !dw_bf16 || C <= 0 || K <= 0 || n_s <= 0 || (bits != 4 && bits != 8)) return AWQK_E_BADARG;
  if (!(group_size == 32 || group_size == 64 || group_size == 128) || (K % group_size) != 0) return AWQK_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(dw_bf16)) & 15u)
    return AWQK_E_ALIGN;
  DeviceGuard guard(w);
  if (guard.status != AWQK_OK) return guard.status;
  return launch_fakequant_delta(w, dtype, C, K, group_size, bits, symmetric != 0, s, n_s,
                                reinterpret_cast<__nv_bfloat16*>(dw_bf16), reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int awqk_sqerr_gemm(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s,
                               double* err, void* stream) {
  if (!x_bf16 || !dw_bf16 || !err || T <= 0 || C <= 0 || K <= 0 || n_s <= 0) return AWQK_E_BADARG;
  if ((K % 8) != 0) return AWQK_E_UNSUPPORTED;    // TMA needs 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(dw_bf16)) & 15u) return AWQK_E_ALIGN;
  if (T > 0x7FFFFFFF || C > 0x7FFFFFFF || K > 0x7FFFFFFF) return AWQK_E_BADARG;
  DeviceGuard guard(x_bf16);
  if (guard.status != AWQK_OK) return guard.status;
  return launch_sqerr_gemm2(x_bf16, dw_bf16, T, C, K, n_s, err, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t awqk_workspace_bytes(int64_t C, int64_t K, int64_t T, int n_grid, int have_s_grid, size_t* minimum) {
  if (C <= 0 || K <= 0 || T <= 0 || n_grid <= 0 || n_grid > 256) {
    if (minimum) *minimum = 0;
    return 0;
  }
  const SearchPlan p = plan_search(C, K, T, n_grid, have_s_grid == 0);
  if (p.rc != AWQK_OK) {
    if (minimum) *minimum = 0;
    return 0;
  }
  if (minimum) *minimum = p.total_min;
  return p.total_pref;
}

extern "C" int awqk_scale_search(const void* w, int dtype, int64_t C, int64_t K, const void* x_bf16, int64_t T,
                                 const float* s_grid_in, int n_grid, int group_size, int bits, int symmetric,
                                 double* err_mean, int32_t* best_idx, float* best_s, int32_t* q_unpacked,
                                 uint32_t* q_packed, void* scales_f16, int32_t* zp, uint32_t* zp_packed,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!w || !x_bf16 || !best_s || !workspace || C <= 0 || K <= 0 || T <= 0 || n_grid <= 0 || n_grid > 256 ||
      (bits != 4 && bits != 8))
    return AWQK_E_BADARG;
  if (dtype != AWQK_BF16 && dtype != AWQK_FP16 && dtype != AWQK_FP32) return AWQK_E_UNSUPPORTED;
  if (!(group_size == 32 || group_size == 64 || group_size == 128) || (K % group_size) != 0 || (K % 64) != 0)
    return AWQK_E_UNSUPPORTED;
  if (T > 0x7FFFFFFF || C > 0x7FFFFFFF || K > 0x7FFFFFFF) return AWQK_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(s_grid_in) |
       reinterpret_cast<uintptr_t>(best_s)) & 15u)
    return AWQK_E_ALIGN;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return AWQK_E_ALIGN;
  DeviceGuard guard(w);
  if (guard.status != AWQK_OK) return guard.status;
  const SearchPlan p = plan_search(C, K, T, n_grid, s_grid_in == nullptr);      // (for the device of w)
  if (p.rc != AWQK_OK) return p.rc;
  if (workspace_bytes < p.total_min) return AWQK_E_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const bool sym = symmetric != 0;

  const float* s_grid = s_grid_in;
  if (s_grid == nullptr) {                       // column statistic and alpha grid from X
    double* colsum = reinterpret_cast<double*>(ws + p.off_colsum);
    float* grid = reinterpret_cast<float*>(ws + p.off_grid);
    AWQK_CUDA(cudaMemsetAsync(colsum, 0, (size_t)K * 8, st));
    int rc = awqk_abs_colsum(x_bf16, AWQK_BF16, T, K, colsum, stream);
    if (rc != AWQK_OK) return rc;
    rc = awqk_alpha_grid(colsum, T, K, n_grid, grid, reinterpret_cast<float*>(ws + p.off_mnmx), stream);
    if (rc != AWQK_OK) return rc;
    s_grid = grid;
  }
  double* err_sum = reinterpret_cast<double*>(ws + p.off_err);
  AWQK_CUDA(cudaMemsetAsync(err_sum, 0, (size_t)n_grid * 8, st));
  const size_t per_depth = (size_t)p.fused.g.smax * p.fused.entry_bytes;
  const int depth = (int)std::min<size_t>((size_t)p.fused.depth_pref, (workspace_bytes - p.off_ring) / per_depth);
  int rc = launch_search_fused(w, dtype, C, K, x_bf16, T, s_grid, n_grid, group_size, bits, sym, err_sum,
                               ws + p.off_sync, ws + p.off_ring, depth, st);
  if (rc != AWQK_OK) return rc;
  search_select_kernel<<<(unsigned)std::min<int64_t>(ceil_div(K, 256), 64), 256, 0, st>>>(
      err_sum, n_grid, (double)T * (double)C, s_grid, K, err_mean, best_idx, best_s);
  AWQK_CUDA(cudaGetLastError());
  if (scales_f16 != nullptr)                      // final AWQ pass: group_quant(fp32(W) * s_best), packed / int32 codes
    return awqk_group_quant(w, dtype, C, K, group_size, bits, symmetric, AWQK_ARITH_FP32, q_unpacked, q_packed,
                            scales_f16, zp, zp_packed, best_s, stream);
  return AWQK_OK;
}
