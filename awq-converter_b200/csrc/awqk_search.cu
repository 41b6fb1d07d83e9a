// K2: activation-aware alpha search.  No reference counterpart: the definition
// is oracle/awq_oracle.py::search_scales, which composes the reference's own group quantizer
// (awq.py:173-250, fp32 arithmetic).
//
//   awqk_abs_colsum      sum_t |X[t,k]| in fp64 (exact for bf16 inputs -> order independent)
//   awqk_alpha_grid      s_i[k] = clamp(m[k]^(i/n), 1e-4) / sqrt(max_k * min_k),  m = colsum / T
//   awqk_fakequant_delta dW_i = bf16( W - dequant(group_quant(W * s_i)) / s_i )   (bandwidth kernel)
//   awqk_sqerr_gemm      err_i = sum_{t,c} ( X . dW_i^T )^2   -- the dense contraction:
//                        tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands staged by TMA into
//                        128-byte-swizzled shared memory through a 4-stage mbarrier ring, double
//                        buffered TMEM accumulators, sum-of-squares epilogue fused on tcgen05.ld.
//
// err uses the delta form ||X (W - W^)^T||^2 (identical in exact arithmetic to ||X W^T - X W^^T||^2):
// half the flops, and rounding dW (not W^) to bf16 keeps the relative error of err ~1e-6.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "awqk_common.cuh"

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
// fake-quant delta: one pass over W for all n_s scale vectors
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 r = ld_stream16(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
  const uint4 r = ld_stream16(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ float dq_fmin_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float dq_fmin3_nan(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float dq_fmax3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float dq_fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// thread = 8 consecutive elements of a row; group = G/8 adjacent lanes (flat layout, K % G == 0)
template <typename T, int G, int BITS>
__global__ void __launch_bounds__(256)
fakequant_delta_kernel(const T* __restrict__ w, int64_t n_elems, int64_t K, bool sym,
                       const float* __restrict__ s_grid, int n_s, __nv_bfloat16* __restrict__ dw) {
  constexpr int LPG = G / 8;
  const int64_t e0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  const bool valid = e0 < n_elems;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  float wv[8];
  if (valid) load8<T>(w + e0, wv);
  else {
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = 0.0f;
  }
  const int64_t k0 = valid ? (e0 % K) : 0;
#pragma unroll 1
  for (int a = 0; a < n_s; ++a) {
    const float* sp = s_grid + (int64_t)a * K + k0;
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fmul_rn(wv[i], sv[i]);                 // Ws = W * s
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      mn = dq_fmin_nan(mn, x[i]);
      mx = dq_fmax_nan(mx, x[i]);
    }
#pragma unroll
    for (int m = 1; m < LPG; m <<= 1) {
      mn = dq_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
      mx = dq_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }
    const FastGroup fg = group_params_fast<AR_F32, BITS>(mn, mx, sym, qmin, qmax);
    float sc = fg.scale, zp = fg.zp;
    float qf[8];
    if (fg.ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = __fadd_rn(div_hoisted(x[i], sc, fg.rcp), zp);
        qf[i] = fminf(fmaxf(rintf(v), qmin), qmax);
      }
    } else {
      const GroupParams gp = group_params<AR_F32>(mn, mx, sym, qmin, qmax);
      sc = gp.scale;
      zp = gp.zp;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float r = rintf(__fadd_rn(__fdiv_rn(x[i], sc), zp));
        qf[i] = (r != r) ? r : fminf(fmaxf(r, qmin), qmax);
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      // W^ = ((q - zp) * scale) / s ;  dW = W - W^
      const float d0 = __fsub_rn(wv[i], __fdiv_rn(__fmul_rn(__fsub_rn(qf[i], zp), sc), sv[i]));
      const float d1 = __fsub_rn(wv[i + 1], __fdiv_rn(__fmul_rn(__fsub_rn(qf[i + 1], zp), sc), sv[i + 1]));
      const __nv_bfloat162 b = __floats2bfloat162_rn(d0, d1);
      o[i / 2] = *reinterpret_cast<const uint32_t*>(&b);
    }
    if (valid) st_stream16(dw + (int64_t)a * n_elems + e0, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// ---- v2: 32 consecutive elements per thread (a group of 128 = 4 lanes), packed fp32x2 math,
// reciprocals of the scale vectors precomputed (exact division by the hoisted-reciprocal sequence).
// CTA = 8 rows x 1024 columns: the 8 warps read the same s / 1/s slab (L1 hits).
__global__ void __launch_bounds__(256)
rcp_grid_kernel(const float* __restrict__ s, int64_t n, float* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) r[i] = refined_rcp(s[i]);
}

template <typename T>
__device__ __forceinline__ void load32(const T* p, float2 (&f)[16]);
template <>
__device__ __forceinline__ void load32<__nv_bfloat16>(const __nv_bfloat16* p, float2 (&f)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      f[4 * c + i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
}
template <>
__device__ __forceinline__ void load32<__half>(const __half* p, float2 (&f)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) f[4 * c + i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
  }
}
template <>
__device__ __forceinline__ void load32<float>(const float* p, float2 (&f)[16]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p) + c);
    f[2 * c] = make_float2(r.x, r.y);
    f[2 * c + 1] = make_float2(r.z, r.w);
  }
}

// The 1024-column slab of s_a and 1/s_a is fetched ONCE per CTA and alpha with coalesced 16-byte
// loads (one per thread and array) into shared memory; every thread then reads its 32 floats with
// conflict-free LDS.128 (chunk c of lane l is stored at position (c + l) % 8 of the lane's 128-byte
// row).  Reading 128 contiguous bytes per lane straight from global memory costs 32 L1 wavefronts per
// load instruction and made the first version of this kernel L1-wavefront bound (741 GB/s written).
template <typename T, int G, int BITS>
__global__ void __launch_bounds__(256, 2)
fakequant_delta_v2(const T* __restrict__ w, int64_t C, int64_t K, bool sym, const float* __restrict__ s_grid,
                   const float* __restrict__ r_grid, int n_s_total, __nv_bfloat16* __restrict__ dw) {
  constexpr int LPG = G / 32;
  // blockIdx.z owns a contiguous slice of the alpha grid (more CTAs in flight for small tensors)
  const int per_z = (n_s_total + (int)gridDim.z - 1) / (int)gridDim.z;
  const int a_begin = (int)blockIdx.z * per_z;
  const int n_s = min(per_z, n_s_total - a_begin);
  if (n_s <= 0) return;
  s_grid += (int64_t)a_begin * K;
  r_grid += (int64_t)a_begin * K;
  dw += (int64_t)a_begin * C * K;
  __shared__ __align__(16) float sm_s[2][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.y * 8 + warp;
  const int64_t col0 = (int64_t)blockIdx.x * 1024;
  const int64_t col = col0 + lane * 32;
  const bool valid = row < C && col < K;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  const float2 magic2 = make_float2(12582912.0f, 12582912.0f), nmagic2 = make_float2(-12582912.0f, -12582912.0f);
  float2 wv[16];
  if (valid) load32<T>(w + row * K + col, wv);
  else {
#pragma unroll
    for (int i = 0; i < 16; ++i) wv[i] = make_float2(0.0f, 0.0f);
  }
  // staging: thread t fetches floats [4t, 4t+4) of the slab -> owner lane t/8, chunk t%8
  const int t = threadIdx.x;
  const int64_t gcol = col0 + 4 * t;
  const bool gvalid = gcol < K;                       // K % 32 == 0 -> a float4 is fully valid or not
  const int st_off = (t >> 3) * 32 + (((t & 7) + (t >> 3)) & 7) * 4;      // float index in the slab buffer
  int ld_off[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) ld_off[c] = lane * 32 + ((c + lane) & 7) * 4;

  __nv_bfloat16* out = dw + row * K + col;
  const int64_t plane = C * K;
  float4 ns = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gvalid) ns = __ldg(reinterpret_cast<const float4*>(s_grid + gcol));
#pragma unroll 1
  for (int a = 0; a < n_s; ++a) {
    const int buf = a & 1;
    *reinterpret_cast<float4*>(&sm_s[buf][st_off]) = ns;
    __syncthreads();                                  // slab a visible; buffer buf^1 is free again after this point
    if (a + 1 < n_s && gvalid)                        // prefetch the next alpha's slab while computing this one
      ns = __ldg(reinterpret_cast<const float4*>(s_grid + (int64_t)(a + 1) * K + gcol));
    float2 sv[16], x[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(&sm_s[buf][ld_off[c]]);
      sv[2 * c] = make_float2(v.x, v.y);
      sv[2 * c + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __fmul2_rn(wv[i], sv[i]);                  // Ws = W * s
    float mn = dq_fmin_nan(x[0].x, x[0].y), mx = dq_fmax_nan(x[0].x, x[0].y);
#pragma unroll
    for (int i = 1; i < 16; ++i) {
      mn = dq_fmin_nan(mn, dq_fmin_nan(x[i].x, x[i].y));
      mx = dq_fmax_nan(mx, dq_fmax_nan(x[i].x, x[i].y));
    }
#pragma unroll
    for (int m = 1; m < LPG; m <<= 1) {
      mn = dq_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
      mx = dq_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }
    const FastGroup fg = group_params_fast<AR_F32, BITS>(mn, mx, sym, qmin, qmax);
    uint32_t o[16];
    if (fg.ok) {
      const float2 r2 = make_float2(fg.rcp, fg.rcp), ns2 = make_float2(-fg.scale, -fg.scale);
      const float2 zp2 = make_float2(fg.zp, fg.zp), nzp2 = make_float2(-fg.zp, -fg.zp);
      const float2 sc2 = make_float2(fg.scale, fg.scale);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * c + h;
          // 1/s on the fly (MUFU + one Newton step): trades two shared-memory slabs for ALU work --
          // the kernel is L1/shared wavefront bound, not issue bound
          const float2 rs = make_float2(refined_rcp(sv[i].x), refined_rcp(sv[i].y));
          const float2 q0 = __fmul2_rn(x[i], r2);
          const float2 q = __ffma2_rn(__ffma2_rn(ns2, q0, x[i]), r2, q0);          // x / scale, exact
          float2 v = __fadd2_rn(q, zp2);
          v.x = fminf(fmaxf(v.x, qmin), qmax);                                     // clamp commutes with rint
          v.y = fminf(fmaxf(v.y, qmin), qmax);
          const float2 qf = __fadd2_rn(__fadd2_rn(v, magic2), nmagic2);            // rint (half-to-even)
          const float2 d = __fmul2_rn(__fadd2_rn(qf, nzp2), sc2);                  // (q - zp) * scale
          const float2 h0 = __fmul2_rn(d, rs);
          const float2 nsv = make_float2(-sv[i].x, -sv[i].y);
          const float2 what = __ffma2_rn(__ffma2_rn(nsv, h0, d), rs, h0);          // deq / s, exact
          const float2 dl = __fadd2_rn(wv[i], make_float2(-what.x, -what.y));      // W - W^
          const __nv_bfloat162 b = __float22bfloat162_rn(dl);
          o[i] = *reinterpret_cast<const uint32_t*>(&b);
        }
      }
    } else {
      const GroupParams gp = group_params<AR_F32>(mn, mx, sym, qmin, qmax);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float d2[2];
        const float xs[2] = {x[i].x, x[i].y}, ss[2] = {sv[i].x, sv[i].y}, ww[2] = {wv[i].x, wv[i].y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float r = rintf(__fadd_rn(__fdiv_rn(xs[h], gp.scale), gp.zp));
          const float qf = (r != r) ? r : fminf(fmaxf(r, qmin), qmax);
          d2[h] = __fsub_rn(ww[h], __fdiv_rn(__fmul_rn(__fsub_rn(qf, gp.zp), gp.scale), ss[h]));
        }
        const __nv_bfloat162 b = __floats2bfloat162_rn(d2[0], d2[1]);
        o[i] = *reinterpret_cast<const uint32_t*>(&b);
      }
    }
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)a * plane);
#pragma unroll
      for (int c = 0; c < 4; ++c) st_stream16(dst + c, make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]));
    }
  }
}

// ---- v4: same scheme as v2 with 16 elements per thread (a group of 128 = 8 lanes): half the
// per-thread state -> <= 80 registers -> 3 CTAs (24 warps) per SM hide the per-alpha barrier and the
// serial group-parameter chain.  CTA = 4 rows x 1024 columns (warps 2r, 2r+1 = the two halves of row r).
template <typename T>
__device__ __forceinline__ void load16(const T* p, float2 (&f)[8]);
template <>
__device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float2 (&f)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      f[4 * c + i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
}
template <>
__device__ __forceinline__ void load16<__half>(const __half* p, float2 (&f)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) f[4 * c + i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
  }
}
template <>
__device__ __forceinline__ void load16<float>(const float* p, float2 (&f)[8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p) + c);
    f[2 * c] = make_float2(r.x, r.y);
    f[2 * c + 1] = make_float2(r.z, r.w);
  }
}

template <typename T, int G, int BITS>
__global__ void __launch_bounds__(256, 3)
fakequant_delta_v4(const T* __restrict__ w, int64_t C, int64_t K, bool sym, const float* __restrict__ s_grid,
                   int n_s_total, __nv_bfloat16* __restrict__ dw) {
  constexpr int LPG = G / 16;
  const int per_z = (n_s_total + (int)gridDim.z - 1) / (int)gridDim.z;
  const int a_begin = (int)blockIdx.z * per_z;
  const int n_s = min(per_z, n_s_total - a_begin);
  if (n_s <= 0) return;
  s_grid += (int64_t)a_begin * K;
  dw += (int64_t)a_begin * C * K;
  __shared__ __align__(16) float sm_s[2][1024];
  __shared__ __align__(16) float sm_r[2][1024];   // refined 1/s, computed once per CTA (4 rows share it)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.y * 4 + (warp >> 1);
  const int64_t col0 = (int64_t)blockIdx.x * 1024;
  const int64_t col = col0 + (warp & 1) * 512 + lane * 16;
  const bool valid = row < C && col < K;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  const float2 magic2 = make_float2(12582912.0f, 12582912.0f), nmagic2 = make_float2(-12582912.0f, -12582912.0f);
  float2 wv[8];
  if (valid) load16<T>(w + row * K + col, wv);
  else {
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = make_float2(0.0f, 0.0f);
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
    __syncthreads();
    if (a + 1 < n_s && gvalid) ns = __ldg(reinterpret_cast<const float4*>(s_grid + (int64_t)(a + 1) * K + gcol));
    float2 sv[8], x[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(&sm_s[buf][ld_off[c]]);
      sv[2 * c] = make_float2(v.x, v.y);
      sv[2 * c + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fmul2_rn(wv[i], sv[i]);                   // Ws = W * s
    float mn = dq_fmin_nan(x[0].x, x[0].y), mx = dq_fmax_nan(x[0].x, x[0].y);
#pragma unroll
    for (int i = 1; i < 8; ++i) {                                  // 3-input FMNMX: one instruction per pair
      mn = dq_fmin3_nan(mn, x[i].x, x[i].y);
      mx = dq_fmax3_nan(mx, x[i].x, x[i].y);
    }
#pragma unroll
    for (int m = 1; m < LPG; m <<= 1) {
      mn = dq_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
      mx = dq_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }
    const FastGroup fg = group_params_fast<AR_F32, BITS>(mn, mx, sym, qmin, qmax);
    uint32_t o[8];
    if (fg.ok) {
      const float2 r2 = make_float2(fg.rcp, fg.rcp), ns2 = make_float2(-fg.scale, -fg.scale);
      // rint(v) - zp = (v + M) - (M + zp): M + zp is an exact integer below 2^24, the difference of two such
      // integers is exact -- one packed add instead of two
      const float2 zp2 = make_float2(fg.zp, fg.zp);
      const float2 nmz2 = make_float2(-(12582912.0f + fg.zp), -(12582912.0f + fg.zp));
      const float2 sc2 = make_float2(fg.scale, fg.scale);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 rr = *reinterpret_cast<const float4*>(&sm_r[buf][ld_off[i >> 1]]);
        const float2 rs = (i & 1) ? make_float2(rr.z, rr.w) : make_float2(rr.x, rr.y);
        const float2 q0 = __fmul2_rn(x[i], r2);
        const float2 q = __ffma2_rn(__ffma2_rn(ns2, q0, x[i]), r2, q0);            // x / scale, exact
        float2 v = __fadd2_rn(q, zp2);
        v.x = fminf(fmaxf(v.x, qmin), qmax);
        v.y = fminf(fmaxf(v.y, qmin), qmax);
        const float2 d = __fmul2_rn(__fadd2_rn(__fadd2_rn(v, magic2), nmz2), sc2); // (rint(v) - zp) * scale
        const float2 h0 = __fmul2_rn(d, rs);
        const float2 nsv = make_float2(-sv[i].x, -sv[i].y);
        const float2 what = __ffma2_rn(__ffma2_rn(nsv, h0, d), rs, h0);            // deq / s, exact
        const float2 dl = __fadd2_rn(wv[i], make_float2(-what.x, -what.y));        // W - W^
        const __nv_bfloat162 b = __float22bfloat162_rn(dl);
        o[i] = *reinterpret_cast<const uint32_t*>(&b);
      }
    } else {
      const GroupParams gp = group_params<AR_F32>(mn, mx, sym, qmin, qmax);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float d2[2];
        const float xs[2] = {x[i].x, x[i].y}, ss[2] = {sv[i].x, sv[i].y}, ww[2] = {wv[i].x, wv[i].y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float r = rintf(__fadd_rn(__fdiv_rn(xs[h], gp.scale), gp.zp));
          const float qf = (r != r) ? r : fminf(fmaxf(r, qmin), qmax);
          d2[h] = __fsub_rn(ww[h], __fdiv_rn(__fmul_rn(__fsub_rn(qf, gp.zp), gp.scale), ss[h]));
        }
        const __nv_bfloat162 b = __floats2bfloat162_rn(d2[0], d2[1]);
        o[i] = *reinterpret_cast<const uint32_t*>(&b);
      }
    }
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)a * plane);
      st_stream16(dst, make_uint4(o[0], o[1], o[2], o[3]));
      st_stream16(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
  }
}

// 32 consecutive elements of a W row, kept in the narrowest register form
template <typename T>
struct Row32;
template <>
struct Row32<__nv_bfloat16> {
  uint32_t r[16];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + c);
      r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0u;
  }
  __device__ __forceinline__ float2 f2(int i) const {
    return make_float2(__uint_as_float(r[i] << 16), __uint_as_float(r[i] & 0xFFFF0000u));
  }
};
template <>
struct Row32<__half> {
  uint32_t r[16];
  __device__ __forceinline__ void load(const __half* p) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + c);
      r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0u;
  }
  __device__ __forceinline__ float2 f2(int i) const { return __half22float2(*reinterpret_cast<const __half2*>(&r[i])); }
};
template <>
struct Row32<float> {
  float2 r[16];
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c);
      r[2 * c] = make_float2(v.x, v.y);
      r[2 * c + 1] = make_float2(v.z, v.w);
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = make_float2(0.0f, 0.0f);
  }
  __device__ __forceinline__ float2 f2(int i) const { return r[i]; }
};

// ---- v3: scale slab in REGISTERS, rows streamed.  A warp keeps s_a / (1/s_a) for its 1024 columns
// in registers and walks `rows_per_warp` rows of W for that alpha: no shared memory (leaves the
// shared-memory bandwidth to a concurrently running GEMM), no block barrier per alpha, the strided
// slab loads amortised over the rows.  W is re-read once per alpha from L2.
template <typename T, int G, int BITS>
__global__ void __launch_bounds__(256, 2)
fakequant_delta_v3(const T* __restrict__ w, int64_t C, int64_t K, bool sym, const float* __restrict__ s_grid,
                   const float* __restrict__ r_grid, int n_s_total, __nv_bfloat16* __restrict__ dw,
                   int rows_per_warp) {
  constexpr int LPG = G / 32;
  const int per_z = (n_s_total + (int)gridDim.z - 1) / (int)gridDim.z;
  const int a_begin = (int)blockIdx.z * per_z;
  const int n_s = min(per_z, n_s_total - a_begin);
  if (n_s <= 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = ((int64_t)blockIdx.y * 8 + warp) * rows_per_warp;
  const int64_t col = (int64_t)blockIdx.x * 1024 + lane * 32;
  const bool cvalid = col < K;
  const int64_t cc = cvalid ? col : 0;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  const float2 magic2 = make_float2(12582912.0f, 12582912.0f), nmagic2 = make_float2(-12582912.0f, -12582912.0f);
  const int64_t plane = C * K;
  int64_t row_end = row0 + rows_per_warp;
  if (row_end > C) row_end = C;
#pragma unroll 1
  for (int a = a_begin; a < a_begin + n_s; ++a) {
    float2 sv[16], rv[16];
    {
      const float4* sp = reinterpret_cast<const float4*>(s_grid + (int64_t)a * K + cc);
      const float4* rp = reinterpret_cast<const float4*>(r_grid + (int64_t)a * K + cc);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 u = __ldg(sp + c), v = __ldg(rp + c);
        sv[2 * c] = make_float2(u.x, u.y); sv[2 * c + 1] = make_float2(u.z, u.w);
        rv[2 * c] = make_float2(v.x, v.y); rv[2 * c + 1] = make_float2(v.z, v.w);
      }
    }
    __nv_bfloat16* outp = dw + (int64_t)a * plane + col;
#pragma unroll 1
    for (int64_t row = row0; row < row_end; ++row) {
      Row32<T> wr;
      if (cvalid) wr.load(w + row * K + col); else wr.zero();
      float mn, mx;
      {
        const float2 x0 = __fmul2_rn(wr.f2(0), sv[0]);
        mn = dq_fmin_nan(x0.x, x0.y);
        mx = dq_fmax_nan(x0.x, x0.y);
#pragma unroll
        for (int i = 1; i < 16; ++i) {
          const float2 x = __fmul2_rn(wr.f2(i), sv[i]);
          mn = dq_fmin_nan(mn, dq_fmin_nan(x.x, x.y));
          mx = dq_fmax_nan(mx, dq_fmax_nan(x.x, x.y));
        }
      }
#pragma unroll
      for (int m = 1; m < LPG; m <<= 1) {
        mn = dq_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
        mx = dq_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
      }
      const FastGroup fg = group_params_fast<AR_F32, BITS>(mn, mx, sym, qmin, qmax);
      uint4* dst = reinterpret_cast<uint4*>(outp + row * K);
      if (fg.ok) {
        const float2 r2 = make_float2(fg.rcp, fg.rcp), ns2 = make_float2(-fg.scale, -fg.scale);
        const float2 zp2 = make_float2(fg.zp, fg.zp), nzp2 = make_float2(-fg.zp, -fg.zp);
        const float2 sc2 = make_float2(fg.scale, fg.scale);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * c + j;
            const float2 wf = wr.f2(i);
            const float2 x = __fmul2_rn(wf, sv[i]);                                   // Ws = W * s
            const float2 q0 = __fmul2_rn(x, r2);
            const float2 q = __ffma2_rn(__ffma2_rn(ns2, q0, x), r2, q0);              // x / scale, exact
            float2 v = __fadd2_rn(q, zp2);
            v.x = fminf(fmaxf(v.x, qmin), qmax);
            v.y = fminf(fmaxf(v.y, qmin), qmax);
            const float2 qf = __fadd2_rn(__fadd2_rn(v, magic2), nmagic2);             // rint (half-to-even)
            const float2 d = __fmul2_rn(__fadd2_rn(qf, nzp2), sc2);                   // (q - zp) * scale
            const float2 h0 = __fmul2_rn(d, rv[i]);
            const float2 nsv = make_float2(-sv[i].x, -sv[i].y);
            const float2 what = __ffma2_rn(__ffma2_rn(nsv, h0, d), rv[i], h0);        // deq / s, exact
            const float2 dl = __fadd2_rn(wf, make_float2(-what.x, -what.y));          // W - W^
            const __nv_bfloat162 b = __float22bfloat162_rn(dl);
            o[j] = *reinterpret_cast<const uint32_t*>(&b);
          }
          if (cvalid) st_stream16(dst + c, make_uint4(o[0], o[1], o[2], o[3]));
        }
      } else {
        const GroupParams gp = group_params<AR_F32>(mn, mx, sym, qmin, qmax);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * c + j;
            float d2[2];
            const float2 wf = wr.f2(i);
            const float ss[2] = {sv[i].x, sv[i].y}, ww[2] = {wf.x, wf.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float xs = __fmul_rn(ww[h], ss[h]);
              const float r = rintf(__fadd_rn(__fdiv_rn(xs, gp.scale), gp.zp));
              const float qf = (r != r) ? r : fminf(fmaxf(r, qmin), qmax);
              d2[h] = __fsub_rn(ww[h], __fdiv_rn(__fmul_rn(__fsub_rn(qf, gp.zp), gp.scale), ss[h]));
            }
            const __nv_bfloat162 b = __floats2bfloat162_rn(d2[0], d2[1]);
            o[j] = *reinterpret_cast<const uint32_t*>(&b);
          }
          if (cvalid) st_stream16(dst + c, make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// tcgen05 GEMM with fused sum-of-squares epilogue
// ------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBN = 256, kBK = 64;          // CTA tile; UMMA 128 x 256 x 16 (x4 per k-block)
constexpr int kGStages = 4;
constexpr int kABytes = kBM * kBK * 2;                 // 16 KiB
constexpr int kBBytes = kBN * kBK * 2;                 // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 192;                      // warp0 TMA, warp1 MMA, warps 2-5 epilogue
constexpr uint32_t kTmemCols = 512;                    // 2 accumulator buffers x 256 fp32 columns

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (1024 B
//   between 8-row groups) | [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kGemmThreads, 1)
sqerr_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dw,
                  int n_s, int m_tiles, int n_tiles, int k_blocks, double* __restrict__ err) {
  extern __shared__ uint8_t gsm_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t raw = g_smem_u32(gsm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gsm = gsm_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + kGStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGStages + 4);
  const uint32_t full0 = g_smem_u32(bars), empty0 = full0 + 8 * kGStages;
  const uint32_t tfull0 = empty0 + 8 * kGStages, tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = n_s * n_tiles * m_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGStages; ++s) {
      g_mbar_init(full0 + 8 * s, 1);
      g_mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      g_mbar_init(tfull0 + 8 * a, 1);
      g_mbar_init(tempty0 + 8 * a, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // one warp owns the TMEM allocation (and the deallocation at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, ph = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles;
        const int rest = tile / m_tiles;
        const int nt = rest % n_tiles;
        const int a = rest / n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          g_mbar_wait(empty0 + 8 * stage, ph);
          g_mbar_expect_tx(full0 + 8 * stage, kStageBytes);
          const uint32_t sa = base + stage * kStageBytes;
          tma_load_2d(sa, &map_x, kb * kBK, mt * kBM, full0 + 8 * stage);
          tma_load_3d(sa + kABytes, &map_dw, kb * kBK, nt * kBN, a, full0 + 8 * stage);
          if (++stage == kGStages) { stage = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    uint32_t stage = 0, ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = (uint32_t)it & 1u;                  // accumulator buffer
      const uint32_t aph = ((uint32_t)it >> 1) & 1u;
      g_mbar_wait(tempty0 + 8 * ab, aph ^ 1u);                // epilogue has drained this buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * kBN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        g_mbar_wait(full0 + 8 * stage, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + stage * kStageBytes;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)                  // +32 bytes (>>4 = 2) per UMMA_K inside the swizzle atom
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, (kb | k) ? 1u : 0u);
          umma_commit(empty0 + 8 * stage);                    // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull0 + 8 * ab);   // accumulator complete
        }
        __syncwarp();
        if (++stage == kGStages) { stage = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: sum of squares of the accumulator =====================
    const uint32_t quarter = (uint32_t)warp & 3u;              // TMEM lanes [32q, 32q+32) for this warp
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int a = (tile / m_tiles) / n_tiles;
      const uint32_t ab = (uint32_t)it & 1u;
      const uint32_t aph = ((uint32_t)it >> 1) & 1u;
      g_mbar_wait(tfull0 + 8 * ab, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * kBN + ((quarter * 32u) << 16);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
      for (int c = 0; c < kBN; c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + c, v0);
        tmem_ld32(taddr + c + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float f0 = __uint_as_float(v0[j]), f1 = __uint_as_float(v1[j]);
          acc0 = __fmaf_rn(f0, f0, acc0);
          acc1 = __fmaf_rn(f1, f1, acc1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) g_mbar_arrive(tempty0 + 8 * ab);          // buffer may be overwritten by the next tile
      double d = (double)acc0 + (double)acc1;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
      if (lane == 0) atomicAdd(err + a, d);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host side: tensor maps through the driver entry point (no libcuda link dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// awqk_search_gemm2.cu: CTA-pair (cta_group::2) version of the same GEMM
int launch_sqerr_gemm2(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s, double* err,
                       cudaStream_t st);

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

template <typename T>
static int launch_delta_v2(const T* w, int64_t C, int64_t K, int g, int bits, bool sym, const float* s, float* r,
                           int n_s, __nv_bfloat16* dw, cudaStream_t st) {
  // default: v2 (shared-memory slab; 1.6 TB/s written).  AWQK_DELTA_V3=1 selects the register-slab
  // variant (no shared memory, slower stand-alone: 1.3 TB/s) for A/B measurements.
  static const bool use_v3 = []() { const char* e = getenv("AWQK_DELTA_V3"); return e && e[0] == '1'; }();
  if (use_v3) {
    const int64_t ns = (int64_t)n_s * K;
    rcp_grid_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, st>>>(s, ns, r);
    const int z = std::min(n_s, 4);
    const int64_t slabs = ceil_div(K, 1024);
    // rows per warp: amortise the slab loads (>= 8 rows) but keep >= ~4 CTAs per SM in flight
    int64_t rpw = 32;
    while (rpw > 8 && slabs * ceil_div(C, 8 * rpw) * z < 600) rpw >>= 1;
    dim3 g3((unsigned)slabs, (unsigned)ceil_div(C, 8 * rpw), (unsigned)z);
    if (g3.y > 65535) return AWQK_E_BADARG;
#define AWQK_DELTA3(GG, BB) fakequant_delta_v3<T, GG, BB><<<g3, 256, 0, st>>>(w, C, K, sym, s, r, n_s, dw, (int)rpw)
    if (bits == 4) {
      if (g == 32) AWQK_DELTA3(32, 4); else if (g == 64) AWQK_DELTA3(64, 4); else AWQK_DELTA3(128, 4);
    } else {
      if (g == 32) AWQK_DELTA3(32, 8); else if (g == 64) AWQK_DELTA3(64, 8); else AWQK_DELTA3(128, 8);
    }
#undef AWQK_DELTA3
    AWQK_CUDA(cudaGetLastError());
    return AWQK_OK;
  }
  static const bool use_v2 = []() { const char* e = getenv("AWQK_DELTA_V2"); return e && e[0] == '1'; }();
  if (!use_v2 && ceil_div(C, 4) <= 65535) {      // default: v4 (16 elements per thread, 3 CTAs/SM)
    // alpha slices: enough CTAs for ~4 waves of 3 x 148, but no more (each slice re-reads its W rows)
    const int64_t base_ctas = ceil_div(K, 1024) * ceil_div(C, 4);
    const int z = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_s, 4), ceil_div(1800, base_ctas)));
    dim3 g4((unsigned)ceil_div(K, 1024), (unsigned)ceil_div(C, 4), (unsigned)z);
#define AWQK_DELTA4(GG, BB) fakequant_delta_v4<T, GG, BB><<<g4, 256, 0, st>>>(w, C, K, sym, s, n_s, dw)
    if (bits == 4) {
      if (g == 32) AWQK_DELTA4(32, 4); else if (g == 64) AWQK_DELTA4(64, 4); else AWQK_DELTA4(128, 4);
    } else {
      if (g == 32) AWQK_DELTA4(32, 8); else if (g == 64) AWQK_DELTA4(64, 8); else AWQK_DELTA4(128, 8);
    }
#undef AWQK_DELTA4
    AWQK_CUDA(cudaGetLastError());
    return AWQK_OK;
  }
  dim3 grid((unsigned)ceil_div(K, 1024), (unsigned)ceil_div(C, 8), (unsigned)std::min(n_s, 4));
  if (grid.y > 65535) return AWQK_E_BADARG;
  // Same shared-memory carve-out as the GEMM (max shared): SMs do not have to be drained and
  // re-configured between the two kernels, so delta(i+1) co-resides with the GEMM of tensor i.
#define AWQK_DELTA2(GG, BB)                                                                                     \
  do {                                                                                                          \
    static std::atomic<uint64_t> cfg{0};                                                                        \
    int dev_ = 0;                                                                                               \
    AWQK_CUDA(cudaGetDevice(&dev_));                                                                            \
    if (!(cfg.load(std::memory_order_acquire) & (1ull << (dev_ & 63)))) {                                       \
      AWQK_CUDA(cudaFuncSetAttribute(fakequant_delta_v2<T, GG, BB>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                     (int)cudaSharedmemCarveoutMaxShared));                                     \
      cfg.fetch_or(1ull << (dev_ & 63), std::memory_order_release);                                             \
    }                                                                                                           \
    fakequant_delta_v2<T, GG, BB><<<grid, 256, 0, st>>>(w, C, K, sym, s, r, n_s, dw);                          \
  } while (0)
  if (bits == 4) {
    if (g == 32) AWQK_DELTA2(32, 4); else if (g == 64) AWQK_DELTA2(64, 4); else AWQK_DELTA2(128, 4);
  } else {
    if (g == 32) AWQK_DELTA2(32, 8); else if (g == 64) AWQK_DELTA2(64, 8); else AWQK_DELTA2(128, 8);
  }
#undef AWQK_DELTA2
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

template <typename T>
static int launch_delta(const T* w, int64_t n, int64_t K, int g, int bits, bool sym, const float* s, int n_s,
                        __nv_bfloat16* dw, cudaStream_t st) {
  const int64_t ctas = ceil_div(n, 256 * 8);
  if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
#define AWQK_DELTA(GG, BB) fakequant_delta_kernel<T, GG, BB><<<(unsigned)ctas, 256, 0, st>>>(w, n, K, sym, s, n_s, dw)
  if (bits == 4) {
    if (g == 32) AWQK_DELTA(32, 4); else if (g == 64) AWQK_DELTA(64, 4); else AWQK_DELTA(128, 4);
  } else {
    if (g == 32) AWQK_DELTA(32, 8); else if (g == 64) AWQK_DELTA(64, 8); else AWQK_DELTA(128, 8);
  }
#undef AWQK_DELTA
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

extern "C" int awqk_fakequant_delta(const void* w, int dtype, int64_t C, int64_t K, int group_size, int bits,
                                    int symmetric, const float* s, int n_s, void* dw_bf16, float* rcp_workspace,
                                    void* stream) {
  if (!w || !s || !dw_bf16 || C <= 0 || K <= 0 || n_s <= 0 || (bits != 4 && bits != 8)) return AWQK_E_BADARG;
  if (!(group_size == 32 || group_size == 64 || group_size == 128) || (K % group_size) != 0) return AWQK_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(dw_bf16)) & 15u)
    return AWQK_E_ALIGN;
  DeviceGuard guard(w);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto out = reinterpret_cast<__nv_bfloat16*>(dw_bf16);
  const bool sym = symmetric != 0;
  if (rcp_workspace != nullptr && (K % 32) == 0 && ceil_div(C, 8) <= 65535) {
    if ((reinterpret_cast<uintptr_t>(rcp_workspace) & 15u) != 0) return AWQK_E_ALIGN;
    if (dtype == AWQK_BF16)
      return launch_delta_v2(reinterpret_cast<const __nv_bfloat16*>(w), C, K, group_size, bits, sym, s, rcp_workspace, n_s, out, st);
    if (dtype == AWQK_FP16)
      return launch_delta_v2(reinterpret_cast<const __half*>(w), C, K, group_size, bits, sym, s, rcp_workspace, n_s, out, st);
    if (dtype == AWQK_FP32)
      return launch_delta_v2(reinterpret_cast<const float*>(w), C, K, group_size, bits, sym, s, rcp_workspace, n_s, out, st);
    return AWQK_E_UNSUPPORTED;
  }
  if (dtype == AWQK_BF16)
    return launch_delta(reinterpret_cast<const __nv_bfloat16*>(w), C * K, K, group_size, bits, sym, s, n_s, out, st);
  if (dtype == AWQK_FP16)
    return launch_delta(reinterpret_cast<const __half*>(w), C * K, K, group_size, bits, sym, s, n_s, out, st);
  if (dtype == AWQK_FP32)
    return launch_delta(reinterpret_cast<const float*>(w), C * K, K, group_size, bits, sym, s, n_s, out, st);
  return AWQK_E_UNSUPPORTED;
}

extern "C" int awqk_sqerr_gemm(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s,
                               double* err, void* stream) {
  if (!x_bf16 || !dw_bf16 || !err || T <= 0 || C <= 0 || K <= 0 || n_s <= 0) return AWQK_E_BADARG;
  if ((K % 8) != 0) return AWQK_E_UNSUPPORTED;    // TMA needs 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(dw_bf16)) & 15u) return AWQK_E_ALIGN;
  if (T > 0x7FFFFFFF || C > 0x7FFFFFFF || K > 0x7FFFFFFF) return AWQK_E_BADARG;
  DeviceGuard guard(x_bf16);
  if (guard.status != AWQK_OK) return guard.status;
  {
    // default: the CTA-pair kernel (tcgen05 cta_group::2, awqk_search_gemm2.cu); AWQK_GEMM_2CTA=0 selects
    // the single-CTA kernel below (kept for A/B measurements; same results)
    static const int two_cta = []() { const char* e = getenv("AWQK_GEMM_2CTA"); return (e && e[0] == '0') ? 0 : 1; }();
    if (two_cta) return launch_sqerr_gemm2(x_bf16, dw_bf16, T, C, K, n_s, err, reinterpret_cast<cudaStream_t>(stream));
  }
  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) return AWQK_E_NODEVICE;

  CUtensorMap map_x, map_dw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {kBK, kBM};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_bf16), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AWQK_E_BADARG;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)C, (cuuint64_t)n_s};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)C * (cuuint64_t)K * 2};
    const cuuint32_t box[3] = {kBK, kBN, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_dw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(dw_bf16), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AWQK_E_BADARG;
  }
  const int m_tiles = (int)ceil_div(T, kBM), n_tiles = (int)ceil_div(C, kBN), k_blocks = (int)ceil_div(K, kBK);
  const int64_t total = (int64_t)n_s * m_tiles * n_tiles;
  if (total > 0x7FFFFFFF) return AWQK_E_BADARG;
  int dev = 0, sms = 0;
  AWQK_CUDA(cudaGetDevice(&dev));
  AWQK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = (size_t)kGStages * kStageBytes + 1024 + 256;
  {
    // the opt-in is per device; set it once per device (atomic bit mask, immutable afterwards)
    static std::atomic<uint64_t> configured{0};
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
      AWQK_CUDA(cudaFuncSetAttribute(sqerr_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured.fetch_or(bit, std::memory_order_release);
    }
  }
  const unsigned grid = (unsigned)std::min<int64_t>(total, sms);
  sqerr_gemm_kernel<<<grid, kGemmThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(map_x, map_dw, n_s, m_tiles,
                                                                                         n_tiles, k_blocks, err);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}
