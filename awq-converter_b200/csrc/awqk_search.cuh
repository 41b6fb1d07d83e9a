// Device code shared by the two producers of the search's delta operand
//   dW_i = bf16( W - dequant(group_quant(W * s_i)) / s_i )        (oracle/awq_oracle.py::fake_quant_delta)
// -- the stand-alone kernel (awqk_search.cu, awqk_fakequant_delta) and the producer warps of the fused
// search kernel (awqk_search_fused.cu).  One lane owns 16 consecutive elements of a W row; a group of G
// elements is G/16 adjacent lanes.  fp32 arithmetic, every operation of the oracle reproduced exactly
// (hoisted-reciprocal IEEE divisions, magic-constant round-half-even), so both producers are bit-identical.
#pragma once
#include "awqk_common.cuh"

namespace awqk {

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

// 16 consecutive weights -> 8 fp32 pairs
template <typename T>
struct Raw16;
template <>
struct Raw16<__nv_bfloat16> {
  uint4 a, b;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    a = __ldg(reinterpret_cast<const uint4*>(p));
    b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() { a = b = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float2 (&f)[8]) const {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
};
template <>
struct Raw16<__half> {
  uint4 a, b;
  __device__ __forceinline__ void load(const __half* p) {
    a = __ldg(reinterpret_cast<const uint4*>(p));
    b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() { a = b = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float2 (&f)[8]) const {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
  }
};
template <>
struct Raw16<float> {
  float4 v[4];
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(p) + c);
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void unpack(float2 (&f)[8]) const {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f[2 * c] = make_float2(v[c].x, v[c].y);
      f[2 * c + 1] = make_float2(v[c].z, v[c].w);
    }
  }
};

// wv: 16 weights (fp32), sv: their column scales s_i[k], rs(i): refined_rcp of pair i of sv (a functor, so that
// the caller decides whether the reciprocals live in registers or are fetched on use).  Must be called by all
// lanes of the warp that share groups (shuffles).  o: the 16 deltas as 8 packed bf16 pairs.
template <int G, int BITS, typename RcpOf>
__device__ __forceinline__ void delta16(const float2 (&wv)[8], const float2 (&sv)[8], RcpOf rs, bool sym,
                                        float qmin, float qmax, uint32_t (&o)[8]) {
  constexpr int LPG = G / 16;
  const float2 magic2 = make_float2(12582912.0f, 12582912.0f);
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = __fmul2_rn(wv[i], sv[i]);                     // Ws = W * s
  float mn = dq_fmin_nan(x[0].x, x[0].y), mx = dq_fmax_nan(x[0].x, x[0].y);
#pragma unroll
  for (int i = 1; i < 8; ++i) {                                    // 3-input FMNMX: one instruction per pair
    mn = dq_fmin3_nan(mn, x[i].x, x[i].y);
    mx = dq_fmax3_nan(mx, x[i].x, x[i].y);
  }
#pragma unroll
  for (int m = 1; m < LPG; m <<= 1) {
    mn = dq_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
    mx = dq_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
  }
  const FastGroup fg = group_params_fast<AR_F32, BITS>(mn, mx, sym, qmin, qmax);
  // asymmetric groups that straddle zero cannot produce a code outside [qmin, qmax] (FastGroup::noclamp): when that
  // holds for every group of the warp -- the normal case -- the four FMNMX per pair of the clamp are skipped
  const bool noclamp = __all_sync(0xFFFFFFFFu, fg.ok && fg.noclamp);
  if (fg.ok) {
    const float2 r2 = make_float2(fg.rcp, fg.rcp), ns2 = make_float2(-fg.scale, -fg.scale);
    // rint(v) - zp = (v + M) - (M + zp): M + zp is an exact integer below 2^24, the difference of two such
    // integers is exact -- one packed add instead of two
    const float2 zp2 = make_float2(fg.zp, fg.zp);
    const float2 nmz2 = make_float2(-(12582912.0f + fg.zp), -(12582912.0f + fg.zp));
    const float2 sc2 = make_float2(fg.scale, fg.scale);
    auto finish = [&](int i, float2 v) {
      const float2 d = __fmul2_rn(__fadd2_rn(__fadd2_rn(v, magic2), nmz2), sc2);   // (rint(v) - zp) * scale
      const float2 ri = rs(i);
      const float2 h0 = __fmul2_rn(d, ri);
      const float2 nsv = make_float2(-sv[i].x, -sv[i].y);
      const float2 what = __ffma2_rn(__ffma2_rn(nsv, h0, d), ri, h0);              // deq / s, exact
      const float2 dl = __fadd2_rn(wv[i], make_float2(-what.x, -what.y));          // W - W^
      const __nv_bfloat162 b = __float22bfloat162_rn(dl);
      o[i] = *reinterpret_cast<const uint32_t*>(&b);
    };
    if (noclamp) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 q0 = __fmul2_rn(x[i], r2);
        const float2 q = __ffma2_rn(__ffma2_rn(ns2, q0, x[i]), r2, q0);            // x / scale, exact
        finish(i, __fadd2_rn(q, zp2));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 q0 = __fmul2_rn(x[i], r2);
        const float2 q = __ffma2_rn(__ffma2_rn(ns2, q0, x[i]), r2, q0);            // x / scale, exact
        float2 v = __fadd2_rn(q, zp2);
        v.x = fminf(fmaxf(v.x, qmin), qmax);                                       // clamp commutes with rint
        v.y = fminf(fmaxf(v.y, qmin), qmax);
        finish(i, v);
      }
    }
  } else {                                                         // non-finite / constant / extreme groups
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
}

// ---- launchers shared between the translation units of the search ------------------------------------------
int launch_sqerr_gemm2(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s, double* err,
                       cudaStream_t st);
int launch_fakequant_delta(const void* w, int dtype, int64_t C, int64_t K, int g, int bits, bool sym, const float* s,
                           int n_s, __nv_bfloat16* dw, cudaStream_t st);

// awqk_search_fused.cu: producer warps + tcgen05 GEMM in one persistent kernel.  The delta operand lives in a
// ring of panel entries (256 W rows x Kc columns, bf16) that stays L2 resident; `sync` holds the per-entry
// ready / done counters (zeroed by the launcher).
struct FusedGeom {                // geometry of one launch (host-computed, identical in planning and launch)
  int K, Kc, panels;              // columns, panel width (2048, or K itself when K <= 2048), ceil(K / Kc)
  int n_grid, mp_tiles, n_tiles;
  int smax, depth, ring;          // slab positions per wave, FIFO depth in panels, ring = smax * depth entries
  int sym;
};
struct FusedPlan {
  FusedGeom g;
  int pairs;                      // CTA pairs of the launch
  int64_t n_entries;              // ring entries over the whole launch (counter arrays)
  size_t sync_bytes, entry_bytes; // counters; one ring entry
  int depth_min, depth_pref;      // ring bytes = g.smax * depth * entry_bytes
};
int fused_plan(int64_t C, int64_t K, int64_t T, int n_grid, FusedPlan* plan);    // queries the current device
int launch_search_fused(const void* w, int dtype, int64_t C, int64_t K, const void* x_bf16, int64_t T,
                        const float* s_grid, int n_grid, int g, int bits, bool sym, double* err_sum, void* sync,
                        void* ring_base, int depth, cudaStream_t st);

}  // namespace awqk
