// K1: fused group quantizer (+ nibble pack).
//
// Replaces the reference's per-group Python loops:
//   AWQQuantizer._quantize_per_group          awq.py:286-374
//   AWQQuantizer._compute_scale_zp_for_group  awq.py:173-213
//   AWQQuantizer._quantize_tensor             awq.py:215-250
//   dtype casts of quantize()                 awq.py:409-412
//
// Two kernels:
//   * group_quant_flat : the bandwidth path.  Valid when K % g == 0 and g in {32,64,128}: the
//     whole [C,K] tensor is then one flat array of C*K/g independent, contiguous groups, so the
//     kernel never needs the row structure.  Each thread owns 8 consecutive elements per "slab"
//     (one 16-byte load for bf16/fp16 -> exactly one packed int4 word), a group is g/8 adjacent
//     lanes, min/max are reduced with xor-shuffles inside that lane segment.  A warp processes
//     UNROLL=4 slabs (1024 elements), a CTA of 256 threads 8192 elements; all 4 loads of a thread
//     are issued before any arithmetic (>= 64 B in flight per thread).
//   * group_quant_generic : one warp per (row, group), any g / ragged K (zero padded) / fp64.
//     Correctness path for shapes the reference accepts but LLM weights never have.
//
// Arithmetic is bit-exact with PyTorch CPU (see awqk_common.cuh): division is IEEE (hoisted
// reciprocal + residual correction on the fast path, __fdiv_rn otherwise), rounding is
// half-to-even, no FMA contraction across reference ops.
#include <cstdlib>

#include "awqk_common.cuh"

namespace awqk {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kSlab = 32 * 8;                      // elements per warp per slab
constexpr int kWarpTile = kSlab * kUnroll;         // 1024
constexpr int kCtaTile = kWarpTile * (kThreads / 32);  // 8192

template <typename T>
struct Vec8;  // 8 consecutive input elements -> 8 floats

template <>
struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = ld_stream16(p); }
  __device__ __forceinline__ void zero() { raw = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <>
struct Vec8<__half> {
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = ld_stream16(p); }
  __device__ __forceinline__ void zero() { raw = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 v = __half22float2(h);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
};
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() {
    a = make_float4(0, 0, 0, 0);
    b = a;
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

__device__ __forceinline__ float fmin_nanprop(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float fmax_nanprop(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

struct QuantOut {
  int32_t* q_unpacked;   // nullable
  uint32_t* q_packed;    // nullable
  __half* scales;
  int32_t* zp;           // nullable
  uint32_t* zp_packed;   // nullable; only written when the flat layout allows it
};

// ------------------------------------------------------------------------------------------
// flat fast path
// ------------------------------------------------------------------------------------------
template <typename InT, int A, int G, int BITS>
__global__ void __launch_bounds__(kThreads)
group_quant_flat(const InT* __restrict__ w, int64_t n_elems, int64_t K, bool sym,
                 const float* __restrict__ col_scale, QuantOut out) {
  constexpr int LPG = G / 8;                     // lanes per group
  constexpr int GPC = kCtaTile / G;              // groups per CTA tile
  constexpr int PER_WORD = 32 / BITS;            // codes per packed word
  constexpr int WORDS = 8 / PER_WORD;            // packed words per 8 elements (1 or 2)
  constexpr uint32_t MAGIC_BITS = 0x4B400000u;   // 1.5 * 2^23
  const float MAGIC = __uint_as_float(MAGIC_BITS);

  __shared__ __half s_scale[GPC];
  __shared__ int32_t s_zp[GPC];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t cta_base = (int64_t)blockIdx.x * kCtaTile;
  const int64_t warp_base = cta_base + (int64_t)warp * kWarpTile;
  const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
  const int iqmin = (int)qmin;

  Vec8<InT> v[kUnroll];
#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int64_t e0 = warp_base + j * kSlab + lane * 8;
    if (e0 < n_elems) v[j].load(w + e0); else v[j].zero();
  }

#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int64_t e0 = warp_base + j * kSlab + lane * 8;
    const bool valid = e0 < n_elems;
    float x[8];
    v[j].unpack(x);
    if (col_scale != nullptr) {  // AWQ per-input-channel scaling: x = float(w) * s[k], fp32
      const int64_t k0 = valid ? (e0 % K) : 0;
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(col_scale + k0));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(col_scale + k0) + 1);
      x[0] = __fmul_rn(x[0], c0.x); x[1] = __fmul_rn(x[1], c0.y);
      x[2] = __fmul_rn(x[2], c0.z); x[3] = __fmul_rn(x[3], c0.w);
      x[4] = __fmul_rn(x[4], c1.x); x[5] = __fmul_rn(x[5], c1.y);
      x[6] = __fmul_rn(x[6], c1.z); x[7] = __fmul_rn(x[7], c1.w);
    }
    float mn = fmin_nanprop(fmin_nanprop(fmin_nanprop(x[0], x[1]), fmin_nanprop(x[2], x[3])),
                            fmin_nanprop(fmin_nanprop(x[4], x[5]), fmin_nanprop(x[6], x[7])));
    float mx = fmax_nanprop(fmax_nanprop(fmax_nanprop(x[0], x[1]), fmax_nanprop(x[2], x[3])),
                            fmax_nanprop(fmax_nanprop(x[4], x[5]), fmax_nanprop(x[6], x[7])));
#pragma unroll
    for (int m = 1; m < LPG; m <<= 1) {
      mn = fmin_nanprop(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
      mx = fmax_nanprop(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }
    const FastGroup fg = group_params_fast<A, BITS>(mn, mx, sym, qmin, qmax);
    float s = fg.scale, zp = fg.zp;

    int n[8];
    uint32_t word[WORDS];
    if (fg.ok) {
      const float r = fg.rcp;
      uint32_t tb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = rnd<A>(div_hoisted(x[i], s, r));
        float b = rnd<A>(__fadd_rn(a, zp));
        float c = fminf(fmaxf(b, qmin), qmax);     // clamp commutes with round-to-integer
        tb[i] = __float_as_uint(__fadd_rn(c, MAGIC));  // RNE to integer in the low mantissa bits
        n[i] = (int)(tb[i] - MAGIC_BITS);
      }
      // word = sum_i (n_i - qmin) << (BITS*i), evaluated mod 2^32 on the raw float bits
#pragma unroll
      for (int wi = 0; wi < WORDS; ++wi) {
        uint32_t acc = 0, bias = 0;
#pragma unroll
        for (int i = 0; i < PER_WORD; ++i) {
          acc += tb[wi * PER_WORD + i] << (BITS * i);
          bias += (MAGIC_BITS + (uint32_t)iqmin) << (BITS * i);
        }
        word[wi] = acc - bias;
      }
    } else {
      const GroupParams gp = group_params<A>(mn, mx, sym, qmin, qmax);   // exact IEEE path
      s = gp.scale;
      zp = gp.zp;
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = quant_exact<A>(x[i], s, zp, qmin, qmax);
#pragma unroll
      for (int wi = 0; wi < WORDS; ++wi) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < PER_WORD; ++i) {
          const int c = n[wi * PER_WORD + i];
          const uint32_t u = (c == INT32_MIN) ? 0u : (uint32_t)(c - iqmin);
          acc |= (u & ((1u << BITS) - 1u)) << (BITS * i);
        }
        word[wi] = acc;
      }
    }

    if (valid) {
      if (out.q_unpacked != nullptr) {
        int4* dst = reinterpret_cast<int4*>(out.q_unpacked + e0);
        st_stream16(dst, make_uint4(n[0], n[1], n[2], n[3]));
        st_stream16(dst + 1, make_uint4(n[4], n[5], n[6], n[7]));
      }
      if (out.q_packed != nullptr) {
        uint32_t* dst = out.q_packed + (e0 / PER_WORD);
        if (WORDS == 1) {
          st_stream4(dst, word[0]);
        } else {
          asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(word[0]),
                       "r"(word[WORDS - 1])
                       : "memory");
        }
      }
    }
    if ((lane % LPG) == 0) {
      const int gi = warp * (kWarpTile / G) + j * (kSlab / G) + lane / LPG;
      s_scale[gi] = __float2half_rn(s);
      s_zp[gi] = f2i_x86(zp);
    }
  }
  __syncthreads();

  // coalesced write-out of the per-group metadata of this CTA tile
  const int64_t g_base = cta_base / G;
  const int64_t n_groups = n_elems / G;
  for (int i = threadIdx.x; i < GPC; i += kThreads) {
    const int64_t gidx = g_base + i;
    if (gidx < n_groups) {
      out.scales[gidx] = s_scale[i];
      if (out.zp != nullptr) out.zp[gidx] = s_zp[i];
    }
  }
  if (out.zp_packed != nullptr) {
    for (int i = threadIdx.x; i < GPC / PER_WORD; i += kThreads) {
      const int64_t g0 = g_base + (int64_t)i * PER_WORD;
      if (g0 < n_groups) {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < PER_WORD; ++k) {
          const int z = s_zp[i * PER_WORD + k];
          const uint32_t u = (z == INT32_MIN || g0 + k >= n_groups) ? 0u : (uint32_t)(z - iqmin);
          acc |= (u & ((1u << BITS) - 1u)) << (BITS * k);
        }
        out.zp_packed[g0 / PER_WORD] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// generic path: one warp per (row, group); fp32-evaluated arithmetic (A) or fp64
// ------------------------------------------------------------------------------------------
template <typename InT>
__device__ __forceinline__ float load_as_float(const InT* p);
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }

__device__ __forceinline__ void emit_code(const QuantOut& out, int64_t row, int64_t K, int64_t k,
                                          int code, int iqmin, int bits, int64_t words_per_row) {
  if (out.q_unpacked != nullptr) out.q_unpacked[row * K + k] = code;
  if (out.q_packed != nullptr) {
    const int per = 32 / bits;
    const uint32_t u = (code == INT32_MIN) ? 0u : ((uint32_t)(code - iqmin) & ((1u << bits) - 1u));
    if (u) atomicOr(out.q_packed + row * words_per_row + k / per, u << (bits * (int)(k % per)));
  }
}

template <typename InT, int A>
__global__ void __launch_bounds__(kThreads)
group_quant_generic(const InT* __restrict__ w, int64_t C, int64_t K, int g, int64_t G, int bits,
                    bool sym, const float* __restrict__ col_scale, QuantOut out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (gw >= C * G) return;
  const int64_t row = gw / G, grp = gw % G;
  const int64_t k0 = grp * g;
  const int64_t k1 = (k0 + g < K) ? (k0 + g) : K;
  const float qmin = sym ? -(float)(1 << (bits - 1)) : 0.0f;
  const float qmax = sym ? (float)((1 << (bits - 1)) - 1) : (float)((1 << bits) - 1);
  const InT* src = w + row * K;

  // zero padding of a ragged last group joins the min/max (awq.py:337-339)
  float mn = (k1 - k0 < g) ? 0.0f : __int_as_float(0x7F800000);
  float mx = (k1 - k0 < g) ? 0.0f : __int_as_float(0xFF800000);
  for (int64_t k = k0 + lane; k < k1; k += 32) {
    float x = load_as_float(src + k);
    if (col_scale != nullptr) x = __fmul_rn(x, col_scale[k]);
    mn = fmin_nanprop(mn, x);
    mx = fmax_nanprop(mx, x);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    mn = fmin_nanprop(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
    mx = fmax_nanprop(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
  }
  const GroupParams gp = group_params<A>(mn, mx, sym, qmin, qmax);
  const int64_t wpr = ceil_div(K * bits, 32);
  for (int64_t k = k0 + lane; k < k1; k += 32) {
    float x = load_as_float(src + k);
    if (col_scale != nullptr) x = __fmul_rn(x, col_scale[k]);
    emit_code(out, row, K, k, quant_exact<A>(x, gp.scale, gp.zp, qmin, qmax), (int)qmin, bits, wpr);
  }
  if (lane == 0) {
    out.scales[row * G + grp] = __float2half_rn(gp.scale);
    const int z = f2i_x86(gp.zp);
    if (out.zp != nullptr) out.zp[row * G + grp] = z;
    if (out.zp_packed != nullptr) {
      const int per = 32 / bits;
      const uint32_t u = (z == INT32_MIN) ? 0u : ((uint32_t)(z - (int)qmin) & ((1u << bits) - 1u));
      if (u) atomicOr(out.zp_packed + row * ceil_div(G * bits, 32) + grp / per, u << (bits * (int)(grp % per)));
    }
  }
}

// fp64 input: the reference computes in double (awq.py does everything in the tensor dtype),
// stores scale/zp through an fp32 buffer (awq.py:327-328) and then fp16 / int32.
__device__ __forceinline__ double dmin_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmin(a, b)); }
__device__ __forceinline__ double dmax_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmax(a, b)); }

__global__ void __launch_bounds__(kThreads)
group_quant_generic_f64(const double* __restrict__ w, int64_t C, int64_t K, int g, int64_t G,
                        int bits, bool sym, QuantOut out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (gw >= C * G) return;
  const int64_t row = gw / G, grp = gw % G;
  const int64_t k0 = grp * g;
  const int64_t k1 = (k0 + g < K) ? (k0 + g) : K;
  const double qmin = sym ? -(double)(1 << (bits - 1)) : 0.0;
  const double qmax = sym ? (double)((1 << (bits - 1)) - 1) : (double)((1 << bits) - 1);
  const double* src = w + row * K;
  double mn = (k1 - k0 < g) ? 0.0 : __longlong_as_double(0x7FF0000000000000LL);
  double mx = (k1 - k0 < g) ? 0.0 : __longlong_as_double(0xFFF0000000000000LL);
  for (int64_t k = k0 + lane; k < k1; k += 32) {
    mn = dmin_nan(mn, src[k]);
    mx = dmax_nan(mx, src[k]);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    mn = dmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
    mx = dmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
  }
  if (sym) {
    double a = dmax_nan(fabs(mn), fabs(mx));
    mn = -a;
    mx = a;
  }
  double s = __ddiv_rn(__dsub_rn(mx, mn), qmax - qmin);
  s = dmax_nan(s, 1e-10);
  double zp = 0.0;
  if (!sym) {
    double r = rint(__dsub_rn(qmin, __ddiv_rn(mn, s)));
    zp = (r != r) ? r : fmin(fmax(r, qmin), qmax);
  }
  const int64_t wpr = ceil_div(K * bits, 32);
  for (int64_t k = k0 + lane; k < k1; k += 32) {
    double r = rint(__dadd_rn(__ddiv_rn(src[k], s), zp));
    int code = (r != r) ? INT32_MIN : __double2int_rz(fmin(fmax(r, qmin), qmax));
    emit_code(out, row, K, k, code, (int)qmin, bits, wpr);
  }
  if (lane == 0) {
    out.scales[row * G + grp] = __float2half_rn(__double2float_rn(s));
    const int z = f2i_x86(__double2float_rn(zp));
    if (out.zp != nullptr) out.zp[row * G + grp] = z;
    if (out.zp_packed != nullptr) {
      const int per = 32 / bits;
      const uint32_t u = (z == INT32_MIN) ? 0u : ((uint32_t)(z - (int)qmin) & ((1u << bits) - 1u));
      if (u) atomicOr(out.zp_packed + row * ceil_div(G * bits, 32) + grp / per, u << (bits * (int)(grp % per)));
    }
  }
}

// packs int32 zero points [C, G] into words when the flat kernel could not (G % per_word != 0)
__global__ void pack_zeros_rows(const int32_t* __restrict__ zp, int64_t C, int64_t G, int bits,
                                int iqmin, uint32_t* __restrict__ out) {
  const int per = 32 / bits;
  const int64_t wpr = ceil_div(G, per);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * wpr) return;
  const int64_t row = idx / wpr, wj = idx % wpr;
  uint32_t acc = 0;
  for (int i = 0; i < per; ++i) {
    const int64_t gi = wj * per + i;
    if (gi < G) {
      const int z = zp[row * G + gi];
      const uint32_t u = (z == INT32_MIN) ? 0u : ((uint32_t)(z - iqmin) & ((1u << bits) - 1u));
      acc |= u << (bits * i);
    }
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------
static bool flat_eligible(int dtype, int64_t K, int g, const void* w) {
  if (dtype == AWQK_FP64) return false;
  if (!(g == 32 || g == 64 || g == 128)) return false;
  if (K % g != 0) return false;
  if ((reinterpret_cast<uintptr_t>(w) & 15u) != 0) return false;
  return true;
}

template <typename InT, int A, int G>
static int launch_flat_bits(const InT* w, int64_t n, int64_t K, int bits, bool sym,
                            const float* cs, QuantOut out, cudaStream_t st) {
  const int64_t ctas = ceil_div(n, kCtaTile);
  if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
  if (bits == 4)
    group_quant_flat<InT, A, G, 4><<<(unsigned)ctas, kThreads, 0, st>>>(w, n, K, sym, cs, out);
  else
    group_quant_flat<InT, A, G, 8><<<(unsigned)ctas, kThreads, 0, st>>>(w, n, K, sym, cs, out);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

template <typename InT, int A>
static int launch_flat(const InT* w, int64_t n, int64_t K, int g, int bits, bool sym,
                       const float* cs, QuantOut out, cudaStream_t st) {
  switch (g) {
    case 32: return launch_flat_bits<InT, A, 32>(w, n, K, bits, sym, cs, out, st);
    case 64: return launch_flat_bits<InT, A, 64>(w, n, K, bits, sym, cs, out, st);
    default: return launch_flat_bits<InT, A, 128>(w, n, K, bits, sym, cs, out, st);
  }
}

template <typename InT, int A>
static int launch_generic(const InT* w, int64_t C, int64_t K, int g, int64_t G, int bits, bool sym,
                          const float* cs, QuantOut out, cudaStream_t st) {
  const int64_t ctas = ceil_div(C * G, kThreads / 32);
  if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
  group_quant_generic<InT, A><<<(unsigned)ctas, kThreads, 0, st>>>(w, C, K, g, G, bits, sym, cs, out);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

// awqk_group_quant_tma.cu
int launch_group_quant_tma(const void* w, int dtype, int64_t n_elems, int g, int bits, bool sym, int arith,
                           uint32_t* q_packed, int32_t* q_unpacked, void* scales, int32_t* zp, uint32_t* zp_packed,
                           int zq_log2, cudaStream_t st);

bool group_quant_tma_cs_eligible(int64_t C, int64_t K);
int launch_group_quant_tma_cs(const void* w, int dtype, int64_t C, int64_t K, int g, bool sym, const float* col_scale,
                              uint32_t* q_packed, int32_t* q_unpacked, void* scales, int32_t* zp, uint32_t* zp_packed,
                              cudaStream_t st);
int launch_group_quant_tma_cs_batch(const awqk_quant_item* items, int n, int dtype, int g, bool sym, cudaStream_t st);
int cs_plan_describe(const int64_t* C, const int64_t* K, int n, int g, bool unpacked, int sms, int64_t* summary, int64_t* rows,
                     int max_items);
constexpr int kCsMaxBatch = 31;   // = kV2MaxTensors

// does awqk_group_quant send this call to the column-slab kernel?
static bool cs_path(int dtype, int64_t C, int64_t K, int group_size, int bits, int arith, const void* w,
                    const uint32_t* q_packed, const int32_t* q_unpacked, const float* col_scale) {
  return bits == 4 && arith == AWQK_ARITH_FP32 && (dtype == AWQK_BF16 || dtype == AWQK_FP16) &&
         (q_packed != nullptr || q_unpacked != nullptr) && col_scale != nullptr &&
         (reinterpret_cast<uintptr_t>(q_packed) & 15u) == 0 && flat_eligible(dtype, K, group_size, w) &&
         group_quant_tma_cs_eligible(C, K);
}

}  // namespace awqk

using namespace awqk;

extern "C" int awqk_group_quant_path(int dtype, int64_t C, int64_t K, int group_size, int bits,
                                     int arith, const void* w) {
  if (C <= 0 || K <= 0 || group_size <= 0 || (bits != 4 && bits != 8)) return AWQK_E_BADARG;
  if (dtype < AWQK_BF16 || dtype > AWQK_FP64) return AWQK_E_BADARG;
  if (arith != AWQK_ARITH_NATIVE && arith != AWQK_ARITH_FP32) return AWQK_E_BADARG;
  return flat_eligible(dtype, K, group_size, w) ? 1 : 0;
}

extern "C" int awqk_group_quant(const void* w, int dtype, int64_t C, int64_t K, int group_size,
                                int bits, int symmetric, int arith, int32_t* q_unpacked,
                                uint32_t* q_packed, void* scales_f16, int32_t* zp,
                                uint32_t* zp_packed, const float* col_scale, void* stream) {
  if (w == nullptr || scales_f16 == nullptr) return AWQK_E_BADARG;
  const int path = awqk_group_quant_path(dtype, C, K, group_size, bits, arith, w);
  if (path < 0) return path;
  if (col_scale != nullptr && (arith != AWQK_ARITH_FP32 || dtype == AWQK_FP64)) return AWQK_E_UNSUPPORTED;
  if (dtype == AWQK_FP64 && arith == AWQK_ARITH_FP32) return AWQK_E_UNSUPPORTED;
  if (q_unpacked && (reinterpret_cast<uintptr_t>(q_unpacked) & 15u)) return AWQK_E_ALIGN;
  if (q_packed && (reinterpret_cast<uintptr_t>(q_packed) & 7u)) return AWQK_E_ALIGN;
  if (col_scale && (reinterpret_cast<uintptr_t>(col_scale) & 15u)) return AWQK_E_ALIGN;

  DeviceGuard guard(w);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool sym = symmetric != 0;
  const int64_t G = ceil_div(K, group_size);
  const int per = 32 / bits;
  const int iqmin = sym ? -(1 << (bits - 1)) : 0;
  QuantOut out{q_unpacked, q_packed, reinterpret_cast<__half*>(scales_f16), zp, zp_packed};

  if (path == 1) {
    const bool tma_plain = (dtype == AWQK_BF16 || dtype == AWQK_FP16 || (dtype == AWQK_FP32 && bits == 4)) &&
                           (q_packed != nullptr || q_unpacked != nullptr) && col_scale == nullptr &&
                           (reinterpret_cast<uintptr_t>(q_packed) & 15u) == 0;
    // rows of 1 / 2 / 4 groups (fewer than a packed word): K1 v2 writes one zero-padded word per row itself
    const int full_log2 = (bits == 4) ? 3 : 2;                 // log2(zero points per word)
    const int zq_log2 = (G % per == 0) ? full_log2 : (G == 1) ? 0 : (G == 2) ? 1 : (G == 4 && bits == 4) ? 2 : full_log2;
    const bool flat_zp = (G % per) == 0 || (tma_plain && zq_log2 < full_log2);
    int32_t* zp_for_pack = zp;
    if (zp_packed != nullptr && !flat_zp) {
      if (zp == nullptr) return AWQK_E_WORKSPACE;  // row-wise zero packing needs the int32 zeros
      out.zp_packed = nullptr;
    }
    const int64_t n = C * K;
    int rc;
    if (tma_plain) {
      // K1 v2: TMA-staged, packed-math kernel (int4 pack path and the reference's int32 code layout)
      rc = launch_group_quant_tma(w, dtype, n, group_size, bits, sym, arith, q_packed, q_unpacked, scales_f16, zp,
                                  out.zp_packed, zq_log2, st);
    } else if (cs_path(dtype, C, K, group_size, bits, arith, w, q_packed, q_unpacked, col_scale)) {
      // K1 v2 CS: column-slab kernel, per-input-channel scales in a shared-memory table (final AWQ pass)
      rc = launch_group_quant_tma_cs(w, dtype, C, K, group_size, sym, col_scale, q_packed, q_unpacked, scales_f16, zp,
                                     out.zp_packed, st);
    } else
    // fp32 arithmetic for any input when arith == FP32; otherwise the input's own dtype
    if (dtype == AWQK_BF16) {
      auto p = reinterpret_cast<const __nv_bfloat16*>(w);
      rc = (arith == AWQK_ARITH_FP32)
               ? launch_flat<__nv_bfloat16, AR_F32>(p, n, K, group_size, bits, sym, col_scale, out, st)
               : launch_flat<__nv_bfloat16, AR_BF16>(p, n, K, group_size, bits, sym, nullptr, out, st);
    } else if (dtype == AWQK_FP16) {
      auto p = reinterpret_cast<const __half*>(w);
      rc = (arith == AWQK_ARITH_FP32)
               ? launch_flat<__half, AR_F32>(p, n, K, group_size, bits, sym, col_scale, out, st)
               : launch_flat<__half, AR_F16>(p, n, K, group_size, bits, sym, nullptr, out, st);
    } else {
      rc = launch_flat<float, AR_F32>(reinterpret_cast<const float*>(w), n, K, group_size, bits, sym,
                                      col_scale, out, st);
    }
    if (rc != AWQK_OK) return rc;
    if (zp_packed != nullptr && !flat_zp) {
      const int64_t words = C * ceil_div(G, per);
      pack_zeros_rows<<<(unsigned)ceil_div(words, 256), 256, 0, st>>>(zp_for_pack, C, G, bits, iqmin, zp_packed);
      AWQK_CUDA(cudaGetLastError());
    }
    return AWQK_OK;
  }

  // generic path: packed outputs are built with atomicOr -> zero them first
  if (q_packed != nullptr)
    AWQK_CUDA(cudaMemsetAsync(q_packed, 0, (size_t)(C * ceil_div(K * bits, 32)) * 4, st));
  if (zp_packed != nullptr)
    AWQK_CUDA(cudaMemsetAsync(zp_packed, 0, (size_t)(C * ceil_div(G * bits, 32)) * 4, st));
  if (dtype == AWQK_FP64) {
    const int64_t ctas = ceil_div(C * G, kThreads / 32);
    if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
    group_quant_generic_f64<<<(unsigned)ctas, kThreads, 0, st>>>(reinterpret_cast<const double*>(w), C, K,
                                                                 group_size, G, bits, sym, out);
    AWQK_CUDA(cudaGetLastError());
    return AWQK_OK;
  }
  if (dtype == AWQK_BF16) {
    auto p = reinterpret_cast<const __nv_bfloat16*>(w);
    return (arith == AWQK_ARITH_FP32)
               ? launch_generic<__nv_bfloat16, AR_F32>(p, C, K, group_size, G, bits, sym, col_scale, out, st)
               : launch_generic<__nv_bfloat16, AR_BF16>(p, C, K, group_size, G, bits, sym, nullptr, out, st);
  }
  if (dtype == AWQK_FP16) {
    auto p = reinterpret_cast<const __half*>(w);
    return (arith == AWQK_ARITH_FP32)
               ? launch_generic<__half, AR_F32>(p, C, K, group_size, G, bits, sym, col_scale, out, st)
               : launch_generic<__half, AR_F16>(p, C, K, group_size, G, bits, sym, nullptr, out, st);
  }
  return launch_generic<float, AR_F32>(reinterpret_cast<const float*>(w), C, K, group_size, G, bits, sym,
                                       col_scale, out, st);
}

extern "C" int awqk_group_quant_batch(const awqk_quant_item* items, int n_items, int dtype, int group_size, int bits,
                                      int symmetric, int arith, void* stream) {
  if (n_items < 0 || (n_items > 0 && items == nullptr)) return AWQK_E_BADARG;
  if (n_items == 0) return AWQK_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // one launch holds tensors with the same set of outputs: bucket = (int32 codes, zp, zp_packed, q_packed)
  awqk_quant_item batch[16][kCsMaxBatch];
  int fill[16] = {0};
  DeviceGuard guard(items[0].w);
  if (guard.status != AWQK_OK) return guard.status;
  auto flush = [&](int k) -> int {
    if (fill[k] == 0) return AWQK_OK;
    const int rc = launch_group_quant_tma_cs_batch(batch[k], fill[k], dtype, group_size, symmetric != 0, st);
    fill[k] = 0;
    return rc;
  };
  for (int i = 0; i < n_items; ++i) {
    const awqk_quant_item& it = items[i];
    if (it.w == nullptr || it.scales_f16 == nullptr) return AWQK_E_BADARG;
    const int path = awqk_group_quant_path(dtype, it.C, it.K, group_size, bits, arith, it.w);
    if (path < 0) return path;
    const bool aligned = !(it.q_unpacked && (reinterpret_cast<uintptr_t>(it.q_unpacked) & 15u)) &&
                         !(it.col_scale && (reinterpret_cast<uintptr_t>(it.col_scale) & 15u));
    if (path == 1 && aligned &&
        cs_path(dtype, it.C, it.K, group_size, bits, arith, it.w, it.q_packed, it.q_unpacked, it.col_scale)) {
      const int k = (it.q_unpacked ? 1 : 0) | (it.zp ? 2 : 0) | (it.zp_packed ? 4 : 0) | (it.q_packed ? 8 : 0);
      batch[k][fill[k]++] = it;
      if (fill[k] == kCsMaxBatch) {
        const int rc = flush(k);
        if (rc != AWQK_OK) return rc;
      }
      continue;
    }
    const int rc = awqk_group_quant(it.w, dtype, it.C, it.K, group_size, bits, symmetric, arith, it.q_unpacked, it.q_packed,
                                    it.scales_f16, it.zp, it.zp_packed, it.col_scale, stream);
    if (rc != AWQK_OK) return rc;
  }
  for (int k = 0; k < 16; ++k) {
    const int rc = flush(k);
    if (rc != AWQK_OK) return rc;
  }
  return AWQK_OK;
}

extern "C" int awqk_group_quant_batch_plan(const int64_t* C, const int64_t* K, int n_tensors, int group_size, int with_int32_codes,
                                           int sms, int64_t* summary5, int64_t* items4, int max_items) {
  if (C == nullptr || K == nullptr || summary5 == nullptr || (max_items > 0 && items4 == nullptr)) return AWQK_E_BADARG;
  if (!(group_size == 32 || group_size == 64 || group_size == 128)) return AWQK_E_BADARG;
  return cs_plan_describe(C, K, n_tensors, group_size, with_int32_codes != 0, sms, summary5, items4, max_items);
}
