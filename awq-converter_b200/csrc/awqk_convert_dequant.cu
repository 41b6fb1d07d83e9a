// K3: bf16 -> fp16 conversion          (replaces convert_bf16_to_fp16, tensor_utils.py:10-22)
// K4: group de-quantizer               (replaces AWQQuantizer.dequantize, awq.py:459-539,252-284)
// Both are pure streaming kernels: 16-byte loads/stores, grid-stride free (one pass, one tile per CTA).
#include "awqk_common.cuh"

namespace awqk {

constexpr int kCvThreads = 256;
constexpr int kCvUnroll = 4;  // 4 x 16 B per thread in flight

__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t w) {
  // bf16 -> fp32 is exact; fp32 -> fp16 is one RNE rounding (overflow -> inf, subnormals kept)
  const float lo = __uint_as_float(w << 16);
  const float hi = __uint_as_float(w & 0xFFFF0000u);
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kCvThreads)
bf16_to_fp16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int64_t n) {
  const int64_t tile = (int64_t)blockIdx.x * (kCvThreads * kCvUnroll * 8);
  uint4 v[kCvUnroll];
  int64_t e[kCvUnroll];
#pragma unroll
  for (int j = 0; j < kCvUnroll; ++j) {
    e[j] = tile + ((int64_t)j * kCvThreads + threadIdx.x) * 8;
    if (e[j] + 8 <= n) v[j] = ld_stream16(in + e[j]);
  }
#pragma unroll
  for (int j = 0; j < kCvUnroll; ++j) {
    if (e[j] + 8 <= n) {
      uint4 o;
      o.x = bf16x2_to_f16x2(v[j].x);
      o.y = bf16x2_to_f16x2(v[j].y);
      o.z = bf16x2_to_f16x2(v[j].z);
      o.w = bf16x2_to_f16x2(v[j].w);
      st_stream16(out + e[j], o);
    } else if (e[j] < n) {  // ragged tail: scalar
      for (int64_t i = e[j]; i < n; ++i) {
        const float f = __uint_as_float((uint32_t)in[i] << 16);
        out[i] = __half_as_ushort(__float2half_rn(f));
      }
    }
  }
}

// scalar fallback for bases that are not 16-byte aligned
__global__ void bf16_to_fp16_scalar(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half_as_ushort(__float2half_rn(__uint_as_float((uint32_t)in[i] << 16)));
}

// out = float( fp16_rn( half(q - zp) * scale ) ).  The reference multiplies an int32 tensor by a
// 0-d fp16 tensor: torch promotes to fp16, evaluates in fp32 and rounds once (awq.py:282).
__device__ __forceinline__ float dequant_one(int q, int zp, float scale_f) {
  const int d = (int)((unsigned)q - (unsigned)zp);          // int32 wrap-around like torch's int32 subtraction
  const float diff = __half2float(__int2half_rn(d));        // int32 -> fp16 (RNE, overflow -> inf)
  return __half2float(__float2half_rn(__fmul_rn(diff, scale_f)));
}

// fast path: K % 4 == 0 and g % 4 == 0 -> the 4 elements of a thread share one group; rows on
// blockIdx.x, 1024-column chunks on blockIdx.y (no 64-bit divisions, 16 B in / 16 B out per thread)
__global__ void __launch_bounds__(256)
dequant_rows_kernel(const int32_t* __restrict__ q, const __half* __restrict__ scales, const int32_t* __restrict__ zp,
                    int64_t K, int g, int64_t G, float* __restrict__ out) {
  const int64_t row = blockIdx.x;
  const uint32_t k0 = (blockIdx.y * 256u + threadIdx.x) * 4u;
  if (k0 >= K) return;
  const int64_t base = row * K + k0;
  const int64_t gi = row * G + k0 / (uint32_t)g;
  const uint4 qv = ld_stream16(q + base);
  const int z = __ldg(zp + gi);
  const float sc = __half2float(__ldg(scales + gi));
  uint4 o;
  o.x = __float_as_uint(dequant_one((int)qv.x, z, sc));
  o.y = __float_as_uint(dequant_one((int)qv.y, z, sc));
  o.z = __float_as_uint(dequant_one((int)qv.z, z, sc));
  o.w = __float_as_uint(dequant_one((int)qv.w, z, sc));
  st_stream16(out + base, o);
}

__global__ void __launch_bounds__(256)
dequant_kernel(const int32_t* __restrict__ q, const __half* __restrict__ scales,
               const int32_t* __restrict__ zp, int64_t C, int64_t K, int g, int64_t G,
               float* __restrict__ out) {
  // generic: one thread per element
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * K) return;
  const int64_t row = idx / K, k = idx % K;
  const int64_t gi = row * G + k / g;
  out[idx] = dequant_one(q[idx], zp[gi], __half2float(scales[gi]));
}

__global__ void __launch_bounds__(256)
dequant_packed_kernel(const uint32_t* __restrict__ qw, const __half* __restrict__ scales,
                      const uint32_t* __restrict__ qz, int64_t C, int64_t K, int g, int64_t G,
                      int bits, int iqmin, float* __restrict__ out) {
  const int per = 32 / bits;
  const int64_t wpr = ceil_div(K, per);
  const int64_t zwpr = ceil_div(G, per);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * wpr) return;
  const int64_t row = idx / wpr;
  const int64_t k0 = (idx % wpr) * per;
  const uint32_t word = qw[idx];
  const uint32_t mask = (1u << bits) - 1u;
  for (int i = 0; i < per && k0 + i < K; ++i) {
    const int64_t grp = (k0 + i) / g;
    const int code = (int)((word >> (bits * i)) & mask) + iqmin;
    const int z = (int)((qz[row * zwpr + grp / per] >> (bits * (int)(grp % per))) & mask) + iqmin;
    out[row * K + k0 + i] = dequant_one(code, z, __half2float(scales[row * G + grp]));
  }
}

}  // namespace awqk

using namespace awqk;

extern "C" int awqk_bf16_to_fp16(const void* in_bf16, void* out_fp16, int64_t n, void* stream) {
  if (n < 0) return AWQK_E_BADARG;
  if (n == 0) return AWQK_OK;
  if (in_bf16 == nullptr || out_fp16 == nullptr) return AWQK_E_BADARG;
  DeviceGuard guard(in_bf16);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out_fp16)) & 15u) == 0;
  if (aligned) {
    const int64_t per_cta = (int64_t)kCvThreads * kCvUnroll * 8;
    const int64_t ctas = ceil_div(n, per_cta);
    if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
    bf16_to_fp16_kernel<<<(unsigned)ctas, kCvThreads, 0, st>>>(
        reinterpret_cast<const uint16_t*>(in_bf16), reinterpret_cast<uint16_t*>(out_fp16), n);
  } else {
    const int64_t ctas = ceil_div(n, 256);
    if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
    bf16_to_fp16_scalar<<<(unsigned)ctas, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(in_bf16),
                                                        reinterpret_cast<uint16_t*>(out_fp16), n);
  }
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

extern "C" int awqk_dequant(const int32_t* q_unpacked, const void* scales_f16, const int32_t* zp,
                            int64_t C, int64_t K, int group_size, float* out, void* stream) {
  if (!q_unpacked || !scales_f16 || !zp || !out || C <= 0 || K <= 0 || group_size <= 0) return AWQK_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(q_unpacked) | reinterpret_cast<uintptr_t>(out)) & 15u) return AWQK_E_ALIGN;
  DeviceGuard guard(q_unpacked);
  if (guard.status != AWQK_OK) return guard.status;
  const int64_t G = ceil_div(K, group_size);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((K % 4) == 0 && (group_size % 4) == 0 && C <= 0x7FFFFFFFLL && ceil_div(K, 1024) <= 65535 && K < (1LL << 32)) {
    dim3 grid((unsigned)C, (unsigned)ceil_div(K, 1024));
    dequant_rows_kernel<<<grid, 256, 0, st>>>(q_unpacked, reinterpret_cast<const __half*>(scales_f16), zp, K, group_size,
                                             G, out);
  } else {
    const int64_t ctas = ceil_div(C * K, 256);
    if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
    dequant_kernel<<<(unsigned)ctas, 256, 0, st>>>(q_unpacked, reinterpret_cast<const __half*>(scales_f16), zp, C, K,
                                                  group_size, G, out);
  }
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

extern "C" int awqk_dequant_packed(const uint32_t* q_packed, const void* scales_f16,
                                   const uint32_t* zp_packed, int64_t C, int64_t K, int group_size,
                                   int bits, int symmetric, float* out, void* stream) {
  if (!q_packed || !scales_f16 || !zp_packed || !out || C <= 0 || K <= 0 || group_size <= 0) return AWQK_E_BADARG;
  if (bits != 4 && bits != 8) return AWQK_E_BADARG;
  DeviceGuard guard(q_packed);
  if (guard.status != AWQK_OK) return guard.status;
  const int64_t G = ceil_div(K, group_size);
  const int64_t threads = C * ceil_div(K, 32 / bits);
  const int64_t ctas = ceil_div(threads, 256);
  if (ctas > 0x7FFFFFFFLL) return AWQK_E_BADARG;
  const int iqmin = symmetric ? -(1 << (bits - 1)) : 0;
  dequant_packed_kernel<<<(unsigned)ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q_packed, reinterpret_cast<const __half*>(scales_f16), zp_packed, C, K, group_size, G, bits, iqmin, out);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}
