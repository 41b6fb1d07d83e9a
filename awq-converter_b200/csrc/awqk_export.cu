// Interop export: K1's row-major packed result -> the AutoAWQ / vLLM "GEMM" checkpoint
// layout.  Pure re-layout of integers (no arithmetic): bit-exact by construction, checked against
// oracle/awq_oracle.py::to_autoawq_gemm.
//
//   in : qweight [C, K/8]  word j of row c = sum_i u[c, 8j+i] << 4i        (K1, asymmetric int4)
//        zp      [C, G] int32,  scales [C, G] fp16
//   out: qweight [K, C/8]  word j of row k = sum_i u[8j + ORDER[i], k] << 4i,  ORDER = {0,2,4,6,1,3,5,7}
//        qzeros  [G, C/8]  same packing of the zero points along C
//        scales  [G, C] fp16
#include "awqk_common.cuh"

namespace awqk {

__device__ __constant__ int kAwqOrder[8] = {0, 2, 4, 6, 1, 3, 5, 7};

// CTA tile: 64 output channels x 256 input features
__global__ void __launch_bounds__(256)
export_qweight_kernel(const uint32_t* __restrict__ qw, int64_t C, int64_t K, uint32_t* __restrict__ out) {
  __shared__ uint32_t tile[64][33];
  const int64_t c0 = (int64_t)blockIdx.y * 64, k0 = (int64_t)blockIdx.x * 256;
  const int64_t wpr = K / 8;
  const int t = threadIdx.x;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = r * 8 + (t >> 5), col = t & 31;
    const int64_t c = c0 + row, wj = k0 / 8 + col;
    tile[row][col] = (c < C && wj < wpr) ? qw[c * wpr + wj] : 0u;
  }
  __syncthreads();
  const int64_t k = k0 + t;
  if (k >= K) return;
  const int sh = 4 * (t & 7), wcol = t >> 3;
  const int64_t opr = C / 8;
#pragma unroll
  for (int jc = 0; jc < 8; ++jc) {
    if (c0 + 8 * jc >= C) break;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc |= ((tile[8 * jc + kAwqOrder[i]][wcol] >> sh) & 15u) << (4 * i);
    out[k * opr + c0 / 8 + jc] = acc;
  }
}

__global__ void __launch_bounds__(256)
export_zeros_scales_kernel(const int32_t* __restrict__ zp, const __half* __restrict__ scales, int64_t C, int64_t G,
                           int iqmin, uint32_t* __restrict__ qzeros, __half* __restrict__ scales_t) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx < G * C) {                                   // scales [C,G] -> [G,C]
    const int64_t g = idx / C, c = idx % C;
    scales_t[idx] = scales[c * G + g];
  }
  const int64_t opr = C / 8;
  if (idx < G * opr) {
    const int64_t g = idx / opr, j = idx % opr;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int z = zp[(8 * j + kAwqOrder[i]) * G + g];
      acc |= ((z == INT32_MIN) ? 0u : ((uint32_t)(z - iqmin) & 15u)) << (4 * i);
    }
    qzeros[idx] = acc;
  }
}

}  // namespace awqk

using namespace awqk;

extern "C" int awqk_export_autoawq(const uint32_t* q_packed, const int32_t* zp, const void* scales_f16, int64_t C,
                                   int64_t K, int64_t G, int symmetric, uint32_t* qweight_out, uint32_t* qzeros_out,
                                   void* scales_out, void* stream) {
  if (!q_packed || !zp || !scales_f16 || !qweight_out || !qzeros_out || !scales_out) return AWQK_E_BADARG;
  if (C <= 0 || K <= 0 || G <= 0) return AWQK_E_BADARG;
  if ((C % 8) != 0 || (K % 8) != 0) return AWQK_E_UNSUPPORTED;
  DeviceGuard guard(q_packed);
  if (guard.status != AWQK_OK) return guard.status;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)ceil_div(K, 256), (unsigned)ceil_div(C, 64));
  if (grid.y > 65535) return AWQK_E_BADARG;
  export_qweight_kernel<<<grid, 256, 0, st>>>(q_packed, C, K, qweight_out);
  const int64_t n = G * C;
  export_zeros_scales_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
      zp, reinterpret_cast<const __half*>(scales_f16), C, G, symmetric ? -8 : 0, qzeros_out,
      reinterpret_cast<__half*>(scales_out));
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}
