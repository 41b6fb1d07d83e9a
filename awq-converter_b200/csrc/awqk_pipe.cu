// Host-buffer pipeline: the end-to-end path behind AWQQuantizer.quantize_model(..., pack=True) and
// bench.py's `e2e` number.  A host-resident (ideally pinned) weight is cut into tile-aligned chunks;
// chunk i+1 is copied H2D while K1 runs on chunk i and the packed outputs of chunk i-1 drain D2H.
// Three private non-blocking streams, kBuf device staging slots, CUDA events between them -- no host
// synchronisation until awqk_pipe_sync().  Replaces the reference's per-tensor
// tensor.to(device) -> quantize -> .cpu() sequence (main.py:300, 374-380; awq.py:402, 410-412).
#include <algorithm>
#include <new>

#include "awqk_common.cuh"

struct awqk_pipe {
  static constexpr int kBuf = 3;
  int device = 0;
  size_t chunk_bytes = 0;      // input bytes per chunk (multiple of 64 KiB)
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  void* d_in[kBuf] = {};
  uint32_t* d_qp[kBuf] = {};   // packed codes      (chunk_elems_max * bits / 8 bytes, sized for 8 bit / 2-byte input)
  void* d_sc[kBuf] = {};       // fp16 scales       (chunk_elems_max / 32 * 2)
  int32_t* d_zp[kBuf] = {};    // int32 zero points (chunk_elems_max / 32 * 4)
  uint32_t* d_zpp[kBuf] = {};  // packed zero points
  int32_t* d_qu[kBuf] = {};    // unpacked codes, allocated on first use (chunk_elems_max * 4)
  cudaEvent_t ev_in[kBuf] = {}, ev_k[kBuf] = {}, ev_out[kBuf] = {};
  bool used[kBuf] = {};
  size_t elems_max = 0;        // chunk_bytes / 2
};

namespace {
struct SetDevice {
  int prev = -1;
  bool ok = true;
  explicit SetDevice(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~SetDevice() { if (prev >= 0 && prev != target) (void)cudaSetDevice(prev); }
  int target = -1;
};
}  // namespace

using namespace awqk;

extern "C" void awqk_pipe_destroy(awqk_pipe* p) {
  if (p == nullptr) return;
  SetDevice sd(p->device);
  if (p->s_in) (void)cudaStreamSynchronize(p->s_in);
  if (p->s_k) (void)cudaStreamSynchronize(p->s_k);
  if (p->s_out) (void)cudaStreamSynchronize(p->s_out);
  for (int b = 0; b < awqk_pipe::kBuf; ++b) {
    (void)cudaFree(p->d_in[b]); (void)cudaFree(p->d_qp[b]); (void)cudaFree(p->d_sc[b]);
    (void)cudaFree(p->d_zp[b]); (void)cudaFree(p->d_zpp[b]); (void)cudaFree(p->d_qu[b]);
    if (p->ev_in[b]) (void)cudaEventDestroy(p->ev_in[b]);
    if (p->ev_k[b]) (void)cudaEventDestroy(p->ev_k[b]);
    if (p->ev_out[b]) (void)cudaEventDestroy(p->ev_out[b]);
  }
  if (p->s_in) (void)cudaStreamDestroy(p->s_in);
  if (p->s_k) (void)cudaStreamDestroy(p->s_k);
  if (p->s_out) (void)cudaStreamDestroy(p->s_out);
  delete p;
}

static int pipe_create_impl(awqk_pipe* p) {
  AWQK_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  AWQK_CUDA(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
  AWQK_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  const size_t e = p->elems_max;
  for (int b = 0; b < awqk_pipe::kBuf; ++b) {
    AWQK_CUDA(cudaMalloc(&p->d_in[b], p->chunk_bytes));
    AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_qp[b]), e));            // 8 bit worst case
    AWQK_CUDA(cudaMalloc(&p->d_sc[b], e / 32 * 2));
    AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_zp[b]), e / 32 * 4));
    AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_zpp[b]), e / 8 + 16));   // row mode: <= one word per group
    AWQK_CUDA(cudaEventCreateWithFlags(&p->ev_in[b], cudaEventDisableTiming));
    AWQK_CUDA(cudaEventCreateWithFlags(&p->ev_k[b], cudaEventDisableTiming));
    AWQK_CUDA(cudaEventCreateWithFlags(&p->ev_out[b], cudaEventDisableTiming));
  }
  return AWQK_OK;
}

extern "C" int awqk_pipe_create(int device, size_t chunk_bytes, awqk_pipe** out) {
  if (out == nullptr || device < 0) return AWQK_E_BADARG;
  *out = nullptr;
  if (chunk_bytes == 0) chunk_bytes = (size_t)32 << 20;
  chunk_bytes = std::max<size_t>((chunk_bytes + 65535) & ~(size_t)65535, 65536);
  SetDevice sd(device);
  if (!sd.ok) return AWQK_E_NODEVICE;
  awqk_pipe* p = new (std::nothrow) awqk_pipe();
  if (p == nullptr) return AWQK_E_WORKSPACE;
  p->device = device;
  p->chunk_bytes = chunk_bytes;
  p->elems_max = chunk_bytes / 2;
  const int rc = pipe_create_impl(p);
  if (rc != AWQK_OK) {
    awqk_pipe_destroy(p);
    return rc;
  }
  *out = p;
  return AWQK_OK;
}

extern "C" int awqk_pipe_sync(awqk_pipe* p) {
  if (p == nullptr) return AWQK_E_BADARG;
  SetDevice sd(p->device);
  AWQK_CUDA(cudaStreamSynchronize(p->s_in));
  AWQK_CUDA(cudaStreamSynchronize(p->s_k));
  AWQK_CUDA(cudaStreamSynchronize(p->s_out));
  return AWQK_OK;
}

extern "C" int awqk_pipe_quant_host(awqk_pipe* p, const void* w_host, int dtype, int64_t C, int64_t K,
                                    int group_size, int bits, int symmetric, int arith,
                                    int32_t* q_unpacked_host, uint32_t* q_packed_host,
                                    void* scales_f16_host, int32_t* zp_host, uint32_t* zp_packed_host) {
  if (p == nullptr || w_host == nullptr || scales_f16_host == nullptr) return AWQK_E_BADARG;
  if (C <= 0 || K <= 0 || (bits != 4 && bits != 8)) return AWQK_E_BADARG;
  if (dtype != AWQK_BF16 && dtype != AWQK_FP16 && dtype != AWQK_FP32) return AWQK_E_UNSUPPORTED;
  if (!(group_size == 32 || group_size == 64 || group_size == 128) || (K % group_size) != 0)
    return AWQK_E_UNSUPPORTED;               // the pipeline handles the flat layout only
  const int per = 32 / bits;
  const int64_t G = K / group_size;
  SetDevice sd(p->device);
  if (!sd.ok) return AWQK_E_NODEVICE;
  const size_t esz = (dtype == AWQK_FP32) ? 4 : 2;
  const int64_t n = C * K;

  if (zp_packed_host != nullptr && (G % per) != 0) {
    // ---- row mode: packed zero points are padded per row, so chunks are whole rows and K1 is called
    // with the row structure (it packs the zero points row-wise itself).
    const int64_t zw = ceil_div(G, per);                       // packed zero words per row
    int64_t rows = std::min<int64_t>((int64_t)(p->chunk_bytes / esz) / K, (int64_t)p->elems_max / K);
    if (rows < 1) return AWQK_E_WORKSPACE;                     // one row does not fit a chunk
    const uint8_t* src = static_cast<const uint8_t*>(w_host);
    int i = 0;
    for (int64_t r0 = 0; r0 < C; r0 += rows, ++i) {
      const int b = i % awqk_pipe::kBuf;
      const int64_t nr = std::min<int64_t>(rows, C - r0);
      const int64_t ne = nr * K, e0 = r0 * K;
      if (q_unpacked_host != nullptr && p->d_qu[b] == nullptr)
        AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_qu[b]), p->elems_max * 4));
      if (p->used[b]) AWQK_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[b], 0));
      AWQK_CUDA(cudaMemcpyAsync(p->d_in[b], src + (size_t)e0 * esz, (size_t)ne * esz, cudaMemcpyHostToDevice, p->s_in));
      AWQK_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
      AWQK_CUDA(cudaStreamWaitEvent(p->s_k, p->ev_in[b], 0));
      const int rc = awqk_group_quant(p->d_in[b], dtype, nr, K, group_size, bits, symmetric, arith,
                                      q_unpacked_host ? p->d_qu[b] : nullptr, q_packed_host ? p->d_qp[b] : nullptr,
                                      p->d_sc[b], p->d_zp[b], p->d_zpp[b], nullptr, p->s_k);
      if (rc != AWQK_OK) return rc;
      AWQK_CUDA(cudaEventRecord(p->ev_k[b], p->s_k));
      AWQK_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_k[b], 0));
      if (q_packed_host)
        AWQK_CUDA(cudaMemcpyAsync(q_packed_host + e0 / per, p->d_qp[b], (size_t)(ne / per) * 4, cudaMemcpyDeviceToHost, p->s_out));
      if (q_unpacked_host)
        AWQK_CUDA(cudaMemcpyAsync(q_unpacked_host + e0, p->d_qu[b], (size_t)ne * 4, cudaMemcpyDeviceToHost, p->s_out));
      AWQK_CUDA(cudaMemcpyAsync(static_cast<uint16_t*>(scales_f16_host) + r0 * G, p->d_sc[b], (size_t)(nr * G) * 2,
                                cudaMemcpyDeviceToHost, p->s_out));
      if (zp_host)
        AWQK_CUDA(cudaMemcpyAsync(zp_host + r0 * G, p->d_zp[b], (size_t)(nr * G) * 4, cudaMemcpyDeviceToHost, p->s_out));
      AWQK_CUDA(cudaMemcpyAsync(zp_packed_host + r0 * zw, p->d_zpp[b], (size_t)(nr * zw) * 4, cudaMemcpyDeviceToHost, p->s_out));
      AWQK_CUDA(cudaEventRecord(p->ev_out[b], p->s_out));
      p->used[b] = true;
    }
    return AWQK_OK;
  }

  // ---- flat mode
  // chunk = whole CTA tiles and whole packed-zero words: multiple of 8192 elements and of per*g
  int64_t chunk_elems = (int64_t)(p->chunk_bytes / esz);
  const int64_t quantum = 8192LL * ((per * group_size + 8191) / 8192);   // = 8192 for all supported (g, bits)
  chunk_elems = std::max<int64_t>(chunk_elems / quantum * quantum, quantum);
  if ((size_t)chunk_elems > p->elems_max) chunk_elems = (int64_t)p->elems_max / quantum * quantum;

  const uint8_t* src = static_cast<const uint8_t*>(w_host);
  int i = 0;
  for (int64_t e0 = 0; e0 < n; e0 += chunk_elems, ++i) {
    const int b = i % awqk_pipe::kBuf;
    const int64_t ne = std::min<int64_t>(chunk_elems, n - e0);
    const int64_t ng = ne / group_size;
    if (q_unpacked_host != nullptr && p->d_qu[b] == nullptr)
      AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_qu[b]), p->elems_max * 4));
    // slot b is reusable once its previous outputs have drained (implies its kernel finished)
    if (p->used[b]) AWQK_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_out[b], 0));
    AWQK_CUDA(cudaMemcpyAsync(p->d_in[b], src + (size_t)e0 * esz, (size_t)ne * esz, cudaMemcpyHostToDevice, p->s_in));
    AWQK_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
    AWQK_CUDA(cudaStreamWaitEvent(p->s_k, p->ev_in[b], 0));
    const int rc = awqk_group_quant(p->d_in[b], dtype, 1, ne, group_size, bits, symmetric, arith,
                                    q_unpacked_host ? p->d_qu[b] : nullptr,
                                    q_packed_host ? p->d_qp[b] : nullptr, p->d_sc[b],
                                    zp_host ? p->d_zp[b] : nullptr, zp_packed_host ? p->d_zpp[b] : nullptr,
                                    nullptr, p->s_k);
    if (rc != AWQK_OK) return rc;
    AWQK_CUDA(cudaEventRecord(p->ev_k[b], p->s_k));
    AWQK_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_k[b], 0));
    if (q_packed_host)
      AWQK_CUDA(cudaMemcpyAsync(q_packed_host + e0 / per, p->d_qp[b], (size_t)(ne / per) * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (q_unpacked_host)
      AWQK_CUDA(cudaMemcpyAsync(q_unpacked_host + e0, p->d_qu[b], (size_t)ne * 4, cudaMemcpyDeviceToHost, p->s_out));
    AWQK_CUDA(cudaMemcpyAsync(static_cast<uint16_t*>(scales_f16_host) + e0 / group_size, p->d_sc[b], (size_t)ng * 2,
                              cudaMemcpyDeviceToHost, p->s_out));
    if (zp_host)
      AWQK_CUDA(cudaMemcpyAsync(zp_host + e0 / group_size, p->d_zp[b], (size_t)ng * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (zp_packed_host)
      AWQK_CUDA(cudaMemcpyAsync(zp_packed_host + e0 / group_size / per, p->d_zpp[b], (size_t)ceil_div(ng, per) * 4,
                                cudaMemcpyDeviceToHost, p->s_out));
    AWQK_CUDA(cudaEventRecord(p->ev_out[b], p->s_out));
    p->used[b] = true;
  }
  return AWQK_OK;
}
