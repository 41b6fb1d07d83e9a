// Host-buffer pipeline: the end-to-end path behind AWQQuantizer.quantize_model(..., pack=True) and
// bench.py's `e2e` number.  A host-resident (ideally pinned) weight is cut into tile-aligned chunks;
// chunk i+1 is copied H2D while K1 runs on chunk i and the packed outputs of chunk i-1 drain D2H.
// Three private non-blocking streams, kBuf device staging slots, CUDA events between them -- no host
// synchronisation until awqk_pipe_sync().  Replaces the reference's per-tensor
// tensor.to(device) -> quantize -> .cpu() sequence (main.py:300, 374-380; awq.py:402, 410-412).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <sys/mman.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <unistd.h>

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
  // gather mode (awqk_pipe_quant_gather): pinned bounce rings, allocated on first use
  void* h_in[kBuf] = {};
  uint8_t* h_out[kBuf] = {};
  size_t h_out_bytes = 0;      // per slot
  cudaEvent_t ev_h2d[kBuf] = {};
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
    if (p->ev_h2d[b]) (void)cudaEventDestroy(p->ev_h2d[b]);
    if (p->h_in[b]) (void)awqk_host_free_pinned(p->h_in[b]);
    if (p->h_out[b]) (void)awqk_host_free_pinned(p->h_out[b]);
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

// ---------------------------------------------------------------------------------------------------
// Gather mode: the model's tensors stay where the loader put them (ordinary pageable memory).  The
// tile-aligned arena of quantization/arena.py exists only virtually: tensor i owns the elements
// [v_i, v_i + numel_i) with v_{i+1} = v_i + roundup(numel_i, 8192).  The calling thread materialises it
// chunk by chunk in a ring of pinned bounce buffers (parallel memcpy), queues H2D -> K1 -> D2H exactly like
// the flat mode above, and a drain thread copies every finished chunk from the pinned output ring to the
// caller's (pageable) result arrays.  Bounded pinned memory (no cudaHostAlloc proportional to the model:
// pinning runs at ~2.3 GB/s on this box, 20x slower than the pipeline), no second pass over the inputs.
// Blocking: returns when every result byte is in place.
// ---------------------------------------------------------------------------------------------------
namespace {

int pipe_threads() {
  static const int n = []() {
    const char* e = getenv("AWQK_PIPE_THREADS");
    int v = e ? atoi(e) : 0;
    if (v <= 0) {
      const unsigned hc = std::thread::hardware_concurrency();
      // measured on a 16-core host, Llama-3-8B end to end (tools/probe_cold.py): 4 -> 0.71 s, 8 -> 0.60 s,
      // 12 -> 0.57 s, 16 -> 0.57 s (the copies compete with the DMA engines for host DRAM, not for cores)
      v = hc >= 16 ? 12 : (hc >= 8 ? 4 : 2);
    }
    return std::min(v, 16);
  }();
  return n;
}

// Streaming copy for the big staging / drain copies: non-temporal 16-byte stores (no read-for-ownership of the
// destination lines, nothing of a 256 MB slot left in the caches the DMA engines and the other copy threads share).
// glibc's memcpy switches to such stores only above a size threshold that a thread's share of a slot (~20 MB) may or
// may not reach; here it is unconditional.  Alone it is slower than memcpy (12 threads: 77 vs 87 GB/s), inside the
// pipeline -- next to the upload DMA, the result DMA and the drain copies -- it is faster: staging copies of a
// Llama-3-8B call 0.50 -> 0.35 s, drain copies 0.35 -> 0.26 s, the call 0.57 -> 0.53 s (same-box A/B; a 32-byte AVX2
// form with software prefetch measured worse: 0.42 s).  Source and destination may have any alignment.
static void stream_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__SSE2__)
  if (n < 4096) {
    memcpy(dst, src, n);
    return;
  }
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head) {
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
  }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
#else
  memcpy(dst, src, n);
#endif
}

// memcpy split over a few threads (a single core moves ~10 GB/s; PCIe wants 50).  One process-wide pool of
// workers, created on first use and never torn down (no destructor-order problems at exit); any number of
// callers (the staging thread and the drain thread of every pipe, awqk_host_copy) may submit concurrently.
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this]() { run(); });
    for (auto& t : threads_) t.detach();
  }
  void copy(void* dst, const void* src, size_t bytes, int ways) {
    const int max_ways = (int)threads_.size() + 1;
    if (ways > max_ways) ways = max_ways;
    if (bytes < ((size_t)1 << 20) || ways <= 1) {
      memcpy(dst, src, bytes);
      return;
    }
    const size_t per = ((bytes + ways - 1) / ways + 4095) & ~(size_t)4095;
    std::atomic<int> pending{0};
    int queued = 0;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (int t = 1; t < ways; ++t) {
        const size_t off = (size_t)t * per;
        if (off >= bytes) break;
        q_.push_back(Task{static_cast<uint8_t*>(dst) + off, static_cast<const uint8_t*>(src) + off,
                          std::min(per, bytes - off), &pending});
        ++queued;
      }
      pending.store(queued, std::memory_order_relaxed);
    }
    if (queued == 1) cv_.notify_one(); else cv_.notify_all();
    stream_copy(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), std::min(per, bytes));
    while (pending.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }

 private:
  struct Task {
    uint8_t* dst;
    const uint8_t* src;
    size_t n;
    std::atomic<int>* pending;
  };
  void run() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this]() { return !q_.empty(); });
        t = q_.front();
        q_.pop_front();
      }
      stream_copy(t.dst, t.src, t.n);
      t.pending->fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Task> q_;
};

void parallel_copy(void* dst, const void* src, size_t bytes, int threads) {
  // workers sleep on a condition variable until used.  Threads do not survive fork(): a child process
  // (e.g. a data-loader worker) gets a pool of its own on first use.
  static std::mutex guard;
  static CopyPool* pool = nullptr;
  static pid_t owner = 0;
  CopyPool* p;
  {
    std::lock_guard<std::mutex> lk(guard);
    if (pool == nullptr || owner != getpid()) {
      pool = new CopyPool(15);
      owner = getpid();
    }
    p = pool;
  }
  p->copy(dst, src, bytes, threads);
}

struct OutLayout {           // byte offsets inside one pinned output slot
  size_t qp = 0, sc = 0, zp = 0, zq = 0, qu = 0, total = 0;
};

// Fresh pageable result arrays are first touched by the drain copies: ask for transparent huge pages on the
// 2 MiB-aligned interior so that the drain thread takes 512x fewer page faults (no-op where THP is off).
void advise_huge(void* ptr, size_t bytes) {
#ifdef MADV_HUGEPAGE
  constexpr uintptr_t kHuge = (uintptr_t)2 << 20;
  if (ptr == nullptr || bytes < 4 * kHuge) return;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(ptr) + kHuge - 1) & ~(kHuge - 1);
  const uintptr_t e = (reinterpret_cast<uintptr_t>(ptr) + bytes) & ~(kHuge - 1);
  if (e > a) (void)madvise(reinterpret_cast<void*>(a), e - a, MADV_HUGEPAGE);
#else
  (void)ptr; (void)bytes;
#endif
}

// page-locked (cudaHostAlloc / cudaHostRegister) host memory?  nullptr counts as "yes" (nothing to copy)
bool is_pinned(const void* ptr) {
  if (ptr == nullptr) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" int awqk_host_copy(void* dst, const void* src, size_t bytes, int threads) {
  if ((dst == nullptr || src == nullptr) && bytes != 0) return AWQK_E_BADARG;
  parallel_copy(dst, src, bytes, threads > 0 ? std::min(threads, 16) : pipe_threads());
  return AWQK_OK;
}

extern "C" int awqk_host_prefault(void* ptr, size_t bytes, int threads) {
  if (ptr == nullptr && bytes != 0) return AWQK_E_BADARG;
  if (bytes == 0) return AWQK_OK;
  advise_huge(ptr, bytes);
  const int n = std::max(1, threads > 0 ? std::min(threads, 16) : pipe_threads());
  const size_t per = ((bytes + n - 1) / n + 4095) & ~(size_t)4095;
  auto touch = [](uint8_t* b, size_t len) {     // write-fault every page without changing its content (atomic | 0)
    for (size_t o = 0; o < len; o += 4096) __atomic_fetch_or(b + o, (uint8_t)0, __ATOMIC_RELAXED);
    if (len) __atomic_fetch_or(b + len - 1, (uint8_t)0, __ATOMIC_RELAXED);
  };
  std::vector<std::thread> ts;
  uint8_t* base = static_cast<uint8_t*>(ptr);
  for (int t = 1; t < n; ++t) {
    const size_t off = (size_t)t * per;
    if (off >= bytes) break;
    ts.emplace_back(touch, base + off, std::min(per, bytes - off));
  }
  touch(base, std::min(per, bytes));
  for (auto& t : ts) t.join();
  return AWQK_OK;
}

// Page-locked staging memory at memcpy speed: cudaHostAlloc takes ~0.1 s per 256 MB on these hosts (2.4 GB/s: it faults
// and pins 4 KiB pages on one thread) -- three upload slots, three result slots and the gather rings were 0.4 s of a
// 1.1 s first call.  Here: a 2 MiB-aligned anonymous mapping, transparent huge pages requested, every page touched by
// the copy threads, then page-locked IN PLACE with cudaHostRegister (~15 ms per 256 MB; tools/probe_register.py).
// The block starts with one header page (kind, mapping size) so that awqk_host_free_pinned needs no global table.
namespace {
constexpr size_t kPinAlign = (size_t)2 << 20;
struct PinHeader {
  uint64_t magic;
  uint64_t map_bytes;
  uint64_t registered;     // 1: mmap + cudaHostRegister, 0: cudaHostAlloc fallback (then map_bytes = 0)
};
constexpr uint64_t kPinMagic = 0x6177716b70696e31ull;   // "awqkpin1"
}  // namespace

extern "C" int awqk_host_alloc_pinned(size_t bytes, void** out) {
  if (out == nullptr || bytes == 0) return AWQK_E_BADARG;
  *out = nullptr;
  const size_t body = (bytes + kPinAlign - 1) & ~(kPinAlign - 1);
  const size_t total = body + 2 * kPinAlign;                  // slack to align the body, header in front of it
  void* raw = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (raw != MAP_FAILED) {
    // body = the first 2 MiB boundary that leaves room for the header page in front of it
    uint8_t* b = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 4096 + kPinAlign - 1) & ~(uintptr_t)(kPinAlign - 1));
    // give back what lies in front of the header page and behind the body
    uint8_t* head = b - 4096;
    if (head > static_cast<uint8_t*>(raw)) (void)munmap(raw, (size_t)(head - static_cast<uint8_t*>(raw)));
    uint8_t* end = b + body;
    uint8_t* raw_end = static_cast<uint8_t*>(raw) + total;
    if (raw_end > end) (void)munmap(end, (size_t)(raw_end - end));
    (void)awqk_host_prefault(b, body, 0);
    if (cudaHostRegister(b, body, cudaHostRegisterPortable) == cudaSuccess) {
      PinHeader* h = reinterpret_cast<PinHeader*>(head);
      h->magic = kPinMagic; h->map_bytes = (uint64_t)(body + 4096); h->registered = 1;
      *out = b;
      return AWQK_OK;
    }
    (void)cudaGetLastError();
    (void)munmap(head, body + 4096);
  }
  // fallback: the runtime's own page-locked allocation (same contract, slower to obtain)
  uint8_t* p = nullptr;
  AWQK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), bytes + 4096, cudaHostAllocPortable));
  PinHeader* h = reinterpret_cast<PinHeader*>(p);
  h->magic = kPinMagic; h->map_bytes = 0; h->registered = 0;
  *out = p + 4096;
  return AWQK_OK;
}

extern "C" int awqk_host_free_pinned(void* ptr) {
  if (ptr == nullptr) return AWQK_OK;
  uint8_t* b = static_cast<uint8_t*>(ptr);
  PinHeader* h = reinterpret_cast<PinHeader*>(b - 4096);
  if (h->magic != kPinMagic) return AWQK_E_BADARG;
  if (h->registered) {
    const size_t map_bytes = (size_t)h->map_bytes;
    (void)cudaHostUnregister(b);
    h->magic = 0;
    (void)munmap(b - 4096, map_bytes);
  } else {
    h->magic = 0;
    (void)cudaFreeHost(b - 4096);
  }
  return AWQK_OK;
}

extern "C" int awqk_pipe_quant_gather(awqk_pipe* p, int n_tensors, const void* const* src, const int64_t* numel,
                                      int64_t row_len, int dtype, int group_size, int bits, int symmetric, int arith,
                                      int32_t* q_unpacked_host, uint32_t* q_packed_host, void* scales_f16_host,
                                      int32_t* zp_host, uint32_t* zp_packed_host) {
  if (p == nullptr || src == nullptr || numel == nullptr || scales_f16_host == nullptr || n_tensors <= 0)
    return AWQK_E_BADARG;
  if (bits != 4 && bits != 8) return AWQK_E_BADARG;
  if (dtype != AWQK_BF16 && dtype != AWQK_FP16 && dtype != AWQK_FP32) return AWQK_E_UNSUPPORTED;
  if (!(group_size == 32 || group_size == 64 || group_size == 128)) return AWQK_E_UNSUPPORTED;
  const int per = 32 / bits;
  const size_t esz = (dtype == AWQK_FP32) ? 4 : 2;
  constexpr int64_t kTile = 8192;
  std::vector<int64_t> voff((size_t)n_tensors + 1, 0);
  if (row_len != 0) {
    // short rows (fewer groups than a packed zero word holds): K1 pads one zero word per row itself
    const int64_t gr = row_len / group_size;
    if (row_len < 0 || row_len % group_size != 0 || gr >= per || per % gr != 0 || kTile % row_len != 0) return AWQK_E_BADARG;
  }
  for (int i = 0; i < n_tensors; ++i) {
    if (src[i] == nullptr || numel[i] <= 0 || numel[i] % group_size != 0) return AWQK_E_BADARG;
    if (row_len != 0 && numel[i] % row_len != 0) return AWQK_E_BADARG;
    if (row_len == 0 && zp_packed_host != nullptr && numel[i] % ((int64_t)group_size * per) != 0) return AWQK_E_BADARG;
    voff[i + 1] = voff[i] + (numel[i] + kTile - 1) / kTile * kTile;
  }
  const int64_t n = voff[n_tensors];
  SetDevice sd(p->device);
  if (!sd.ok) return AWQK_E_NODEVICE;

  int64_t chunk_elems = (int64_t)(p->chunk_bytes / esz) / kTile * kTile;
  if ((size_t)chunk_elems > p->elems_max) chunk_elems = (int64_t)p->elems_max / kTile * kTile;
  if (chunk_elems < kTile) return AWQK_E_WORKSPACE;
  const size_t E = (size_t)chunk_elems;
  OutLayout lay;
  lay.qp = 0;
  lay.sc = lay.qp + E;                      // packed codes: E * bits / 8 <= E bytes
  lay.zp = lay.sc + E / 32 * 2;
  lay.zq = lay.zp + E / 32 * 4;
  lay.qu = lay.zq + E / 32 * 4;
  lay.total = lay.qu + (q_unpacked_host ? E * 4 : 0);
  constexpr int kBuf = awqk_pipe::kBuf;
  // results that already live in page-locked memory are written by the D2H copies themselves (no bounce, no
  // drain thread): the caller keeps a cache of pinned result arenas when it converts model after model
  const bool direct = is_pinned(q_unpacked_host) && is_pinned(q_packed_host) && is_pinned(scales_f16_host) &&
                      is_pinned(zp_host) && is_pinned(zp_packed_host);
  for (int b = 0; b < kBuf; ++b) {
    if (p->h_in[b] == nullptr) {
      const int rc = awqk_host_alloc_pinned(p->chunk_bytes, &p->h_in[b]);
      if (rc != AWQK_OK) return rc;
    }
    if (p->ev_h2d[b] == nullptr) AWQK_CUDA(cudaEventCreateWithFlags(&p->ev_h2d[b], cudaEventDisableTiming));
    if (q_unpacked_host != nullptr && p->d_qu[b] == nullptr)
      AWQK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_qu[b]), p->elems_max * 4));
  }
  if (!direct) {
    advise_huge(q_packed_host, (size_t)(n / per) * 4);
    advise_huge(q_unpacked_host, (size_t)n * 4);
  }
  if (!direct && p->h_out_bytes < lay.total) {
    for (int b = 0; b < kBuf; ++b) {
      if (p->h_out[b]) { AWQK_CUDA(cudaStreamSynchronize(p->s_out)); (void)awqk_host_free_pinned(p->h_out[b]); p->h_out[b] = nullptr; }
      const int rc = awqk_host_alloc_pinned(lay.total, reinterpret_cast<void**>(&p->h_out[b]));
      if (rc != AWQK_OK) return rc;
    }
    p->h_out_bytes = lay.total;
  }
  // everything queued earlier on this pipe must have left the device slots before they are rewired
  AWQK_CUDA(cudaStreamSynchronize(p->s_out));

  const int threads = pipe_threads();
  const int64_t n_chunks = (n + chunk_elems - 1) / chunk_elems;
  // packed zero words: flat = one per `per` groups; short rows = one per row
  auto zq_off = [=](int64_t e) { return row_len ? e / row_len : e / group_size / per; };
  auto zq_words = [=](int64_t ne_) { return row_len ? ne_ / row_len : (ne_ / group_size + per - 1) / per; };
  // ---- drain thread: chunk c is copied out once its D2H event fired; then its output slot is free ----
  std::mutex mu;
  std::condition_variable cv;
  int64_t queued = 0, drained = 0;          // chunks whose GPU work is enqueued / whose results are in place
  int status = AWQK_OK;
  bool stop = false;
  const int device = p->device;
  std::thread drainer([&]() {
    (void)cudaSetDevice(device);
    for (int64_t c = 0; c < n_chunks && !direct; ++c) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return queued > c || stop; });
        if (queued <= c) return;            // stopped before this chunk was queued
      }
      const int b = (int)(c % kBuf);
      if (cudaEventSynchronize(p->ev_out[b]) != cudaSuccess) {
        std::lock_guard<std::mutex> lk(mu);
        status = AWQK_E_CUDA;
        drained = c + 1;
        cv.notify_all();
        continue;
      }
      const int64_t e0 = c * chunk_elems, ne = std::min<int64_t>(chunk_elems, n - e0), ng = ne / group_size;
      const uint8_t* o = p->h_out[b];
      if (q_packed_host) parallel_copy(q_packed_host + e0 / per, o + lay.qp, (size_t)(ne / per) * 4, threads);
      if (q_unpacked_host) parallel_copy(q_unpacked_host + e0, o + lay.qu, (size_t)ne * 4, threads);
      memcpy(static_cast<uint16_t*>(scales_f16_host) + e0 / group_size, o + lay.sc, (size_t)ng * 2);
      if (zp_host) memcpy(zp_host + e0 / group_size, o + lay.zp, (size_t)ng * 4);
      if (zp_packed_host) memcpy(zp_packed_host + zq_off(e0), o + lay.zq, (size_t)zq_words(ne) * 4);
      {
        std::lock_guard<std::mutex> lk(mu);
        drained = c + 1;
      }
      cv.notify_all();
    }
  });
  auto finish = [&](int rc) {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    drainer.join();
    return rc != AWQK_OK ? rc : status;
  };
#define AWQK_CUDA_G(expr)                                                \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) {                                             \
      ::awqk::set_cuda_error(_e, #expr, __FILE__, __LINE__);             \
      return finish(AWQK_E_CUDA);                                        \
    }                                                                    \
  } while (0)

  int ti = 0;                               // first tensor that may overlap the current chunk
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int b = (int)(c % kBuf);
    const int64_t e0 = c * chunk_elems, ne = std::min<int64_t>(chunk_elems, n - e0), e1 = e0 + ne;
    const int64_t ng = ne / group_size;
    if (c >= kBuf) {
      AWQK_CUDA_G(cudaEventSynchronize(p->ev_h2d[b]));          // pinned input slot: its H2D has finished
      if (!direct) {
        std::unique_lock<std::mutex> lk(mu);                     // pinned output slot: chunk c - kBuf is drained
        cv.wait(lk, [&]() { return drained >= c - kBuf + 1; });
      }
      AWQK_CUDA_G(cudaStreamWaitEvent(p->s_in, p->ev_k[b], 0)); // device input slot: K1 of chunk c - kBuf has read it
    }
    // ---- stage: gather the pieces of the virtual arena that fall into [e0, e1) ----
    uint8_t* hin = static_cast<uint8_t*>(p->h_in[b]);
    while (ti < n_tensors && voff[ti + 1] <= e0) ++ti;
    for (int i = ti; i < n_tensors && voff[i] < e1; ++i) {
      const int64_t t0 = std::max<int64_t>(voff[i], e0), t1 = std::min<int64_t>(voff[i] + numel[i], e1);
      if (t1 > t0)
        parallel_copy(hin + (size_t)(t0 - e0) * esz, static_cast<const uint8_t*>(src[i]) + (size_t)(t0 - voff[i]) * esz,
                      (size_t)(t1 - t0) * esz, threads);
      const int64_t z0 = std::max<int64_t>(voff[i] + numel[i], e0), z1 = std::min<int64_t>(voff[i + 1], e1);
      if (z1 > z0) memset(hin + (size_t)(z0 - e0) * esz, 0, (size_t)(z1 - z0) * esz);     // tile padding
    }
    // ---- queue H2D -> K1 -> D2H (into the pinned output slot) ----
    AWQK_CUDA_G(cudaMemcpyAsync(p->d_in[b], hin, (size_t)ne * esz, cudaMemcpyHostToDevice, p->s_in));
    AWQK_CUDA_G(cudaEventRecord(p->ev_h2d[b], p->s_in));
    AWQK_CUDA_G(cudaStreamWaitEvent(p->s_k, p->ev_h2d[b], 0));
    if (c >= kBuf) AWQK_CUDA_G(cudaStreamWaitEvent(p->s_k, p->ev_out[b], 0));   // device output slot drained to the host
    const int rc = awqk_group_quant(p->d_in[b], dtype, row_len ? ne / row_len : 1, row_len ? row_len : ne, group_size, bits,
                                    symmetric, arith, q_unpacked_host ? p->d_qu[b] : nullptr,
                                    q_packed_host ? p->d_qp[b] : nullptr, p->d_sc[b], zp_host ? p->d_zp[b] : nullptr,
                                    zp_packed_host ? p->d_zpp[b] : nullptr, nullptr, p->s_k);
    if (rc != AWQK_OK) return finish(rc);
    AWQK_CUDA_G(cudaEventRecord(p->ev_k[b], p->s_k));
    AWQK_CUDA_G(cudaStreamWaitEvent(p->s_out, p->ev_k[b], 0));
    uint8_t* o = p->h_out[b];
    void* dst_qp = direct ? static_cast<void*>(q_packed_host + e0 / per) : o + lay.qp;
    void* dst_qu = direct ? static_cast<void*>(q_unpacked_host + e0) : o + lay.qu;
    void* dst_sc = direct ? static_cast<void*>(static_cast<uint16_t*>(scales_f16_host) + e0 / group_size) : o + lay.sc;
    void* dst_zp = direct ? static_cast<void*>(zp_host + e0 / group_size) : o + lay.zp;
    void* dst_zq = direct ? static_cast<void*>(zp_packed_host + zq_off(e0)) : o + lay.zq;
    if (q_packed_host)
      AWQK_CUDA_G(cudaMemcpyAsync(dst_qp, p->d_qp[b], (size_t)(ne / per) * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (q_unpacked_host)
      AWQK_CUDA_G(cudaMemcpyAsync(dst_qu, p->d_qu[b], (size_t)ne * 4, cudaMemcpyDeviceToHost, p->s_out));
    AWQK_CUDA_G(cudaMemcpyAsync(dst_sc, p->d_sc[b], (size_t)ng * 2, cudaMemcpyDeviceToHost, p->s_out));
    if (zp_host) AWQK_CUDA_G(cudaMemcpyAsync(dst_zp, p->d_zp[b], (size_t)ng * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (zp_packed_host)
      AWQK_CUDA_G(cudaMemcpyAsync(dst_zq, p->d_zpp[b], (size_t)zq_words(ne) * 4, cudaMemcpyDeviceToHost, p->s_out));
    AWQK_CUDA_G(cudaEventRecord(p->ev_out[b], p->s_out));
    p->used[b] = true;
    {
      std::lock_guard<std::mutex> lk(mu);
      queued = c + 1;
    }
    cv.notify_all();
  }
  if (direct) AWQK_CUDA_G(cudaStreamSynchronize(p->s_out));
#undef AWQK_CUDA_G
  if (!direct) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&]() { return drained >= n_chunks; });
  }
  drainer.join();
  return status;
}
