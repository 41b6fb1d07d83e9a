// K2 GEMM, CTA-pair version: tcgen05.mma.cta_group::2 (UMMA M = 256 across two SMs, N = 256).
//
// Same contraction and fused sum-of-squares epilogue as sqerr_gemm_kernel (awqk_search.cu), but each
// thread-block cluster of 2 CTAs computes a 256 x 256 tile: CTA r holds rows [128r, 128r+128) of the A
// tile and rows [128r, 128r+128) of the B tile (the N dimension of B is split across the pair), so per
// k-block each SM stages 16 KiB of A + 16 KiB of B instead of 16 + 32: one third less L2->SMEM
// traffic per flop (the 1-CTA kernel pulls ~15 TB/s out of L2), and 6 ring stages instead of 4.
//
//   barriers (per CTA unless noted):
//     full[s]   leader only is waited on; expects the bytes of BOTH CTAs (peer's TMA signals the
//               leader's barrier: cp.async.bulk.tensor ... .cta_group::2 with the peer bit cleared)
//     empty[s]  one per CTA, released by the leader's tcgen05.commit.cta_group::2 multicast
//     tfull[a]  one per CTA (accumulator ready), multicast commit
//     tempty[a] leader only, count 8: the 4 epilogue warps of both CTAs arrive (remote arrive from CTA 1)
#include <cuda.h>

#include <algorithm>
#include <atomic>

#include "awqk_common.cuh"

namespace awqk {

constexpr int k2BM = 128, k2BN = 256, k2BNh = 128, k2BK = 64;   // per-CTA A rows, pair N, per-CTA B rows
constexpr int k2Stages = 6;
constexpr int k2ABytes = k2BM * k2BK * 2;    // 16 KiB
constexpr int k2BBytes = k2BNh * k2BK * 2;   // 16 KiB
constexpr int k2StageBytes = k2ABytes + k2BBytes;
constexpr int k2Threads = 192;
constexpr uint32_t k2TmemCols = 512;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (pair clusters)

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
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
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mb_arrive_cluster(uint32_t local_bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                             uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16: D = F32, A = B = BF16, K-major both, N = 256, M = 256 (pair)
constexpr uint32_t k2Idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(k2BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(k2Idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {   // arrive on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
sqerr_gemm2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dw,
                   int n_s, int mp_tiles, int n_tiles, int k_blocks, double* __restrict__ err) {
  extern __shared__ uint8_t g2_raw[];
  const uint32_t raw = s2u(g2_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;               // identical offset in both CTAs of the pair
  uint8_t* gsm = g2_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + k2Stages * k2StageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k2Stages + 4);
  const uint32_t full0 = s2u(bars), empty0 = full0 + 8 * k2Stages;
  const uint32_t tfull0 = empty0 + 8 * k2Stages, tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                        // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total_tiles = n_s * n_tiles * mp_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      mb_init(full0 + 8 * s, 1);
      mb_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mb_init(tfull0 + 8 * a, 1);
      mb_init(tempty0 + 8 * a, 8);                             // 4 epilogue warps x 2 CTAs (used on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the same logical warp in both CTAs allocates (and later frees) the pair's TMEM
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "r"(k2TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                          // peer barriers are initialised before any remote signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion lands on the LEADER's full barrier) ==========
    if (lane == 0) {
      uint32_t stage = 0, ph = 1;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int mp = tile % mp_tiles;
        const int rest = tile / mp_tiles;
        const int nt = rest % n_tiles;
        const int a = rest / n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mb_wait(empty0 + 8 * stage, ph);                     // own slot free (multicast commit from the leader)
          const uint32_t lbar = (full0 + 8 * stage) & kPeerMask;
          if (rank == 0) mb_expect_tx(full0 + 8 * stage, 2 * k2StageBytes);
          const uint32_t sa = base + stage * k2StageBytes;
          tma2_load_2d(sa, &map_x, kb * k2BK, mp * 256 + (int)rank * k2BM, lbar);
          tma2_load_3d(sa + k2ABytes, &map_dw, kb * k2BK, nt * k2BN + (int)rank * k2BNh, a, lbar);
          if (++stage == k2Stages) { stage = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the leader CTA =====================
    if (rank == 0) {
      uint32_t stage = 0, ph = 0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const uint32_t ab = (uint32_t)it & 1u;
        const uint32_t aph = ((uint32_t)it >> 1) & 1u;
        mb_wait(tempty0 + 8 * ab, aph ^ 1u);                   // both CTAs' epilogues drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + ab * k2BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mb_wait(full0 + 8 * stage, ph);                      // both CTAs' A and B halves have landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t sa = base + stage * k2StageBytes;
            const uint64_t adesc = desc_sw128(sa);
            const uint64_t bdesc = desc_sw128(sa + k2ABytes);
#pragma unroll
            for (int k = 0; k < k2BK / 16; ++k) umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, (kb | k) ? 1u : 0u);
            umma2_commit_mc(empty0 + 8 * stage);               // frees the slot in BOTH CTAs
            if (kb == k_blocks - 1) umma2_commit_mc(tfull0 + 8 * ab);
          }
          __syncwarp();
          if (++stage == k2Stages) { stage = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): sum of squares of this CTA's 128 x 256 half ==========
    const uint32_t quarter = (uint32_t)warp & 3u;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int a = (tile / mp_tiles) / n_tiles;
      const uint32_t ab = (uint32_t)it & 1u;
      const uint32_t aph = ((uint32_t)it >> 1) & 1u;
      mb_wait(tfull0 + 8 * ab, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ab * k2BN + ((quarter * 32u) << 16);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
      for (int c = 0; c < k2BN; c += 64) {
        uint32_t v0[32], v1[32];
        tm_ld32(taddr + c, v0);
        tm_ld32(taddr + c + 32, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float f0 = __uint_as_float(v0[j]), f1 = __uint_as_float(v1[j]);
          acc0 = __fmaf_rn(f0, f0, acc0);
          acc1 = __fmaf_rn(f1, f1, acc1);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive_cluster(tempty0 + 8 * ab, 0);   // tell the leader this half is drained
      double d = (double)acc0 + (double)acc1;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
      if (lane == 0) atomicAdd(err + a, d);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                          // nobody leaves while the peer may still signal us
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(k2TmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns AWQK_OK, or an error; caller falls back to nothing (errors are reported)
int launch_sqerr_gemm2(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s, double* err,
                       cudaStream_t st) {
  static EncodeTiledFn2 encode = []() -> EncodeTiledFn2 {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn2>(p);
  }();
  if (encode == nullptr) return AWQK_E_NODEVICE;
  CUtensorMap map_x, map_dw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {k2BK, k2BM};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_bf16), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AWQK_E_BADARG;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)C, (cuuint64_t)n_s};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)C * (cuuint64_t)K * 2};
    const cuuint32_t box[3] = {k2BK, k2BNh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_dw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(dw_bf16), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AWQK_E_BADARG;
  }
  const int mp_tiles = (int)ceil_div(T, 256), n_tiles = (int)ceil_div(C, k2BN), k_blocks = (int)ceil_div(K, k2BK);
  const int64_t total = (int64_t)n_s * mp_tiles * n_tiles;
  if (total > 0x7FFFFFFF) return AWQK_E_BADARG;
  int dev = 0, sms = 0;
  AWQK_CUDA(cudaGetDevice(&dev));
  AWQK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = (size_t)k2Stages * k2StageBytes + 1024 + 256;
  {
    static std::atomic<uint64_t> configured{0};
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
      AWQK_CUDA(cudaFuncSetAttribute(sqerr_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured.fetch_or(bit, std::memory_order_release);
    }
  }
  const unsigned pairs = (unsigned)std::min<int64_t>(total, sms / 2);
  sqerr_gemm2_kernel<<<pairs * 2, k2Threads, smem, st>>>(map_x, map_dw, n_s, mp_tiles, n_tiles, k_blocks, err);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

}  // namespace awqk
