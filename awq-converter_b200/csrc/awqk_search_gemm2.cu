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
#include <algorithm>
#include <atomic>

#include "awqk_search.cuh"
#include "awqk_tc.cuh"

namespace awqk {

constexpr int k2BM = 128, k2BN = 256, k2BNh = 128, k2BK = 64;   // per-CTA A rows, pair N, per-CTA B rows
constexpr int k2Stages = 7;
constexpr int k2ABytes = k2BM * k2BK * 2;    // 16 KiB
constexpr int k2BBytes = k2BNh * k2BK * 2;   // 16 KiB
constexpr int k2StageBytes = k2ABytes + k2BBytes;
constexpr int k2Threads = 192;
constexpr uint32_t k2TmemCols = 512;

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
    if (elect_one()) {
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
    // ===================== MMA issuer: ONE elected thread of the leader CTA runs the whole loop ==========
    // (the loop must stay short: it is the pacing instruction stream of the tensor pipe -- 4 UMMAs per
    // 512 tensor cycles; descriptors are advanced by adds, nothing is recomputed per k-block)
    if (rank == 0 && elect_one()) {
      uint32_t stage = 0, ph = 0;
      const uint32_t lo0 = desc_lo_sw128(base);
      uint32_t alo = lo0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const uint32_t ab = (uint32_t)it & 1u;
        const uint32_t aph = ((uint32_t)it >> 1) & 1u;
        mb_wait(tempty0 + 8 * ab, aph ^ 1u);                   // both CTAs' epilogues drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + ab * k2BN;
        uint32_t acc = 0u;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mb_wait(full0 + 8 * stage, ph);                      // both CTAs' A and B halves have landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t blo = alo + (k2ABytes >> 4);
          umma2_bf16_lo(d_tmem, alo, blo, acc);
          umma2_bf16_lo(d_tmem, alo + 2, blo + 2, 1u);
          umma2_bf16_lo(d_tmem, alo + 4, blo + 4, 1u);
          umma2_bf16_lo(d_tmem, alo + 6, blo + 6, 1u);
          acc = 1u;
          umma2_commit_mc(empty0 + 8 * stage);                 // frees the slot in BOTH CTAs
          alo += (k2StageBytes >> 4);
          if (++stage == k2Stages) { stage = 0; ph ^= 1u; alo = lo0; }
        }
        umma2_commit_mc(tfull0 + 8 * ab);
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

// returns AWQK_OK, or an error code
int launch_sqerr_gemm2(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K, int n_s, double* err,
                       cudaStream_t st) {
  if (tensor_map_encoder() == nullptr) return AWQK_E_NODEVICE;
  CUtensorMap map_x, map_dw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    if (!encode_bf16_sw128(&map_x, x_bf16, 2, dims, strides, k2BM)) return AWQK_E_BADARG;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)C, (cuuint64_t)n_s};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)C * (cuuint64_t)K * 2};
    if (!encode_bf16_sw128(&map_dw, dw_bf16, 3, dims, strides, k2BNh)) return AWQK_E_BADARG;
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
