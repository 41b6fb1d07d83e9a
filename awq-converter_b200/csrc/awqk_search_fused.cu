// K2 fused: the scores  err_i = sum_{t,c} ( X . dW_i^T )^2  of one tensor for the whole alpha grid in ONE persistent
// kernel -- the fake-quant delta operand is produced by warps of the same CTAs that run the tcgen05 GEMM, while
// the tensor pipe is busy, instead of by a separate kernel that writes [n_grid, C, K] bf16 to HBM first.
//
//   CTA pair (cluster of 2, one per SM pair, persistent), 15 warps per CTA placed by scheduler (warp % 4):
//     9 PRODUCER warps (0-2, 4-6, 8-10)  dW units: 8 rows x 512 columns, the same arithmetic as the stand-alone
//                             kernel (delta16, awqk_search.cuh) -> ring entry in global memory (L2 resident)
//                             -> ready[entry] += 1 (release)
//     warp 3   MMA            leader CTA only, ONE elected thread: tcgen05.mma.cta_group::2 (M = 256 across the
//                             pair, N = 256), TMEM double buffered; when the last k-block of a panel has LANDED in
//                             shared memory it adds 1 to done[entry] (release)
//     warp 7   TMA            waits ready[entry] (acquire + cross-proxy fence), then streams X and dW k-blocks
//                             into the 7-stage SWIZZLE_128B ring (cp.async.bulk.tensor, cta_group::2)
//     warps 11-14 EPILOGUE    tcgen05.ld of the accumulator, sum of squares, one fp64 atomic per warp and tile
//   The MMA thread's instruction stream paces the tensor pipe (4 UMMAs per 512 tensor cycles).  It shares its
//   scheduler with the TMA warp and one epilogue warp only -- never with a producer warp.
//
//   Work decomposition.  slab q = n_tile * n_grid + alpha (256 W rows x K for one alpha; alpha fastest, so the W
//   rows stay hot for the whole grid);  tile t = q * m_tiles + m_tile;  WAVE w = tiles [w * pairs, (w+1) * pairs):
//   pair p runs tile w * pairs + p, so all pairs move from wave to wave together.  A slab is cut into PANELS of
//   Kc <= 2048 columns.  A ring ENTRY is one panel of one slab of one wave (256 x Kc bf16); entries are produced in
//   the order the tensor pipe needs them -- wave, then panel, then slab -- into a FIFO of `depth` panels per slab
//   position, and an entry's slot is reused once every tile of its slab in that wave has loaded the panel.  So the
//   pairs of a wave can never drift further apart than `depth` panels: the live data -- depth panels of the ~10
//   slabs in flight plus the same k-window of X -- stays inside L2 for any K (whole-slab rings of the first
//   version spilled 3.4 TB/s to HBM at K = 14336 and pulled the clock to 0.86 GHz under the power cap).  A slab
//   that straddles two waves is produced once per wave (+ ~10 % producer work, no cross-wave state).
//
// Progress: entries are produced in a total order; entry e waits only for the consumers of entry e - ring, which
// wait only for entries <= e - ring and for their pair's previous wave.  The lowest unfinished entry can therefore
// always progress as long as all CTAs are resident -- the kernel is launched cooperatively, with at most
// cudaOccupancyMaxActiveClusters pairs.  Every wait is bounded (trap after 8 s) so that a protocol bug faults
// instead of hanging the device.
#include <algorithm>
#include <atomic>

#include "awqk_search.cuh"
#include "awqk_tc.cuh"

namespace awqk {

constexpr int kfBM = 128, kfBN = 256, kfBNh = 128, kfBK = 64;   // per-CTA A rows, pair N, per-CTA B rows
constexpr int kfStages = 7;
constexpr int kfABytes = kfBM * kfBK * 2;
constexpr int kfBBytes = kfBNh * kfBK * 2;
constexpr int kfStageBytes = kfABytes + kfBBytes;
constexpr int kfProducerWarps = 9;
constexpr int kfThreads = 15 * 32;                              // 480: see the role table in the kernel
constexpr uint32_t kfTmemCols = 512;
constexpr int kfSlabRows = 256;
constexpr int kfUnitRows = 8, kfUnitCols = 512;
constexpr int kfCtr = 32;                                      // one counter per 128-byte line (uint32 stride):
                                                               // pollers of different slabs hit different L2 lines

// ---- bounded waits (a protocol bug must fault, not hang the device) --------------------------------------------
constexpr unsigned long long kfWaitLimitNs = 8000000000ull;    // 8 s
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mb_wait_bounded(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  unsigned long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (!ok) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > kfWaitLimitNs) __trap();
    }
  } while (!ok);
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t target, unsigned sleep_ns) {
  if (ld_acquire_gpu(p) >= target) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(sleep_ns);
    if (global_ns() - t0 > kfWaitLimitNs) __trap();
  }
}
__device__ __forceinline__ void producer_bar() {               // named barrier 1: the producer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kfProducerWarps * 32) : "memory");
}

template <typename WT, int G, int BITS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kfThreads, 1)
search_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_ring,
                    const WT* __restrict__ w, const float* __restrict__ s_grid, __nv_bfloat16* __restrict__ ring_base,
                    int64_t C, const FusedGeom g, uint32_t* __restrict__ ready, uint32_t* __restrict__ done,
                    double* __restrict__ err) {
  extern __shared__ uint8_t kf_raw[];
  const uint32_t raw = s2u(kf_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;               // identical offset in both CTAs of the pair
  uint8_t* gsm = kf_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + kfStages * kfStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kfStages + 4);
  const uint32_t full0 = s2u(bars), empty0 = full0 + 8 * kfStages;
  const uint32_t tfull0 = empty0 + 8 * kfStages, tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                        // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int K = g.K, Kc = g.Kc, panels = g.panels, n_grid = g.n_grid, mp_tiles = g.mp_tiles;
  const int n_slabs = g.n_tiles * n_grid;
  const int total_tiles = n_slabs * mp_tiles;
  const int n_waves = (total_tiles + n_pairs - 1) / n_pairs;
  const int cb_max = (Kc + kfUnitCols - 1) / kfUnitCols;       // unit columns per panel
  const int upe = (kfSlabRows / kfUnitRows) * cb_max;          // unit slots per entry (a ragged last panel fills fewer)
  // entry (wave, panel, slab position) -> index in production order
  auto entry_of = [&](int wv, int j, int sp) { return (wv * panels + j) * g.smax + sp; };
  auto wave_q_lo = [&](int wv) { return (wv * n_pairs) / mp_tiles; };
  auto wave_nq = [&](int wv) {                                 // slabs touched by the wave
    const int t_end = min(total_tiles, (wv + 1) * n_pairs);
    return (t_end - 1) / mp_tiles - (wv * n_pairs) / mp_tiles + 1;
  };
  auto tiles_of = [&](int wv, int q) {                         // tiles of slab q inside wave wv
    const int lo = max(q * mp_tiles, wv * n_pairs), hi = min((q + 1) * mp_tiles, min(total_tiles, (wv + 1) * n_pairs));
    return hi - lo;
  };
  auto panel_cols = [&](int j) { return min(Kc, K - j * Kc); };
  // warp roles by scheduler (a warp runs on sub-partition warp % 4): the MMA issuer and the TMA warp share
  // sub-partition 3 with one epilogue warp only, so no producer warp ever competes with them for issue slots
  //   warp  0 1 2 | 3   | 4 5 6 | 7   | 8 9 10 | 11  12 13 14
  //   role  P P P | MMA | P P P | TMA | P P P  | epilogue (TMEM lane quarter = warp % 4)
  constexpr int kMmaWarp = 3, kTmaWarp = 7;
  const bool is_producer = (warp & 3) < 3 && warp < 11;
  const int pw = (warp >> 2) * 3 + (warp & 3);                 // producer index 0..8

  if (threadIdx.x == 0) {
    for (int s = 0; s < kfStages; ++s) {
      mb_init(full0 + 8 * s, 1);
      mb_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mb_init(tfull0 + 8 * a, 1);
      mb_init(tempty0 + 8 * a, 8);                             // 4 epilogue warps x 2 CTAs (used on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {   // the same logical warp in both CTAs allocates (and later frees) the pair's TMEM
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)),
                 "r"(kfTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                          // peer barriers are initialised before any remote signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (is_producer) {
    // ===================== delta producers =====================
    const bool sym = g.sym != 0;
    const float qmin = sym ? -(float)(1 << (BITS - 1)) : 0.0f;
    const float qmax = sym ? (float)((1 << (BITS - 1)) - 1) : (float)((1 << BITS) - 1);
    const int n_entries = n_waves * panels * g.smax;
    const int64_t n_units = (int64_t)n_entries * upe;
    const int64_t stride = (int64_t)gridDim.x * kfProducerWarps;
    const int ring = g.ring;
    int checked = ring - 1;                                    // entries <= checked may be written (their slot is free)
    // every producer warp of the CTA runs the same number of rounds (a warp without a unit in the last round only
    // joins the barrier): ONE lane per CTA polls the ring, not one per warp
    auto wait_slot = [&](int e) {                              // the entry that used e's slot has been loaded by all its tiles
      const int ep = e - ring;                                 // same slab position, `depth` panels earlier
      const int sp = ep % g.smax, wj = ep / g.smax;
      const int wv = wj / panels;
      if (sp >= wave_nq(wv)) return;                           // an unused position: nothing was ever written there
      wait_counter(done + (int64_t)ep * kfCtr, (uint32_t)tiles_of(wv, wave_q_lo(wv) + sp), 500);
    };
#pragma unroll 1
    for (int64_t v0 = (int64_t)blockIdx.x * kfProducerWarps; v0 < n_units; v0 += stride) {
      const int64_t v = v0 + pw;
      const int64_t v_hi = min(v0 + kfProducerWarps - 1, n_units - 1);
      const int e_lo = (int)(v0 / upe), e_hi = (int)(v_hi / upe);                          // <= 2 distinct entries
      if (e_hi > checked) {                                    // CTA uniform
        if (pw == 0 && lane == 0) {
          if (e_lo > checked) wait_slot(e_lo);
          if (e_hi != e_lo) wait_slot(e_hi);
        }
        producer_bar();
        checked = e_hi;
      }
      if (v >= n_units) continue;
      const int e = (int)(v / upe);
      const int u = (int)(v - (int64_t)e * upe);
      const int sp = e % g.smax, wj = e / g.smax;
      const int wv = wj / panels, j = wj - wv * panels;
      if (sp >= wave_nq(wv)) continue;                         // unused slab position of this wave
      const int rb = u / cb_max, cb = u - rb * cb_max;
      const int pcols = panel_cols(j);
      if (cb * kfUnitCols >= pcols) continue;                  // ragged last panel: unit slot past its columns
      const int q = wave_q_lo(wv) + sp;
      const int nt = q / n_grid, a = q - nt * n_grid;
      const int slot = e % ring;
      const int pcol = cb * kfUnitCols + lane * 16;            // column inside the panel
      const int col = j * Kc + pcol;                           // column of W / s_grid
      const bool cvalid = pcol < pcols;
      float2 sv[8], rs[8];
      if (cvalid) {
        const float4* sp = reinterpret_cast<const float4*>(s_grid + (int64_t)a * K + col);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 t = __ldg(sp + c);
          sv[2 * c] = make_float2(t.x, t.y);
          sv[2 * c + 1] = make_float2(t.z, t.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sv[i] = make_float2(1.0f, 1.0f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) rs[i] = make_float2(refined_rcp(sv[i].x), refined_rcp(sv[i].y));
      const int64_t row0 = (int64_t)nt * kfSlabRows + rb * kfUnitRows;
      const WT* wp = w + row0 * K + col;
      __nv_bfloat16* dst = ring_base + ((int64_t)slot * kfSlabRows + rb * kfUnitRows) * Kc + pcol;
      Raw16<WT> nxt;
      if (cvalid && row0 < C) nxt.load(wp); else nxt.zero();
#pragma unroll 1
      for (int r = 0; r < kfUnitRows; ++r) {
        const Raw16<WT> cur = nxt;
        if (r + 1 < kfUnitRows) {
          if (cvalid && row0 + r + 1 < C) nxt.load(wp + (int64_t)(r + 1) * K); else nxt.zero();
        }
        uint32_t o[8];
        if (row0 + r < C) {                                    // warp uniform
          float2 wv[8];
          cur.unpack(wv);
          delta16<G, BITS>(wv, sv, [&](int i) { return rs[i]; }, sym, qmin, qmax, o);
        } else {                                               // rows past C: the slot is reused, write the zeros
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = 0u;
        }
        if (cvalid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + (int64_t)r * Kc);
          d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
          d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      // generic-proxy writes -> visible to the TMA (async proxy) reads of other SMs: fence, then release
      __threadfence();
      fence_proxy_async_global();
      __syncwarp();
      if (lane == 0) red_release_gpu_add(ready + (int64_t)e * kfCtr, 1u);
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA (both CTAs; completion lands on the LEADER's full barrier) ==========
    if (elect_one()) {
      uint32_t stage = 0, ph = 1;
      int wv = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++wv) {
        const int q = tile / mp_tiles;
        const int mp = tile - q * mp_tiles;
        const int sp = q - wave_q_lo(wv);
        auto units_of = [&](int j) {
          return (uint32_t)((kfSlabRows / kfUnitRows) * ((panel_cols(j) + kfUnitCols - 1) / kfUnitCols));
        };
        // the counter of the NEXT panel is read while this panel's loads are issued (the load is in flight behind
        // the TMA issue loop; an acquire round trip through a busy L2 otherwise opens a bubble at every panel start)
        uint32_t seen = ld_acquire_gpu(ready + (int64_t)entry_of(wv, 0, sp) * kfCtr);
        for (int j = 0; j < panels; ++j) {
          const int e = entry_of(wv, j, sp);
          const int pcols = panel_cols(j);
          if (seen < units_of(j)) wait_counter(ready + (int64_t)e * kfCtr, units_of(j), 100);   // the whole panel has been produced
          fence_proxy_async_global();
          const int slot = e % g.ring;
          const int kbs = pcols / kfBK;
          for (int kb = 0; kb < kbs; ++kb) {
            if (kb == 0 && j + 1 < panels) seen = ld_acquire_gpu(ready + (int64_t)entry_of(wv, j + 1, sp) * kfCtr);
            mb_wait_bounded(empty0 + 8 * stage, ph);           // own slot free (multicast commit from the leader)
            const uint32_t lbar = (full0 + 8 * stage) & kPeerMask;
            if (rank == 0) mb_expect_tx(full0 + 8 * stage, 2 * kfStageBytes);
            const uint32_t sa = base + stage * kfStageBytes;
            tma2_load_2d(sa, &map_x, j * Kc + kb * kfBK, mp * 256 + (int)rank * kfBM, lbar);
            tma2_load_3d(sa + kfABytes, &map_ring, kb * kfBK, (int)rank * kfBNh, slot, lbar);
            if (++stage == kfStages) { stage = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: ONE elected thread of the leader CTA runs the whole loop ==========
    // (this instruction stream paces the tensor pipe -- 4 UMMAs per 512 tensor cycles -- so it is kept short:
    // descriptors advance by adds, nothing is recomputed per k-block, and the warp sits on a scheduler it
    // shares with no producer warp, see the role table above)
    if (rank == 0 && elect_one()) {
      uint32_t stage = 0, ph = 0;
      const uint32_t lo0 = desc_lo_sw128(base);
      uint32_t alo = lo0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const int q = tile / mp_tiles;
        const int sp = q - wave_q_lo(it);                      // it == the wave
        const uint32_t ab = (uint32_t)it & 1u;
        const uint32_t aph = ((uint32_t)it >> 1) & 1u;
        mb_wait_bounded(tempty0 + 8 * ab, aph ^ 1u);           // both CTAs' epilogues drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + ab * kfBN;
        uint32_t acc = 0u;
        for (int j = 0; j < panels; ++j) {
          const int kbs = panel_cols(j) / kfBK;
          for (int kb = 0; kb < kbs; ++kb) {
            mb_wait_bounded(full0 + 8 * stage, ph);            // both CTAs' A and B halves have landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t blo = alo + (kfABytes >> 4);
            umma2_bf16_lo(d_tmem, alo, blo, acc);
            umma2_bf16_lo(d_tmem, alo + 2, blo + 2, 1u);
            umma2_bf16_lo(d_tmem, alo + 4, blo + 4, 1u);
            umma2_bf16_lo(d_tmem, alo + 6, blo + 6, 1u);
            acc = 1u;
            umma2_commit_mc(empty0 + 8 * stage);               // frees the slot in BOTH CTAs
            alo += (kfStageBytes >> 4);
            if (++stage == kfStages) { stage = 0; ph ^= 1u; alo = lo0; }
          }
          // the panel's last k-block is in shared memory (observed through its mbarrier): this tile no longer
          // needs the ring entry.  Relaxed: the TMA reads have completed, this thread has nothing of its own to
          // publish, and a release fence here would stall the thread that paces the tensor pipe once per panel
          red_relaxed_gpu_add(done + (int64_t)entry_of(it, j, sp) * kfCtr, 1u);
        }
        umma2_commit_mc(tfull0 + 8 * ab);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): sum of squares of this CTA's 128 x 256 half ==========
    const uint32_t quarter = (uint32_t)warp & 3u;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int a = (tile / mp_tiles) % n_grid;
      const uint32_t ab = (uint32_t)it & 1u;
      const uint32_t aph = ((uint32_t)it >> 1) & 1u;
      mb_wait_bounded(tfull0 + 8 * ab, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ab * kfBN + ((quarter * 32u) << 16);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
      for (int c = 0; c < kfBN; c += 64) {
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
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kfTmemCols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
constexpr size_t kfSmemBytes = (size_t)kfStages * kfStageBytes + 1024 + 256;

// co-resident CTA pairs on the current device (cooperative launch); cached per device, write-once
static int fused_pairs_cap() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int v = cache[dev].load(std::memory_order_acquire);
  if (v > 0) return v;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  auto kernel = search_fused_kernel<float, 128, 8>;            // the instance with the most registers
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kfSmemBytes) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kfThreads, 1, 1);
  cfg.dynamicSmemBytes = kfSmemBytes;
  cfg.gridDim = dim3((unsigned)(sms / 2) * 2, 1, 1);
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) return 0;
  v = std::min(sms / 2, max_clusters);
  if (v > 0) cache[dev].store(v, std::memory_order_release);
  return v;
}

// geometry + workspace needs of one search; depth = panels of lookahead per slab position (>= 2)
int fused_plan(int64_t C, int64_t K, int64_t T, int n_grid, FusedPlan* p) {
  const int cap = fused_pairs_cap();
  if (cap <= 0) return AWQK_E_NODEVICE;
  const int64_t n_tiles = ceil_div(C, kfSlabRows), mp_tiles = ceil_div(T, 256);
  const int64_t n_slabs = n_tiles * n_grid, total = n_slabs * mp_tiles;
  if (total > 0x3FFFFFFF || K > 0x3FFFFFFF) return AWQK_E_BADARG;
  const int pairs = (int)std::min<int64_t>(cap, total);
  const int Kc = K <= 2048 ? (int)K : 2048;
  const int panels = (int)ceil_div(K, Kc);
  const int smax = (int)std::min<int64_t>(n_slabs, (pairs - 1) / mp_tiles + 2);
  const int64_t n_waves = ceil_div(total, pairs);
  const int64_t n_entries = n_waves * panels * smax;
  if (n_entries > 0x3FFFFFFF / ((kfSlabRows / kfUnitRows) * 4)) return AWQK_E_BADARG;   // unit index stays far below 2^63, entry index below 2^31
  p->pairs = pairs;
  p->n_entries = n_entries;
  p->sync_bytes = (size_t)n_entries * 2 * kfCtr * sizeof(uint32_t);
  p->entry_bytes = (size_t)kfSlabRows * Kc * 2;
  p->depth_min = (int)std::min<int64_t>(2, n_waves * panels);
  // measured (tools/probe_fused.py --ring): the depth does not limit K <= 4096; for larger K every extra panel is
  // L2 footprint (4096x14336: depth 2 -> 1163, depth 6 -> 1125 TFLOP/s)
  p->depth_pref = (int)std::min<int64_t>(panels > 2 ? 2 : 3, n_waves * panels);
  p->g = FusedGeom{(int)K, Kc, panels, n_grid, (int)mp_tiles, (int)n_tiles, smax, 0, 0, 0};
  return AWQK_OK;
}

template <typename WT, int G, int BITS>
static int launch_fused_t(const CUtensorMap& map_x, const CUtensorMap& map_ring, const WT* w, const float* s_grid,
                          __nv_bfloat16* ring_base, int64_t C, const FusedGeom& g, int pairs, uint32_t* ready,
                          uint32_t* done, double* err, cudaStream_t st) {
  auto kernel = search_fused_kernel<WT, G, BITS>;
  AWQK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kfSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kfThreads, 1, 1);
  cfg.dynamicSmemBytes = kfSmemBytes;
  cfg.stream = st;
  cfg.gridDim = dim3((unsigned)pairs * 2, 1, 1);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;                 // all CTAs resident: they wait on one another
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AWQK_CUDA(cudaLaunchKernelEx(&cfg, kernel, map_x, map_ring, w, s_grid, ring_base, C, g, ready, done, err));
  return AWQK_OK;
}

int launch_search_fused(const void* w, int dtype, int64_t C, int64_t K, const void* x_bf16, int64_t T,
                        const float* s_grid, int n_grid, int g, int bits, bool sym, double* err_sum, void* sync,
                        void* ring_base, int depth, cudaStream_t st) {
  if (tensor_map_encoder() == nullptr) return AWQK_E_NODEVICE;
  FusedPlan plan;
  const int rc = fused_plan(C, K, T, n_grid, &plan);
  if (rc != AWQK_OK) return rc;
  if (depth < plan.depth_min) return AWQK_E_WORKSPACE;
  FusedGeom geom = plan.g;
  geom.depth = std::min(depth, 8);
  geom.ring = geom.smax * geom.depth;
  geom.sym = sym ? 1 : 0;
  CUtensorMap map_x, map_ring;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    if (!encode_bf16_sw128(&map_x, x_bf16, 2, dims, strides, kfBM)) return AWQK_E_BADARG;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)geom.Kc, (cuuint64_t)kfSlabRows, (cuuint64_t)geom.ring};
    const cuuint64_t strides[2] = {(cuuint64_t)geom.Kc * 2, (cuuint64_t)plan.entry_bytes};
    if (!encode_bf16_sw128(&map_ring, ring_base, 3, dims, strides, kfBNh)) return AWQK_E_BADARG;
  }
  uint32_t* ready = reinterpret_cast<uint32_t*>(sync);
  uint32_t* done = ready + plan.n_entries * kfCtr;
  AWQK_CUDA(cudaMemsetAsync(sync, 0, plan.sync_bytes, st));
  auto* ring_bf = reinterpret_cast<__nv_bfloat16*>(ring_base);
#define AWQK_FUSED(WT, GG, BB)                                                                                     \
  return launch_fused_t<WT, GG, BB>(map_x, map_ring, reinterpret_cast<const WT*>(w), s_grid, ring_bf, C, geom,    \
                                    plan.pairs, ready, done, err_sum, st)
#define AWQK_FUSED_G(WT, BB)                                  \
  do {                                                        \
    if (g == 32) AWQK_FUSED(WT, 32, BB);                      \
    if (g == 64) AWQK_FUSED(WT, 64, BB);                      \
    AWQK_FUSED(WT, 128, BB);                                  \
  } while (0)
  if (dtype == AWQK_BF16) { if (bits == 4) AWQK_FUSED_G(__nv_bfloat16, 4); else AWQK_FUSED_G(__nv_bfloat16, 8); }
  if (dtype == AWQK_FP16) { if (bits == 4) AWQK_FUSED_G(__half, 4); else AWQK_FUSED_G(__half, 8); }
  if (dtype == AWQK_FP32) { if (bits == 4) AWQK_FUSED_G(float, 4); else AWQK_FUSED_G(float, 8); }
#undef AWQK_FUSED_G
#undef AWQK_FUSED
  return AWQK_E_UNSUPPORTED;
}

}  // namespace awqk
