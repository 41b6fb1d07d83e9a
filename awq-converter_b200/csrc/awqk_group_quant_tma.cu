// K1 v2: the bandwidth path for int4 packing of bf16 / fp16 weights (the headline configuration).
//
// Same arithmetic as group_quant_flat (awq.py:173-250 bit-for-bit), restructured so that the SM's
// issue slots stop being the limiter (v1 measured 26.6 instructions per element, 84 % issue-active,
// DRAM 32 % -- profiles/r01_k1_v1.md):
//   * HBM -> shared memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), a ring of
//     kStages x 16 KiB per CTA filled by one producer thread; no per-thread load instructions, no
//     address arithmetic in the consumers, 32-64 KiB per CTA in flight.
//   * each consumer thread owns 32 consecutive elements (64 B, read as 4 conflict-free LDS.128), so
//     a group of 128 is 4 lanes (2 shuffle steps), of 64 two lanes, of 32 one lane, and its 32
//     codes are exactly 4 packed words = one 16-byte store (512 B contiguous per warp).
//   * min/max on packed bf16x2/fp16x2 words (HMNMX2.NAN), the exact hoisted IEEE division as
//     packed fp32x2 math (FMUL2/FFMA2/FADD2: two elements per issue slot), round-to-integer by a
//     magic-constant add, clamp + rebase in ONE integer instruction per pair (VIADDMNMX.S16x2.RELU),
//     nibble merge by one IMAD per pair + 3 PRMT per word.
// A warp tile is 1024 elements = 8/16/32 groups, i.e. a whole number of packed zero-point words.
#include <atomic>
#include <cstring>

#include "awqk_common.cuh"

namespace awqk {

constexpr int kV2ConsumerWarps = 8;
constexpr int kV2Threads = (kV2ConsumerWarps + 1) * 32;   // + 1 producer warp
constexpr int kV2WarpTile = 1024;                         // elements
constexpr int kV2CtaTile = kV2WarpTile * kV2ConsumerWarps;  // 8192 elements = 16 KiB
constexpr int kV2Stages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(1000000u)   // suspend-time hint (ns): sleep in hardware, do not spin
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- per-input-type packed helpers ---------------------------------------------------------
template <typename InT>
struct Packed;

template <>
struct Packed<__nv_bfloat16> {
  using V2 = __nv_bfloat162;
  static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) {
    V2 r = __hmin2_nan(*reinterpret_cast<V2*>(&a), *reinterpret_cast<V2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    V2 r = __hmax2_nan(*reinterpret_cast<V2*>(&a), *reinterpret_cast<V2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float2 to_f2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  }
};
template <>
struct Packed<__half> {
  using V2 = __half2;
  static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) {
    V2 r = __hmin2_nan(*reinterpret_cast<V2*>(&a), *reinterpret_cast<V2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    V2 r = __hmax2_nan(*reinterpret_cast<V2*>(&a), *reinterpret_cast<V2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float2 to_f2(uint32_t w) {
    return __half22float2(*reinterpret_cast<V2*>(&w));
  }
};

// fp32 input never takes the packed 16-bit paths; these only keep the discarded branches well-formed
template <>
struct Packed<float> {
  static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t) { return a; }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t) { return a; }
  static __device__ __forceinline__ float2 to_f2(uint32_t w) { return make_float2(__uint_as_float(w), 0.0f); }
};

__device__ __forceinline__ float v2_fmin_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float v2_fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// pair of clamped unsigned codes (lo | hi << 16), each in [0, 2^BITS - 1], from a packed pair of
// exact quotients q = x / s.  A selects the reference's arithmetic dtype.
template <int A, int QMIN, int BITS>
struct PairQuant;

template <int QMIN, int BITS>
struct PairQuant<AR_F32, QMIN, BITS> {
  // v = q + zp (fp32), t = v + 1.5*2^23 -> low 16 bits of each t hold round(v) as an s16
  // (|v| < 2^14 guaranteed by the fast-group predicate)
  // the magic constant carries -QMIN, so the low half IS the unsigned code before clamping
  // (adding an integer to the magic does not move rounding ties: 1.5*2^23 - QMIN stays even-aligned)
  float2 zp2, magic2;
  __device__ __forceinline__ void prepare() { magic2 = make_float2(12582912.0f - (float)QMIN, 12582912.0f - (float)QMIN); }
  __device__ __forceinline__ void init(float zp) { zp2 = make_float2(zp, zp); }
  __device__ __forceinline__ uint32_t run(float2 q) const {
    const float2 t = __fadd2_rn(__fadd2_rn(q, zp2), magic2);
    const uint32_t both = __byte_perm(__float_as_uint(t.x), __float_as_uint(t.y), 0x5410);
    return __vimin_s16x2_relu(both, (uint32_t)((1 << BITS) - 1) * 0x00010001u);
  }
  // int4: 8 quotients (elements 0..7 = q[0].x, q[0].y, q[1].x, ...) -> one packed word.  Element j is paired with
  // element j + 4 (low / high s16 lane), so the clamped pair r_j only has to move left by 4j bits: the nibble merge
  // is 3 IMADs, no byte gather.  sel = 0x5410, or 0x1054 to swap the two halves of the word (fp32 input, odd lanes).
  __device__ __forceinline__ uint32_t pack8(const float2 (&q)[4], uint32_t sel) const {
    float2 t[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) t[p] = __fadd2_rn(__fadd2_rn(q[p], zp2), magic2);
    const uint32_t r0 = __vimin_s16x2_relu(__byte_perm(__float_as_uint(t[0].x), __float_as_uint(t[2].x), sel), 0x000F000Fu);
    const uint32_t r1 = __vimin_s16x2_relu(__byte_perm(__float_as_uint(t[0].y), __float_as_uint(t[2].y), sel), 0x000F000Fu);
    const uint32_t r2 = __vimin_s16x2_relu(__byte_perm(__float_as_uint(t[1].x), __float_as_uint(t[3].x), sel), 0x000F000Fu);
    const uint32_t r3 = __vimin_s16x2_relu(__byte_perm(__float_as_uint(t[1].y), __float_as_uint(t[3].y), sel), 0x000F000Fu);
    return (r3 * 16u + r2) * 256u + (r1 * 16u + r0);
  }
  // the same without the clamp (FastGroup::noclamp for every group of the warp): the low halves are the codes
  __device__ __forceinline__ uint32_t pack8_noclamp(const float2 (&q)[4], uint32_t sel) const {
    float2 t[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) t[p] = __fadd2_rn(__fadd2_rn(q[p], zp2), magic2);
    const uint32_t r0 = __byte_perm(__float_as_uint(t[0].x), __float_as_uint(t[2].x), sel);
    const uint32_t r1 = __byte_perm(__float_as_uint(t[0].y), __float_as_uint(t[2].y), sel);
    const uint32_t r2 = __byte_perm(__float_as_uint(t[1].x), __float_as_uint(t[3].x), sel);
    const uint32_t r3 = __byte_perm(__float_as_uint(t[1].y), __float_as_uint(t[3].y), sel);
    return (r3 * 16u + r2) * 256u + (r1 * 16u + r0);
  }
};
template <int QMIN>
struct PairQuant<AR_BF16, QMIN, 4> {
  // a = bf16(q); b = bf16(a + zp); t = bf16(b + 192): [128,256) has ulp 1 -> bits = 0x4340 + round(b)
  __nv_bfloat162 zp2, magic2;
  uint32_t rebase;
  __device__ __forceinline__ void prepare() {
    magic2 = __float2bfloat162_rn(192.0f);
    rebase = (uint32_t)((-(0x4340 + QMIN)) & 0xFFFF) * 0x00010001u;
    asm volatile("" : "+r"(rebase));                 // one live register, not one re-materialisation per use
  }
  __device__ __forceinline__ void init(float zp) { zp2 = __float2bfloat162_rn(zp); }
  __device__ __forceinline__ uint32_t run(float2 q) const {
    const __nv_bfloat162 a = __float22bfloat162_rn(q);
    const __nv_bfloat162 t = __hadd2(__hadd2(a, zp2), magic2);
    return __viaddmin_s16x2_relu(*reinterpret_cast<const uint32_t*>(&t), rebase, 0x000F000Fu);
  }
  __device__ __forceinline__ uint32_t pair(float lo, float hi) const {
    const __nv_bfloat162 t = __hadd2(__hadd2(__floats2bfloat162_rn(lo, hi), zp2), magic2);
    return __viaddmin_s16x2_relu(*reinterpret_cast<const uint32_t*>(&t), rebase, 0x000F000Fu);
  }
  __device__ __forceinline__ uint32_t pack8_noclamp(const float2 (&q)[4], uint32_t s) const { return pack8(q, s); }
  __device__ __forceinline__ uint32_t pack8(const float2 (&q)[4], uint32_t) const {   // (see PairQuant<AR_F32>)
    const uint32_t r0 = pair(q[0].x, q[2].x), r1 = pair(q[0].y, q[2].y), r2 = pair(q[1].x, q[3].x), r3 = pair(q[1].y, q[3].y);
    return (r3 * 16u + r2) * 256u + (r1 * 16u + r0);
  }
};
template <int QMIN>
struct PairQuant<AR_F16, QMIN, 4> {
  // fp16: [1024,2048) has ulp 1 -> magic 1536 = 0x6600
  __half2 zp2, magic2;
  uint32_t rebase;
  __device__ __forceinline__ void prepare() {
    magic2 = __float2half2_rn(1536.0f);
    rebase = (uint32_t)((-(0x6600 + QMIN)) & 0xFFFF) * 0x00010001u;
    asm volatile("" : "+r"(rebase));
  }
  __device__ __forceinline__ void init(float zp) { zp2 = __float2half2_rn(zp); }
  __device__ __forceinline__ uint32_t run(float2 q) const {
    const __half2 a = __float22half2_rn(q);
    const __half2 t = __hadd2(__hadd2(a, zp2), magic2);
    return __viaddmin_s16x2_relu(*reinterpret_cast<const uint32_t*>(&t), rebase, 0x000F000Fu);
  }
  __device__ __forceinline__ uint32_t pair(float lo, float hi) const {
    const __half2 t = __hadd2(__hadd2(__floats2half2_rn(lo, hi), zp2), magic2);
    return __viaddmin_s16x2_relu(*reinterpret_cast<const uint32_t*>(&t), rebase, 0x000F000Fu);
  }
  __device__ __forceinline__ uint32_t pack8_noclamp(const float2 (&q)[4], uint32_t s) const { return pack8(q, s); }
  __device__ __forceinline__ uint32_t pack8(const float2 (&q)[4], uint32_t) const {
    const uint32_t r0 = pair(q[0].x, q[2].x), r1 = pair(q[0].y, q[2].y), r2 = pair(q[1].x, q[3].x), r3 = pair(q[1].y, q[3].y);
    return (r3 * 16u + r2) * 256u + (r1 * 16u + r0);
  }
};

// 8 bit in the reference's bf16 / fp16 arithmetic: a = A(q), b = A(a + zp) as above, but codes up to 255 leave
// the window in which a bf16 / fp16 magic constant has ulp 1 -- b is widened to fp32 (exact) and rounded with the
// fp32 magic instead (rint of an A-representable value is the same number in either format).
template <int QMIN>
struct PairQuant<AR_BF16, QMIN, 8> {
  __device__ __forceinline__ uint32_t pack8(const float2 (&)[4], uint32_t) const { return 0u; }   // int4 only
  __device__ __forceinline__ uint32_t pack8_noclamp(const float2 (&)[4], uint32_t) const { return 0u; }
  __nv_bfloat162 zp2;
  float2 magic2;
  __device__ __forceinline__ void prepare() { magic2 = make_float2(12582912.0f - (float)QMIN, 12582912.0f - (float)QMIN); }
  __device__ __forceinline__ void init(float zp) { zp2 = __float2bfloat162_rn(zp); }
  __device__ __forceinline__ uint32_t run(float2 q) const {
    const __nv_bfloat162 b = __hadd2(__float22bfloat162_rn(q), zp2);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&b);
    const float2 t = __fadd2_rn(make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u)), magic2);
    const uint32_t both = __byte_perm(__float_as_uint(t.x), __float_as_uint(t.y), 0x5410);
    return __vimin_s16x2_relu(both, 0x00FF00FFu);
  }
};
template <int QMIN>
struct PairQuant<AR_F16, QMIN, 8> {
  __device__ __forceinline__ uint32_t pack8(const float2 (&)[4], uint32_t) const { return 0u; }   // int4 only
  __device__ __forceinline__ uint32_t pack8_noclamp(const float2 (&)[4], uint32_t) const { return 0u; }
  __half2 zp2;
  float2 magic2;
  __device__ __forceinline__ void prepare() { magic2 = make_float2(12582912.0f - (float)QMIN, 12582912.0f - (float)QMIN); }
  __device__ __forceinline__ void init(float zp) { zp2 = __float2half2_rn(zp); }
  __device__ __forceinline__ uint32_t run(float2 q) const {
    const __half2 b = __hadd2(__float22half2_rn(q), zp2);
    const float2 t = __fadd2_rn(__half22float2(b), magic2);
    const uint32_t both = __byte_perm(__float_as_uint(t.x), __float_as_uint(t.y), 0x5410);
    return __vimin_s16x2_relu(both, 0x00FF00FFu);
  }
};

struct V2Out {
  uint32_t* q_packed;   // nullable when q_unpacked is given
  int32_t* q_unpacked;  // nullable; only the UNPACKED instantiations look at it
  __half* scales;
  int32_t* zp;          // nullable
  uint32_t* zp_packed;  // nullable (flat layout only)
  int zq_log2;          // log2(groups per packed zero-point word): 3 = eight consecutive groups of the flat
                        // layout; 0/1/2 = rows of 1/2/4 groups, one zero-padded word per row ([C, 1] qzeros)
};

// UNPACKED: additionally emits the reference's int32 code tensor (awq.py:329, 4 B per element).  The 32
// codes of a thread (128 B) would cost 32 L1 wavefronts per store instruction if written directly, so
// each warp transposes its 1024 codes through a private 4 KiB shared staging area (XOR-swizzled, both
// directions conflict-free) and stores 512 contiguous bytes per instruction.
constexpr int kV2UnpStageBytes = kV2ConsumerWarps * kV2WarpTile * 4;   // 32 KiB per CTA

__device__ __forceinline__ float fmin3_nan(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// One consumer thread of K1 v2: its invariants and the work on one warp tile (32 consecutive elements per
// thread, already in registers).  Shared by the flat kernel and the column-slab kernel below.
//   CS: x = fp32(w) * s[k] with the thread's 32 column scales read from a shared-memory table laid out
//   [chunk j = 0..7][lane] x 16 B (conflict-free LDS.128 at immediate offsets j * 512 from `cs_thr`).
template <typename InT, int A, int G, bool SYM, bool UNPACKED, bool CS, int BITS>
struct V2Consumer {
  static constexpr bool F32IN = sizeof(InT) == 4;
  static constexpr int NLD = F32IN ? 8 : 4;                // LDS.128 per thread and tile
  static constexpr int NW = BITS;                          // packed words per thread: 32 codes * BITS / 32
  static constexpr uint32_t CMAX = (1u << BITS) - 1u;
  static constexpr int LPG = G / 32;                       // lanes per group (4, 2, 1)
  static constexpr int QMIN = SYM ? -(1 << (BITS - 1)) : 0;
  static constexpr float FQMIN = (float)QMIN, FQMAX = (float)(QMIN + (int)CMAX);

  int lane, rot, odd_shift, LPW;
  uint32_t half_sel, zq_shift, zq_mask, zq_lane_mask;
  bool leader, zq_writer;
  uint32_t stg_w[8], stg_r[8];
  PairQuant<A, QMIN, BITS> pq;

  // stg0: this warp's 4 KiB of the UNPACKED staging area (shared-window address)
  __device__ __forceinline__ void init(int lane_, int zq_log2, uint32_t stg0) {
    lane = lane_;
    // rot: bf16/fp16 -> slot c holds chunk (c + rot) & 3; fp32 -> word slot wp holds word wp ^ rot
    rot = (lane >> 1) & 3;
    LPW = LPG << zq_log2;                // lanes per packed zero-point word (8 groups: 32, 16, 8)
    // fp32 input: half-word merge selector (even lanes: slot 2w is the low half of word w; odd lanes: the high half)
    half_sel = (F32IN && (lane & 1)) ? 0x1054u : 0x5410u;
    odd_shift = (F32IN && (lane & 1)) ? 2 : 0;
    leader = (lane % LPG) == 0;
    zq_writer = (lane & (LPW - 1)) == 0;
    zq_lane_mask = leader ? CMAX : 0u;
    zq_shift = (uint32_t)BITS * ((uint32_t)(lane / LPG) & ((1u << zq_log2) - 1u));
    zq_mask = (LPW == 32) ? 0xFFFFFFFFu : (((1u << LPW) - 1u) << (lane & ~(LPW - 1)));
    // UNPACKED: warp-private staging of 256 x 16 B chunks.  Thread l owns logical chunks 8l..8l+7 (4 codes
    // each); chunk j is stored at 8l + (j ^ (l & 7)); store instruction t reads logical chunk 32t + l.
    if (UNPACKED) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        stg_w[j] = stg0 + (uint32_t)((8 * lane + (j ^ (lane & 7))) << 4);
        const int o = 4 * j + (lane >> 3);                       // owner lane of logical chunk 32j + lane
        stg_r[j] = stg0 + (uint32_t)((8 * o + ((lane & 7) ^ (o & 7))) << 4);
      }
    }
    pq.prepare();
  }

  // wds: the thread's 32 elements.  valid: they exist.  *_dst: where this thread's results of this tile go
  // (qu_dst: the lane's first coalesced 16-byte chunk of the warp tile, qu_left: int32 codes from the start of the
  // warp tile to the end of the tensor).
  __device__ __forceinline__ void tile(const uint32_t (&wds)[4 * NLD], uint32_t cs_thr, bool valid, uint8_t* q_dst,
                                       uint8_t* s_dst, uint8_t* z_dst, uint8_t* zq_dst, uint8_t* qu_dst, int64_t qu_left,
                                       bool has_qp, bool has_zp, bool has_zq) {
    // ---- group min / max: packed tree over 16 words, then fold halves, then LPG lanes --------
    float mn, mx;
    float2 xv[(CS || F32IN) ? 16 : 1];
    if (CS || F32IN) {
      // CS: x = fp32(w) * s[k]; fp32 input: x = w.  The group statistics are taken on x (3-input FMNMX)
      if (CS) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {           // chunk ch = 2c + h: scales of words 4c + 2h, 4c + 2h + 1
          float4 sv;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w)
                       : "r"(cs_thr + (uint32_t)ch * 512u));
          xv[2 * ch] = __fmul2_rn(Packed<InT>::to_f2(wds[(2 * ch) % (4 * NLD)]), make_float2(sv.x, sv.y));
          xv[2 * ch + 1] = __fmul2_rn(Packed<InT>::to_f2(wds[(2 * ch + 1) % (4 * NLD)]), make_float2(sv.z, sv.w));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          xv[i] = make_float2(__uint_as_float(wds[(2 * i) % (4 * NLD)]), __uint_as_float(wds[(2 * i + 1) % (4 * NLD)]));
      }
      // 16 three-input operations per statistic, as four independent chains of depth 3 (one per LDS chunk of W:
      // a chain starts as soon as its chunk is there) and a merge of depth 4 -- not one chain of depth 16
      float pmn[4], pmx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pmn[c] = fmin3_nan(xv[4 * c].x, xv[4 * c].y, xv[4 * c + 1].x);
        pmx[c] = fmax3_nan(xv[4 * c].x, xv[4 * c].y, xv[4 * c + 1].x);
        pmn[c] = fmin3_nan(pmn[c], xv[4 * c + 1].y, xv[4 * c + 2].x);
        pmx[c] = fmax3_nan(pmx[c], xv[4 * c + 1].y, xv[4 * c + 2].x);
        pmn[c] = fmin3_nan(pmn[c], xv[4 * c + 2].y, xv[4 * c + 3].x);
        pmx[c] = fmax3_nan(pmx[c], xv[4 * c + 2].y, xv[4 * c + 3].x);
      }
      mn = fmin3_nan(pmn[0], pmn[1], pmn[2]);
      mx = fmax3_nan(pmx[0], pmx[1], pmx[2]);
      mn = fmin3_nan(mn, pmn[3], xv[3].y);
      mx = fmax3_nan(mx, pmx[3], xv[3].y);
      mn = fmin3_nan(mn, xv[7].y, xv[11].y);
      mx = fmax3_nan(mx, xv[7].y, xv[11].y);
      mn = v2_fmin_nan(mn, xv[15].y);
      mx = v2_fmax_nan(mx, xv[15].y);
    } else {
      uint32_t mn2 = wds[0], mx2 = wds[0];
#pragma unroll
      for (int i = 1; i < 16; ++i) {
        mn2 = Packed<InT>::min2(mn2, wds[i]);
        mx2 = Packed<InT>::max2(mx2, wds[i]);
      }
      const float2 mnf = Packed<InT>::to_f2(mn2), mxf = Packed<InT>::to_f2(mx2);
      mn = v2_fmin_nan(mnf.x, mnf.y);
      mx = v2_fmax_nan(mxf.x, mxf.y);
    }
#pragma unroll
    for (int m = 1; m < LPG; m <<= 1) {
      mn = v2_fmin_nan(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, m));
      mx = v2_fmax_nan(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, m));
    }

    const FastGroup fg = group_params_fast<A, BITS>(mn, mx, SYM, FQMIN, FQMAX);
    float sc = fg.scale;
    int zi;                          // the zero point as the reference's int32 (NaN -> INT32_MIN: slow path only)
    uint32_t words[NW];              // BITS = 4: slot wi -> words[wi]; BITS = 8: slot wi -> words[2 wi], words[2 wi + 1]
    uint32_t nanmask = 0;            // UNPACKED: bit e <=> code of logical element e is NaN (INT32_MIN)
    // tighter than group_params_fast: the packed 16-bit clamp needs round(x/s + zp) inside the
    // range where the magic-constant add is linear and the s16 rebase cannot wrap
    //   fp32 : |x|/s < 2^14 (low half of the fp32 magic sum is an s16)
    //   bf16 : b + 192 must stay in [128, 256)  -> |x|/s < 48
    //   fp16 : b + 1536 must stay in [1024, 2048) -> |x|/s < 400
    //   8 bit: every mode rounds through the fp32 magic -> 2^14
    constexpr float LIM = (A == AR_F32 || BITS == 8) ? 16384.0f : ((A == AR_BF16) ? 48.0f : 400.0f);
    const bool fast = fg.ok && (fmaxf(fabsf(mn), fabsf(mx)) < sc * LIM);
    // asymmetric int4 in fp32 arithmetic: when no group of the warp needs the final clamp (the normal case: every
    // group straddles zero) the 16 VIMNMX of the tile are skipped -- the ALU pipe is the busiest one of this kernel
    constexpr bool NOCLAMP_OK = !SYM && A == AR_F32 && BITS == 4;
    const bool noclamp = NOCLAMP_OK && __all_sync(0xFFFFFFFFu, fast && fg.noclamp);
    if (fast) {
      pq.init(fg.zp);
      zi = __float2int_rz(fg.zp);
      const float2 r2 = make_float2(fg.rcp, fg.rcp);
      const float2 ns2 = make_float2(-sc, -sc);
      auto quotients = [&](int wi, float2 (&qv)[4]) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 x = (CS || F32IN) ? xv[(CS || F32IN) ? 4 * wi + p : 0] : Packed<InT>::to_f2(wds[4 * wi + p]);
          const float2 q0 = __fmul2_rn(x, r2);
          const float2 e = __ffma2_rn(ns2, q0, x);
          qv[p] = __ffma2_rn(e, r2, q0);                          // correctly rounded x / s
        }
      };
      if (NOCLAMP_OK && noclamp) {
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          float2 qv[4];
          quotients(wi, qv);
          words[wi * (NW / 4)] = pq.pack8_noclamp(qv, half_sel);
        }
      } else {
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          float2 qv[4];
          quotients(wi, qv);
          if (BITS == 4) {
            words[wi * (NW / 4)] = pq.pack8(qv, half_sel);
          } else {                                                // int8: bytes 0 and 2 of each pair (u_lo | u_hi << 16)
            words[wi * (NW / 4)] = __byte_perm(pq.run(qv[0]), pq.run(qv[1]), 0x6420);
            words[wi * (NW / 4) + (NW / 4 - 1)] = __byte_perm(pq.run(qv[2]), pq.run(qv[3]), 0x6420);
          }
        }
      }
    } else {
      const GroupParams gp = group_params<A>(mn, mx, SYM, FQMIN, FQMAX);   // exact IEEE path
      sc = gp.scale;
      const float zp = gp.zp;
      zi = f2i_x86(zp);
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        uint32_t acc = 0, acc2 = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 x = (CS || F32IN) ? xv[(CS || F32IN) ? 4 * wi + p : 0] : Packed<InT>::to_f2(wds[4 * wi + p]);
          const int c0 = quant_exact<A>(x.x, sc, zp, FQMIN, FQMAX);
          const int c1 = quant_exact<A>(x.y, sc, zp, FQMIN, FQMAX);
          const uint32_t u0 = (c0 == INT32_MIN) ? 0u : (uint32_t)(c0 - QMIN);
          const uint32_t u1 = (c1 == INT32_MIN) ? 0u : (uint32_t)(c1 - QMIN);
          if (BITS == 4) {
            acc |= (u0 & 15u) << (8 * (p ^ odd_shift));
            acc |= (u1 & 15u) << (8 * (p ^ odd_shift) + 4);
          } else if (p < 2) {
            acc |= ((u0 & 255u) | ((u1 & 255u) << 8)) << (16 * p);
          } else {
            acc2 |= ((u0 & 255u) | ((u1 & 255u) << 8)) << (16 * (p - 2));
          }
          if (UNPACKED) {            // slot wi holds logical 16-byte chunk (wi + rot) & 3  (fp32: word wi ^ rot)
            const int e = F32IN ? 8 * (wi ^ rot) + 2 * (p ^ odd_shift) : 8 * ((wi + rot) & 3) + 2 * p;
            nanmask |= (c0 == INT32_MIN ? 1u : 0u) << e;
            nanmask |= (c1 == INT32_MIN ? 1u : 0u) << (e + 1);
          }
        }
        words[wi * (NW / 4)] = acc;
        if (BITS == 8) words[wi * (NW / 4) + (NW / 4 - 1)] = acc2;
      }
    }

    // ---- stores ------------------------------------------------------------------------------
    // slot c holds chunk (c + rot) & 3  ->  chunk k sits in slot (k - rot) & 3: rotate left by rot
    // (a slot is one word for int4, two for int8: ow[h][k] = word h of chunk k)
    uint32_t ow[NW / 4][4];
#pragma unroll
    for (int h = 0; h < NW / 4; ++h) {
      uint32_t o0 = words[h], o1 = words[(NW / 4) + h], o2 = words[2 * (NW / 4) + h], o3 = words[3 * (NW / 4) + h];
      if (F32IN) {                 // word slot wp holds word wp ^ rot
        if (rot & 1) { uint32_t t = o0; o0 = o1; o1 = t; t = o2; o2 = o3; o3 = t; }
      } else {
        if (rot & 1) { const uint32_t t = o3; o3 = o2; o2 = o1; o1 = o0; o0 = t; }
      }
      if (rot & 2) { uint32_t t = o0; o0 = o2; o2 = t; t = o1; o1 = o3; o3 = t; }
      ow[h][0] = o0; ow[h][1] = o1; ow[h][2] = o2; ow[h][3] = o3;
    }
    if (valid && (!UNPACKED || has_qp)) {
      if (BITS == 4) {
        st_stream16(q_dst, make_uint4(ow[0][0], ow[0][1], ow[0][2], ow[0][3]));
      } else {
        st_stream16(q_dst, make_uint4(ow[0][0], ow[NW / 4 - 1][0], ow[0][1], ow[NW / 4 - 1][1]));
        st_stream16(q_dst + 16, make_uint4(ow[0][2], ow[NW / 4 - 1][2], ow[0][3], ow[NW / 4 - 1][3]));
      }
    }
    if (UNPACKED) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int c[8];
        if (BITS == 4) {
          // 8 nibbles -> 8 int32 codes: even nibbles / odd nibbles to bytes, then one PRMT per code
          const uint32_t ev = ow[0][k] & 0x0F0F0F0Fu, od = (ow[0][k] >> 4) & 0x0F0F0F0Fu;
          c[0] = (int)__byte_perm(ev, 0u, 0x4440) + QMIN; c[1] = (int)__byte_perm(od, 0u, 0x4440) + QMIN;
          c[2] = (int)__byte_perm(ev, 0u, 0x4441) + QMIN; c[3] = (int)__byte_perm(od, 0u, 0x4441) + QMIN;
          c[4] = (int)__byte_perm(ev, 0u, 0x4442) + QMIN; c[5] = (int)__byte_perm(od, 0u, 0x4442) + QMIN;
          c[6] = (int)__byte_perm(ev, 0u, 0x4443) + QMIN; c[7] = (int)__byte_perm(od, 0u, 0x4443) + QMIN;
        } else {
          const uint32_t w0 = ow[0][k], w1 = ow[NW / 4 - 1][k];
          c[0] = (int)__byte_perm(w0, 0u, 0x4440) + QMIN; c[1] = (int)__byte_perm(w0, 0u, 0x4441) + QMIN;
          c[2] = (int)__byte_perm(w0, 0u, 0x4442) + QMIN; c[3] = (int)__byte_perm(w0, 0u, 0x4443) + QMIN;
          c[4] = (int)__byte_perm(w1, 0u, 0x4440) + QMIN; c[5] = (int)__byte_perm(w1, 0u, 0x4441) + QMIN;
          c[6] = (int)__byte_perm(w1, 0u, 0x4442) + QMIN; c[7] = (int)__byte_perm(w1, 0u, 0x4443) + QMIN;
        }
        if (nanmask != 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if ((nanmask >> (8 * k + i)) & 1u) c[i] = INT32_MIN;
        }
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg_w[2 * k]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg_w[2 * k + 1]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]) : "memory");
      }
      __syncwarp();
      // whole 16-byte chunks are valid or not: the tensor ends on a group (>= 32 element) boundary
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg_r[t]));
        if (128 * t + 4 * lane < qu_left) st_stream16(qu_dst + 512 * t, v);
      }
      __syncwarp();                                    // staging is reused by the next tile
    }
    if (valid && leader) {
      *reinterpret_cast<__half*>(s_dst) = __float2half_rn(sc);
      if (has_zp) *reinterpret_cast<int32_t*>(z_dst) = zi;
    }
    if (has_zq) {
      // (a NaN zero point -- INT32_MIN, asymmetric only, QMIN = 0 -- packs as 0 by itself: its low bits are 0)
      const uint32_t uz = valid ? ((uint32_t)(zi - QMIN) & zq_lane_mask) : 0u;
      // (a compile-time full mask lets the compiler drop the re-convergence sequence around REDUX: the common
      // g = 128 / int4 case packs one word per warp)
      const uint32_t wordz = (LPW == 32) ? __reduce_or_sync(0xFFFFFFFFu, uz << zq_shift)
                                         : __reduce_or_sync(zq_mask, uz << zq_shift);
      if (valid && zq_writer) *reinterpret_cast<uint32_t*>(zq_dst) = wordz;
    }
  }
};

// ============================================================================================================
// flat mode: the tensor (or a whole arena of tensors) is one run of contiguous groups
// ============================================================================================================
template <typename InT, int A, int G, bool SYM, bool UNPACKED, int BITS>
__global__ void __launch_bounds__(kV2Threads, (UNPACKED || sizeof(InT) == 4) ? 2 : 3)
group_quant_tma(const InT* __restrict__ w, int64_t n_elems, int64_t n_tiles, V2Out out) {
  static_assert(BITS == 4 || BITS == 8, "int4 or int8 codes");
  // fp32 input: a thread's 32 elements are 128 B = 8 LDS.128; stages are 32 KiB, so 3 of them (2 with the
  // UNPACKED staging) keep 2 CTAs per SM.  Chunk c of the thread's span is read into slot c ^ (lane & 7)
  // (conflict-free: a quarter warp covers all 8 bank groups); slots 2w, 2w+1 still hold the two halves of one
  // packed word, swapped for odd lanes.
  using Cons = V2Consumer<InT, A, G, SYM, UNPACKED, false, BITS>;
  constexpr bool F32IN = Cons::F32IN;
  static_assert(!F32IN || (A == AR_F32 && BITS == 4), "fp32 input: fp32 arithmetic, int4");
  constexpr int STAGES = F32IN ? (UNPACKED ? 2 : 3) : kV2Stages;
  constexpr uint32_t STAGE_BYTES = kV2CtaTile * sizeof(InT);
  constexpr int NLD = Cons::NLD;

  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + (UNPACKED ? kV2UnpStageBytes : 0));
  uint32_t full0 = smem_u32(bars);
  uint32_t empty0 = smem_u32(bars + STAGES);
  asm volatile("" : "+r"(full0), "+r"(empty0));   // keep the shared-window addresses in registers

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, kV2ConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // CTA b takes tiles b, b + grid, ...
  const int64_t tile0 = (int64_t)blockIdx.x;
  const int64_t tile_step = (int64_t)gridDim.x;
  const int64_t n_iters = (n_tiles > tile0) ? (n_tiles - tile0 + tile_step - 1) / tile_step : 0;

  if (warp == kV2ConsumerWarps) {
    // ================= producer: one thread streams CTA tiles into the ring =================
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(w) + tile0 * STAGE_BYTES;
      const int64_t src_stride = (int64_t)gridDim.x * STAGE_BYTES;
      int64_t left = n_elems * (int64_t)sizeof(InT) - tile0 * STAGE_BYTES;          // bytes from this tile to the end
      uint32_t stage = 0, ph = 1;                                   // first pass over the ring: slots are free
      for (int64_t it = 0; it < n_iters; ++it) {
        mbar_wait(empty0 + 8 * stage, ph);
        const uint32_t bytes = (left < STAGE_BYTES) ? (uint32_t)left : (uint32_t)STAGE_BYTES;
        mbar_expect_tx(full0 + 8 * stage, bytes);
        bulk_g2s(smem_u32(smem) + stage * STAGE_BYTES, src, bytes, full0 + 8 * stage);
        src += src_stride;
        left -= src_stride;
        if (++stage == STAGES) { stage = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ================================ consumers =================================================
  // per-thread invariants.  Output addresses are (thread base) + it * (byte stride): one IMAD.WIDE
  // per store, no loop-carried 64-bit pointers.
  const int64_t warp_e_first = tile0 * kV2CtaTile + (int64_t)warp * kV2WarpTile;
  const int64_t e_first = warp_e_first + (int64_t)lane * 32;
  const int64_t e_stride = tile_step * kV2CtaTile;
  const uint32_t iters = (uint32_t)n_iters;
  // iterations in which this thread's 32 elements exist (whole groups are valid or not)
  const uint32_t valid_iters = (n_elems > e_first) ? (uint32_t)((n_elems - e_first + e_stride - 1) / e_stride) : 0u;
  const uint32_t smem_thr = smem_u32(smem) + (uint32_t)warp * (kV2WarpTile * (uint32_t)sizeof(InT)) +
                            (uint32_t)lane * (32u * (uint32_t)sizeof(InT));
  Cons cons;
  cons.init(lane, out.zq_log2, smem_u32(smem) + STAGES * STAGE_BYTES + (uint32_t)warp * (kV2WarpTile * 4));
  uint32_t ld_off[NLD];
#pragma unroll
  for (int c = 0; c < NLD; ++c) {
    ld_off[c] = smem_thr + (uint32_t)((F32IN ? (c ^ (lane & 7)) : ((c + cons.rot) & 3)) << 4);
    asm volatile("" : "+r"(ld_off[c]));             // (no per-tile re-derivation from SR_CgaCtaId)
  }
  uint8_t* const q_base = reinterpret_cast<uint8_t*>(out.q_packed) + e_first * BITS / 8;
  uint8_t* const s_base = reinterpret_cast<uint8_t*>(out.scales + e_first / G);
  uint8_t* const z_base = reinterpret_cast<uint8_t*>(out.zp + e_first / G);
  uint8_t* const zq_base = reinterpret_cast<uint8_t*>(out.zp_packed + ((e_first / G) >> out.zq_log2));
  uint8_t* const qu_base = UNPACKED ? reinterpret_cast<uint8_t*>(out.q_unpacked + warp_e_first + 4 * lane) : nullptr;
  const uint32_t q_step = (uint32_t)(e_stride * BITS / 8);         // bytes per iteration (< 2^32: grid <= 3*SMs)
  const uint32_t s_step = (uint32_t)(e_stride / G) * 2u;
  const uint32_t z_step = (uint32_t)(e_stride / G) * 4u;
  const uint32_t zq_step = (uint32_t)((e_stride / G) >> out.zq_log2) * 4u;
  const uint32_t qu_step = (uint32_t)e_stride * 4u;              // bytes per iteration (host keeps it < 2^32)
  const bool has_zp = out.zp != nullptr, has_zq = out.zp_packed != nullptr, has_qp = out.q_packed != nullptr;

  uint32_t stage = 0, ph = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    const bool valid = it < valid_iters;
    mbar_wait(full0 + 8 * stage, ph);
    // 64 B per thread as 4 x LDS.128.  Register slot c holds 16-byte chunk (c + rot) & 3 of the
    // thread's span: rotating the chunk order by lane/2 makes every quarter-warp hit 8 distinct
    // bank groups.  Min/max and the per-word quantization are order independent; only the final
    // 16-byte store has to rotate the 4 result words back (8 SELs).
    uint32_t wds[4 * NLD];
#pragma unroll
    for (int c = 0; c < NLD; ++c) {
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(wds[4 * c]), "=r"(wds[4 * c + 1]), "=r"(wds[4 * c + 2]), "=r"(wds[4 * c + 3])
                   : "r"(ld_off[c] + stage * STAGE_BYTES));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * stage);   // slot is free as soon as it sits in registers
    if (++stage == STAGES) { stage = 0; ph ^= 1u; }
    // (threads past the end of the tensor compute on stale shared memory and store nothing)
    cons.tile(wds, 0u, valid, q_base + (uint64_t)it * q_step, s_base + (uint64_t)it * s_step, z_base + (uint64_t)it * z_step,
              zq_base + (uint64_t)it * zq_step, UNPACKED ? qu_base + (uint64_t)it * qu_step : nullptr,
              UNPACKED ? n_elems - (warp_e_first + (int64_t)it * e_stride) : 0, has_qp, has_zp, has_zq);
  }
}

// ============================================================================================================
// CS ("column scaled", the final AWQ pass: quantize fp32(w) * s[k], awqk.h col_scale), one launch for a BATCH of
// tensors.  Every tensor [C, K] (K % 1024 == 0) is cut into column slabs of 1024 (one warp tile wide) and row chunks
// of R rows; a UNIT is one row chunk of one slab, streamed 8 rows per stage (8 bulk copies of 2 KiB, one per consumer
// warp), so that a thread's 32 columns -- and their scales -- stay the same for the whole unit.  Units are numbered
// (tensor, row chunk, slab) with the slab fastest and dealt round-robin to the persistent CTAs: the CTAs that run at
// the same time sweep the same rows of neighbouring slabs, which keeps the HBM pages of a row open.  The producer
// warp also brings in the unit's 1024 column scales: all 32 lanes load them (8 x 16 B each, prefetched one unit
// ahead) and store them transposed -- [chunk][lane] -- into one of two 4 KiB tables, handed over with a pair of
// mbarriers per table.  Output offsets are (unit base) + tile * (constant stride) in the flat row-major arrays,
// exactly as in the flat mode.
// ============================================================================================================
constexpr int kV2MaxBatch = 32;      // items of a launch (a tensor is one item, or two when its rows are split)
constexpr int kV2MaxTensors = 31;
struct V2BatchItem {
  const void* w;
  const float* s;       // [K]
  uint32_t* q_packed;   // nullable when q_unpacked is given
  int32_t* q_unpacked;  // nullable (UNPACKED instantiations only)
  __half* scales;
  int32_t* zp;          // nullable
  uint32_t* zp_packed;  // nullable
  int32_t C, K;
  uint32_t unit_begin, unit_end;   // this item's units: [unit_begin, unit_end)
  uint32_t n_slabs;                // K / 1024
  uint32_t magic;                  // floor(2^32 / n_slabs) + 1 (n_slabs > 1): unit / n_slabs = umulhi(unit, magic)
  uint32_t rows_per_unit;          // R, a multiple of 8 (tall units first, short ones for the tail of the launch)
  uint32_t pad_;
};
struct V2Batch {
  V2BatchItem it[kV2MaxBatch];
  uint32_t n_units;
  uint32_t has_zp, has_zq, has_qp;   // the same for every item of a launch
};

struct V2Unit {
  uint32_t slab, row0, rows;         // rows: the unit's height (R, or what is left of the item)
};
__device__ __forceinline__ V2Unit v2_locate(const V2Batch& b, uint32_t u, uint32_t& t) {
  while (u >= b.it[t].unit_end) ++t;                 // units are visited in increasing order
  const uint32_t local = u - b.it[t].unit_begin;
  const uint32_t ns = b.it[t].n_slabs;
  const uint32_t R = b.it[t].rows_per_unit;
  const uint32_t rc = (ns == 1) ? local : __umulhi(local, b.it[t].magic);
  const uint32_t row0 = rc * R;
  const uint32_t left = (uint32_t)b.it[t].C - row0;
  return V2Unit{local - rc * ns, row0, left < R ? left : R};
}

template <typename InT, int G, bool SYM, bool UNPACKED>
__global__ void __launch_bounds__(kV2Threads, UNPACKED ? 2 : 3)
group_quant_tma_cs(const __grid_constant__ V2Batch b) {
  using Cons = V2Consumer<InT, AR_F32, G, SYM, UNPACKED, true, 4>;
  static_assert(sizeof(InT) == 2, "column-slab mode: bf16 / fp16 weights");
  constexpr int STAGES = kV2Stages;
  constexpr uint32_t STAGE_BYTES = kV2CtaTile * 2;
  constexpr uint32_t ROW_BYTES = kV2WarpTile * 2;          // one warp tile = one row of the slab
  constexpr uint32_t TABLE_BYTES = kV2WarpTile * 4;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t table0 = smem_u32(smem) + STAGES * STAGE_BYTES + (UNPACKED ? kV2UnpStageBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + (UNPACKED ? kV2UnpStageBytes : 0) + 2 * TABLE_BYTES);
  uint32_t full0 = smem_u32(bars);
  uint32_t empty0 = smem_u32(bars + STAGES);
  uint32_t tfull0 = smem_u32(bars + 2 * STAGES);
  uint32_t tempty0 = smem_u32(bars + 2 * STAGES + 2);
  asm volatile("" : "+r"(full0), "+r"(empty0), "+r"(tfull0), "+r"(tempty0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, kV2ConsumerWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, kV2ConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t u = blockIdx.x;
  if (u >= b.n_units) return;

  if (warp == kV2ConsumerWarps) {
    // ================= producer warp: scale tables (all lanes) + W stages (lane 0) =================
    const int rot0 = (lane >> 1) & 3;
    float4 tab[8];
    // lane l: its 32 scales of the slab, in the rotated chunk order of the W loads (slot c <- chunk (c + rot) & 3)
    auto load_table = [&](const float* s, uint32_t slab) {
      const float* sp0 = s + (size_t)slab * kV2WarpTile + lane * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int h = 0; h < 2; ++h) tab[2 * c + h] = __ldg(reinterpret_cast<const float4*>(sp0 + 8 * ((c + rot0) & 3) + 4 * h));
      }
    };
    auto store_table = [&](uint32_t slot) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(table0 + slot * TABLE_BYTES + (uint32_t)j * 512u + (uint32_t)lane * 16u),
                     "f"(tab[j].x), "f"(tab[j].y), "f"(tab[j].z), "f"(tab[j].w)
                     : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tfull0 + 8 * slot);
    };
    uint32_t t = 0, tn = 0;                      // item cursors of the current / the next unit
    V2Unit cur = v2_locate(b, u, t);
    load_table(b.it[t].s, cur.slab);
    store_table(0);
    uint32_t stage = 0, ph = 1;                  // first pass over the ring: slots are free
    for (uint32_t i = 0;; ++i) {
      const uint32_t un = u + gridDim.x;
      const bool has_next = un < b.n_units;
      V2Unit nxt{0, 0, 0};
      if (has_next) {
        tn = t;
        nxt = v2_locate(b, un, tn);
        load_table(b.it[tn].s, nxt.slab);        // in flight while lane 0 issues this unit's stages
      }
      if (lane == 0) {
        const int64_t row_bytes = (int64_t)b.it[t].K * 2;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(b.it[t].w) + (int64_t)cur.row0 * row_bytes + (int64_t)cur.slab * ROW_BYTES;
        for (int rows_left = (int)cur.rows; rows_left > 0; rows_left -= kV2ConsumerWarps) {
          mbar_wait(empty0 + 8 * stage, ph);
          const int rows = rows_left < kV2ConsumerWarps ? rows_left : kV2ConsumerWarps;
          mbar_expect_tx(full0 + 8 * stage, (uint32_t)rows * ROW_BYTES);
          const uint32_t dst = smem_u32(smem) + stage * STAGE_BYTES;
          for (int r = 0; r < rows; ++r) bulk_g2s(dst + r * ROW_BYTES, src + r * row_bytes, ROW_BYTES, full0 + 8 * stage);
          src += kV2ConsumerWarps * row_bytes;
          if (++stage == STAGES) { stage = 0; ph ^= 1u; }
        }
      }
      __syncwarp();
      if (!has_next) break;
      const uint32_t slot = (i + 1) & 1u;
      mbar_wait(tempty0 + 8 * slot, (((i + 1) >> 1) & 1u) ^ 1u);   // the consumers are done with this table's previous unit
      store_table(slot);
      u = un;
      t = tn;
      cur = nxt;
    }
    return;
  }

  // ================================ consumers =================================================
  const uint32_t smem_thr = smem_u32(smem) + (uint32_t)warp * ROW_BYTES + (uint32_t)lane * 64u;
  Cons cons;
  cons.init(lane, 3, smem_u32(smem) + STAGES * STAGE_BYTES + (uint32_t)warp * (kV2WarpTile * 4));
  uint32_t ld_off[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ld_off[c] = smem_thr + (uint32_t)(((c + cons.rot) & 3) << 4);
    asm volatile("" : "+r"(ld_off[c]));
  }
  const uint32_t cs_lane = table0 + (uint32_t)lane * 16u;

  const bool has_zp = b.has_zp != 0, has_zq = b.has_zq != 0, has_qp = b.has_qp != 0;
  uint32_t t = 0, stage = 0, ph = 0;
  for (uint32_t i = 0; u < b.n_units; u += gridDim.x, ++i) {
    const V2Unit un = v2_locate(b, u, t);
    const V2BatchItem& item = b.it[t];
    const int64_t K = item.K;
    const int row = (int)un.row0 + warp;                        // this warp's row in the unit's first tile
    const int64_t warp_e_first = (int64_t)row * K + (int64_t)un.slab * kV2WarpTile;
    const int64_t e_first = warp_e_first + (int64_t)lane * 32;
    uint8_t* const q_base = reinterpret_cast<uint8_t*>(item.q_packed) + e_first / 2;
    uint8_t* const s_base = reinterpret_cast<uint8_t*>(item.scales + e_first / G);
    uint8_t* const z_base = reinterpret_cast<uint8_t*>(item.zp + e_first / G);
    uint8_t* const zq_base = reinterpret_cast<uint8_t*>(item.zp_packed + ((e_first / G) >> 3));
    uint8_t* const qu_base = UNPACKED ? reinterpret_cast<uint8_t*>(item.q_unpacked + warp_e_first + 4 * lane) : nullptr;
    // tile j (8 rows further down): q + j * 4K bytes, scales + j * 16 K/G, zp + j * 32 K/G, zp_packed + j * 4 K/G,
    // int32 codes + j * 32K -- one IMAD.WIDE per store
    const uint32_t kq = (uint32_t)K * 4u, kg = (uint32_t)(K / G) * 4u;
    const uint32_t slot = i & 1u;
    const uint32_t cs_thr = cs_lane + slot * TABLE_BYTES;
    const uint32_t n_tiles = (un.rows + kV2ConsumerWarps - 1) / kV2ConsumerWarps;
    // tiles in which this warp's row exists (the warps past the last row of a ragged unit only keep the ring going)
    const uint32_t my_tiles = ((int)un.rows > warp) ? (un.rows - (uint32_t)warp + kV2ConsumerWarps - 1) / kV2ConsumerWarps : 0u;
    mbar_wait(tfull0 + 8 * slot, (i >> 1) & 1u);
    for (uint32_t j = 0; j < n_tiles; ++j) {
      mbar_wait(full0 + 8 * stage, ph);
      uint32_t wds[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(wds[4 * c]), "=r"(wds[4 * c + 1]), "=r"(wds[4 * c + 2]), "=r"(wds[4 * c + 3])
                     : "r"(ld_off[c] + stage * STAGE_BYTES));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);   // slot is free as soon as it sits in registers
      if (++stage == STAGES) { stage = 0; ph ^= 1u; }
      if (j < my_tiles)                                 // (warp-uniform)
        cons.tile(wds, cs_thr, true, q_base + (uint64_t)j * kq, s_base + (uint64_t)j * (kg * 4u), z_base + (uint64_t)j * (kg * 8u),
                  zq_base + (uint64_t)j * kg, UNPACKED ? qu_base + (uint64_t)j * kq * 8u : nullptr, (int64_t)kV2WarpTile,
                  has_qp, has_zp, has_zq);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty0 + 8 * slot);     // this warp no longer reads the table
  }
}

static int v2_sms(int* sms) {
  int dev = 0;
  AWQK_CUDA(cudaGetDevice(&dev));
  AWQK_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return dev;
}

template <typename InT, int A, int G, bool UNPACKED, int BITS>
static int launch_v2_sym(const InT* w, int64_t n, bool sym, V2Out out, cudaStream_t st) {
  int sms = 0;
  const int dev = v2_sms(&sms);
  if (dev < 0) return dev;
  const int64_t want = (int64_t)sms * ((UNPACKED || sizeof(InT) == 4) ? 2 : 3);   // resident CTAs per SM (smem / register bound)
  const int64_t n_tiles = ceil_div(n, kV2CtaTile);
  const unsigned grid = (unsigned)(n_tiles < want ? n_tiles : want);
  constexpr bool F32IN = sizeof(InT) == 4;                         // (same constants as in the kernel)
  constexpr int STAGES = F32IN ? (UNPACKED ? 2 : 3) : kV2Stages;
  const size_t smem = (size_t)STAGES * kV2CtaTile * sizeof(InT) + (UNPACKED ? kV2UnpStageBytes : 0) + 2 * STAGES * sizeof(uint64_t);
  // the dynamic-smem opt-in is per (kernel instantiation, device): set once, then immutable
  static std::atomic<uint64_t> configured[2] = {{0}, {0}};
  const uint64_t bit = 1ull << (dev & 63);
  const bool need = !(configured[sym ? 1 : 0].load(std::memory_order_acquire) & bit);
  if (sym) {
    auto k = group_quant_tma<InT, A, G, true, UNPACKED, BITS>;
    if (need) AWQK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kV2Threads, smem, st>>>(w, n, n_tiles, out);
  } else {
    auto k = group_quant_tma<InT, A, G, false, UNPACKED, BITS>;
    if (need) AWQK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kV2Threads, smem, st>>>(w, n, n_tiles, out);
  }
  if (need) configured[sym ? 1 : 0].fetch_or(bit, std::memory_order_release);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

template <typename InT, int A, int BITS>
static int launch_v2_g(const InT* w, int64_t n, int g, bool sym, V2Out out, cudaStream_t st) {
  if (out.q_unpacked != nullptr) {
    switch (g) {
      case 32: return launch_v2_sym<InT, A, 32, true, BITS>(w, n, sym, out, st);
      case 64: return launch_v2_sym<InT, A, 64, true, BITS>(w, n, sym, out, st);
      default: return launch_v2_sym<InT, A, 128, true, BITS>(w, n, sym, out, st);
    }
  }
  switch (g) {
    case 32: return launch_v2_sym<InT, A, 32, false, BITS>(w, n, sym, out, st);
    case 64: return launch_v2_sym<InT, A, 64, false, BITS>(w, n, sym, out, st);
    default: return launch_v2_sym<InT, A, 128, false, BITS>(w, n, sym, out, st);
  }
}

template <typename InT, int A>
static int launch_v2_bits(const InT* w, int64_t n, int g, int bits, bool sym, V2Out out, cudaStream_t st) {
  return bits == 8 ? launch_v2_g<InT, A, 8>(w, n, g, sym, out, st) : launch_v2_g<InT, A, 4>(w, n, g, sym, out, st);
}

// Entry used by awqk_group_quant: int4 / int8, bf16/fp16 input, flat layout (K % g == 0, g in {32,64,128},
// 16-byte aligned base).  zp_packed: zq_log2 = log2(32 / bits) when a row is a whole number of packed words,
// smaller when a row is 1 / 2 / 4 groups (one zero-padded word per row); otherwise it must be null.
int launch_group_quant_tma(const void* w, int dtype, int64_t n_elems, int g, int bits, bool sym, int arith,
                           uint32_t* q_packed, int32_t* q_unpacked, void* scales, int32_t* zp, uint32_t* zp_packed,
                           int zq_log2, cudaStream_t st) {
  V2Out out{q_packed, q_unpacked, reinterpret_cast<__half*>(scales), zp, zp_packed, zq_log2};
  if (dtype == AWQK_BF16) {
    auto p = reinterpret_cast<const __nv_bfloat16*>(w);
    return arith == AWQK_ARITH_FP32 ? launch_v2_bits<__nv_bfloat16, AR_F32>(p, n_elems, g, bits, sym, out, st)
                                    : launch_v2_bits<__nv_bfloat16, AR_BF16>(p, n_elems, g, bits, sym, out, st);
  }
  if (dtype == AWQK_FP32)             // int4 only (the dispatcher keeps 8-bit fp32 input on the register path)
    return launch_v2_g<float, AR_F32, 4>(reinterpret_cast<const float*>(w), n_elems, g, sym, out, st);
  auto p = reinterpret_cast<const __half*>(w);
  return arith == AWQK_ARITH_FP32 ? launch_v2_bits<__half, AR_F32>(p, n_elems, g, bits, sym, out, st)
                                  : launch_v2_bits<__half, AR_F16>(p, n_elems, g, bits, sym, out, st);
}

// ---- column-slab batches -----------------------------------------------------------------------------------
// K % 1024 == 0, and the per-tile output strides (8 rows) below 2^32 bytes; anything else stays on the register path.
bool group_quant_tma_cs_eligible(int64_t C, int64_t K) {
  if (!(C > 0 && C < ((int64_t)1 << 31) && K > 0 && K % kV2WarpTile == 0 && K <= ((int64_t)1 << 21))) return false;
  // unit -> (row chunk, slab) uses a multiply-high division that is exact while (units of the tensor) x slabs < 2^32;
  // the shortest units are 8 rows
  const int64_t slabs = K / kV2WarpTile;
  return ceil_div(C, 8) * slabs < ((int64_t)1 << 32) / slabs;
}

template <typename InT, int G, bool UNPACKED>
static int launch_cs_sym(const V2Batch& b, unsigned grid, bool sym, int dev, cudaStream_t st) {
  const size_t smem = (size_t)kV2Stages * kV2CtaTile * 2 + (UNPACKED ? kV2UnpStageBytes : 0) + 2 * kV2WarpTile * 4 +
                      (2 * kV2Stages + 4) * sizeof(uint64_t);
  static std::atomic<uint64_t> configured[2] = {{0}, {0}};
  const uint64_t bit = 1ull << (dev & 63);
  const bool need = !(configured[sym ? 1 : 0].load(std::memory_order_acquire) & bit);
  if (sym) {
    auto k = group_quant_tma_cs<InT, G, true, UNPACKED>;
    if (need) AWQK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kV2Threads, smem, st>>>(b);
  } else {
    auto k = group_quant_tma_cs<InT, G, false, UNPACKED>;
    if (need) AWQK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kV2Threads, smem, st>>>(b);
  }
  if (need) configured[sym ? 1 : 0].fetch_or(bit, std::memory_order_release);
  AWQK_CUDA(cudaGetLastError());
  return AWQK_OK;
}

template <typename InT, bool UNPACKED>
static int launch_cs_g(const V2Batch& b, unsigned grid, int g, bool sym, int dev, cudaStream_t st) {
  switch (g) {
    case 32: return launch_cs_sym<InT, 32, UNPACKED>(b, grid, sym, dev, st);
    case 64: return launch_cs_sym<InT, 64, UNPACKED>(b, grid, sym, dev, st);
    default: return launch_cs_sym<InT, 128, UNPACKED>(b, grid, sym, dev, st);
  }
}

// One launch over up to kV2MaxTensors eligible tensors (same dtype / group size / symmetry, and the same set of
// outputs).  Units are dealt round-robin, so the launch takes (units per CTA) x (unit height): tall units keep the
// hand-over cost (new table, new output bases: ~2 rows' worth) small, short units keep the last round short.  The
// plan therefore uses units of r_main rows for as many whole rounds over the persistent grid as there are, and
// units of r_tail rows for the rows that are left (the end of the tensor list); both heights are chosen per launch
// by the cost model  rounds_main * (r_main + 2) + rounds_tail * (r_tail + 2).
struct CsPlan {
  V2Batch b;
  int64_t cost;
  int r_main, r_tail;
  int n_items;
  int tensor_of[kV2MaxBatch];       // launch item -> index into the caller's tensor list
  int64_t row_begin[kV2MaxBatch];   // first row of the tensor the item covers
};

// r_tail == 0: one height for everything
static bool cs_plan(const awqk_quant_item* items, int n, int g, int64_t want, int r_main, int r_tail, CsPlan* plan) {
  V2Batch& b = plan->b;
  plan->r_main = r_main;
  plan->r_tail = r_tail;
  memset(&b, 0, sizeof(b));
  int64_t units_main_all = 0;
  for (int i = 0; i < n; ++i) units_main_all += (items[i].K / kV2WarpTile) * ceil_div(items[i].C, r_main);
  const int64_t main_target = r_tail ? (units_main_all / want) * want : units_main_all;
  if (r_tail && main_target == 0) return false;                 // less than one round: single-height plans cover it
  int n_it = 0;
  int64_t u = 0, units_main = 0;
  bool tail = false;
  for (int i = 0; i < n; ++i) {
    const awqk_quant_item& s = items[i];
    const int64_t slabs = s.K / kV2WarpTile;
    int64_t split = s.C;                                        // rows [0, split) in tall units, [split, C) in short ones
    if (tail) {
      split = 0;
    } else if (r_tail) {
      const int64_t chunks = ceil_div(s.C, r_main);
      if (units_main + slabs * chunks > main_target) {
        split = ((main_target - units_main) / slabs) * r_main;
        tail = true;
      }
    }
    for (int part = 0; part < 2; ++part) {
      const int64_t r0 = part ? split : 0, r1 = part ? s.C : split;
      if (r1 <= r0) continue;
      if (n_it == kV2MaxBatch) return false;
      const int r = part ? r_tail : r_main;
      const int64_t G = s.K / g;
      plan->tensor_of[n_it] = i;
      plan->row_begin[n_it] = r0;
      V2BatchItem& d = b.it[n_it++];
      d.w = reinterpret_cast<const uint8_t*>(s.w) + r0 * s.K * 2;
      d.s = s.col_scale;
      d.q_packed = s.q_packed ? s.q_packed + r0 * s.K / 8 : nullptr;
      d.q_unpacked = s.q_unpacked ? s.q_unpacked + r0 * s.K : nullptr;
      d.scales = reinterpret_cast<__half*>(s.scales_f16) + r0 * G;
      d.zp = s.zp ? s.zp + r0 * G : nullptr;
      d.zp_packed = s.zp_packed ? s.zp_packed + r0 * G / 8 : nullptr;
      d.C = (int32_t)(r1 - r0);
      d.K = (int32_t)s.K;
      d.n_slabs = (uint32_t)slabs;
      d.magic = slabs > 1 ? (uint32_t)(((uint64_t)1 << 32) / (uint64_t)slabs) + 1u : 0u;
      d.rows_per_unit = (uint32_t)r;
      d.unit_begin = (uint32_t)u;
      const int64_t nu = slabs * ceil_div(r1 - r0, r);
      u += nu;
      if (!part) units_main += nu;
      d.unit_end = (uint32_t)u;
      if (u >= ((int64_t)1 << 31)) return false;
    }
  }
  for (int i = n_it; i < kV2MaxBatch; ++i) b.it[i].unit_begin = b.it[i].unit_end = 0xFFFFFFFFu;   // sentinels: the item scan stops before
  b.n_units = (uint32_t)u;
  plan->n_items = n_it;
  const int64_t units_tail = u - units_main;
  plan->cost = ceil_div(units_main, want) * (r_main + 2) + (r_tail ? ceil_div(units_tail, want) * (r_tail + 2) : 0);
  return u > 0;
}

static bool cs_best_plan(const awqk_quant_item* items, int n, int g, int64_t want, CsPlan* best) {
  CsPlan cand;
  bool have = false;
  for (int r_main = 8; r_main <= 128; r_main *= 2) {
    for (int r_tail = 0; r_tail < r_main; r_tail = r_tail ? r_tail * 2 : 8) {
      if (!cs_plan(items, n, g, want, r_main, r_tail, &cand)) continue;
      if (!have || cand.cost < best->cost || (cand.cost == best->cost && r_tail == 0)) {
        *best = cand;
        have = true;
      }
    }
  }
  return have;
}

// introspection for tests (awqk_group_quant_batch_plan): no device needed
int cs_plan_describe(const int64_t* C, const int64_t* K, int n, int g, bool unpacked, int sms, int64_t* summary, int64_t* rows,
                     int max_items) {
  if (n <= 0 || n > kV2MaxTensors || sms <= 0) return AWQK_E_BADARG;
  awqk_quant_item items[kV2MaxTensors];
  memset(items, 0, sizeof(items));
  for (int i = 0; i < n; ++i) {
    if (!group_quant_tma_cs_eligible(C[i], K[i]) || K[i] % g != 0) return AWQK_E_UNSUPPORTED;
    items[i].C = C[i];
    items[i].K = K[i];
  }
  CsPlan best;
  if (!cs_best_plan(items, n, g, (int64_t)sms * (unpacked ? 2 : 3), &best)) return AWQK_E_BADARG;
  summary[0] = best.r_main; summary[1] = best.r_tail; summary[2] = best.b.n_units; summary[3] = best.n_items; summary[4] = best.cost;
  for (int i = 0; i < best.n_items && i < max_items; ++i) {
    rows[4 * i] = best.tensor_of[i];
    rows[4 * i + 1] = best.row_begin[i];
    rows[4 * i + 2] = best.b.it[i].C;
    rows[4 * i + 3] = best.b.it[i].rows_per_unit;
  }
  return AWQK_OK;
}

int launch_group_quant_tma_cs_batch(const awqk_quant_item* items, int n, int dtype, int g, bool sym, cudaStream_t st) {
  if (n <= 0 || n > kV2MaxTensors) return AWQK_E_BADARG;
  int sms = 0;
  const int dev = v2_sms(&sms);
  if (dev < 0) return dev;
  const bool unpacked = items[0].q_unpacked != nullptr;
  const bool has_zp = items[0].zp != nullptr, has_zq = items[0].zp_packed != nullptr, has_qp = items[0].q_packed != nullptr;
  for (int i = 1; i < n; ++i)
    if ((items[i].q_unpacked != nullptr) != unpacked || (items[i].zp != nullptr) != has_zp ||
        (items[i].zp_packed != nullptr) != has_zq || (items[i].q_packed != nullptr) != has_qp)
      return AWQK_E_BADARG;
  const int64_t want = (int64_t)sms * (unpacked ? 2 : 3);
  CsPlan best;
  if (!cs_best_plan(items, n, g, want, &best)) return AWQK_E_BADARG;
  V2Batch& b = best.b;
  b.has_zp = has_zp; b.has_zq = has_zq; b.has_qp = has_qp;
  const unsigned grid = (unsigned)((int64_t)b.n_units < want ? (int64_t)b.n_units : want);
  if (dtype == AWQK_BF16)
    return unpacked ? launch_cs_g<__nv_bfloat16, true>(b, grid, g, sym, dev, st) : launch_cs_g<__nv_bfloat16, false>(b, grid, g, sym, dev, st);
  return unpacked ? launch_cs_g<__half, true>(b, grid, g, sym, dev, st) : launch_cs_g<__half, false>(b, grid, g, sym, dev, st);
}

int launch_group_quant_tma_cs(const void* w, int dtype, int64_t C, int64_t K, int g, bool sym, const float* col_scale,
                              uint32_t* q_packed, int32_t* q_unpacked, void* scales, int32_t* zp, uint32_t* zp_packed,
                              cudaStream_t st) {
  const awqk_quant_item item{w, C, K, col_scale, q_unpacked, q_packed, scales, zp, zp_packed};
  return launch_group_quant_tma_cs_batch(&item, 1, dtype, g, sym, st);
}

}  // namespace awqk
