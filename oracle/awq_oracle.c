/*
 * awq_oracle.c -- plain-C restatement of the AWQ-Converter group quantizer.  TEST INFRASTRUCTURE
 * ONLY (checker and CPU baseline): nothing in the product path links or calls this file.
 *
 * Follows the reference's arithmetic (src/awq_quantizer/quantization/awq.py):
 *   stats      awq.py:192-211   min / max -> (symmetric fold) -> scale = (max-min)/(qmax-qmin),
 *                               clamp(min=1e-10), zp = clamp(round(qmin - min/scale))
 *   codes      awq.py:245-248   clamp(round(x/scale + zp), qmin, qmax), round = half-to-even
 *   layout     awq.py:306-368   rows = dim0, groups along the flattened rest, zero padding joins min/max
 *   casts      awq.py:409-412   scale -> fp32 -> fp16, zp/codes -> int32 (NaN -> INT32_MIN on x86)
 * with PyTorch-CPU semantics for low-precision tensors: every op is evaluated in fp32 and its result
 * is rounded to the tensor dtype ("native" arithmetic), or everything stays fp32 (arith = 1, which is
 * what the reference computes on w.float()).
 *
 * Pinned by tests/test_oracle_golden.py::test_c_oracle_* against the fixtures frozen from the
 * reference (tests/golden/) and against oracle/awq_oracle.py.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/build_oracle.py)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { DT_BF16 = 0, DT_FP16 = 1, DT_FP32 = 2 };

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float bf16_to_f(uint16_t h) { return bits2f((uint32_t)h << 16); }
static inline float round_bf16(float x) {           /* fp32 -> bf16 (RNE) -> fp32 */
  uint32_t u = f2bits(x);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return bits2f((u | 0x00400000u) & 0xFFFF0000u); /* NaN */
  u += 0x7FFFu + ((u >> 16) & 1u);
  return bits2f(u & 0xFFFF0000u);
}
static inline float f16_to_f(uint16_t h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
static inline uint16_t f_to_f16bits(float x) { _Float16 v = (_Float16)x; uint16_t h; memcpy(&h, &v, 2); return h; }
static inline float round_f16(float x) { return (float)(_Float16)x; }

static inline float rnd(float x, int a) { return a == DT_BF16 ? round_bf16(x) : (a == DT_FP16 ? round_f16(x) : x); }
static inline float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
static inline float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
static inline int32_t f2i(float v) { return (v != v) ? INT32_MIN : (int32_t)v; }

static inline float load(const void* w, int dtype, int64_t i) {
  if (dtype == DT_BF16) return bf16_to_f(((const uint16_t*)w)[i]);
  if (dtype == DT_FP16) return f16_to_f(((const uint16_t*)w)[i]);
  return ((const float*)w)[i];
}

/* one group: elements [k0, k1) of a row, zero padded to g */
static void quant_group(const void* w, int dtype, int a, int64_t base, int64_t k0, int64_t k1, int g,
                        float qmin, float qmax, int sym, int32_t* q, uint16_t* scale_out, int32_t* zp_out) {
  float mn, mx;
  if (k1 - k0 < g) { mn = 0.0f; mx = 0.0f; } else { mn = INFINITY; mx = -INFINITY; }
  for (int64_t k = k0; k < k1; ++k) {
    float x = load(w, dtype, base + k);
    mn = min_nan(mn, x);
    mx = max_nan(mx, x);
  }
  if (sym) {
    float am = max_nan(fabsf(mn), fabsf(mx));
    mn = -am; mx = am;
  }
  float floor_ = (a == DT_BF16) ? round_bf16(1e-10f) : (a == DT_FP16 ? 0.0f : 1e-10f);
  float s = rnd(rnd(mx - mn, a) / (qmax - qmin), a);
  s = max_nan(s, floor_);
  float zp = 0.0f;
  if (!sym) {
    float r = rintf(rnd(qmin - rnd(mn / s, a), a));
    zp = (r != r) ? r : fminf(fmaxf(r, qmin), qmax);
  }
  for (int64_t k = k0; k < k1; ++k) {
    float x = load(w, dtype, base + k);
    float r = rintf(rnd(rnd(x / s, a) + zp, a));
    q[base + k] = (r != r) ? INT32_MIN : (int32_t)fminf(fmaxf(r, qmin), qmax);
  }
  *scale_out = f_to_f16bits(s);
  *zp_out = f2i(zp);
}

/* w: [C, K] row-major; q int32 [C,K]; scales fp16 bits [C,G]; zp int32 [C,G]; returns 0 */
int awq_oracle_group_quant(const void* w, int dtype, int64_t C, int64_t K, int g, int bits, int symmetric,
                           int arith_fp32, int32_t* q, uint16_t* scales, int32_t* zp, int nthreads) {
  if (dtype < 0 || dtype > 2 || C <= 0 || K <= 0 || g <= 0 || (bits != 4 && bits != 8)) return -1;
  const int a = arith_fp32 ? DT_FP32 : dtype;
  const float qmin = symmetric ? -(float)(1 << (bits - 1)) : 0.0f;
  const float qmax = symmetric ? (float)((1 << (bits - 1)) - 1) : (float)((1 << bits) - 1);
  const int64_t G = (K + g - 1) / g;
  const int64_t total = C * G;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t t = 0; t < total; ++t) {
    const int64_t row = t / G, grp = t % G;
    const int64_t k0 = grp * g;
    const int64_t k1 = (k0 + g < K) ? k0 + g : K;
    quant_group(w, dtype, a, row * K, k0, k1, g, qmin, qmax, symmetric, q, scales + t, zp + t);
  }
  return 0;
}

/* 8 nibbles (or 4 bytes) per uint32 word along each row; pad codes are 0; NaN codes pack as 0 */
int awq_oracle_pack_rows(const int32_t* codes, int64_t R, int64_t N, int bits, int qmin, uint32_t* words,
                         int nthreads) {
  const int per = 32 / bits;
  const int64_t wpr = (N + per - 1) / per;
  const uint32_t mask = (1u << bits) - 1u;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t r = 0; r < R; ++r)
    for (int64_t j = 0; j < wpr; ++j) {
      uint32_t acc = 0;
      for (int i = 0; i < per; ++i) {
        const int64_t k = j * per + i;
        if (k < N && codes[r * N + k] != INT32_MIN) acc |= (((uint32_t)(codes[r * N + k] - qmin)) & mask) << (bits * i);
      }
      words[r * wpr + j] = acc;
    }
  return 0;
}

/* out = float(fp16_rn(half(q - zp) * scale))  (awq.py:282 with torch's int32 * 0-d fp16 promotion) */
int awq_oracle_dequant(const int32_t* q, const uint16_t* scales, const int32_t* zp, int64_t C, int64_t K,
                       int g, float* out, int nthreads) {
  const int64_t G = (K + g - 1) / g;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t r = 0; r < C; ++r)
    for (int64_t k = 0; k < K; ++k) {
      const int64_t gi = r * G + k / g;
      const int32_t d = (int32_t)((uint32_t)q[r * K + k] - (uint32_t)zp[gi]);
      const float dh = round_f16((float)d);
      out[r * K + k] = round_f16(dh * f16_to_f(scales[gi]));
    }
  return 0;
}

/* bf16 -> fp16 (tensor_utils.py:10-22) */
int awq_oracle_bf16_to_fp16(const uint16_t* in, uint16_t* out, int64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < n; ++i) out[i] = f_to_f16bits(bf16_to_f(in[i]));
  return 0;
}
