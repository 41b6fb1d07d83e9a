/*
 * awqk.h -- C ABI of the B200 (sm_100a) kernels behind AWQ-Converter's quantization hot path.
 *
 * The reference (shanefitch/AWQ-Converter) is pure Python: it has no FFI.  The boundary it
 * does have is the method surface of `AWQQuantizer` (src/awq_quantizer/quantization/awq.py)
 * and `convert_bf16_to_fp16` (src/awq_quantizer/utils/tensor_utils.py).  Each entry point
 * below replaces the arithmetic of one of those methods; the Python mirror in
 * awq-converter_b200/awq_quantizer binds them with ctypes (see INTEGRATION.md for the stub
 * a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types.  Device pointers unless the name says host.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - return value: 0 = ok, negative = AWQK_E_* (awqk_error_string() names it).  No exceptions,
 *     no allocation inside (except awqk_pipe_*, which owns its staging buffers).  Every call is
 *     re-entrant and may be issued concurrently from several host threads (the reference calls
 *     quantize() from a ThreadPoolExecutor, main.py:609-621).  The only process-wide state is
 *     write-once caches (the driver entry point of cuTensorMapEncodeTiled, per-device "opt-in shared
 *     memory size was set" bits) and the lazily created copy-thread pool of awqk_pipe_*; all are
 *     initialised thread-safely and never change afterwards.
 *   - the device the pointers live on is made current for the duration of the call.
 */
#ifndef AWQK_H_
#define AWQK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AWQK_VERSION 210

#if defined(__GNUC__)
#define AWQK_API __attribute__((visibility("default")))
#else
#define AWQK_API
#endif

/* element types of weight tensors */
enum { AWQK_BF16 = 0, AWQK_FP16 = 1, AWQK_FP32 = 2, AWQK_FP64 = 3 };

/* arithmetic contract (DESIGN.md "arithmetic-dtype contract"):
 *   NATIVE: every op rounded to the input dtype, as the reference does on a tensor of that
 *           dtype (awq.py:192-211, 245-248)  -> equals ref.quantize(w)
 *   FP32  : every op in fp32                 -> equals ref.quantize(w.float())            */
enum { AWQK_ARITH_NATIVE = 0, AWQK_ARITH_FP32 = 1 };

enum {
  AWQK_OK = 0,
  AWQK_E_BADARG = -1,      /* null pointer, non-positive size, unsupported bits/dtype/arith   */
  AWQK_E_ALIGN = -2,       /* pointer not aligned as documented                                */
  AWQK_E_CUDA = -3,        /* a CUDA runtime call failed; see awqk_last_cuda_error()           */
  AWQK_E_UNSUPPORTED = -4, /* combination not implemented (e.g. col_scale with NATIVE)         */
  AWQK_E_WORKSPACE = -5,   /* workspace missing or too small                                   */
  AWQK_E_NODEVICE = -6     /* no sm_100 device                                                 */
};

AWQK_API int awqk_version(void);
AWQK_API const char* awqk_error_string(int code);
/* thread-local text of the last CUDA failure seen by this thread ("" if none) */
AWQK_API const char* awqk_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------
 * K1  fused group quantizer (+ pack).   Replaces AWQQuantizer._quantize_per_group +
 *     _compute_scale_zp_for_group + _quantize_tensor            (awq.py:286-374,173-213,215-250)
 *     and the dtype casts of quantize()                          (awq.py:409-412).
 *
 *   w          [C, K] row-major, `dtype`; groups of `group_size` run along K; the last group
 *              of a row is zero-padded (the zeros take part in min/max, awq.py:337-339).
 *   q_unpacked nullable, int32 [C, K]          -- the reference's `tensor_q`
 *   q_packed   nullable, uint32 [C, ceil(K*bits/32)]: word j of a row = sum_i (q[8j+i]-qmin) << 4i
 *              (bits=8: 4 codes per word, << 8i); tail codes are 0.
 *   scales     fp16 [C, G], G = ceil(K / group_size)  -- the reference's `scales`
 *   zp         nullable, int32 [C, G]                  -- the reference's `zero_points`
 *   zp_packed  nullable, uint32 [C, ceil(G*bits/32)], same packing of (zp - qmin) along G
 *   col_scale  nullable, fp32 [K]: quantize (float(w) * col_scale[k]) instead of w (AWQ
 *              per-input-channel scaling); requires arith == AWQK_ARITH_FP32.
 *   bits       4 or 8.  symmetric: qrange [-2^(b-1), 2^(b-1)-1] else [0, 2^b-1] (awq.py:121-128)
 *
 *   Non-finite arithmetic follows the reference on x86: a NaN code / zero point becomes
 *   INT32_MIN (0 in the packed forms).
 * ------------------------------------------------------------------------------------- */
AWQK_API int awqk_group_quant(const void* w, int dtype, int64_t C, int64_t K, int group_size, int bits,
                     int symmetric, int arith, int32_t* q_unpacked, uint32_t* q_packed,
                     void* scales_f16, int32_t* zp, uint32_t* zp_packed, const float* col_scale,
                     void* stream);

/* K1 over a BATCH of tensors in as few launches as possible: the same results as awqk_group_quant(item, ...) for
 * every item, on `stream`.  Items that qualify for the column-slab kernel (col_scale given, bf16 / fp16 weights,
 * bits = 4, arith = FP32, K % 1024 == 0, 16-byte aligned q_packed) are quantized by ONE persistent launch per 32
 * tensors -- the final AWQ pass of a whole wave of linears (awq.py:435-457 loops over the model tensor by tensor) --
 * which is what lets small tensors (a 1024 x 4096 k_proj is 2 us of HBM time) run at the bandwidth of large ones.
 * Anything else falls back to one awqk_group_quant call per item.  `items` is a HOST array, read before returning. */
typedef struct awqk_quant_item {
  const void* w;          /* [C, K] */
  int64_t C, K;
  const float* col_scale; /* nullable fp32 [K] */
  int32_t* q_unpacked;    /* nullable */
  uint32_t* q_packed;     /* nullable */
  void* scales_f16;
  int32_t* zp;            /* nullable */
  uint32_t* zp_packed;    /* nullable */
} awqk_quant_item;
AWQK_API int awqk_group_quant_batch(const awqk_quant_item* items, int n_items, int dtype, int group_size, int bits,
                           int symmetric, int arith, void* stream);

/* Introspection (tests, tools; no device needed): the unit plan of ONE column-slab launch over tensors [C[i], K[i]]
 * (n_tensors <= 31, all eligible) on a device with `sms` SMs.  A tensor is cut into column slabs of 1024 and row chunks;
 * a unit = one row chunk of one slab; units are dealt round-robin to 3 (2 with int32 codes) persistent CTAs per SM.
 *   summary5: tall unit height, short unit height (0 = none), units, launch items, modelled cost (rows per CTA)
 *   items4  : per launch item (tensor index, first row, rows, unit height) -- the items partition every tensor's rows;
 *             short units cover the END of the tensor list, so that the last round over the grid stays short. */
AWQK_API int awqk_group_quant_batch_plan(const int64_t* C, const int64_t* K, int n_tensors, int group_size,
                                int with_int32_codes, int sms, int64_t* summary5, int64_t* items4, int max_items);

/* which kernel awqk_group_quant would pick: 1 = flat fast path, 0 = generic path, <0 error */
AWQK_API int awqk_group_quant_path(int dtype, int64_t C, int64_t K, int group_size, int bits, int arith,
                          const void* w);

/* ---------------------------------------------------------------------------------------
 * K4  group de-quantizer.  Replaces AWQQuantizer.dequantize / _dequantize_tensor
 *     (awq.py:459-539, 252-284): out = float(fp16_rn(half(q - zp) * scale)), fp32 [C, K].
 *     Either q_unpacked (int32 [C,K]) or q_packed/zp_packed (+qmin via symmetric/bits) is given.
 * ------------------------------------------------------------------------------------- */
AWQK_API int awqk_dequant(const int32_t* q_unpacked, const void* scales_f16, const int32_t* zp, int64_t C,
                 int64_t K, int group_size, float* out, void* stream);
AWQK_API int awqk_dequant_packed(const uint32_t* q_packed, const void* scales_f16, const uint32_t* zp_packed,
                        int64_t C, int64_t K, int group_size, int bits, int symmetric, float* out,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * K3  bf16 -> fp16 (round to nearest even; overflow -> inf; subnormals kept).
 *     Replaces convert_bf16_to_fp16 (tensor_utils.py:10-22).
 * ------------------------------------------------------------------------------------- */
AWQK_API int awqk_bf16_to_fp16(const void* in_bf16, void* out_fp16, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * K2  activation-aware scale search.  NO reference counterpart: awq.py only cites the AWQ paper
 *     (awq.py:1-7; scale_method is stored at awq.py:36,66 and never read).  Definition =
 *     oracle/awq_oracle.py::search_scales, which composes the reference's own group quantizer
 *     (awq.py:173-250) -- "parity unpinned" for everything but that quantizer.
 *
 *       m      = mean_t |X[t,:]|                                   (fp64 accumulation)
 *       s_i    = clamp(m^(i/n_grid), 1e-4) / sqrt(max*min)          i = 0 .. n_grid-1
 *       dW_i   = bf16( W - dequant(group_quant(W*s_i)) / s_i )      fp32 arithmetic
 *       err_i  = mean_{t,c} (X . dW_i^T)^2                          tcgen05 bf16 GEMM, fp32 accumulate in TMEM
 *       best   = first minimum of err;  final result = group_quant(fp32(W) * s_best)
 * ------------------------------------------------------------------------------------- */

/* The whole search for ONE tensor, queued on `stream` without any host synchronisation:
 *   w          [C, K] bf16/fp16/fp32;   x_bf16 [T, K] bf16 calibration activations.
 *   s_grid     nullable fp32 [n_grid, K]: scale vectors to score (e.g. cached per activation tensor with
 *              awqk_abs_colsum + awqk_alpha_grid).  NULL: computed from x_bf16 inside the workspace.
 *   err_mean   nullable fp64 [n_grid]: the scores;  best_idx nullable int32[1];  best_s fp32 [K] (required).
 *   q_unpacked / q_packed / scales_f16 / zp / zp_packed: outputs of the final pass, exactly as in
 *              awqk_group_quant(..., arith = FP32, col_scale = best_s); scales_f16 == NULL skips the final pass.
 *   workspace  device memory, 256-byte aligned, >= the minimum of awqk_workspace_bytes (the preferred size
 *              gives the panel ring one more panel of lookahead where that helps).  Must not be shared by two
 *              searches that may run at the same time.  The sizes depend on the CURRENT device (number of
 *              co-resident CTA pairs): query them with the device of `w` current.
 *   Requires group_size in {32,64,128}, K % group_size == 0, K % 64 == 0, 16-byte aligned bases.
 * Inside: the score kernel is ONE persistent launch in which producer warps compute the dW operand of the
 * tensor-core GEMM tile by tile (awqk_search_fused.cu); the argmin, the winning scale vector and the final
 * column-scaled K1 pass follow on the same stream. */
AWQK_API int awqk_scale_search(const void* w, int dtype, int64_t C, int64_t K, const void* x_bf16, int64_t T,
                      const float* s_grid, int n_grid, int group_size, int bits, int symmetric,
                      double* err_mean, int32_t* best_idx, float* best_s, int32_t* q_unpacked,
                      uint32_t* q_packed, void* scales_f16, int32_t* zp, uint32_t* zp_packed,
                      void* workspace, size_t workspace_bytes, void* stream);
/* preferred workspace size of awqk_scale_search for this problem on the current device; *minimum (nullable) gets
 * the smallest size it accepts.  have_s_grid != 0: the caller passes s_grid.  0 on bad arguments or without a
 * usable device.  (A few tens of MB: counters + a ring of 2-3 panels of 256 x 2048 bf16 per slab in flight.) */
AWQK_API size_t awqk_workspace_bytes(int64_t C, int64_t K, int64_t T, int n_grid, int have_s_grid, size_t* minimum);

/* ---- the stages of the search as separate calls (tests, tools) ---- */

/* column statistic: sum_t |X[t,k]| accumulated in fp64.  X is [T, K] bf16/fp16/fp32.
 * `colsum` fp64 [K] must be zeroed by the caller (or accumulate over several calls). */
AWQK_API int awqk_abs_colsum(const void* x, int dtype, int64_t T, int64_t K, double* colsum, void* stream);

/* s_grid[i, k] = normalised clamp((colsum[k]/T)^(i/n_grid), 1e-4), fp32 [n_grid, K].
 * workspace: 2 * n_grid floats (min / max per grid point). */
AWQK_API int awqk_alpha_grid(const double* colsum, int64_t T, int64_t K, int n_grid, float* s_grid,
                    float* workspace_2n, void* stream);

/* dW[i] = bf16( W - dequant(group_quant(W * s_i)) / s_i ), i = 0..n_s-1; fp32 arithmetic.
 *   w [C,K] bf16/fp16/fp32;  s [n_s, K] fp32;  dw bf16 [n_s, C, K].  K % group_size == 0.
 * Stand-alone form of the producer inside awqk_scale_search (same device code, bit-identical). */
AWQK_API int awqk_fakequant_delta(const void* w, int dtype, int64_t C, int64_t K, int group_size, int bits,
                         int symmetric, const float* s, int n_s, void* dw_bf16, void* stream);

/* err[i] += sum over [T, C] of (X . dW_i^T)^2, tcgen05 bf16 GEMM with fp32 TMEM accumulators and a
 * fused sum-of-squares epilogue.  X [T,K] bf16, dW [n_s, C, K] bf16, err fp64 [n_s] (zeroed by the
 * caller).  Requires K % 8 == 0 and 16-byte aligned bases.  Stand-alone form of the consumer inside
 * awqk_scale_search. */
AWQK_API int awqk_sqerr_gemm(const void* x_bf16, const void* dw_bf16, int64_t T, int64_t C, int64_t K,
                    int n_s, double* err, void* stream);

/* ---------------------------------------------------------------------------------------
 * Interop export: K1's packed int4 result -> the AutoAWQ / vLLM "GEMM" checkpoint layout.
 *   in : q_packed [C, K/8] (K1 packing), zp int32 [C, G], scales fp16 [C, G]
 *   out: qweight [K, C/8] with the 0,2,4,6,1,3,5,7 nibble interleave along C, qzeros [G, C/8] (same
 *        packing of zp - qmin), scales fp16 [G, C].   Requires C % 8 == 0 and K % 8 == 0.
 * The closest the reference gets is its non-functional examples/load_quantized_model.py.
 * ------------------------------------------------------------------------------------- */
AWQK_API int awqk_export_autoawq(const uint32_t* q_packed, const int32_t* zp, const void* scales_f16, int64_t C,
                        int64_t K, int64_t G, int symmetric, uint32_t* qweight_out, uint32_t* qzeros_out,
                        void* scales_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Host-buffer pipeline (the e2e path): quantize+pack a host-resident weight through chunked,
 * double-buffered H2D -> K1 -> D2H on private streams.  Host buffers should be pinned
 * (cudaHostAlloc / torch pin_memory) for the copies to overlap.
 * ------------------------------------------------------------------------------------- */
typedef struct awqk_pipe awqk_pipe;
AWQK_API int awqk_pipe_create(int device, size_t chunk_bytes, awqk_pipe** out);
AWQK_API void awqk_pipe_destroy(awqk_pipe* p);
AWQK_API int awqk_pipe_quant_host(awqk_pipe* p, const void* w_host, int dtype, int64_t C, int64_t K,
                         int group_size, int bits, int symmetric, int arith,
                         int32_t* q_unpacked_host, uint32_t* q_packed_host, void* scales_f16_host,
                         int32_t* zp_host, uint32_t* zp_packed_host);
/* Gather mode: the same pipeline over a VIRTUAL arena of n_tensors host tensors (ordinary pageable memory is
 * fine): tensor i owns elements [v_i, v_i + numel[i]) with v_0 = 0, v_{i+1} = v_i + roundup(numel[i], 8192).
 * The pipe stages it chunk by chunk through its own bounded ring of pinned buffers (AWQK_PIPE_THREADS memcpy
 * threads, default by core count) and a drain thread copies finished chunks into the result arrays, which are
 * flat over the virtual arena exactly as in awqk_pipe_quant_host (C = 1, K = v_n).  Every numel[i] must be a
 * multiple of group_size (of group_size * 32 / bits when zp_packed_host is given, so that packed zero words
 * never straddle tensors).  row_len = 0: that flat layout.  row_len > 0: every tensor is made of rows of row_len
 * elements holding 1, 2 or 4 groups (fewer than a packed zero word): zp_packed gets ONE zero-padded word per row
 * (K = 512 at g = 128, e.g. OPT's embed_tokens / project_in).  BLOCKING: returns when all results are in place.  Replaces the reference's
 * per-tensor tensor.to(device) / .cpu() round trips (main.py:300, 374-380) for a whole model. */
AWQK_API int awqk_pipe_quant_gather(awqk_pipe* p, int n_tensors, const void* const* src, const int64_t* numel,
                           int64_t row_len, int dtype, int group_size, int bits, int symmetric, int arith,
                           int32_t* q_unpacked_host, uint32_t* q_packed_host, void* scales_f16_host,
                           int32_t* zp_host, uint32_t* zp_packed_host);
/* memcpy split over `threads` host threads (0 = the pipe's default: AWQK_PIPE_THREADS or by core count).  One
 * core moves ~10 GB/s, a PCIe 5 x16 link wants 50: the staging copies between pageable tensors and the pinned
 * rings of the Python-side streams (quantization/search.py) go through this. */
AWQK_API int awqk_host_copy(void* dst, const void* src, size_t bytes, int threads);
/* first-touch a fresh pageable host buffer from `threads` threads (0 = default), after asking for transparent huge
 * pages on it: result arrays that the drain copies would otherwise fault in page by page (~10 GB/s instead of
 * memcpy speed).  Content is not changed (atomic OR with 0), so it may overlap with writes into the buffer. */
AWQK_API int awqk_host_prefault(void* ptr, size_t bytes, int threads);
/* Page-locked host memory for staging rings, obtained ~7x faster than cudaHostAlloc (which runs at ~2.4 GB/s on these
 * hosts): a 2 MiB-aligned mapping with transparent huge pages, first-touched by the copy threads, then page-locked in
 * place (cudaHostRegister, portable).  Falls back to cudaHostAlloc where registration fails.  *out is 2 MiB aligned;
 * free with awqk_host_free_pinned (after the copies that use it have finished). */
AWQK_API int awqk_host_alloc_pinned(size_t bytes, void** out);
AWQK_API int awqk_host_free_pinned(void* ptr);
/* wait for everything queued on the pipe */
AWQK_API int awqk_pipe_sync(awqk_pipe* p);

#ifdef __cplusplus
}
#endif
#endif /* AWQK_H_ */
