/*
 * b200q.h — C ABI of libb200q.so: the B200 (sm_100a) INT4 dequantize-linear / INT4 MoE hot path.
 *
 * This is the drop-in boundary.  Every entry point below replaces one piece of the
 * reference's Python/pybind surface (citations are paths under the reference repo):
 *
 *   b200q_linear_fwd        <- fused_quant_linear_cuda.forward      csrc/quantized_linear.cpp:22-28,
 *                                                                    csrc/quantized_linear_kernel.cu:293-378
 *   b200q_quantize_rows     <- quantize_weights                     python/quantize.py:38-124
 *   b200q_quantize_rows_given / b200q_minmax
 *                           <- quantize_weights_moe                 python/moe_int4_module.py:19-80
 *   b200q_dequantize_rows   <- dequantize_weights                   python/quantize.py:127-173
 *   b200q_moe_topk          <- simulate_routing softmax/topk/renorm benchmark/moe_grouped_gemm/routing.py:72-76
 *   b200q_moe_permute       <- simulate_routing histogram/offsets   routing.py:79-86
 *                              + create_expert_inputs               routing.py:96-149
 *   b200q_moe_gather_rows   <- create_expert_inputs row gather      routing.py:137-147
 *   b200q_moe_grouped_fwd   <- moe_int4_cuda.forward                csrc/moe_int4_kernel.cu:93-141
 *                              + QuantizedMoE.forward               benchmark/moe_grouped_gemm/moe_int4_module.py:123-125
 *   b200q_moe_grouped_fwd_ranges <- moe_int4_cuda.forward (input_offsets, tokens_per_expert)   csrc/moe_int4_kernel.cu:112-123
 *   b200q_moe_silu_mul, b200q_moe_grouped_gated_fwd
 *                           <- (north_star extension: gated MLP, silu(x w1^T) * (x w3^T))
 *   b200q_moe_combine       <- combine_expert_outputs               routing.py:152-189
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers unless
 *     the parameter name starts with `h_`.
 *   - every function returns 0 on success or a negative B200Q_E* code; it never throws, never
 *     allocates device memory, never synchronises the device.  Work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - outputs and workspaces are caller-owned.  A workspace must be zero-filled ONCE by the
 *     caller when it is allocated and is then private to ONE entry point (b200q_linear_fwd and the
 *     two b200q_moe_grouped_fwd* may share one: their layouts reserve the same 4 KB of ticket counters);
 *     the kernels leave the counters zeroed again on exit.  Alignment: 128 bytes.
 *   - re-entrant / thread-safe: the only global state is an immutable per-device property cache
 *     and a thread-local last-error string.
 *   - there is no CPU path.  On a machine without an sm_100 GPU the compute entry points
 *     return B200Q_EARCH / B200Q_ECUDA.
 *
 * Packed layout (identical to python/quantize.py:120-122): packed[n, b] holds column 2b in the
 * low nibble and column 2b+1 in the high nibble; scales[n], zero_points[n] are fp32 per row and
 * w[n,k] = (q[n,k] - zero_points[n]) * scales[n].
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_VERSION 200 /* major*10000 + minor*100 + patch */

/* error codes */
#define B200Q_OK 0
#define B200Q_EINVAL (-1)     /* bad argument (null pointer, negative size, unsupported dtype ...) */
#define B200Q_EALIGN (-2)     /* pointer or shape violates an alignment rule (see each function)   */
#define B200Q_EARCH (-3)      /* current device is not compute capability 10.x                     */
#define B200Q_ECUDA (-4)      /* a CUDA runtime call failed; see b200q_last_error_string()          */
#define B200Q_EWORKSPACE (-5) /* workspace too small                                                */
#define B200Q_ENCCL (-6)      /* NCCL could not be loaded, or an NCCL call failed (b200q_ep_*)       */

/* element types of activations / outputs */
#define B200Q_F32 0
#define B200Q_F16 1
#define B200Q_BF16 2

/* flags for b200q_linear_fwd */
#define B200Q_FLAG_NONE 0u
/* The packed weights / scales / zero points are NOT written by the kernel that precedes this call
 * on `stream` (true for any inference module whose weights are fixed buffers).  Allows the kernel
 * to be launched with programmatic dependent launch and to start streaming weights into shared
 * memory while its predecessor drains; activations are only read after the dependency resolves. */
#define B200Q_FLAG_STATIC_WEIGHTS 1u

int b200q_version(void);
/* Thread-local description of the last non-zero return value on this thread ("" if none). */
const char* b200q_last_error_string(void);
/* 0 if `device` (or the current device when device < 0) can run this library (sm_100). */
int b200q_device_check(int device);
/* Number of SMs of the current device (148 on B200); negative error code on failure. */
int b200q_sm_count(void);

/* ---- quantisation format ------------------------------------------------------------------ */

/* Per-row asymmetric INT4 quantisation, bit-exact with python/quantize.py:38-124.
 *   w [N,K] f32 row-major (K even, K <= 2^20) -> packed [N,K/2] u8, scales [N] f32, zps [N] f32 */
int b200q_quantize_rows(const float* w, int64_t N, int64_t K, uint8_t* packed, float* scales,
                        float* zps, void* stream);

/* Quantise with caller-supplied per-row scale / zero point (python/moe_int4_module.py:56-76:
 * q = clamp(round(w/scale + zp), 0, 15)); scales/zps are [N] f32 and are only read. */
int b200q_quantize_rows_given(const float* w, int64_t N, int64_t K, const float* scales,
                              const float* zps, uint8_t* packed, void* stream);

/* min and max over `count` f32 values -> out_minmax[0], out_minmax[1] (for the per-expert scalar
 * scale of python/moe_int4_module.py:46-50).  ws: >= b200q_minmax_ws_bytes() bytes. */
size_t b200q_minmax_ws_bytes(void);
int b200q_minmax(const float* v, int64_t count, float* out_minmax, void* ws, size_t ws_bytes,
                 void* stream);

/* python/quantize.py:127-173: out[n,k] = (q[n,k] - zps[n]) * scales[n], fp32, bit-exact. */
int b200q_dequantize_rows(const uint8_t* packed, const float* scales, const float* zps, int64_t N,
                          int64_t K, float* out, void* stream);

/* ---- fused dequantize + linear ------------------------------------------------------------ */

/* y[M,N] = x[M,K] @ dequant(packed[N,K/2], scales[N], zps[N])^T
 *   x_dtype / y_dtype: B200Q_F32 | B200Q_F16 | B200Q_BF16 (the reference API is f32/f32).
 *   x, y row-major and contiguous; x 16-byte aligned, packed 16-byte aligned, K even.
 *   M may be any value >= 0.  K % 256 == 0 with 16-byte aligned pointers: M <= 2 takes the exact-integer decode
 *   kernel (TMA tensor boxes, IMMA over the raw packed bytes), 3 <= M <= 24 (32 for K <= 8192 when one wave of CTAs holds the rows) the mid-batch decode kernel (eight tokens per
 *   tensor instruction: three-digit IMMA form for fp32 activations -- wants x 32-byte aligned --, fp16 HMMA form for 16-bit
 *   ones), both with the CTA's weight rows resident in shared memory; larger batches the tcgen05 GEMM; anything else a
 *   generic SIMT kernel.
 *   Rows of x that contain NaN / Inf give the values dequantize + F.linear gives (python/quantize.py:172, 202).
 *   ws: >= b200q_linear_ws_bytes(M,N,K) bytes (may be NULL when that is 0). */
size_t b200q_linear_ws_bytes(int64_t M, int64_t N, int64_t K);
int b200q_linear_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                     const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                     void* ws, size_t ws_bytes, unsigned flags, void* stream);

/* Same, plus a hint for decoders that know which fused linear follows on this stream: the packed
 * weights of the NEXT call (`next_packed`, `next_bytes`, 16-byte aligned; NULL / 0 = none).  The decode
 * kernel pulls them into L2 behind its own weight stream (cp.async.bulk.prefetch.L2), so HBM keeps
 * streaming through its epilogue and the next call's prologue.  Purely a performance hint: results are
 * identical, and the next call may be anything. */
int b200q_linear_fwd_next(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream,
                          const uint8_t* next_packed, size_t next_bytes);

/* Same, plus a per-output bias: y[m,n] += bias[n] (bias [N] f32, NULL = none).  The reference asserts
 * `linear.bias is None` (python/module.py:84); this lifts that restriction.  The decode kernel adds the
 * bias in its epilogue; the other paths run a small element-wise pass behind the GEMM. */
int b200q_linear_bias_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream,
                          const uint8_t* next_packed, size_t next_bytes);

/* Group-wise scales (the reference's format is per-row only, python/quantize.py:73-80): one scale / zero
 * point per `group_size` consecutive columns, scales / zps [N, K/group_size] f32 row-major, packed as above.  The groups are
 * quantised with the per-row formulas on W viewed as [N K / group_size, group_size] (b200q_quantize_rows on that view gives
 * exactly this layout; b200q_dequantize_rows inverts it).  group_size: a multiple of 8 that divides K.  Decode-sized
 * batches (M <= 32) with group_size % 128 == 0 run on the mid-batch decode kernel (K % 256 == 0, x 32-byte (fp32) / 16-byte
 * aligned); everything else on the format's reference-speed kernel (w = (q - zp) * s, fp32 multiply-add, the reference's
 * arithmetic order).  The prefill GEMM takes per-row scales only. */
int b200q_linear_groupwise_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                               int64_t group_size, void* y, int y_dtype, int64_t M, int64_t N, int64_t K, void* stream);
/* Same with an optional bias [N] f32 (fused on the decode kernel), the B200Q_FLAG_* flags and the next-layer L2 hint of
 * b200q_linear_bias_fwd. */
int b200q_linear_groupwise_bias_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                                    const float* bias, int64_t group_size, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                                    unsigned flags, void* stream, const uint8_t* next_packed, size_t next_bytes);

/* Fused gate + up pair of a gated MLP (`down(silu(gate(x)) * up(x))` in two launches):
 *   h[m,f] = silu(x[m,:] . W[2f,:]) * (x[m,:] . W[2f+1,:])
 * packed13 [2F,K/2], scales13 / zps13 [2F]: the rows of the gate and the up projection INTERLEAVED (2f: gate, 2f+1: up),
 * h [M,F].  M <= 16: the decode kernel with the gate in its epilogue (both rows of a column live in one CTA); larger M:
 * the tcgen05 GEMM with the gated epilogue.  Needs K % 128 == 0 and 16-byte aligned buffers.
 * ws: >= b200q_linear_ws_bytes(M, 2F, K). */
int b200q_linear_gated_fwd(const void* x, int x_dtype, const uint8_t* packed13, const float* scales13, const float* zps13,
                           void* h, int h_dtype, int64_t M, int64_t F, int64_t K, void* ws, size_t ws_bytes, unsigned flags,
                           void* stream, const uint8_t* next_packed, size_t next_bytes);

/* Same with HOST activations: enqueues H2D copy of x (h_x -> the caller's device staging buffer d_x), the
 * fused dequantize-linear, and the D2H copy of the result (d_y -> h_y) on `stream`; h_x / h_y should be
 * pinned.  This is the call a host-resident caller of the reference's QuantizedLinear.forward maps to
 * (python/module.py:100-118 with x on the CPU and the module on the GPU).  Capturable in a CUDA graph.
 * For M <= 8 and a device-addressable pinned h_y the kernel stores the result straight into h_y (no copy node);
 * d_y is then unused. */
int b200q_linear_fwd_host(const void* h_x, int x_dtype, void* d_x, const uint8_t* packed, const float* scales,
                          const float* zps, void* d_y, void* h_y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream);

/* Bench / tuning hook: override a launch heuristic process-wide; value < 0 restores the default.  Keys:
 *   force_path (1 generic SIMT, 2 ring decode kernel, 3 tcgen05 GEMM, 6 exact-integer decode kernel, 7 mid-batch decode
 *   kernel), hm_min_m, hm_max_m, hm_i3, hm_waves, moe_dec_hm, moe_dec_compact, gemv_early, gemv_pf,
 *   gemv_tma3d, gemv_xprep, gemv_warps, gemv_slabs, gemv_stages, gemv_pdl, gemv_ctas, gemv_occ2, gemv_debug,
 *   gemm_bn (32 / 64 / 128 / 192 / 256), gemm_sk, gemm_mt_major (tile order: 0 weight-tile-major, 1 token-tile-major),
 *   gemm_debug (needs a -DB200Q_PROF build), host_direct.
 * Their meaning is documented next to the Tuning struct in csrc/internal.h.  Not needed by callers. */
int b200q_tune_set(const char* key, int value);

/* ---- MoE ---------------------------------------------------------------------------------- */

/* softmax over E logits -> top-k (ties: lowest expert index first) -> renormalise
 * (routing.py:72-76).  logits [T,E] f32, E <= 256, k <= 8.  idx [T,k] i32, weights [T,k] f32. */
int b200q_moe_topk(const float* logits, int64_t T, int E, int k, int32_t* idx, float* weights,
                   void* stream);

/* Histogram + exclusive offsets + stable counting sort of the T*k (token,slot) assignments by
 * expert (routing.py:79-86, 121-132).
 *   idx [T*k] i32 -> counts [E] i32, offsets [E+1] i32 (offsets[E] = T*k),
 *   sorted_slot [T*k] i32 : flat assignment index (t*k+s) of every sorted position, ascending
 *                           inside an expert (the stable order; the reference's argsort is
 *                           unstable, so only the per-expert SET is comparable),
 *   inv_perm [T*k] i32    : sorted position of flat assignment t*k+s.
 *   ws: >= b200q_moe_permute_ws_bytes(T,E,k) bytes. */
size_t b200q_moe_permute_ws_bytes(int64_t T, int E, int k);
int b200q_moe_permute(const int32_t* idx, int64_t T, int E, int k, int32_t* counts,
                      int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws,
                      size_t ws_bytes, void* stream);

/* Same, sorted by remap[expert] instead of the expert id (remap [E] i32 on the device, a permutation of 0..E-1;
 * NULL = identity): offsets / sorted_slot / inv_perm follow the remapped order, counts stay per ORIGINAL expert id.
 * Expert-parallel ranks sort by (destination rank, expert) this way (see b200q_ep_plan). */
int b200q_moe_permute_mapped(const int32_t* idx, const int32_t* remap, int64_t T, int E, int k, int32_t* counts,
                             int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws,
                             size_t ws_bytes, void* stream);

/* b200q_moe_topk followed by b200q_moe_permute_mapped in one call (one host transition for the whole router). */
int b200q_moe_route(const float* logits, const int32_t* remap, int64_t T, int E, int k, int32_t* idx, float* weights,
                    int32_t* counts, int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws, size_t ws_bytes,
                    void* stream);

/* xs[p,:] = x[sorted_slot[p] / k, :]  (routing.py:137-147).  d % 4 == 0 (f32) / 8 (16-bit). */
int b200q_moe_gather_rows(const void* x, int dtype, const int32_t* sorted_slot, int64_t rows,
                          int k, int64_t d, void* xs, void* stream);

/* Grouped fused dequantize-linear over rows already grouped by expert:
 *   y[p,:] = xs[p,:] @ dequant(packed[e], scales[e], zps[e])^T   for offsets[e] <= p < offsets[e+1]
 *   packed [E,N,K/2] u8, scales/zps [E,N] f32, xs [R,K], y [R,N]; offsets [E+1] i32 ON DEVICE
 *   (no host sync).  Rows outside every group are zero-filled (moe_int4_kernel.cu:109).
 *   ws: >= b200q_moe_grouped_ws_bytes(R,E,N,K). */
size_t b200q_moe_grouped_ws_bytes(int64_t R, int E, int64_t N, int64_t K);
int b200q_moe_grouped_fwd(const void* xs, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, const int32_t* offsets, int E, void* y, int y_dtype,
                          int64_t R, int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream);

/* Same, for groups given as explicit ranges [starts[e], ends[e]) (both [E] i32 on device) that
 * need not be contiguous -- the (input_offsets, tokens_per_expert) form of
 * csrc/moe_int4_kernel.cu:112-123 with ends = input_offsets + tokens_per_expert.  Rows covered by no
 * range are NOT written (the caller zero-fills y, as moe_int4_kernel.cu:109 does). */
int b200q_moe_grouped_fwd_ranges(const void* xs, int x_dtype, const uint8_t* packed,
                                 const float* scales, const float* zps, const int32_t* starts,
                                 const int32_t* ends, int E, void* y, int y_dtype, int64_t R,
                                 int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream);

/* Gated first half of the expert MLP in ONE grouped GEMM (north-star extension; semantics composed from the
 * reference's primitives):  h[p,f] = silu(xs[p,:] . w1_e[f,:]) * (xs[p,:] . w3_e[f,:])  for
 * offsets[e] <= p < offsets[e+1].  packed13 [E,2F,K/2], scales13 / zps13 [E,2F] hold the rows of w1 and w3
 * INTERLEAVED (row 2f = w1[f], row 2f+1 = w3[f]) so that a tile's epilogue has both projections of an h column;
 * h [R,F].  Needs K % 128 == 0 and 16-byte aligned buffers (B200Q_EINVAL otherwise: use b200q_moe_grouped_fwd on
 * the concatenated w1||w3 followed by b200q_moe_silu_mul).  ws: >= b200q_moe_grouped_ws_bytes(R,E,2F,K). */
int b200q_moe_grouped_gated_fwd(const void* xs, int x_dtype, const uint8_t* packed13, const float* scales13,
                                const float* zps13, const int32_t* offsets, int E, void* h, int h_dtype,
                                int64_t R, int64_t F, int64_t K, void* ws, size_t ws_bytes, void* stream);

/* h[p,f] = silu(a[p,f]) * b[p,f] where a = gu[p, 0:F], b = gu[p, F:2F]  (gated-MLP extension). */
int b200q_moe_silu_mul(const void* gu, int dtype, int64_t R, int64_t F, void* h, void* stream);

/* Same family, ranges with an explicit weight expert: rows [starts[v], ends[v]) use expert range_expert[v] of
 * packed [n_experts,N,K/2] (all three arrays [n_ranges] i32 on the device; empty ranges are skipped; rows covered by
 * no range are not written).  gated != 0: the fused SiLU-gate form (N = 2F interleaved rows, y is h [R, N/2]).
 * This is what an expert-parallel rank runs on the rows it received, which arrive ordered (source rank, expert):
 * one range per (local expert, source) pair instead of a re-sort (see b200q_ep_plan). */
int b200q_moe_grouped_fwd_mapped(const void* xs, int x_dtype, const uint8_t* packed, const float* scales,
                                 const float* zps, const int32_t* starts, const int32_t* ends,
                                 const int32_t* range_expert, int n_ranges, int n_experts, int gated, void* y, int y_dtype,
                                 int64_t R, int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream);

/* The whole routed gated layer for DECODE-sized calls (T <= 16 tokens) behind one C call:
 *   out[t,:] = sum_s p[t,s] * w2_e( silu(w1_e x[t,:]) * (w3_e x[t,:]) ),  e = expert of slot s of token t  (fp32)
 * Four launches: routing (softmax / top-k / counting sort in one CTA), the fused gate / up GEMV (reads x in place through
 * the token map) and the down GEMV -- both grouped over the experts with device-side row offsets on the decode kernels
 * (exact-integer kernel below 1.5 rows per expert on average, mid-batch kernel above; experts without tokens exit at
 * once; HBM-bound: only the experts that were hit are read) -- and
 * b200q_moe_combine.  E <= 256, k <= 8.  packed13 [E,2F,d/2] (w1 / w3 interleaved), packed2
 * [E,d,F/2]; d and F multiples of 256.  ws: >= b200q_moe_decode_ws_bytes, 256-byte aligned (no zero-fill needed).
 * B200Q_EINVAL when the shape does not fit the decode kernel (use the grouped tcgen05 path then). */
size_t b200q_moe_decode_ws_bytes(int64_t T, int E, int k, int64_t d, int64_t F);
int b200q_moe_decode_fwd(const void* x, int x_dtype, const float* logits, int64_t T, int E, int k,
                         const uint8_t* packed13, const float* scales13, const float* zps13,
                         const uint8_t* packed2, const float* scales2, const float* zps2, int64_t d, int64_t F,
                         float* out, void* ws, size_t ws_bytes, void* stream);

/* out[t,:] = sum_s weights[t,s] * y[inv_perm[t*k+s], :]   (routing.py:175-187). */
int b200q_moe_combine(const void* y, int dtype, const int32_t* inv_perm, const float* weights,
                      int64_t T, int k, int64_t F, void* out, int out_dtype, void* stream);

/* ---- expert parallelism (not in the reference: it is single-GPU; north-star extension) -------------------------
 * One process per GPU.  Rank r owns experts [r E/world, (r+1) E/world); experts flagged in `replicated` are held by
 * every rank and serve their tokens where those live (hot-expert replication for skewed routing).  Every rank sorts
 * its (token, slot) assignments by (destination rank, expert) -- b200q_moe_permute on remapped expert ids -- and:
 *
 * b200q_ep_plan: counts_all [world,E] i32 (the all-gathered tokens-per-expert histograms, original expert ids),
 *   replicated [E] i32 or NULL, local_index [E] i32 (index of expert e in this rank's weight tensor, -1 if absent), all on
 *   the device  ->  splits [2 world] i32 (rows sent to / received from every rank), and for the received rows (ordered
 *   source rank, then expert) range_starts / range_ends / range_expert [E world] i32 = the arguments of
 *   b200q_moe_grouped_fwd_mapped (entry local_index * world + source).  One small kernel, no host work.
 * b200q_ep_unique_id / b200q_ep_comm_create / b200q_ep_comm_destroy: an NCCL communicator for the exchange (the id is
 *   128 bytes of HOST memory made on one rank and handed to the others by the caller, e.g. a torch.distributed
 *   broadcast).  libnccl.so.2 is resolved with dlopen; B200Q_ENCCL if it cannot be found or a call fails.
 * b200q_ep_allgather_i32: recv [world * count] <- send [count] of every rank (the histogram exchange).
 * b200q_ep_exchange: variable-split all-to-all of rows of row_bytes bytes: h_send_rows / h_recv_rows [world] i64 on the
 *   HOST (NCCL's send / recv sizes are host arguments: the one place the plan's splits are needed on the host). */
int b200q_ep_plan(const int32_t* counts_all, int world, int rank, int E, const int32_t* replicated,
                  const int32_t* local_index, int32_t* splits, int32_t* range_starts, int32_t* range_ends,
                  int32_t* range_expert, void* stream);
int b200q_ep_unique_id(void* h_id128);
int b200q_ep_comm_create(const void* h_id128, int rank, int world, void** comm);
int b200q_ep_comm_destroy(void* comm);
int b200q_ep_allgather_i32(void* comm, const int32_t* send, int32_t* recv, int64_t count, void* stream);
int b200q_ep_exchange(void* comm, int world, const void* send, const int64_t* h_send_rows, void* recv,
                      const int64_t* h_recv_rows, int64_t row_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
