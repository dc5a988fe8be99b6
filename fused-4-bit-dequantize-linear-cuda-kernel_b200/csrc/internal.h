// Host-side internals shared by the translation units of libb200q.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "../../include/b200q.h"

struct CUtensorMap_st;

namespace b200q {

// ---- error plumbing (thread-local message, C ABI never throws)
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define B200Q_CUDA(expr)                                        \
    do {                                                        \
        int _rc = ::b200q::check_cuda((expr), #expr);           \
        if (_rc != 0) return _rc;                               \
    } while (0)

// ---- immutable per-device cache
struct DeviceInfo {
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    int max_smem_optin = 0;
};
// Fills `out` for the current device; returns 0 or an error code.  Fails with B200Q_EARCH when the
// device is not compute capability 10.x.
int current_device(DeviceInfo* out);

// ---- tuning overrides (bench hook, see b200q_tune_set)
struct Tuning {
    int gemv_warps = -1;    // consumer warps per CTA (8 or 16)
    int gemv_slabs = -1;    // number of K slabs (CTAs splitting K)
    int gemv_stages = -1;   // cap on ring depth
    int gemv_pdl = -1;      // 0 disables programmatic dependent launch
    int gemv_ctas = -1;     // cap on the number of CTAs (default: SM count)
    int gemm_bn = -1;       // token-tile height of the tcgen05 GEMM (128 / 192 / 256), default: heuristic
    int host_direct = -1;   // 0: b200q_linear_fwd_host always copies the result back instead of storing into pinned memory
    int gemm_debug = -1;    // bench-only ablations of the tcgen05 GEMM (1: no weight loads, 2: no activation loads)
    int gemm_mt_major = -1; // tcgen05 GEMM tile order: -1 auto (token-tile-major for grouped calls, weight-tile-major for the dense linear), 0 weight-tile-major, 1 token-tile-major
    int gemm_sk = -1;       // stream-K in the tcgen05 GEMM: -1 heuristic, 0 off, 1 whenever possible
    int gemv_res = 1;       // 0: never use the resident-slab decode kernel
    int gemv_early = -1;    // tiles requested before the x loads (-1: all)
    int gemv_xprep = 0;     // 0: x operand built inside every CTA (default); 1: by a separate preparation kernel
    int gemv_pf = 3;        // next-layer L2 prefetch: 0 off; resident kernel: 1 behind the last own request, 2 before the own requests, 3 after the operand build (default: the prologue is sensitive to memory traffic)
    int gemv_occ2 = 0;      // 1: 8-warp CTAs sized so that two launches share an SM (cross-layer prefetch)
    int gemv_debug = -1;    // bench-only ablations: 1 = skip the mma work, 2 = skip the weight loads
    int force_path = -1;    // 0 auto, 1 generic SIMT, 2 ring decode kernel, 3 tcgen05 gemm, 6 resident decode kernel (gemv_dec), 7 fp16 HMMA decode kernel (gemv_hm)
    int gemv_bufs = 0;      // resident decode kernel: cap on the tile buffers of a CTA (0 = as many as fit; tests force the ring with it)
    int hm_waves = 0;       // gemv_hm.cu: most waves of CTAs (0 = 4); 1 = only shapes whose rows fit with one CTA per SM
    int moe_dec_hm = -1;    // b200q_moe_decode_fwd: -1 the mid-batch kernel from 1.5 rows per expert on average, 0 never, 1 always
    int moe_dec_compact = 1;   // b200q_moe_decode_fwd: grid rows = min(E, T k) with device-side expert ranks (0: one grid row per expert)
    int hm_i3 = 1;          // gemv_hm.cu, fp32 activations, 1 = three-digit IMMA form, 0 = fp16 hi / lo HMMA form
    int hm_max_m = 32;      // largest batch on the mid-batch decode kernel (<= 32: four passes of eight tokens)
    int hm_min_m = 3;       // smallest batch that goes to the fp16 HMMA decode kernel (gemv_hm.cu)
    int gemv_slots = 1;     // resident decode kernel, M <= 2: 0 = pipelined cross-warp reduction instead of per-warp slots
};
const Tuning& tuning();

// ---- kernel launchers (each returns 0 / error code, enqueues on stream, no sync)
int launch_quantize_rows(const float* w, int64_t N, int64_t K, const float* given_scales,
                         const float* given_zps, uint8_t* packed, float* scales, float* zps,
                         cudaStream_t st);
int launch_dequantize_rows(const uint8_t* packed, const float* scales, const float* zps, int64_t N,
                           int64_t K, float* out, cudaStream_t st);
int launch_minmax(const float* v, int64_t count, float* out, void* ws, cudaStream_t st);

// generic SIMT fused dequant-linear (any even K).  Grouped mode: rows [starts[e], ends[e]) use
// expert e of packed [E,N,K/2]; starts == nullptr is the plain linear.  zero_outside: also zero the
// rows before starts[0] and after ends[E-1] (the offsets[E+1] form of the C ABI).
int launch_linear_generic(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          const int32_t* starts, const int32_t* ends, int E, int zero_outside,
                          cudaStream_t st, const int32_t* emap = nullptr, int kgroup = 0);

// decode GEMV, M <= 16, K % 128 == 0.  Returns B200Q_EINVAL if the shape is not supported so
// the dispatcher can fall through.
bool gemv_supported(int64_t M, int64_t N, int64_t K, int x_dtype, const DeviceInfo* dev = nullptr);   // dev == nullptr: the current device
size_t gemv_ws_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemv(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                const float* scales, const float* zps, void* y, int y_dtype, int64_t M, int64_t N,
                int64_t K, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st,
                const uint8_t* next_packed = nullptr, size_t next_bytes = 0);

// y[m, n] += bias[n] (paths without a fused bias)
int launch_bias_add(void* y, int y_dtype, const float* bias, int64_t M, int64_t N, cudaStream_t st);

// Rows of x that contain NaN / Inf are recomputed in the reference's order (w = (q - zp) * s, then fp32 multiply-add;
// python/quantize.py:172, 202) so that non-finite inputs propagate as in dequantize + F.linear.  nf_flags: per-row
// flags written by the activation preparation (nullptr: the kernel scans x itself).  Grouped: starts / ends / E as in
// launch_gemm_tc; gated: y is h [R, N/2] = silu(row 2f) * (row 2f+1).
int launch_nonfinite_fixup(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                           const int* nf_flags, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                           const int32_t* starts, const int32_t* ends, int E, int gated, cudaStream_t st,
                           const int32_t* emap = nullptr);

// rows [0, *first) and [*last, R) of y <- 0 (the offsets[E+1] form of b200q_moe_grouped_fwd: first = offsets, last = offsets + E)
int launch_zero_rows_outside(void* y, int y_dtype, int64_t R, int64_t N, const int32_t* first, const int32_t* last,
                             cudaStream_t st);

// device-addressable pinned host memory -> device buffer, `bytes` % 16 == 0 (one small kernel instead of a copy node)
int launch_stage_host(const void* src_mapped, void* dst, size_t bytes, cudaStream_t st);

// decode GEMV with the CTA's rows resident in shared memory (gemv_dec.cu): M <= 16, K % 256 == 0, K <= 16384; a CTA whose rows do not
// fit in its tile buffers refills them as a ring; no workspace; optional bias [N] f32; gated: fused SiLU-gate of
// interleaved gate / up rows; offsets != nullptr: grouped over n_experts experts (M = the largest group, <= 16)
bool gemv_dec_supported(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated = 0);
bool gemv_dec_resident(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated = 0);     // every tile of a CTA has its own buffer
int launch_gemv_dec(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                    const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                    unsigned flags, cudaStream_t st, const uint8_t* next_packed, size_t next_bytes, int gated = 0,
                    const int32_t* offsets = nullptr, int n_experts = 1, const int32_t* row_map = nullptr, int max_groups = 0);
// 3-D tensor map over packed [N, K/2] viewed as [row][128-byte column][byte]: boxes of [16 rows][chunk columns][128 B]
// land in shared memory as [column][row][128 B] with the 128-byte swizzle (cached per (pointer, shape, chunk))
int dec_weight_map(const uint8_t* packed, int64_t N, int64_t K, int chunk, CUtensorMap_st* out);
// decode GEMV on fp16 HMMA (gemv_hm.cu): M <= 16, every tile of a CTA resident in shared memory; tokens in passes of
// eight; fp32 activations as fp16 hi + lo parts; optional bias; gated: fused SiLU-gate of interleaved gate / up rows
bool gemv_hm_supported(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated = 0);
int launch_gemv_hm(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                   const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                   unsigned flags, cudaStream_t st, const uint8_t* next_packed, size_t next_bytes, int gated = 0,
                   const int32_t* offsets = nullptr, int n_experts = 1, const int32_t* row_map = nullptr, int kgroup = 0);
// decode-sized routing in one launch (moe.cu): T <= 16, E <= 256, k <= 8; src_token (optional): token of every sorted position
int moe_route_small(const float* logits, int64_t T, int E, int k, int32_t* idx, float* weights, int32_t* counts, int32_t* offsets,
                    int32_t* sorted_slot, int32_t* inv_perm, int32_t* src_token, cudaStream_t st);

// prefill / grouped path on tcgen05 tensor cores (M >= 17 rows, K % 128 == 0, N % 16 == 0).
// starts == nullptr: plain linear; else grouped over E experts (packed [E,N,K/2]).
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int x_dtype, int y_dtype);
size_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemm_tc(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                   const float* scales, const float* zps, void* y, int y_dtype, int64_t M,
                   int64_t N, int64_t K, const int32_t* starts, const int32_t* ends, int E,
                   void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st, int gated = 0,
                   const int32_t* emap = nullptr, int n_wexperts = 0);

}  // namespace b200q
