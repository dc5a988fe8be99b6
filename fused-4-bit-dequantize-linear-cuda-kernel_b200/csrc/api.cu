// C ABI entry points of libb200q.so (declared in include/b200q.h): argument validation and
// dispatch to the kernel launchers.  Nothing here allocates device memory or synchronises.
#include "internal.h"

using namespace b200q;

namespace {

inline int elem_size(int dtype) {
    switch (dtype) {
        case B200Q_F32: return 4;
        case B200Q_F16: return 2;
        case B200Q_BF16: return 2;
    }
    return 0;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace

extern "C" {

int b200q_quantize_rows(const float* w, int64_t N, int64_t K, uint8_t* packed, float* scales,
                        float* zps, void* stream) {
    if (N < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "quantize_rows: need N >= 0 and even K >= 0 (N=%lld K=%lld)", (long long)N, (long long)K);
    if (K > (1 << 20)) return set_error(B200Q_EINVAL, "quantize_rows: K > 2^20");
    if (N == 0) return 0;
    if (!scales || !zps || (K > 0 && (!w || !packed))) return set_error(B200Q_EINVAL, "quantize_rows: null pointer");
    if (!aligned(w, 4) || !aligned(scales, 4) || !aligned(zps, 4)) return set_error(B200Q_EALIGN, "quantize_rows: fp32 pointers must be 4-byte aligned");
    if (K == 0) return set_error(B200Q_EINVAL, "quantize_rows: K == 0 has no min/max");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    return launch_quantize_rows(w, N, K, nullptr, nullptr, packed, scales, zps, static_cast<cudaStream_t>(stream));
}

int b200q_quantize_rows_given(const float* w, int64_t N, int64_t K, const float* scales,
                              const float* zps, uint8_t* packed, void* stream) {
    if (N < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "quantize_rows_given: need N >= 0 and even K >= 0");
    if (N == 0 || K == 0) return 0;
    if (!w || !scales || !zps || !packed) return set_error(B200Q_EINVAL, "quantize_rows_given: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    return launch_quantize_rows(w, N, K, scales, zps, packed, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int b200q_minmax(const float* v, int64_t count, float* out_minmax, void* ws, size_t ws_bytes,
                 void* stream) {
    if (count <= 0 || !v || !out_minmax) return set_error(B200Q_EINVAL, "minmax: need count > 0 and non-null pointers");
    if (!ws || ws_bytes < b200q_minmax_ws_bytes()) return set_error(B200Q_EWORKSPACE, "minmax: workspace too small");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    return launch_minmax(v, count, out_minmax, ws, static_cast<cudaStream_t>(stream));
}

int b200q_dequantize_rows(const uint8_t* packed, const float* scales, const float* zps, int64_t N,
                          int64_t K, float* out, void* stream) {
    if (N < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "dequantize_rows: need N >= 0 and even K >= 0");
    if (N == 0 || K == 0) return 0;
    if (!packed || !scales || !zps || !out) return set_error(B200Q_EINVAL, "dequantize_rows: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    return launch_dequantize_rows(packed, scales, zps, N, K, out, static_cast<cudaStream_t>(stream));
}

size_t b200q_linear_ws_bytes(int64_t M, int64_t N, int64_t K) {
    size_t a = gemv_ws_bytes(M, N, K);
    size_t b = gemm_tc_ws_bytes(M, N, K);
    return b > a ? b : a;
}

int b200q_linear_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                     const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                     void* ws, size_t ws_bytes, unsigned flags, void* stream) {
    return b200q_linear_bias_fwd(x, x_dtype, packed, scales, zps, nullptr, y, y_dtype, M, N, K, ws, ws_bytes, flags, stream,
                                 nullptr, 0);
}

int b200q_linear_fwd_next(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream,
                          const uint8_t* next_packed, size_t next_bytes) {
    return b200q_linear_bias_fwd(x, x_dtype, packed, scales, zps, nullptr, y, y_dtype, M, N, K, ws, ws_bytes, flags, stream,
                                 next_packed, next_bytes);
}

int b200q_linear_bias_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream,
                          const uint8_t* next_packed, size_t next_bytes) {
    if (M < 0 || N < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "linear_fwd: need M,N >= 0 and even K >= 0 (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
    if (!elem_size(x_dtype) || !elem_size(y_dtype)) return set_error(B200Q_EINVAL, "linear_fwd: unsupported dtype (x=%d y=%d)", x_dtype, y_dtype);
    if (M == 0 || N == 0) return 0;
    if (!y || !scales || !zps || (K > 0 && (!x || !packed))) return set_error(B200Q_EINVAL, "linear_fwd: null pointer");
    if (!aligned(x, elem_size(x_dtype)) || !aligned(y, elem_size(y_dtype))) return set_error(B200Q_EALIGN, "linear_fwd: x / y not aligned to their element size");
    if (bias && !aligned(bias, 4)) return set_error(B200Q_EALIGN, "linear_fwd: bias must be 4-byte aligned");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int force = tuning().force_path;
    const bool vec_ok = aligned(x, 16) && aligned(packed, 16) && aligned(y, 16);
    // decode (M <= 16): the CTA-resident IMMA kernel (gemv_dec.cu) whenever ceil(N / SMs) rows fit in shared memory
    // (wide K with many batch rows would need > 4 passes of two rows: the tcgen05 GEMM is faster there)
    // (K > 8192 with M >= 3 and fewer tile buffers than tiles -- Mixtral's down projection -- refills only in the last
    // pass over a buffer: measured slower than the ring kernel below, 32 vs 28 us at M = 4)
    // (measured crossovers with the tcgen05 GEMM, tools/dec_tune.py: 11008 -> 4096 M = 9 27.8 vs 31.9 us, M = 12 34.6 vs 32.1)
    // 3 <= M <= 16 with every tile resident: the fp16 HMMA decode kernel (gemv_hm.cu; tuning key hm_min_m moves the crossover)
    if ((force <= 0 || force == 7) && vec_ok && (x_dtype != B200Q_F32 || aligned(x, 32)) && M <= tuning().hm_max_m && (force == 7 || M >= tuning().hm_min_m) && gemv_hm_supported(d, M, N, K))
        return launch_gemv_hm(d, x, x_dtype, packed, scales, zps, bias, y, y_dtype, M, N, K, flags, st, next_packed, next_bytes);
    const bool dec_res = gemv_dec_resident(d, M, N, K);
    if ((force <= 0 || force == 6) && vec_ok && gemv_dec_supported(d, M, N, K) &&
        (force == 6 || ((M <= 8 || K <= 8192 || (M <= 10 && dec_res)) && (M <= 2 || K <= 8192 || dec_res))))
        return launch_gemv_dec(d, x, x_dtype, packed, scales, zps, bias, y, y_dtype, M, N, K, flags, st, next_packed, next_bytes);
    int rc = 0;
    bool done = false;
    // ... else the ring kernel (M <= 8: e.g. Mixtral's 14336-wide projections, 198 KB of weights per SM)
    // (14336 -> 4096: 24 / 28 / 40 / 49 us at M = 3 / 4 / 5 / 8 against 38 us for the 32-token tiles of the tcgen05 GEMM)
    if (!done && (force <= 0 || force == 2) && vec_ok && gemv_supported(M, N, K, x_dtype, &d) &&
        (force == 2 || M <= 4 || K <= 8192 || !gemm_tc_supported(M, N, K, x_dtype, y_dtype))) {
        rc = launch_gemv(d, x, x_dtype, packed, scales, zps, y, y_dtype, M, N, K, ws, ws_bytes, flags, st, next_packed, next_bytes);
        if (!rc) rc = launch_nonfinite_fixup(x, x_dtype, packed, scales, zps, nullptr, y, y_dtype, M, N, K, nullptr, nullptr, 1, 0, st);
        done = true;
    }
    // prefill: tcgen05 GEMM (rows with NaN / Inf are recomputed by its fix-up pass)
    if (!done && (force <= 0 || force == 3) && vec_ok && gemm_tc_supported(M, N, K, x_dtype, y_dtype)) {
        rc = launch_gemm_tc(d, x, x_dtype, packed, scales, zps, y, y_dtype, M, N, K, nullptr, nullptr, 1, ws, ws_bytes, flags, st);
        done = true;
    }
    if (!done && force > 1) return set_error(B200Q_EINVAL, "linear_fwd: forced path %d does not support this shape / alignment", force);
    if (!done) rc = launch_linear_generic(x, x_dtype, packed, scales, zps, y, y_dtype, M, N, K, nullptr, nullptr, 1, 0, st);
    if (!rc && bias) rc = launch_bias_add(y, y_dtype, bias, M, N, st);
    return rc;
}

int b200q_linear_groupwise_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                               int64_t group_size, void* y, int y_dtype, int64_t M, int64_t N, int64_t K, void* stream) {
    return b200q_linear_groupwise_bias_fwd(x, x_dtype, packed, scales, zps, nullptr, group_size, y, y_dtype, M, N, K, 0u, stream, nullptr, 0);
}

int b200q_linear_groupwise_bias_fwd(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                                    const float* bias, int64_t group_size, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                                    unsigned flags, void* stream, const uint8_t* next_packed, size_t next_bytes) {
    if (M < 0 || N < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "linear_groupwise_fwd: need M,N >= 0 and even K >= 0");
    if (group_size <= 0 || (group_size & 7) || K % group_size) return set_error(B200Q_EINVAL, "linear_groupwise_fwd: group_size must be a multiple of 8 that divides K (got %lld, K=%lld)", (long long)group_size, (long long)K);
    if (!elem_size(x_dtype) || !elem_size(y_dtype)) return set_error(B200Q_EINVAL, "linear_groupwise_fwd: unsupported dtype");
    if (M == 0 || N == 0) return 0;
    if (!y || !scales || !zps || !x || !packed) return set_error(B200Q_EINVAL, "linear_groupwise_fwd: null pointer");
    if (bias && !aligned(bias, 4)) return set_error(B200Q_EALIGN, "linear_groupwise_fwd: bias must be 4-byte aligned");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    // decode-sized batches with groups of 128, 256, ... columns: the mid-batch decode kernel (the two halves of every
    // 256-column pair meet their own scale / zero point); everything else the reference-speed SIMT kernel
    if (tuning().force_path != 1 && M <= 32 && group_size % 128 == 0 && aligned(x, x_dtype == B200Q_F32 ? 32 : 16) && aligned(packed, 16) &&
        gemv_hm_supported(d, M <= 8 ? M : 8, N, K))
        return launch_gemv_hm(d, x, x_dtype, packed, scales, zps, bias, y, y_dtype, M, N, K, flags,
                              static_cast<cudaStream_t>(stream), next_packed, next_bytes, 0, nullptr, 1, nullptr, (int)group_size);
    int rc = launch_linear_generic(x, x_dtype, packed, scales, zps, y, y_dtype, M, N, K, nullptr, nullptr, 1, 0,
                                   static_cast<cudaStream_t>(stream), nullptr, (int)group_size);
    if (!rc && bias) rc = launch_bias_add(y, y_dtype, bias, M, N, static_cast<cudaStream_t>(stream));
    return rc;
}

int b200q_linear_gated_fwd(const void* x, int x_dtype, const uint8_t* packed13, const float* scales13, const float* zps13,
                           void* h, int h_dtype, int64_t M, int64_t F, int64_t K, void* ws, size_t ws_bytes, unsigned flags,
                           void* stream, const uint8_t* next_packed, size_t next_bytes) {
    if (M < 0 || F < 0 || K < 0 || (K & 1)) return set_error(B200Q_EINVAL, "linear_gated_fwd: need M,F >= 0 and even K >= 0");
    if (!elem_size(x_dtype) || !elem_size(h_dtype)) return set_error(B200Q_EINVAL, "linear_gated_fwd: unsupported dtype");
    if (M == 0 || F == 0) return 0;
    if (!h || !scales13 || !zps13 || !x || !packed13) return set_error(B200Q_EINVAL, "linear_gated_fwd: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!(aligned(x, 16) && aligned(packed13, 16) && aligned(h, 16)))
        return set_error(B200Q_EALIGN, "linear_gated_fwd: x, packed13 and h must be 16-byte aligned");
    const int force = tuning().force_path;
    if ((force <= 0 || force == 7) && (x_dtype != B200Q_F32 || aligned(x, 32)) && M <= tuning().hm_max_m && (force == 7 || M >= tuning().hm_min_m) && gemv_hm_supported(d, M, 2 * F, K, 1))
        return launch_gemv_hm(d, x, x_dtype, packed13, scales13, zps13, nullptr, h, h_dtype, M, 2 * F, K, flags, st, next_packed, next_bytes, 1);
    if ((force <= 0 || force == 6) && gemv_dec_supported(d, M, 2 * F, K, 1) && (force == 6 || M <= 8 || K <= 8192))
        return launch_gemv_dec(d, x, x_dtype, packed13, scales13, zps13, nullptr, h, h_dtype, M, 2 * F, K, flags, st, next_packed, next_bytes, 1);
    if (force != 6 && gemm_tc_supported(M, 2 * F, K, x_dtype, h_dtype))
        return launch_gemm_tc(d, x, x_dtype, packed13, scales13, zps13, h, h_dtype, M, 2 * F, K, nullptr, nullptr, 1, ws, ws_bytes, flags, st, 1);
    return set_error(B200Q_EINVAL, "linear_gated_fwd: needs K %% 128 == 0 (otherwise: b200q_linear_fwd on the concatenated rows, then b200q_moe_silu_mul)");
}

// ---- MoE decode (T <= 16 tokens): the whole routed gated layer behind one C call --------------------------------
namespace {
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
struct DecodeWs {
    size_t idx, wts, counts, offsets, sorted_slot, inv_perm, perm_ws, h, y, total;
};
DecodeWs decode_ws(int64_t T, int E, int k, int64_t d, int64_t F, int esz) {
    DecodeWs w{};
    const size_t A = (size_t)T * k;
    size_t o = 0;
    w.idx = o; o += up256(A * 4);
    w.wts = o; o += up256(A * 4);
    w.counts = o; o += up256((size_t)E * 4);
    w.offsets = o; o += up256((size_t)(E + 1) * 4);
    w.sorted_slot = o; o += up256(A * 4);
    w.inv_perm = o; o += up256(A * 4);
    w.perm_ws = o; o += up256(A * 4);                         // token of every sorted position
    w.h = o; o += up256(A * F * esz);
    w.y = o; o += up256(A * d * esz);
    w.total = o;
    return w;
}
}  // namespace

size_t b200q_moe_decode_ws_bytes(int64_t T, int E, int k, int64_t d, int64_t F) {
    if (T <= 0 || E <= 0 || k <= 0 || d <= 0 || F <= 0) return 0;
    return decode_ws(T, E, k, d, F, 4).total;
}

int b200q_moe_decode_fwd(const void* x, int x_dtype, const float* logits, int64_t T, int E, int k,
                         const uint8_t* packed13, const float* scales13, const float* zps13,
                         const uint8_t* packed2, const float* scales2, const float* zps2, int64_t d, int64_t F,
                         float* out, void* ws, size_t ws_bytes, void* stream) {
    if (T <= 0 || T > 16 || E <= 0 || E > 256 || k <= 0 || k > 8 || k > E || d <= 0 || F <= 0)
        return set_error(B200Q_EINVAL, "moe_decode_fwd: need 1 <= T <= 16, 1 <= k <= min(8, E), E <= 256, d, F > 0");
    if (!elem_size(x_dtype)) return set_error(B200Q_EINVAL, "moe_decode_fwd: unsupported dtype");
    if (!x || !logits || !packed13 || !scales13 || !zps13 || !packed2 || !scales2 || !zps2 || !out || !ws) return set_error(B200Q_EINVAL, "moe_decode_fwd: null pointer");
    const DecodeWs w = decode_ws(T, E, k, d, F, elem_size(x_dtype));
    if (ws_bytes < w.total) return set_error(B200Q_EWORKSPACE, "moe_decode_fwd: workspace too small (%zu < %zu)", ws_bytes, w.total);
    if (!aligned(x, 16) || !aligned(ws, 256) || !aligned(packed13, 16) || !aligned(packed2, 16)) return set_error(B200Q_EALIGN, "moe_decode_fwd: x / packed must be 16-byte, ws 256-byte aligned");
    DeviceInfo dv;
    if (int rc = current_device(&dv)) return rc;
    // both expert GEMVs run on the resident decode kernel, grouped over the experts (device-side row offsets)
    if (!gemv_dec_supported(dv, T, 2 * F, d, 1) || !gemv_dec_supported(dv, T, d, F, 0))
        return set_error(B200Q_EINVAL, "moe_decode_fwd: shape not supported by the decode kernel (d, F multiples of 256, <= 16384)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* base = static_cast<uint8_t*>(ws);
    int32_t* idx = reinterpret_cast<int32_t*>(base + w.idx);
    float* wts = reinterpret_cast<float*>(base + w.wts);
    int32_t* counts = reinterpret_cast<int32_t*>(base + w.counts);
    int32_t* offsets = reinterpret_cast<int32_t*>(base + w.offsets);
    int32_t* sorted_slot = reinterpret_cast<int32_t*>(base + w.sorted_slot);
    int32_t* inv_perm = reinterpret_cast<int32_t*>(base + w.inv_perm);
    int32_t* src_token = reinterpret_cast<int32_t*>(base + w.perm_ws);
    void* h = base + w.h;
    void* y = base + w.y;
    // three launches + the combine: routing (one CTA), gate / up GEMV reading x in place through the token map, down GEMV
    if (int rc = moe_route_small(logits, T, E, k, idx, wts, counts, offsets, sorted_slot, inv_perm, src_token, st)) return rc;
    // from 1.5 rows per expert on average the mid-batch kernel (eight tokens per tensor instruction, gemv_hm.cu) serves the
    // groups; below that the exact-integer kernel (two rows per pass).  Mixtral layer, tools/moe_decode_crossover.py:
    // T = 4: 0.219 / 0.216 ms, T = 6: 0.250 / 0.221, T = 8: 0.337 / 0.258, T = 16: 0.552 / 0.298 (grouped tcgen05 GEMMs: 0.488)
    const int hm_mode = tuning().moe_dec_hm;
    const bool use_hm = hm_mode != 0 && (hm_mode > 0 || 2 * T * k >= 3 * (int64_t)E) && (x_dtype != B200Q_F32 || aligned(x, 32)) &&
                        gemv_hm_supported(dv, T < 8 ? T : 8, 2 * F, d, 1) && gemv_hm_supported(dv, T < 8 ? T : 8, d, F, 0);   // (the plan does not depend on the rows beyond the pass count)
    if (use_hm) {
        if (int rc = launch_gemv_hm(dv, x, x_dtype, packed13, scales13, zps13, nullptr, h, x_dtype, T, 2 * F, d, B200Q_FLAG_STATIC_WEIGHTS, st,
                                    nullptr, 0, 1, offsets, E, src_token)) return rc;
        if (int rc = launch_gemv_hm(dv, h, x_dtype, packed2, scales2, zps2, nullptr, y, x_dtype, T, d, F, B200Q_FLAG_STATIC_WEIGHTS, st,
                                    nullptr, 0, 0, offsets, E, nullptr)) return rc;
        return b200q_moe_combine(y, x_dtype, inv_perm, wts, T, k, d, out, B200Q_F32, stream);
    }
    // (T k routed rows hit at most T k experts: the grid has that many rows of CTAs, not E)
    const int max_groups = tuning().moe_dec_compact != 0 ? (int)std::min<int64_t>(E, T * k) : 0;
    if (int rc = launch_gemv_dec(dv, x, x_dtype, packed13, scales13, zps13, nullptr, h, x_dtype, T, 2 * F, d, B200Q_FLAG_STATIC_WEIGHTS, st,
                                 nullptr, 0, 1, offsets, E, src_token, max_groups)) return rc;
    if (int rc = launch_gemv_dec(dv, h, x_dtype, packed2, scales2, zps2, nullptr, y, x_dtype, T, d, F, B200Q_FLAG_STATIC_WEIGHTS, st,
                                 nullptr, 0, 0, offsets, E, nullptr, max_groups)) return rc;
    return b200q_moe_combine(y, x_dtype, inv_perm, wts, T, k, d, out, B200Q_F32, stream);
}

int b200q_linear_fwd_host(const void* h_x, int x_dtype, void* d_x, const uint8_t* packed, const float* scales,
                          const float* zps, void* d_y, void* h_y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          void* ws, size_t ws_bytes, unsigned flags, void* stream) {
    if (M < 0 || N < 0 || K < 0) return set_error(B200Q_EINVAL, "linear_fwd_host: negative size");
    if (!elem_size(x_dtype) || !elem_size(y_dtype)) return set_error(B200Q_EINVAL, "linear_fwd_host: unsupported dtype");
    if (M == 0 || N == 0) return 0;
    if (!h_x || !d_x || !d_y || !h_y) return set_error(B200Q_EINVAL, "linear_fwd_host: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Decode-sized activations (<= 256 KB) in device-addressable pinned memory: a small kernel pulls them over PCIe
    // into d_x (a copy node costs ~8 us of launch latency for 16 KB); everything else goes through cudaMemcpyAsync.
    const size_t xbytes = (size_t)M * K * elem_size(x_dtype);
    bool staged = false;
    if (M <= 16 && tuning().host_direct != 0 && xbytes <= (1u << 20) && xbytes % 16 == 0 && aligned(d_x, 16)) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, h_x) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
            aligned(at.devicePointer, 16)) {
            if (int rc = launch_stage_host(at.devicePointer, d_x, xbytes, st)) return rc;
            staged = true;
        } else {
            (void)cudaGetLastError();
        }
    }
    if (!staged) B200Q_CUDA(cudaMemcpyAsync(d_x, h_x, xbytes, cudaMemcpyHostToDevice, st));
    // Decode-sized results (M <= 8: tens of KB): when h_y is pinned memory that the device can address (any
    // cudaHostAlloc / cudaHostRegister allocation under unified addressing), the kernel's epilogue stores straight
    // into it over PCIe -- posted writes, ~1 us -- instead of a separate ~10 us copy node behind the kernel.
    if (M <= 16 && tuning().host_direct != 0) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, h_y) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
            aligned(at.devicePointer, 16))
            return b200q_linear_fwd(d_x, x_dtype, packed, scales, zps, at.devicePointer, y_dtype, M, N, K, ws, ws_bytes, flags, stream);
        (void)cudaGetLastError();       // pageable memory: not an error here, take the copy
    }
    if (int rc = b200q_linear_fwd(d_x, x_dtype, packed, scales, zps, d_y, y_dtype, M, N, K, ws, ws_bytes, flags, stream)) return rc;
    B200Q_CUDA(cudaMemcpyAsync(h_y, d_y, (size_t)M * N * elem_size(y_dtype), cudaMemcpyDeviceToHost, st));
    return 0;
}

size_t b200q_moe_grouped_ws_bytes(int64_t R, int E, int64_t N, int64_t K) {
    (void)E;
    return gemm_tc_ws_bytes(R, N, K);
}

static int grouped_impl(const void* xs, int x_dtype, const uint8_t* packed, const float* scales,
                        const float* zps, const int32_t* starts, const int32_t* ends, int E,
                        int zero_outside, void* y, int y_dtype, int64_t R, int64_t N, int64_t K,
                        void* ws, size_t ws_bytes, void* stream) {
    if (R < 0 || N < 0 || K < 0 || (K & 1) || E <= 0) return set_error(B200Q_EINVAL, "moe_grouped_fwd: need R,N >= 0, even K >= 0, E > 0");
    if (!elem_size(x_dtype) || !elem_size(y_dtype)) return set_error(B200Q_EINVAL, "moe_grouped_fwd: unsupported dtype");
    if (R == 0 || N == 0) return 0;
    if (!y || !scales || !zps || !starts || !ends || (K > 0 && (!xs || !packed))) return set_error(B200Q_EINVAL, "moe_grouped_fwd: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int force = tuning().force_path;
    const bool vec_ok = aligned(xs, 16) && aligned(packed, 16) && aligned(y, 16);
    if (force != 1 && vec_ok && gemm_tc_supported(R, N, K, x_dtype, y_dtype)) {
        // the GEMM only writes rows inside a group: honour "rows outside every group are zero-filled" first
        if (zero_outside)
            if (int rc = launch_zero_rows_outside(y, y_dtype, R, N, starts, ends + (E - 1), st)) return rc;
        return launch_gemm_tc(d, xs, x_dtype, packed, scales, zps, y, y_dtype, R, N, K, starts, ends, E, ws, ws_bytes, 0u, st);
    }
    return launch_linear_generic(xs, x_dtype, packed, scales, zps, y, y_dtype, R, N, K, starts, ends, E, zero_outside, st);
}

int b200q_moe_grouped_fwd(const void* xs, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, const int32_t* offsets, int E, void* y, int y_dtype,
                          int64_t R, int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream) {
    return grouped_impl(xs, x_dtype, packed, scales, zps, offsets, offsets ? offsets + 1 : nullptr, E, 1,
                        y, y_dtype, R, N, K, ws, ws_bytes, stream);
}

int b200q_moe_grouped_gated_fwd(const void* xs, int x_dtype, const uint8_t* packed13, const float* scales13,
                                const float* zps13, const int32_t* offsets, int E, void* h, int h_dtype,
                                int64_t R, int64_t F, int64_t K, void* ws, size_t ws_bytes, void* stream) {
    if (R < 0 || F < 0 || K < 0 || E <= 0) return set_error(B200Q_EINVAL, "moe_grouped_gated_fwd: need R,F,K >= 0, E > 0");
    if (!elem_size(x_dtype) || !elem_size(h_dtype)) return set_error(B200Q_EINVAL, "moe_grouped_gated_fwd: unsupported dtype");
    if (R == 0 || F == 0) return 0;
    if (!h || !scales13 || !zps13 || !offsets || !xs || !packed13) return set_error(B200Q_EINVAL, "moe_grouped_gated_fwd: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    if (!(aligned(xs, 16) && aligned(packed13, 16) && aligned(h, 16)) || !gemm_tc_supported(R, 2 * F, K, x_dtype, h_dtype))
        return set_error(B200Q_EINVAL, "moe_grouped_gated_fwd: needs K %% 128 == 0 and 16-byte aligned buffers "
                                       "(otherwise: b200q_moe_grouped_fwd on w1||w3, then b200q_moe_silu_mul)");
    return launch_gemm_tc(d, xs, x_dtype, packed13, scales13, zps13, h, h_dtype, R, 2 * F, K, offsets, offsets + 1, E, ws, ws_bytes,
                          0u, static_cast<cudaStream_t>(stream), 1);
}

int b200q_moe_grouped_fwd_ranges(const void* xs, int x_dtype, const uint8_t* packed,
                                 const float* scales, const float* zps, const int32_t* starts,
                                 const int32_t* counts_end, int E, void* y, int y_dtype, int64_t R,
                                 int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream) {
    return grouped_impl(xs, x_dtype, packed, scales, zps, starts, counts_end, E, 0, y, y_dtype, R, N, K,
                        ws, ws_bytes, stream);
}

}  // extern "C"
