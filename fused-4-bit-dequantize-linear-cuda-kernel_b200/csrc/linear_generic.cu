// Generic SIMT fused dequantize + linear: any even K, any M / N, any supported dtype, optional
// grouped (per-expert) mode.  This is the shape-agnostic path of libb200q (odd alignments, tiny
// problems); the Llama / Mixtral shapes are served by gemv.cu and gemm_tc.cu.
//
// Math mirrors the reference exactly (python/quantize.py:172, 202): w = (q - zp) * s in fp32
// (subtract, then multiply), y = sum_k w * x accumulated in fp32.
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

namespace {

constexpr int GW = 8;   // warps per CTA, one weight row each
constexpr int MT = 8;   // x rows per CTA

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

struct GroupArgs {
    const int32_t* starts;   // [E] first row of every group (device), or nullptr for the plain linear
    const int32_t* ends;     // [E] one past the last row of every group
    int E;
    int zero_outside;        // offsets form: rows before the first / after the last group are zeroed
    const int32_t* emap;     // weight expert of every group (nullptr: group e uses expert e)
    int kgroup;              // 0: one scale / zero point per weight row; G > 0: one per G columns (scales [N, K/G])
    int64_t wstride;   // packed bytes per expert
    int64_t sstride;   // scales per expert
};

// Tile -> (expert, first row, row count).  expert == -1: rows outside every group (zero fill).
__device__ __forceinline__ bool locate_tile(const GroupArgs& g, int64_t R, int tile, int& expert,
                                            int64_t& m0, int& rows) {
    if (!g.starts) {
        m0 = (int64_t)tile * MT;
        if (m0 >= R) return false;
        rows = (int)min((int64_t)MT, R - m0);
        expert = 0;
        return true;
    }
    int t = tile;
    // leading rows [0, offsets[0]) and trailing rows [offsets[E], R) belong to no expert
    for (int e = -1; e <= g.E; ++e) {
        int64_t lo, hi;
        if (e == -1) { if (!g.zero_outside) continue; lo = 0; hi = min((int64_t)max(g.starts[0], 0), R); }
        else if (e == g.E) { if (!g.zero_outside) continue; lo = min((int64_t)max(g.ends[g.E - 1], 0), R); hi = R; }
        else { lo = g.starts[e]; hi = g.ends[e]; lo = max(lo, (int64_t)0); hi = min(hi, R); }
        int64_t cnt = hi - lo;
        if (cnt <= 0) continue;
        int ntile = (int)((cnt + MT - 1) / MT);
        if (t < ntile) {
            expert = (e == g.E) ? -1 : e;
            m0 = lo + (int64_t)t * MT;
            rows = (int)min((int64_t)MT, hi - m0);
            return true;
        }
        t -= ntile;
    }
    return false;
}

template <typename XT, typename YT, bool VEC>
__global__ void __launch_bounds__(GW * 32)
linear_generic_kernel(const XT* __restrict__ x, const uint8_t* __restrict__ packed,
                      const float* __restrict__ scales, const float* __restrict__ zps,
                      YT* __restrict__ y, int64_t R, int64_t N, int64_t K, GroupArgs g) {
    int expert, rows;
    int64_t m0;
    if (!locate_tile(g, R, blockIdx.y, expert, m0, rows)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * GW + warp;
    if (n >= N) return;
    if (expert < 0) {
        if (lane < rows) y[(m0 + lane) * N + n] = from_f32<YT>(0.0f);
        return;
    }
    if (g.emap) expert = g.emap[expert];
    const uint8_t* pr = packed + expert * g.wstride + n * (K / 2);
    const int64_t ngrp = g.kgroup ? K / g.kgroup : 1;
    const float* srow = scales + (expert * g.sstride + n) * ngrp;
    const float* zrow = zps + (expert * g.sstride + n) * ngrp;
    float s = srow[0], z = zrow[0];
    const XT* xb = x + m0 * K;

    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.0f;

    if (VEC) {
        const uint32_t* p4 = reinterpret_cast<const uint32_t*>(pr);
        for (int64_t i = lane; i < K / 8; i += 32) {
            const uint32_t wd = __ldg(p4 + i);
            if (g.kgroup) { s = srow[(8 * i) / g.kgroup]; z = zrow[(8 * i) / g.kgroup]; }
            float wf[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                wf[j] = __fmul_rn(__fsub_rn(static_cast<float>((wd >> (4 * j)) & 15u), z), s);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (m < rows) {
                    float xv[8];
                    load8<XT>(xb + m * K + 8 * i, xv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[m] = fmaf(wf[j], xv[j], acc[m]);
                }
            }
        }
    } else {
        for (int64_t i = lane; i < K / 2; i += 32) {
            const uint32_t b = pr[i];
            if (g.kgroup) { s = srow[(2 * i) / g.kgroup]; z = zrow[(2 * i) / g.kgroup]; }
            const float w0 = __fmul_rn(__fsub_rn(static_cast<float>(b & 15u), z), s);
            const float w1 = __fmul_rn(__fsub_rn(static_cast<float>(b >> 4), z), s);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (m < rows) {
                    acc[m] = fmaf(w0, to_f32<XT>(xb[m * K + 2 * i]), acc[m]);
                    acc[m] = fmaf(w1, to_f32<XT>(xb[m * K + 2 * i + 1]), acc[m]);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
        if (lane == m && m < rows) y[(m0 + m) * N + n] = from_f32<YT>(acc[m]);
}

template <typename XT, typename YT>
int launch_typed(const void* x, const uint8_t* packed, const float* scales, const float* zps,
                 void* y, int64_t R, int64_t N, int64_t K, const int32_t* starts, const int32_t* ends,
                 int E, int zero_outside, cudaStream_t st, const int32_t* emap, int kgroup) {
    GroupArgs g{starts, ends, E, zero_outside, emap, kgroup, N * (K / 2), N};
    int64_t tiles = (R + MT - 1) / MT + (starts ? E + 2 : 0);
    if (tiles > 65535) return set_error(B200Q_EINVAL, "generic path: too many row tiles (%lld)", (long long)tiles);
    dim3 grid(static_cast<unsigned>((N + GW - 1) / GW), static_cast<unsigned>(tiles));
    const bool vec = (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(packed) & 3) == 0);
    if (vec)
        linear_generic_kernel<XT, YT, true><<<grid, GW * 32, 0, st>>>(
            static_cast<const XT*>(x), packed, scales, zps, static_cast<YT*>(y), R, N, K, g);
    else
        linear_generic_kernel<XT, YT, false><<<grid, GW * 32, 0, st>>>(
            static_cast<const XT*>(x), packed, scales, zps, static_cast<YT*>(y), R, N, K, g);
    return check_cuda(cudaGetLastError(), "linear_generic launch");
}

template <typename XT>
int launch_x(const void* x, const uint8_t* packed, const float* scales, const float* zps, void* y,
             int y_dtype, int64_t R, int64_t N, int64_t K, const int32_t* starts, const int32_t* ends,
             int E, int zero_outside, cudaStream_t st, const int32_t* emap, int kgroup) {
    switch (y_dtype) {
        case B200Q_F32: return launch_typed<XT, float>(x, packed, scales, zps, y, R, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
        case B200Q_F16: return launch_typed<XT, __half>(x, packed, scales, zps, y, R, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
        case B200Q_BF16: return launch_typed<XT, __nv_bfloat16>(x, packed, scales, zps, y, R, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
    }
    return set_error(B200Q_EINVAL, "bad y_dtype %d", y_dtype);
}

}  // namespace

int launch_linear_generic(const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                          const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                          const int32_t* starts, const int32_t* ends, int E, int zero_outside,
                          cudaStream_t st, const int32_t* emap, int kgroup) {
    if (M == 0 || N == 0) return 0;
    switch (x_dtype) {
        case B200Q_F32: return launch_x<float>(x, packed, scales, zps, y, y_dtype, M, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
        case B200Q_F16: return launch_x<__half>(x, packed, scales, zps, y, y_dtype, M, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
        case B200Q_BF16: return launch_x<__nv_bfloat16>(x, packed, scales, zps, y, y_dtype, M, N, K, starts, ends, E, zero_outside, st, emap, kgroup);
    }
    return set_error(B200Q_EINVAL, "bad x_dtype %d", x_dtype);
}

}  // namespace b200q
