// Quantisation format kernels: per-row asymmetric INT4 quantise / pack, dequantise, min-max.
//
// Bit-exactness contract (reference python/quantize.py:38-124, 127-173): every arithmetic step
// below is the IEEE-754 round-to-nearest fp32 operation torch performs on the CPU, in the same
// order, with no FMA contraction and no fast-math substitutions -- this TU must never be built
// with --use_fast_math.
#include <cfloat>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

namespace {

constexpr int QTHREADS = 256;

__device__ __forceinline__ float clamp_like_torch(float v, float lo, float hi) {
    // torch.clamp keeps -0.0 (python/quantize.py:101 on an all-zero row yields zp = -0.0)
    return v < lo ? lo : (v > hi ? hi : v);
}

__device__ __forceinline__ uint32_t quant1(float w, float scale, float zp) {
    // python/quantize.py:106-109: clamp(round(w / scale + zp), 0, 15)
    float q = rintf(__fadd_rn(__fdiv_rn(w, scale), zp));
    q = clamp_like_torch(q, 0.0f, 15.0f);
    return static_cast<uint32_t>(q);
}

// torch.min / torch.max return NaN when the row holds one (python/quantize.py:74-75); fminf / fmaxf would drop it
__device__ __forceinline__ float min_nan(float a, float b) {
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One CTA per weight row.  Pass 1: row min / max.  Pass 2 (row is L2/L1 resident): quantise, pack.
template <bool GIVEN>
__global__ void __launch_bounds__(QTHREADS)
quantize_rows_kernel(const float* __restrict__ w, int64_t K, const float* __restrict__ gscales,
                     const float* __restrict__ gzps, uint8_t* __restrict__ packed,
                     float* __restrict__ scales, float* __restrict__ zps) {
    const int64_t row = blockIdx.x;
    const float* wr = w + row * K;
    uint8_t* pr = packed + row * (K / 2);
    __shared__ float s_min[QTHREADS / 32], s_max[QTHREADS / 32];
    __shared__ float s_scale, s_zp;
    const bool vec = (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(wr) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(pr) & 3) == 0);

    if (!GIVEN) {
        float mn = INFINITY, mx = -INFINITY;       // (a row of +inf has min = +inf)
        if (vec) {
            const float4* w4 = reinterpret_cast<const float4*>(wr);
            for (int64_t i = threadIdx.x; i < K / 4; i += QTHREADS) {
                float4 v = w4[i];
                mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
                mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
            }
        } else {
            for (int64_t i = threadIdx.x; i < K; i += QTHREADS) {
                float v = wr[i];
                mn = min_nan(mn, v);
                mx = max_nan(mx, v);
            }
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) {
            s_min[threadIdx.x >> 5] = mn;
            s_max[threadIdx.x >> 5] = mx;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < QTHREADS / 32; ++i) {
                mn = min_nan(mn, s_min[i]);
                mx = max_nan(mx, s_max[i]);
            }
            // python/quantize.py:80-94
            float scale = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
            if (mx == mn) scale = __fdiv_rn(fmaxf(fabsf(mx), 1.0f), 15.0f);
            scale = scale < 1e-8f ? 1e-8f : scale;          // torch.clamp(min=1e-8): NaN stays NaN (:94)
            // python/quantize.py:100-101
            float zp = rintf(__fdiv_rn(-mn, scale));
            zp = clamp_like_torch(zp, 0.0f, 15.0f);
            s_scale = scale;
            s_zp = zp;
            scales[row] = scale;
            zps[row] = zp;
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) {
            s_scale = gscales[row];
            s_zp = gzps[row];
        }
        __syncthreads();
    }
    const float scale = s_scale, zp = s_zp;

    if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        uint32_t* p4 = reinterpret_cast<uint32_t*>(pr);
        for (int64_t i = threadIdx.x; i < K / 8; i += QTHREADS) {
            float4 a = w4[2 * i], b = w4[2 * i + 1];
            uint32_t word = quant1(a.x, scale, zp) | (quant1(a.y, scale, zp) << 4) |
                            (quant1(a.z, scale, zp) << 8) | (quant1(a.w, scale, zp) << 12) |
                            (quant1(b.x, scale, zp) << 16) | (quant1(b.y, scale, zp) << 20) |
                            (quant1(b.z, scale, zp) << 24) | (quant1(b.w, scale, zp) << 28);
            p4[i] = word;
        }
    } else {
        for (int64_t i = threadIdx.x; i < K / 2; i += QTHREADS) {
            uint32_t lo = quant1(wr[2 * i], scale, zp), hi = quant1(wr[2 * i + 1], scale, zp);
            pr[i] = static_cast<uint8_t>((hi << 4) | lo);  // python/quantize.py:120-122
        }
    }
}

// python/quantize.py:152-172: ((float)nibble - zp) * scale, subtract then multiply, fp32.
__global__ void __launch_bounds__(QTHREADS)
dequantize_rows_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ scales,
                       const float* __restrict__ zps, int64_t K, float* __restrict__ out) {
    const int64_t row = blockIdx.y;
    const float s = scales[row], z = zps[row];
    const uint8_t* pr = packed + row * (K / 2);
    float* o = out + row * K;
    const bool vec = (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(pr) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
    if (vec) {
        const uint32_t* p4 = reinterpret_cast<const uint32_t*>(pr);
        float4* o4 = reinterpret_cast<float4*>(o);
        for (int64_t i = blockIdx.x * (int64_t)QTHREADS + threadIdx.x; i < K / 8;
             i += (int64_t)gridDim.x * QTHREADS) {
            uint32_t wd = p4[i];
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = __fmul_rn(__fsub_rn(static_cast<float>((wd >> (4 * j)) & 15u), z), s);
            o4[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            o4[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)QTHREADS + threadIdx.x; i < K / 2;
             i += (int64_t)gridDim.x * QTHREADS) {
            uint32_t b = pr[i];
            o[2 * i] = __fmul_rn(__fsub_rn(static_cast<float>(b & 15u), z), s);
            o[2 * i + 1] = __fmul_rn(__fsub_rn(static_cast<float>(b >> 4), z), s);
        }
    }
}

// Global min / max.  ws layout: float part_min[1024], part_max[1024], uint32 counter (zeroed).
constexpr int MM_MAX_BLOCKS = 1024;

__global__ void __launch_bounds__(QTHREADS)
minmax_kernel(const float* __restrict__ v, int64_t count, float* __restrict__ out, float* ws) {
    float* part_min = ws;
    float* part_max = ws + MM_MAX_BLOCKS;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + 2 * MM_MAX_BLOCKS);
    __shared__ float s_min[QTHREADS / 32], s_max[QTHREADS / 32];
    __shared__ bool s_last;
    float mn = INFINITY, mx = -INFINITY;       // (a row of +inf has min = +inf)
    for (int64_t i = blockIdx.x * (int64_t)QTHREADS + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * QTHREADS) {
        float x = v[i];
        mn = min_nan(mn, x);
        mx = max_nan(mx, x);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = mn;
        s_max[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < QTHREADS / 32; ++i) {
            mn = min_nan(mn, s_min[i]);
            mx = max_nan(mx, s_max[i]);
        }
        part_min[blockIdx.x] = mn;
        part_max[blockIdx.x] = mx;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        mn = INFINITY;
        mx = -INFINITY;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += QTHREADS) {
            mn = min_nan(mn, __ldcg(part_min + i));
            mx = max_nan(mx, __ldcg(part_max + i));
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) {
            s_min[threadIdx.x >> 5] = mn;
            s_max[threadIdx.x >> 5] = mx;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < QTHREADS / 32; ++i) {
                mn = min_nan(mn, s_min[i]);
                mx = max_nan(mx, s_max[i]);
            }
            out[0] = mn;
            out[1] = mx;
            *counter = 0u;  // leave the workspace zeroed for the next call
        }
    }
}

}  // namespace

int launch_quantize_rows(const float* w, int64_t N, int64_t K, const float* given_scales,
                         const float* given_zps, uint8_t* packed, float* scales, float* zps,
                         cudaStream_t st) {
    if (N == 0 || K == 0) return 0;
    if (N > 0x7fffffffLL) return set_error(B200Q_EINVAL, "N too large");
    dim3 grid(static_cast<unsigned>(N));
    if (given_scales)
        quantize_rows_kernel<true><<<grid, QTHREADS, 0, st>>>(w, K, given_scales, given_zps, packed,
                                                              nullptr, nullptr);
    else
        quantize_rows_kernel<false><<<grid, QTHREADS, 0, st>>>(w, K, nullptr, nullptr, packed,
                                                               scales, zps);
    return check_cuda(cudaGetLastError(), "quantize_rows launch");
}

int launch_dequantize_rows(const uint8_t* packed, const float* scales, const float* zps, int64_t N,
                           int64_t K, float* out, cudaStream_t st) {
    if (N == 0 || K == 0) return 0;
    int64_t per_row = (K / 8 + QTHREADS - 1) / QTHREADS;
    if (per_row < 1) per_row = 1;
    if (per_row > 8) per_row = 8;
    for (int64_t n0 = 0; n0 < N; n0 += 65535) {
        int64_t rows = N - n0 < 65535 ? N - n0 : 65535;
        dim3 grid(static_cast<unsigned>(per_row), static_cast<unsigned>(rows));
        dequantize_rows_kernel<<<grid, QTHREADS, 0, st>>>(packed + n0 * (K / 2), scales + n0,
                                                          zps + n0, K, out + n0 * K);
    }
    return check_cuda(cudaGetLastError(), "dequantize_rows launch");
}

int launch_minmax(const float* v, int64_t count, float* out, void* ws, cudaStream_t st) {
    int64_t blocks = (count + QTHREADS * 8 - 1) / (QTHREADS * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > MM_MAX_BLOCKS) blocks = MM_MAX_BLOCKS;
    minmax_kernel<<<static_cast<unsigned>(blocks), QTHREADS, 0, st>>>(v, count, out,
                                                                      static_cast<float*>(ws));
    return check_cuda(cudaGetLastError(), "minmax launch");
}

}  // namespace b200q

extern "C" size_t b200q_minmax_ws_bytes(void) {
    return (2 * b200q::MM_MAX_BLOCKS + 4) * sizeof(float);
}
