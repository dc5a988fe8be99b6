// Small companions of the fast paths (all HBM / latency-bound element-wise work, no host synchronisation):
//   * nonfinite_fixup: rows of x that contain NaN / Inf are recomputed in the reference's arithmetic order
//     (w = (q - zp) * s in fp32, then fp32 multiply-add: python/quantize.py:172, 202), because the fast paths
//     compute sum q X - zp sum X, which turns +-Inf into NaN and (on the integer path) would drop NaN;
//   * bias_add: y[m, n] += bias[n] for the paths without a fused bias (python/module.py:84 forbids a bias; lifted here);
//   * zero_rows_outside: the "rows outside every group are zero-filled" clause of b200q_moe_grouped_fwd
//     (csrc/moe_int4_kernel.cu:109 zero-initialises the whole output).
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

namespace {

__device__ __forceinline__ float ld1(const void* x, int dtype, int64_t i) {
    if (dtype == B200Q_F32) return static_cast<const float*>(x)[i];
    if (dtype == B200Q_F16) return __half2float(static_cast<const __half*>(x)[i]);
    return __bfloat162float(static_cast<const __nv_bfloat16*>(x)[i]);
}
__device__ __forceinline__ void st1(void* y, int dtype, int64_t i, float v) {
    if (dtype == B200Q_F32) static_cast<float*>(y)[i] = v;
    else if (dtype == B200Q_F16) static_cast<__half*>(y)[i] = __float2half_rn(v);
    else static_cast<__nv_bfloat16*>(y)[i] = __float2bfloat16_rn(v);
}

struct FixParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    const int* nf;             // per-row flags or nullptr (scan x)
    void* y;
    const int32_t* starts;
    const int32_t* ends;
    const int32_t* emap;       // weight expert of every range (nullptr: identity)
    int E, gated, x_dtype, y_dtype;
    int64_t M, N, K;
};

// one warp: dot(dequant(row n of expert e), x[m, :]) in the reference's order per lane, lanes combined at the end
__device__ __forceinline__ float ref_dot(const FixParams& p, int e, int64_t n, int64_t m, int lane) {
    const int64_t rb = p.K >> 1;
    const uint8_t* wr = p.packed + ((int64_t)e * p.N + n) * rb;
    const float sc = p.scales[(int64_t)e * p.N + n], zp = p.zps[(int64_t)e * p.N + n];
    float acc = 0.0f;
    for (int64_t kb = lane; kb < rb; kb += 32) {
        const unsigned int b = wr[kb];
        const float w0 = __fmul_rn(__fsub_rn((float)(b & 15u), zp), sc), w1 = __fmul_rn(__fsub_rn((float)(b >> 4), zp), sc);
        acc = fmaf(w0, ld1(p.x, p.x_dtype, m * p.K + 2 * kb), acc);
        acc = fmaf(w1, ld1(p.x, p.x_dtype, m * p.K + 2 * kb + 1), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

__global__ void __launch_bounds__(256) nonfinite_fixup_kernel(const FixParams p) {
    // launched with programmatic stream serialisation: the kernel after this one may set itself up meanwhile; this
    // one only touches memory once the kernel whose rows it repairs has completed
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t m = blockIdx.x; m < p.M; m += gridDim.x) {
        int bad;
        if (p.nf) {
            bad = p.nf[m];                                   // uniform
        } else {
            int mine = 0;
            for (int64_t k = threadIdx.x; k < p.K; k += 256) {
                const float v = ld1(p.x, p.x_dtype, m * p.K + k);
                mine |= ((__float_as_uint(v) & 0x7fffffffu) >= 0x7f800000u) ? 1 : 0;
            }
            bad = __syncthreads_or(mine);
        }
        if (!bad) continue;
        int e = 0;
        if (p.starts) {
            e = -1;
            for (int i = 0; i < p.E; ++i)
                if (m >= p.starts[i] && m < p.ends[i]) { e = p.emap ? p.emap[i] : i; break; }
            if (e < 0) continue;
        }
        if (p.gated) {
            const int64_t F = p.N >> 1;
            for (int64_t f = warp; f < F; f += 8) {
                const float gv = ref_dot(p, e, 2 * f, m, lane), uv = ref_dot(p, e, 2 * f + 1, m, lane);
                if (lane == 0) st1(p.y, p.y_dtype, m * F + f, gv / (1.0f + __expf(-gv)) * uv);
            }
        } else {
            for (int64_t n = warp; n < p.N; n += 8) {
                const float v = ref_dot(p, e, n, m, lane);
                if (lane == 0) st1(p.y, p.y_dtype, m * p.N + n, v);
            }
        }
    }
}

__global__ void __launch_bounds__(256) bias_add_kernel(void* y, int y_dtype, const float* __restrict__ bias, int64_t M, int64_t N) {
    const int64_t total = M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        st1(y, y_dtype, i, ld1(y, y_dtype, i) + __ldg(bias + (i % N)));
}

__global__ void __launch_bounds__(256) zero_rows_outside_kernel(void* y, int y_dtype, int64_t R, int64_t N,
                                                                const int32_t* __restrict__ first, const int32_t* __restrict__ last) {
    const int64_t lo = min((int64_t)max(*first, 0), R), hi = min((int64_t)max(*last, 0), R);
    const int64_t head = lo * N, tail = (R - hi) * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < head + tail; i += (int64_t)gridDim.x * blockDim.x)
        st1(y, y_dtype, i < head ? i : hi * N + (i - head), 0.0f);
}

}  // namespace

int launch_nonfinite_fixup(const void* x, int x_dtype, const uint8_t* packed, const float* scales, const float* zps,
                           const int* nf_flags, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                           const int32_t* starts, const int32_t* ends, int E, int gated, cudaStream_t st,
                           const int32_t* emap) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    FixParams p{x, packed, scales, zps, nf_flags, y, starts, ends, emap, E, gated, x_dtype, y_dtype, M, N, K};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(M < 1184 ? M : 1184));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, nonfinite_fixup_kernel, p), "nonfinite_fixup launch");
}

int launch_bias_add(void* y, int y_dtype, const float* bias, int64_t M, int64_t N, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    int64_t blocks = (M * N + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    bias_add_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, y_dtype, bias, M, N);
    return check_cuda(cudaGetLastError(), "bias_add launch");
}

int launch_zero_rows_outside(void* y, int y_dtype, int64_t R, int64_t N, const int32_t* first, const int32_t* last,
                             cudaStream_t st) {
    if (R <= 0 || N <= 0) return 0;
    zero_rows_outside_kernel<<<148, 256, 0, st>>>(y, y_dtype, R, N, first, last);
    return check_cuda(cudaGetLastError(), "zero_rows_outside launch");
}

}  // namespace b200q
