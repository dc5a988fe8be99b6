// Device helpers shared by the decode kernels (gemv_dec.cu: exact-integer IMMA form, gemv_hm.cu: fp16 HMMA form).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

__device__ __forceinline__ void load8f(const void* x, int dtype, int64_t idx, float (&v)[8]) {
    if (dtype == B200Q_F32) {
        const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + idx);
        const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + idx + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (dtype == B200Q_F16) {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const __half*>(x) + idx);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    } else {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(x) + idx);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}

__device__ __forceinline__ float load1f(const void* x, int dtype, int64_t idx) {
    if (dtype == B200Q_F32) return static_cast<const float*>(x)[idx];
    if (dtype == B200Q_F16) return __half2float(static_cast<const __half*>(x)[idx]);
    return __bfloat162float(static_cast<const __nv_bfloat16*>(x)[idx]);
}

__device__ __forceinline__ void store_out(void* y, int dtype, int64_t idx, float v) {
    if (dtype == B200Q_F32) static_cast<float*>(y)[idx] = v;
    else if (dtype == B200Q_F16) static_cast<__half*>(y)[idx] = __float2half_rn(v);
    else static_cast<__nv_bfloat16*>(y)[idx] = __float2bfloat16_rn(v);
}

// D(16x8,s32) += A(16x32,u8,row) * B(32x8,s8,col)      SASS: IMMA.16832.U8.S8
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                     uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ... unsigned x unsigned
__device__ __forceinline__ void imma_uu(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void red_add_s32(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// TMA tensor-box loads with an L2 eviction hint (the weights are read once)
__device__ __forceinline__ void tma_box_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_box_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol)
        : "memory");
}

// A fragment of IMMA m16n8k32 straight from the swizzled tile: four 8 x 16-byte matrices (rows 0-7 / 8-15 of two
// adjacent 16-byte columns) land in a0..a3 of every lane.  SASS: LDSM.16.M88.4
__device__ __forceinline__ void ldsm_x4(uint32_t (&a)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                 : "r"(addr));
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

}  // namespace b200q
