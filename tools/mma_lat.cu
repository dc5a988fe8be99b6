// Dependent-chain latency of legacy mma.sync on sm_100a (one warp per SM, one accumulator), plus the latency of
// a red.shared.add and of an mbarrier try_wait on a completed phase.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k_lat(int iters, int* out, long long* cyc) {
    int ci[4] = {0, 0, 0, 0};
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
    __shared__ int s[64];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    }
    s[threadIdx.x] = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(ci[0]), "+r"(ci[1]), "+r"(ci[2]), "+r"(ci[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        else if (KIND == 1)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        else if (KIND == 2) {
            asm volatile("red.shared.add.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s[threadIdx.x])), "r"(it) : "memory");
        } else if (KIND == 3) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(0u) : "memory");
            ci[0] += ok;
        } else {
            int v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(&s[(threadIdx.x + ci[0]) & 31])));
            ci[0] += v;
        }
    }
    long long t1 = clock64();
    if (ci[0] + ci[1] + ci[2] + ci[3] + (int)(c[0] + c[1] + c[2] + c[3]) == 123456) out[0] = 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    int* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    const char* names[] = {"IMMA m16n8k32 u8*s8 dependent", "HMMA m16n8k16 f32 dependent", "red.shared.add.s32 (issue)", "mbarrier.try_wait (done phase)", "ld.shared dependent"};
    for (int kind = 0; kind < 5; ++kind) {
        long long h = 0;
        if (kind == 0) k_lat<0><<<148, 32>>>(iters, out, cyc);
        if (kind == 1) k_lat<1><<<148, 32>>>(iters, out, cyc);
        if (kind == 2) k_lat<2><<<148, 32>>>(iters, out, cyc);
        if (kind == 3) k_lat<3><<<148, 32>>>(iters, out, cyc);
        if (kind == 4) k_lat<4><<<148, 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %.1f cycles each\n", names[kind], (double)h / iters);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
