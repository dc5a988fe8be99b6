// HMMA m16n8k16 issue rate with the register traffic of the decode kernel: distinct A per instruction (LOP3 of a
// loaded word), B from a register array, CH independent accumulator chains.  cycles per HMMA per SMSP at 4 warps / SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH, bool LOPS>
__global__ void __launch_bounds__(512) k(int iters, const uint32_t* src, float* out, long long* cyc) {
    float c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t b[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { b[i][0] = src[threadIdx.x + i]; b[i][1] = src[threadIdx.x + 32 + i]; }
    uint32_t w0 = src[threadIdx.x], w1 = src[threadIdx.x + 7];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t a0, a1, a2, a3;
            if (LOPS) {
                const uint32_t v0 = (w0 >> (i & 1 ? 8 : 0)) + i, v1 = (w1 >> (i & 1 ? 8 : 0)) + it;
                a0 = v0 & 0x000f000f; a1 = v1 & 0x000f000f; a2 = v0 & 0x00f000f0; a3 = v1 & 0x00f000f0;
            } else { a0 = w0; a1 = w1; a2 = w0; a3 = w1; }
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i % CH][0]), "+f"(c[i % CH][1]), "+f"(c[i % CH][2]), "+f"(c[i % CH][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b[i][0]), "r"(b[i][1]));
        }
        w0 = w0 * 3 + 1; w1 ^= w0;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH, bool LOPS>
void run(const char* name, const uint32_t* src, float* out, long long* cyc) {
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        long long h = 0;
        k<CH, LOPS><<<148, warps * 32>>>(iters, src, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SM %2d: %.2f clk per HMMA per SMSP\n", name, warps, (double)h / ((double)iters * 16 * (warps / 4.0)));
    }
}

int main() {
    uint32_t* src; float* out; long long* cyc;
    cudaMalloc(&src, 4096 * 4); cudaMemset(src, 0x11, 4096 * 4); cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    run<8, false>("8 chains, same A", src, out, cyc);
    run<4, false>("4 chains, same A", src, out, cyc);
    run<2, false>("2 chains, same A", src, out, cyc);
    run<8, true>("8 chains, LOP3 A", src, out, cyc);
    run<4, true>("4 chains, LOP3 A", src, out, cyc);
    run<2, true>("2 chains, LOP3 A", src, out, cyc);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
