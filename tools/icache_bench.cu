// Does straight-line code pay instruction-fetch latency on every launch?  Kernels of N unrolled
// dependent FFMAs, one warp per SM x 4 or 16 warps, launched back to back; cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void k_line(float* out, long long* cyc, float a, float b) {
    float x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fmaf(x, a, b + i);      // distinct immediates: N distinct instructions
    long long t1 = clock64();
    if (x == 1.2345f) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// the streaming kernel in between evicts / does not evict the instruction lines from L2
__global__ void k_stream(const uint4* __restrict__ src, size_t n, uint4* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(src + i);
        acc.x ^= v.x;
    }
    if (acc.x == 0x12345u) sink[0] = acc;
}

template <int N>
void run(int warps, uint4* big, size_t nbig, uint4* sink, float* out, long long* cyc) {
    long long h[6];
    for (int rep = 0; rep < 6; ++rep) {
        if (big && rep >= 3) k_stream<<<1184, 256>>>(big, nbig, sink);
        k_line<N><<<148, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(&h[rep], cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("N=%5d (%3d KB) warps=%2d: cycles/instr per launch: %.2f %.2f %.2f | after 512MB stream: %.2f %.2f %.2f\n", N, N * 16 / 1024,
           warps, (double)h[0] / N, (double)h[1] / N, (double)h[2] / N, (double)h[3] / N, (double)h[4] / N, (double)h[5] / N);
}

int main() {
    float* out; long long* cyc; uint4 *big, *sink;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8); cudaMalloc(&sink, 64);
    size_t nbig = (512u << 20) / 16;
    cudaMalloc(&big, nbig * 16); cudaMemset(big, 1, nbig * 16);
    for (int warps : {1, 16}) {
        run<256>(warps, big, nbig, sink, out, cyc);
        run<1024>(warps, big, nbig, sink, out, cyc);
        run<2048>(warps, big, nbig, sink, out, cyc);
        run<4096>(warps, big, nbig, sink, out, cyc);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
