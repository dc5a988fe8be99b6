// Launch-overhead calibration: empty / near-empty kernels, back to back in a CUDA graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/launch_bench tools/launch_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void k_empty(int* out) {
    extern __shared__ char sm[];
    if (threadIdx.x == 1023 && out) sm[0] = 1;
}
__global__ void k_pdl(int* out) {
    extern __shared__ char sm[];
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 1023 && out) sm[0] = 1;
}
// touches `bytes` of global memory per CTA with plain loads (bandwidth reference for a persistent grid)
__global__ void k_stream(const uint4* __restrict__ src, size_t n16_per_cta, uint4* sink) {
    const uint4* p = src + (size_t)blockIdx.x * n16_per_cta;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = threadIdx.x; i < n16_per_cta; i += blockDim.x) {
        uint4 v = __ldg(p + i);
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if (acc.x == 0x12345678u) sink[0] = acc;
}

template <typename F>
float time_graph(F launch, int per_graph, int reps) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    launch(st);
    cudaStreamSynchronize(st);
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < per_graph; ++i) launch(st);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / (reps * per_graph);
}

int main() {
    int smems[] = {0, 48 * 1024, 100 * 1024, 160 * 1024, 227 * 1024};
    int threads[] = {128, 512, 544};
    for (int sm : smems)
        for (int th : threads) {
            cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
            float a = time_graph([&](cudaStream_t st) { k_empty<<<148, th, sm, st>>>(nullptr); }, 24, 50);
            float b = time_graph([&](cudaStream_t st) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(148); cfg.blockDim = dim3(th); cfg.dynamicSmemBytes = sm; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, k_pdl, (int*)nullptr);
            }, 24, 50);
            printf("smem %3d KB threads %3d: empty %.2f us/launch, pdl %.2f us/launch\n", sm / 1024, th, a, b);
        }
    // streaming reference: 24 distinct 22.5 MB buffers, one persistent CTA per SM, plain LDG
    size_t bytes = 22544384, n16 = bytes / 16 / 148;
    std::vector<uint4*> bufs(24);
    for (auto& b : bufs) { cudaMalloc(&b, bytes); cudaMemset(b, 1, bytes); }
    uint4* sink; cudaMalloc(&sink, 64);
    for (int th : {256, 512, 1024}) {
        int i = 0;
        float t = time_graph([&](cudaStream_t st) { k_stream<<<148, th, 0, st>>>(bufs[(i++) % 24], n16, sink); }, 24, 50);
        printf("stream 22.5MB persistent 148 CTAs x %4d thr: %.2f us/launch = %.0f GB/s\n", th, t, bytes / t / 1e3);
        int j = 0;
        float t2 = time_graph([&](cudaStream_t st) { k_stream<<<148 * 8, th / 4 < 128 ? 128 : th / 4, 0, st>>>(bufs[(j++) % 24], n16 / 8, sink); }, 24, 50);
        printf("stream 22.5MB 1184 CTAs: %.2f us/launch = %.0f GB/s\n", t2, bytes / t2 / 1e3);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
