// Error plumbing, device cache and tuning overrides of libb200q.so.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include "internal.h"

namespace b200q {

static thread_local char g_err[512] = {0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    // clear the sticky-less error so that later calls report their own failure
    (void)cudaGetLastError();
    return set_error(B200Q_ECUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

static std::mutex g_dev_mu;
static DeviceInfo g_dev[64];
static bool g_dev_valid[64] = {false};

int current_device(DeviceInfo* out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return set_error(B200Q_EINVAL, "device index %d out of range", dev);
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        if (g_dev_valid[dev]) {
            *out = g_dev[dev];
            return 0;
        }
    }
    DeviceInfo d;
    d.device = dev;
    B200Q_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    B200Q_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    B200Q_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    B200Q_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (d.cc_major != 10)
        return set_error(B200Q_EARCH, "device %d is sm_%d%d; libb200q is built for sm_100a only", dev,
                         d.cc_major, d.cc_minor);
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        g_dev[dev] = d;
        g_dev_valid[dev] = true;
    }
    *out = d;
    return 0;
}

static Tuning g_tuning;
const Tuning& tuning() { return g_tuning; }

// ---- pinned host memory -> device staging by a kernel (b200q_linear_fwd_host, decode-sized activations)
__global__ void __launch_bounds__(256) stage_host_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");      // dst may still be read by the preceding kernel
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

int launch_stage_host(const void* src_mapped, void* dst, size_t bytes, cudaStream_t st) {
    const int n16 = (int)(bytes / 16);
    int blocks = (n16 + 255) / 256;
    if (blocks > 8) blocks = 8;
    if (blocks < 1) blocks = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, stage_host_kernel, static_cast<const uint4*>(src_mapped), static_cast<uint4*>(dst), n16),
                      "stage_host launch");
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_version(void) { return B200Q_VERSION; }

const char* b200q_last_error_string(void) { return g_err; }

int b200q_device_check(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDeviceCount");
    if (n == 0) return set_error(B200Q_ECUDA, "no CUDA device");
    int dev = device;
    if (dev < 0) B200Q_CUDA(cudaGetDevice(&dev));
    if (dev >= n) return set_error(B200Q_EINVAL, "device %d >= device count %d", dev, n);
    int major = 0;
    B200Q_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return set_error(B200Q_EARCH, "device %d is not compute capability 10.x", dev);
    return 0;
}

int b200q_sm_count(void) {
    DeviceInfo d;
    int rc = current_device(&d);
    if (rc) return rc;
    return d.sm_count;
}

int b200q_tune_set(const char* key, int value) {
    if (!key) return set_error(B200Q_EINVAL, "null key");
    if (!strcmp(key, "gemv_warps")) g_tuning.gemv_warps = value;
    else if (!strcmp(key, "gemv_slabs")) g_tuning.gemv_slabs = value;
    else if (!strcmp(key, "gemv_stages")) g_tuning.gemv_stages = value;
    else if (!strcmp(key, "gemv_pdl")) g_tuning.gemv_pdl = value;
    else if (!strcmp(key, "gemv_ctas")) g_tuning.gemv_ctas = value;
    else if (!strcmp(key, "gemv_debug")) g_tuning.gemv_debug = value;
    else if (!strcmp(key, "gemm_bn")) g_tuning.gemm_bn = value;
    else if (!strcmp(key, "host_direct")) g_tuning.host_direct = value;
    else if (!strcmp(key, "gemm_debug")) g_tuning.gemm_debug = value;
    else if (!strcmp(key, "gemm_sk")) g_tuning.gemm_sk = value;
    else if (!strcmp(key, "gemm_mt_major")) g_tuning.gemm_mt_major = value;
    else if (!strcmp(key, "gemv_res")) g_tuning.gemv_res = value < 0 ? 1 : value;
    else if (!strcmp(key, "gemv_early")) g_tuning.gemv_early = value;
    else if (!strcmp(key, "gemv_xprep")) g_tuning.gemv_xprep = value < 0 ? 0 : value;
    else if (!strcmp(key, "gemv_pf")) g_tuning.gemv_pf = value < 0 ? 3 : value;
    else if (!strcmp(key, "gemv_occ2")) g_tuning.gemv_occ2 = value < 0 ? 0 : value;
    else if (!strcmp(key, "force_path")) g_tuning.force_path = value;
    else if (!strcmp(key, "gemv_bufs")) g_tuning.gemv_bufs = value < 0 ? 0 : value;
    else if (!strcmp(key, "hm_waves")) g_tuning.hm_waves = value < 0 ? 0 : value;
    else if (!strcmp(key, "moe_dec_hm")) g_tuning.moe_dec_hm = value;
    else if (!strcmp(key, "moe_dec_compact")) g_tuning.moe_dec_compact = value;
    else if (!strcmp(key, "hm_i3")) g_tuning.hm_i3 = value < 0 ? 1 : value;
    else if (!strcmp(key, "hm_max_m")) g_tuning.hm_max_m = (value < 0 || value > 32) ? 32 : value;
    else if (!strcmp(key, "hm_min_m")) g_tuning.hm_min_m = value < 0 ? 3 : value;
    else if (!strcmp(key, "gemv_slots")) g_tuning.gemv_slots = value < 0 ? 1 : value;
    else return set_error(B200Q_EINVAL, "unknown tuning key '%s'", key);
    return 0;
}

}  // extern "C"
