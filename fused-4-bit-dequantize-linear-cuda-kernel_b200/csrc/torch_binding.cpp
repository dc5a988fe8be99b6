// Thin torch binding over the C ABI of libb200q.so (include/b200q.h): the compiled counterpart of the reference's
// csrc/quantized_linear.cpp:22-28 (`fused_quant_linear_cuda.forward`) for callers that pay per call -- one C++ function
// does the checks, the device guard, the output / workspace allocation and the current-stream lookup that the ctypes
// path (_lib.py) does in ~25 us of Python.  No kernels here: every computation stays behind the C ABI.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <mutex>
#include <unordered_map>
#include "../../include/b200q.h"

namespace {

int dtype_code(at::ScalarType t) {
    switch (t) {
        case at::kFloat: return B200Q_F32;
        case at::kHalf: return B200Q_F16;
        case at::kBFloat16: return B200Q_BF16;
        default: TORCH_CHECK(false, "unsupported dtype ", t, "; expected float32, float16 or bfloat16");
    }
    return -1;
}

// zero-initialised scratch, private to (device, stream): the kernels leave their ticket counters zeroed again
struct WsKey {
    int device;
    void* stream;
    bool operator==(const WsKey& o) const { return device == o.device && stream == o.stream; }
};
struct WsHash {
    size_t operator()(const WsKey& k) const { return std::hash<void*>()(k.stream) ^ (size_t)k.device * 0x9E3779B97F4A7C15ull; }
};
std::mutex g_ws_mu;
std::unordered_map<WsKey, at::Tensor, WsHash> g_ws;

at::Tensor workspace(const at::Device& dev, void* stream, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsKey key{dev.index(), stream};
    auto it = g_ws.find(key);
    if (it == g_ws.end() || (size_t)it->second.numel() < bytes) {
        const int64_t n = std::max<int64_t>((int64_t)bytes, 1 << 20);
        at::Tensor t = at::zeros({n}, at::TensorOptions().dtype(at::kByte).device(dev));
        g_ws[key] = t;
        return t;
    }
    return it->second;
}

// y[..., N] = x[..., K] @ dequant(packed, scales, zero_points)^T (+ bias)
at::Tensor linear_forward(const at::Tensor& input, const at::Tensor& packed_weights, const at::Tensor& scales,
                          const at::Tensor& zero_points, const c10::optional<at::Tensor>& bias,
                          const c10::optional<at::ScalarType>& out_dtype, int64_t flags,
                          const c10::optional<at::Tensor>& next_packed) {
    // the reference's checks (csrc/quantized_linear_kernel.cu:311-335), plus contiguity of scales / zero points
    TORCH_CHECK(input.is_cuda(), "input must be a CUDA tensor");
    TORCH_CHECK(packed_weights.is_cuda(), "packed_weights must be a CUDA tensor");
    TORCH_CHECK(scales.is_cuda(), "scales must be a CUDA tensor");
    TORCH_CHECK(zero_points.is_cuda(), "zero_points must be a CUDA tensor");
    TORCH_CHECK(input.is_contiguous(), "input must be contiguous");
    TORCH_CHECK(packed_weights.is_contiguous(), "packed_weights must be contiguous");
    TORCH_CHECK(scales.is_contiguous(), "scales must be contiguous");
    TORCH_CHECK(zero_points.is_contiguous(), "zero_points must be contiguous");
    TORCH_CHECK(packed_weights.scalar_type() == at::kByte, "packed_weights must be uint8");
    TORCH_CHECK(scales.scalar_type() == at::kFloat, "scales must be float32");
    TORCH_CHECK(zero_points.scalar_type() == at::kFloat, "zero_points must be float32");
    TORCH_CHECK(input.dim() >= 1 && packed_weights.dim() == 2, "input must have >= 1 dims, packed_weights 2");
    const int64_t K = input.size(-1), N = packed_weights.size(0);
    TORCH_CHECK(K % 2 == 0 && packed_weights.size(1) == K / 2, "packed_weights dim 1 must be input_dim / 2");
    TORCH_CHECK(scales.numel() == N && zero_points.numel() == N, "scales / zero_points must have one entry per output row");
    const int64_t M = K > 0 ? input.numel() / K : 0;
    const at::ScalarType ydt = out_dtype.has_value() ? *out_dtype : input.scalar_type();
    std::vector<int64_t> shape(input.sizes().begin(), input.sizes().end());
    shape.back() = N;
    const float* bias_p = nullptr;
    if (bias.has_value() && bias->defined()) {
        TORCH_CHECK(bias->is_cuda() && bias->is_contiguous() && bias->scalar_type() == at::kFloat && bias->numel() == N,
                    "bias must be a contiguous float32 CUDA tensor with one entry per output row");
        bias_p = bias->data_ptr<float>();
    }
    const c10::cuda::CUDAGuard guard(input.device());
    at::Tensor y = at::empty(shape, input.options().dtype(ydt));
    if (M == 0 || N == 0) return y;
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(input.device().index()).stream();
    const size_t ws_bytes = b200q_linear_ws_bytes(M, N, K);
    at::Tensor ws;
    if (ws_bytes) ws = workspace(input.device(), stream, ws_bytes);
    const uint8_t* next_p = nullptr;
    size_t next_n = 0;
    if (next_packed.has_value() && next_packed->defined()) {
        next_p = next_packed->data_ptr<uint8_t>();
        next_n = (size_t)next_packed->numel();
    }
    const int rc = b200q_linear_bias_fwd(input.data_ptr(), dtype_code(input.scalar_type()), packed_weights.data_ptr<uint8_t>(),
                                         scales.data_ptr<float>(), zero_points.data_ptr<float>(), bias_p, y.data_ptr(), dtype_code(ydt),
                                         M, N, K, ws_bytes ? ws.data_ptr() : nullptr, ws_bytes ? (size_t)ws.numel() : 0, (unsigned)flags,
                                         stream, next_p, next_n);
    TORCH_CHECK(rc == 0, "b200q_linear_fwd failed (code ", rc, "): ", b200q_last_error_string());
    return y;
}

// Host-resident caller (python/module.py:100-118 with x on the CPU and the module on the GPU): pinned x in, pinned y out;
// the staging copies and the kernel are enqueued by one C call (b200q_linear_fwd_host).  x_stage / y_stage: the caller's
// device staging buffers ([M, K] / [M, N] of the activations' dtype).
at::Tensor linear_forward_host(const at::Tensor& x_host, const at::Tensor& x_stage, const at::Tensor& packed_weights,
                               const at::Tensor& scales, const at::Tensor& zero_points, const at::Tensor& y_stage,
                               const at::Tensor& y_host, int64_t flags) {
    TORCH_CHECK(!x_host.is_cuda() && !y_host.is_cuda(), "x_host / y_host must be host tensors (pinned)");
    TORCH_CHECK(x_host.is_contiguous() && y_host.is_contiguous() && x_stage.is_contiguous() && y_stage.is_contiguous(),
                "activations must be contiguous");
    TORCH_CHECK(packed_weights.is_cuda() && scales.is_cuda() && zero_points.is_cuda() && x_stage.is_cuda() && y_stage.is_cuda(),
                "weights and staging buffers must be CUDA tensors");
    TORCH_CHECK(packed_weights.is_contiguous() && scales.is_contiguous() && zero_points.is_contiguous(), "weights must be contiguous");
    TORCH_CHECK(packed_weights.scalar_type() == at::kByte && scales.scalar_type() == at::kFloat && zero_points.scalar_type() == at::kFloat,
                "packed_weights must be uint8, scales / zero_points float32");
    TORCH_CHECK(x_host.dim() == 2 && packed_weights.dim() == 2, "x_host must be [M, K], packed_weights [N, K / 2]");
    const int64_t M = x_host.size(0), K = x_host.size(1), N = packed_weights.size(0);
    TORCH_CHECK(K % 2 == 0 && packed_weights.size(1) == K / 2, "packed_weights dim 1 must be input_dim / 2");
    TORCH_CHECK(x_stage.numel() >= M * K && y_stage.numel() >= M * N && y_host.numel() == M * N, "staging / output buffers too small");
    TORCH_CHECK(x_stage.scalar_type() == x_host.scalar_type() && y_stage.scalar_type() == y_host.scalar_type(), "staging dtypes must match");
    const c10::cuda::CUDAGuard guard(packed_weights.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(packed_weights.device().index()).stream();
    const size_t ws_bytes = b200q_linear_ws_bytes(M, N, K);
    at::Tensor ws;
    if (ws_bytes) ws = workspace(packed_weights.device(), stream, ws_bytes);
    const int rc = b200q_linear_fwd_host(x_host.data_ptr(), dtype_code(x_host.scalar_type()), x_stage.data_ptr(),
                                         packed_weights.data_ptr<uint8_t>(), scales.data_ptr<float>(), zero_points.data_ptr<float>(),
                                         y_stage.data_ptr(), y_host.data_ptr(), dtype_code(y_host.scalar_type()), M, N, K,
                                         ws_bytes ? ws.data_ptr() : nullptr, ws_bytes ? (size_t)ws.numel() : 0, (unsigned)flags, stream);
    TORCH_CHECK(rc == 0, "b200q_linear_fwd_host failed (code ", rc, "): ", b200q_last_error_string());
    return y_host;
}

// the reference's extension surface: forward(input [K] or [M,K] f32, packed, scales, zero_points) -> [N] or [M,N] f32
at::Tensor forward(const at::Tensor& input, const at::Tensor& packed_weights, const at::Tensor& scales,
                   const at::Tensor& zero_points) {
    TORCH_CHECK(input.scalar_type() == at::kFloat, "input must be float32");
    TORCH_CHECK(input.dim() == 1 || input.dim() == 2, "input must be 1-D or 2-D, packed_weights 2-D");
    return linear_forward(input, packed_weights, scales, zero_points, c10::nullopt, c10::nullopt, 0, c10::nullopt);
}

}  // namespace

PYBIND11_MODULE(b200q_torch, m) {
    m.doc() = "compiled torch binding of libb200q.so (C ABI: include/b200q.h)";
    m.def("forward", &forward, "fused INT4 dequantize + linear (reference signature, csrc/quantized_linear.cpp:22-28)",
          py::arg("input"), py::arg("packed_weights"), py::arg("scales"), py::arg("zero_points"));
    m.def("linear_forward", &linear_forward, "fused INT4 dequantize + linear: any leading dims, f32 / f16 / bf16, bias, flags, next-layer hint",
          py::arg("input"), py::arg("packed_weights"), py::arg("scales"), py::arg("zero_points"), py::arg("bias") = py::none(),
          py::arg("out_dtype") = py::none(), py::arg("flags") = 0, py::arg("next_packed") = py::none());
    m.def("linear_forward_host", &linear_forward_host, "pinned host activations in, pinned host result out (one C call enqueues copies + kernel)",
          py::arg("x_host"), py::arg("x_stage"), py::arg("packed_weights"), py::arg("scales"), py::arg("zero_points"), py::arg("y_stage"),
          py::arg("y_host"), py::arg("flags") = 0);
    m.def("version", []() { return b200q_version(); });
}
