// Prefill / grouped path on tcgen05 tensor cores -- placeholder until the UMMA kernel lands:
// reports "unsupported" so the dispatcher uses the generic SIMT path.
#include "internal.h"

namespace b200q {

bool gemm_tc_supported(int64_t, int64_t, int64_t, int, int) { return false; }
size_t gemm_tc_ws_bytes(int64_t, int64_t, int64_t) { return 0; }
int launch_gemm_tc(const DeviceInfo&, const void*, int, const uint8_t*, const float*, const float*,
                   void*, int, int64_t, int64_t, int64_t, const int32_t*, const int32_t*, int, void*, size_t,
                   unsigned, cudaStream_t) {
    return set_error(B200Q_EINVAL, "gemm_tc: not built");
}

}  // namespace b200q
