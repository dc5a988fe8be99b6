// Expert parallelism over the GPUs of one NVSwitch box (one process per GPU): the device-side dispatch plan and
// the NCCL exchange behind b200q_ep_* (include/b200q.h).  Not in the reference (single GPU);
// required by the north star.
//
//   * b200q_ep_plan: from the all-gathered [world][E] tokens-per-expert histogram, ONE small kernel derives the
//     all-to-all split sizes and, for the rows this rank will receive (ordered source rank, then expert), the
//     [start, end) range of every (local expert, source) pair -- the form b200q_moe_grouped_fwd_mapped consumes, so the
//     received rows are never re-sorted.  Replicated ("hot") experts are served where their tokens live.
//   * b200q_ep_comm_* / b200q_ep_allgather_i32 / b200q_ep_exchange: NCCL (the library torch ships, resolved with
//     dlopen: libb200q.so has no link-time dependency on it), variable-split all-to-all as grouped ncclSend / ncclRecv.
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "internal.h"

namespace b200q {

namespace {

// ------------------------------------------------------------------------------------------------ plan
// dest_s(e) = s if expert e is replicated, else its owner e / (E / world).  Rank s sends its rows sorted by
// (dest_s(e), e); this rank therefore receives, from every source s in turn, the experts with dest_s(e) == rank in
// ascending e.
__global__ void ep_plan_kernel(const int32_t* __restrict__ counts_all, int world, int rank, int E,
                               const int32_t* __restrict__ replicated, const int32_t* __restrict__ local_index,
                               int32_t* __restrict__ splits, int32_t* __restrict__ range_starts,
                               int32_t* __restrict__ range_ends, int32_t* __restrict__ range_expert) {
    const int per = E / world;
    for (int v = threadIdx.x; v < E * world; v += blockDim.x) {
        range_starts[v] = 0;
        range_ends[v] = 0;
        range_expert[v] = v / world;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        // rows this rank sends to rank r
        const int r = threadIdx.x;
        int send = 0;
        for (int e = 0; e < E; ++e) {
            const int dest = (replicated && replicated[e]) ? rank : e / per;
            if (dest == r) send += counts_all[rank * E + e];
        }
        splits[r] = send;
    }
    if (threadIdx.x == 0) {
        int offset = 0;
        for (int s = 0; s < world; ++s) {
            int recv = 0;
            for (int e = 0; e < E; ++e) {
                const int dest = (replicated && replicated[e]) ? s : e / per;
                if (dest != rank) continue;
                const int c = counts_all[s * E + e];
                const int li = local_index[e];
                if (li >= 0) {
                    range_starts[li * world + s] = offset;
                    range_ends[li * world + s] = offset + c;
                }
                offset += c;
                recv += c;
            }
            splits[world + s] = recv;
        }
        // ranges of one expert that happen to be adjacent in the receive buffer (always the case with one expert per
        // rank) are merged: the grouped GEMM tiles rows per range, and many short ranges mean many ragged last tiles
        for (int li = 0; li < E; ++li) {
            int cur = -1;
            for (int s = 0; s < world; ++s) {
                const int v = li * world + s;
                if (range_ends[v] <= range_starts[v]) continue;
                if (cur >= 0 && range_ends[cur] == range_starts[v]) {
                    range_ends[cur] = range_ends[v];
                    range_starts[v] = 0;
                    range_ends[v] = 0;
                } else {
                    cur = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ NCCL through dlopen
typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*FnGetUniqueId)(NcclUniqueId*);
typedef int (*FnCommInitRank)(void**, int, NcclUniqueId, int);
typedef int (*FnCommDestroy)(void*);
typedef int (*FnGroup)(void);
typedef int (*FnSendRecv)(void*, size_t, int, int, void*, cudaStream_t);      // ncclSend(const void*, ...) / ncclRecv
typedef int (*FnAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*FnErrStr)(int);

struct Nccl {
    void* h = nullptr;
    FnGetUniqueId get_id = nullptr;
    FnCommInitRank init = nullptr;
    FnCommDestroy destroy = nullptr;
    FnGroup gstart = nullptr, gend = nullptr;
    FnSendRecv send = nullptr, recv = nullptr;
    FnAllGather allgather = nullptr;
    FnErrStr errstr = nullptr;
    bool ok = false;
};

const Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);            // the copy torch already loaded, if any
            if (!n.h) n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) return;
        n.get_id = (FnGetUniqueId)dlsym(n.h, "ncclGetUniqueId");
        n.init = (FnCommInitRank)dlsym(n.h, "ncclCommInitRank");
        n.destroy = (FnCommDestroy)dlsym(n.h, "ncclCommDestroy");
        n.gstart = (FnGroup)dlsym(n.h, "ncclGroupStart");
        n.gend = (FnGroup)dlsym(n.h, "ncclGroupEnd");
        n.send = (FnSendRecv)dlsym(n.h, "ncclSend");
        n.recv = (FnSendRecv)dlsym(n.h, "ncclRecv");
        n.allgather = (FnAllGather)dlsym(n.h, "ncclAllGather");
        n.errstr = (FnErrStr)dlsym(n.h, "ncclGetErrorString");
        n.ok = n.get_id && n.init && n.destroy && n.gstart && n.gend && n.send && n.recv && n.allgather;
    });
    return n;
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return 0;
    const Nccl& n = nccl();
    return set_error(B200Q_ENCCL, "%s: NCCL error %d (%s)", what, rc, n.errstr ? n.errstr(rc) : "?");
}

int need_nccl() {
    if (!nccl().ok) return set_error(B200Q_ENCCL, "libnccl.so.2 could not be loaded (import torch first, or put NCCL on the library path)");
    return 0;
}

constexpr int NCCL_INT8 = 0, NCCL_INT32 = 2;       // ncclDataType_t

}  // namespace
}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_ep_plan(const int32_t* counts_all, int world, int rank, int E, const int32_t* replicated,
                  const int32_t* local_index, int32_t* splits, int32_t* range_starts, int32_t* range_ends,
                  int32_t* range_expert, void* stream) {
    if (world <= 0 || rank < 0 || rank >= world || E <= 0 || E % world != 0 || world > 1024)
        return set_error(B200Q_EINVAL, "ep_plan: need 0 <= rank < world <= 1024 and E %% world == 0 (world=%d rank=%d E=%d)", world, rank, E);
    if (!counts_all || !local_index || !splits || !range_starts || !range_ends || !range_expert)
        return set_error(B200Q_EINVAL, "ep_plan: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    ep_plan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(counts_all, world, rank, E, replicated, local_index, splits,
                                                                      range_starts, range_ends, range_expert);
    return check_cuda(cudaGetLastError(), "ep_plan launch");
}

int b200q_ep_unique_id(void* h_id128) {
    if (!h_id128) return set_error(B200Q_EINVAL, "ep_unique_id: null pointer");
    if (int rc = need_nccl()) return rc;
    return nccl_check(nccl().get_id(static_cast<NcclUniqueId*>(h_id128)), "ncclGetUniqueId");
}

int b200q_ep_comm_create(const void* h_id128, int rank, int world, void** comm) {
    if (!h_id128 || !comm || world <= 0 || rank < 0 || rank >= world) return set_error(B200Q_EINVAL, "ep_comm_create: bad argument");
    if (int rc = need_nccl()) return rc;
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    NcclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    return nccl_check(nccl().init(comm, world, id, rank), "ncclCommInitRank");
}

int b200q_ep_comm_destroy(void* comm) {
    if (!comm) return 0;
    if (int rc = need_nccl()) return rc;
    return nccl_check(nccl().destroy(comm), "ncclCommDestroy");
}

int b200q_ep_allgather_i32(void* comm, const int32_t* send, int32_t* recv, int64_t count, void* stream) {
    if (!comm || !send || !recv || count < 0) return set_error(B200Q_EINVAL, "ep_allgather_i32: bad argument");
    if (int rc = need_nccl()) return rc;
    return nccl_check(nccl().allgather(send, recv, (size_t)count, NCCL_INT32, comm, static_cast<cudaStream_t>(stream)), "ncclAllGather");
}

int b200q_ep_exchange(void* comm, int world, const void* send, const int64_t* h_send_rows, void* recv,
                      const int64_t* h_recv_rows, int64_t row_bytes, void* stream) {
    if (!comm || world <= 0 || !h_send_rows || !h_recv_rows || row_bytes <= 0) return set_error(B200Q_EINVAL, "ep_exchange: bad argument");
    if (int rc = need_nccl()) return rc;
    const Nccl& n = nccl();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = nccl_check(n.gstart(), "ncclGroupStart")) return rc;
    int64_t so = 0, ro = 0;
    int first = 0;
    for (int r = 0; r < world; ++r) {
        if (h_send_rows[r] < 0 || h_recv_rows[r] < 0) { first = set_error(B200Q_EINVAL, "ep_exchange: negative row count"); break; }
        if (h_send_rows[r] > 0 && !first)
            first = nccl_check(n.send(const_cast<char*>(static_cast<const char*>(send)) + so * row_bytes, (size_t)(h_send_rows[r] * row_bytes), NCCL_INT8, r, comm, st), "ncclSend");
        if (h_recv_rows[r] > 0 && !first)
            first = nccl_check(n.recv(static_cast<char*>(recv) + ro * row_bytes, (size_t)(h_recv_rows[r] * row_bytes), NCCL_INT8, r, comm, st), "ncclRecv");
        so += h_send_rows[r];
        ro += h_recv_rows[r];
    }
    const int rc_end = nccl_check(n.gend(), "ncclGroupEnd");
    return first ? first : rc_end;
}

int b200q_moe_grouped_fwd_mapped(const void* xs, int x_dtype, const uint8_t* packed, const float* scales,
                                 const float* zps, const int32_t* starts, const int32_t* ends,
                                 const int32_t* range_expert, int n_ranges, int n_experts, int gated, void* y, int y_dtype,
                                 int64_t R, int64_t N, int64_t K, void* ws, size_t ws_bytes, void* stream) {
    if (R < 0 || N < 0 || K < 0 || (K & 1) || n_ranges <= 0 || n_experts <= 0) return set_error(B200Q_EINVAL, "moe_grouped_fwd_mapped: need R,N >= 0, even K >= 0, ranges > 0, experts > 0");
    if (x_dtype < 0 || x_dtype > 2 || y_dtype < 0 || y_dtype > 2) return set_error(B200Q_EINVAL, "moe_grouped_fwd_mapped: unsupported dtype");
    if (gated && (N & 1)) return set_error(B200Q_EINVAL, "moe_grouped_fwd_mapped: the gated form needs an even number of weight rows");
    if (R == 0 || N == 0) return 0;
    if (!y || !scales || !zps || !starts || !ends || !range_expert || (K > 0 && (!xs || !packed))) return set_error(B200Q_EINVAL, "moe_grouped_fwd_mapped: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec_ok = !((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(y)) & 15);
    if (tuning().force_path != 1 && vec_ok && gemm_tc_supported(R, N, K, x_dtype, y_dtype))
        return launch_gemm_tc(d, xs, x_dtype, packed, scales, zps, y, y_dtype, R, N, K, starts, ends, n_ranges, ws, ws_bytes, 0u, st,
                              gated, range_expert, n_experts);
    if (gated) return set_error(B200Q_EINVAL, "moe_grouped_fwd_mapped: the gated form needs K %% 128 == 0 and 16-byte aligned buffers");
    return launch_linear_generic(xs, x_dtype, packed, scales, zps, y, y_dtype, R, N, K, starts, ends, n_ranges, 0, st, range_expert);
}

}  // extern "C"
