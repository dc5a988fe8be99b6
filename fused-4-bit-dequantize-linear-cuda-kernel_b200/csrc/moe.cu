// MoE routing kernels: softmax/top-k/renormalise, histogram + stable counting-sort permutation,
// row gather (dispatch), SiLU-gate, weighted combine.  All HBM/latency-bound integer and
// element-wise work: coalesced 16-byte accesses, no host synchronisation, device-side offsets.
//
// Reference semantics: benchmark/moe_grouped_gemm/routing.py:72-86 (router), :121-147 (dispatch),
// :175-187 (combine).
#include <cfloat>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

namespace {

// ------------------------------------------------------------------------------- top-k router
// One warp per token.  E <= 256 -> up to 8 logits per lane.  Ties: lowest expert index wins.
constexpr int TOPK_WARPS = 8;
constexpr int MAX_E_PER_LANE = 8;

// one warp: softmax over the E logits of a token, top-k, renormalise; lane s < k returns slot s in (isel, wsel)
__device__ __forceinline__ void topk_token(const float* __restrict__ lr, int E, int k, int lane, int& isel_out, float& wsel_out) {
    float v[MAX_E_PER_LANE];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAX_E_PER_LANE; ++j) {
        int e = lane + 32 * j;
        v[j] = e < E ? lr[e] : -INFINITY;
        mx = fmaxf(mx, v[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // routing.py:72 softmax: exp(x - max) / sum
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < MAX_E_PER_LANE; ++j) {
        int e = lane + 32 * j;
        v[j] = e < E ? expf(v[j] - mx) : 0.0f;
        sum += v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    // A NaN / +Inf logit makes every probability of the token NaN (as torch.softmax does).  The selection below must
    // still return k valid, distinct experts (the permutation / gather kernels index with them): NaN probabilities
    // are ranked as equal keys (ties -> lowest expert index), and the NaN is carried by the routing weights, so it
    // reaches the layer output exactly as in the reference (routing.py:72-76, 186-187).
    bool bad = false;
#pragma unroll
    for (int j = 0; j < MAX_E_PER_LANE; ++j) {
        int e = lane + 32 * j;
        v[j] = e < E ? v[j] * inv : -1.0f;   // probabilities are >= 0; -1 marks "absent / taken"
        if (v[j] != v[j]) { bad = true; v[j] = 2.0f; }
    }
    bad = __any_sync(0xffffffffu, bad);
    // routing.py:73 top-k (descending probability), then :76 renormalise over the k winners
    float wsel = 0.0f, wsum = 0.0f;
    int isel = 0;
    for (int s = 0; s < k; ++s) {
        float best = -2.0f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < MAX_E_PER_LANE; ++j) {
            int e = lane + 32 * j;
            if (v[j] > best) { best = v[j]; besti = e; }   // ascending e inside a lane: first max wins
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if ((besti & 31) == lane) {
#pragma unroll
            for (int j = 0; j < MAX_E_PER_LANE; ++j)
                if (lane + 32 * j == besti) v[j] = -1.0f;
        }
        wsum += best;
        if (lane == s) { wsel = best; isel = besti; }
    }
    isel_out = isel;
    wsel_out = bad ? __int_as_float(0x7fc00000) : wsel / wsum;
}

__global__ void __launch_bounds__(TOPK_WARPS * 32)
moe_topk_kernel(const float* __restrict__ logits, int64_t T, int E, int k,
                int32_t* __restrict__ idx, float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * TOPK_WARPS + (threadIdx.x >> 5);
    if (t >= T) return;
    int isel;
    float wsel;
    topk_token(logits + t * E, E, k, lane, isel, wsel);
    if (lane < k) {
        idx[t * k + lane] = isel;
        weights[t * k + lane] = wsel;
    }
}

// Decode-sized routing (T <= 16 tokens) in ONE launch of one CTA: warp t routes token t, then the T k assignments are
// counting-sorted by expert (stable: equal experts keep token order) -- the same outputs as the four kernels above.
// src_token (optional): token of every sorted position, so a consumer can read x in place instead of a gathered copy.
constexpr int ROUTE_SMALL_T = 16;
__global__ void __launch_bounds__(ROUTE_SMALL_T * 32)
moe_route_small_kernel(const float* __restrict__ logits, int T, int E, int k, int32_t* __restrict__ idx,
                       float* __restrict__ weights, int32_t* __restrict__ counts, int32_t* __restrict__ offsets,
                       int32_t* __restrict__ sorted_slot, int32_t* __restrict__ inv_perm, int32_t* __restrict__ src_token) {
    __shared__ int32_t s_idx[ROUTE_SMALL_T * 8];
    __shared__ int32_t s_off[32 * MAX_E_PER_LANE + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = T * k;
    if (warp < T) {
        int isel;
        float wsel;
        topk_token(logits + (int64_t)warp * E, E, k, lane, isel, wsel);
        if (lane < k) {
            idx[warp * k + lane] = isel;
            weights[warp * k + lane] = wsel;
            s_idx[warp * k + lane] = isel;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        int c = 0;
        for (int a = 0; a < A; ++a) c += s_idx[a] == e ? 1 : 0;
        counts[e] = c;
        s_off[e] = c;
    }
    __syncthreads();
    if (warp == 0) {                                         // exclusive scan of the E <= 256 counts
        int v[MAX_E_PER_LANE], tot = 0;
#pragma unroll
        for (int j = 0; j < MAX_E_PER_LANE; ++j) {
            const int e = lane * MAX_E_PER_LANE + j;
            v[j] = e < E ? s_off[e] : 0;
            tot += v[j];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        int run = incl - tot;
#pragma unroll
        for (int j = 0; j < MAX_E_PER_LANE; ++j) {
            const int e = lane * MAX_E_PER_LANE + j;
            if (e < E) { s_off[e] = run; offsets[e] = run; }
            run += v[j];
        }
        if (lane == 31) offsets[E] = incl;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const int e = s_idx[a];
        int pos = s_off[e];
        for (int b = 0; b < a; ++b) pos += s_idx[b] == e ? 1 : 0;
        sorted_slot[pos] = a;
        inv_perm[a] = pos;
        if (src_token) src_token[pos] = a / k;
    }
}

// ------------------------------------------------------------------------------- permutation
// Stable counting sort of A = T*k assignments by expert id.
//   pass 1: per-block histogram            -> ws[b*E + e]
//   pass 2: single CTA: counts, exclusive offsets, per-block bases (in place)
//   pass 3: per-block stable scatter
constexpr int PERM_THREADS = 256;
constexpr int PERM_CHUNK = 2048;   // assignments per block

__global__ void __launch_bounds__(PERM_THREADS)
perm_hist_kernel(const int32_t* __restrict__ idx, int64_t A, int E, int32_t* __restrict__ ws, const int32_t* __restrict__ remap) {
    extern __shared__ int32_t s_hist[];
    for (int e = threadIdx.x; e < E; e += PERM_THREADS) s_hist[e] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * PERM_CHUNK;
    for (int i = threadIdx.x; i < PERM_CHUNK; i += PERM_THREADS) {
        int64_t a = base + i;
        if (a < A) {
            int e = idx[a];
            if (e >= 0 && e < E) atomicAdd(&s_hist[remap ? remap[e] : e], 1);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += PERM_THREADS) ws[(int64_t)blockIdx.x * E + e] = s_hist[e];
}

__global__ void __launch_bounds__(PERM_THREADS)
perm_scan_kernel(int32_t* __restrict__ ws, int nblocks, int E, int32_t* __restrict__ counts,
                 int32_t* __restrict__ offsets, int32_t* __restrict__ sorted_slot, int64_t A, const int32_t* __restrict__ remap) {
    extern __shared__ int32_t s_cnt[];   // [E] totals then exclusive offsets
    for (int e = threadIdx.x; e < E; e += PERM_THREADS) {
        int tot = 0;
        for (int b = 0; b < nblocks; ++b) tot += ws[(int64_t)b * E + e];
        s_cnt[e] = tot;
        if (!remap) counts[e] = tot;
    }
    __syncthreads();
    // sorted by a remapped id (expert-parallel send order): the histogram is still reported per ORIGINAL expert id
    if (remap)
        for (int e = threadIdx.x; e < E; e += PERM_THREADS) counts[e] = s_cnt[remap[e]];
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int e = 0; e < E; ++e) {
            int c = s_cnt[e];
            s_cnt[e] = run;
            offsets[e] = run;
            run += c;
        }
        offsets[E] = run;
        s_cnt[E] = run;
    }
    __syncthreads();
    // assignments with an expert id outside [0, E) (only possible with caller-made routing) are dropped: the tail
    // of sorted_slot they leave unused is marked -1 (gather writes zero rows, combine skips them)
    if (sorted_slot)
        for (int64_t i = s_cnt[E] + (int64_t)threadIdx.x; i < A; i += PERM_THREADS) sorted_slot[i] = -1;
    for (int e = threadIdx.x; e < E; e += PERM_THREADS) {
        int run = s_cnt[e];
        for (int b = 0; b < nblocks; ++b) {
            int c = ws[(int64_t)b * E + e];
            ws[(int64_t)b * E + e] = run;
            run += c;
        }
    }
}

__global__ void __launch_bounds__(PERM_THREADS)
perm_scatter_kernel(const int32_t* __restrict__ idx, int64_t A, int E,
                    const int32_t* __restrict__ ws, int32_t* __restrict__ sorted_slot,
                    int32_t* __restrict__ inv_perm, const int32_t* __restrict__ remap) {
    extern __shared__ int32_t smem[];
    int32_t* s_run = smem;            // [E] next free position per expert
    int32_t* s_warp = smem + E;       // [8][E] per-warp counts of the current round
    constexpr int NW = PERM_THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < E; e += PERM_THREADS) s_run[e] = ws[(int64_t)blockIdx.x * E + e];
    const int64_t base = (int64_t)blockIdx.x * PERM_CHUNK;
    for (int r = 0; r < PERM_CHUNK / PERM_THREADS; ++r) {
        for (int i = threadIdx.x; i < NW * E; i += PERM_THREADS) s_warp[i] = 0;
        __syncthreads();
        const int64_t a = base + (int64_t)r * PERM_THREADS + threadIdx.x;
        int e = -1;
        if (a < A) {
            e = idx[a];
            if (e < 0 || e >= E) e = -1;
            else if (remap) e = remap[e];
        }
        const unsigned peers = __match_any_sync(0xffffffffu, e);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (e >= 0 && rank == 0) s_warp[warp * E + e] = __popc(peers);
        __syncthreads();
        if (e >= 0) {
            int pos = s_run[e] + rank;
            for (int w = 0; w < warp; ++w) pos += s_warp[w * E + e];
            sorted_slot[pos] = (int32_t)a;
            inv_perm[a] = pos;
        } else if (a < A) {
            inv_perm[a] = -1;
        }
        __syncthreads();
        for (int ee = threadIdx.x; ee < E; ee += PERM_THREADS) {
            int add = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) add += s_warp[w * E + ee];
            s_run[ee] += add;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------- row gather
// xs[p,:] = x[sorted_slot[p] / k, :], 16-byte vectors, one warp per row chunk.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ sorted_slot,
                   int64_t rows, int k, int64_t vec_per_row, uint4* __restrict__ xs) {
    const int64_t p = blockIdx.x;
    if (p >= rows) return;
    const int slot = sorted_slot[p];
    uint4* d = xs + p * vec_per_row;
    if (slot < 0) {                                       // unused tail position (dropped assignment)
        for (int64_t i = threadIdx.x; i < vec_per_row; i += blockDim.x) d[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const uint4* s = x + (int64_t)(slot / k) * vec_per_row;
    for (int64_t i = threadIdx.x; i < vec_per_row; i += blockDim.x) d[i] = __ldg(s + i);
}

// ------------------------------------------------------------------------------- SiLU gate
// h[p,f] = silu(gu[p,f]) * gu[p,F+f].  HBM-bound: 16-byte loads / stores when F allows it.
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__global__ void __launch_bounds__(256)
silu_mul_kernel(const T* __restrict__ gu, int64_t R, int64_t F, T* __restrict__ h) {
    constexpr int V = Vec16<T>::N;
    if (F % V == 0 && (reinterpret_cast<uintptr_t>(gu) & 15) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0) {
        const int64_t fv = F / V, total = R * fv;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t p = i / fv, f = (i - p * fv) * V;
            const uint4 av = *reinterpret_cast<const uint4*>(gu + p * 2 * F + f);
            const uint4 bv = *reinterpret_cast<const uint4*>(gu + p * 2 * F + F + f);
            uint4 ov;
            const T* a = reinterpret_cast<const T*>(&av);
            const T* b = reinterpret_cast<const T*>(&bv);
            T* o = reinterpret_cast<T*>(&ov);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float x = to_f32<T>(a[j]);
                o[j] = from_f32<T>(x / (1.0f + expf(-x)) * to_f32<T>(b[j]));
            }
            *reinterpret_cast<uint4*>(h + p * F + f) = ov;
        }
        return;
    }
    const int64_t total = R * F;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / F, f = i - p * F;
        const float a = to_f32<T>(gu[p * 2 * F + f]);
        const float b = to_f32<T>(gu[p * 2 * F + F + f]);
        const float sg = a / (1.0f + expf(-a));
        h[i] = from_f32<T>(sg * b);
    }
}

// ------------------------------------------------------------------------------- combine
// out[t,f] = sum_s weights[t,s] * y[inv_perm[t*k+s], f]   (routing.py:183-187: product rounded to
// fp32, then summed over the k slots in slot order)
template <typename YT, typename OT>
__global__ void __launch_bounds__(256)
combine_kernel(const YT* __restrict__ y, const int32_t* __restrict__ inv_perm,
               const float* __restrict__ weights, int k, int64_t F, OT* __restrict__ out) {
    const int64_t t = blockIdx.x;
    for (int64_t f = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; f < F; f += (int64_t)gridDim.y * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < k; ++s) {
            const int64_t p = inv_perm[t * k + s];
            const float prod = p < 0 ? 0.0f : __fmul_rn(to_f32<YT>(y[p * F + f]), weights[t * k + s]);
            acc = s == 0 ? prod : __fadd_rn(acc, prod);
        }
        out[t * F + f] = from_f32<OT>(acc);
    }
}

// Same arithmetic, 16 bytes of y per thread and slot (V = 8 halves / 4 floats): the scalar form moves 2-byte elements and
// runs at a third of the HBM rate on the 0.5 GB of a 16384-token Mixtral step.  k <= 8; F % V == 0; y / out 16-byte aligned.
template <typename YT, typename OT>
__global__ void __launch_bounds__(256)
combine_vec_kernel(const YT* __restrict__ y, const int32_t* __restrict__ inv_perm,
                   const float* __restrict__ weights, int k, int64_t F, OT* __restrict__ out) {
    constexpr int V = 16 / (int)sizeof(YT);
    const int64_t t = blockIdx.x;
    int64_t row[8];
    float w[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        row[s] = s < k ? (int64_t)inv_perm[t * k + s] : -1;
        w[s] = s < k ? weights[t * k + s] : 0.0f;
    }
    for (int64_t f = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * V; f < F; f += (int64_t)gridDim.y * blockDim.x * V) {
        float acc[V];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (s < k) {
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (row[s] >= 0) raw = __ldcs(reinterpret_cast<const uint4*>(y + row[s] * F + f));     // (read once)
                const YT* e = reinterpret_cast<const YT*>(&raw);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float prod = row[s] < 0 ? 0.0f : __fmul_rn(to_f32<YT>(e[i]), w[s]);
                    acc[i] = s == 0 ? prod : __fadd_rn(acc[i], prod);
                }
            }
        }
        alignas(16) OT o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = from_f32<OT>(acc[i]);
        OT* dst = out + t * F + f;
        if constexpr (V * sizeof(OT) >= 16) {
#pragma unroll
            for (int c = 0; c < (int)(V * sizeof(OT) / 16); ++c) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(o)[c];
        } else {
            *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(o);                      // 4 floats -> 4 halves
        }
    }
}

template <typename YT>
int combine_out(const void* y, const int32_t* inv_perm, const float* weights, int64_t T, int k,
                int64_t F, void* out, int out_dtype, cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(YT);
    if (k <= 8 && F % V == 0 && !(reinterpret_cast<uintptr_t>(y) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) &&
        (out_dtype == B200Q_F32 || out_dtype == B200Q_F16 || out_dtype == B200Q_BF16)) {
        const int64_t per_block = 256 * V;
        int64_t ny = T >= 256 ? 1 : (256 + T - 1) / T;      // few tokens (decode): spread the columns of a token over several CTAs as well
        if (ny > (F + per_block - 1) / per_block) ny = (F + per_block - 1) / per_block;
        dim3 grid(static_cast<unsigned>(T), static_cast<unsigned>(ny));
        if (out_dtype == B200Q_F32) combine_vec_kernel<YT, float><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<float*>(out));
        else if (out_dtype == B200Q_F16) combine_vec_kernel<YT, __half><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<__half*>(out));
        else combine_vec_kernel<YT, __nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<__nv_bfloat16*>(out));
        return check_cuda(cudaGetLastError(), "moe_combine launch");
    }
    // few tokens (decode): spread the columns of a token over several CTAs as well
    int64_t ny = T >= 256 ? 1 : (256 + T - 1) / T;
    if (ny > (F + 255) / 256) ny = (F + 255) / 256;
    dim3 grid(static_cast<unsigned>(T), static_cast<unsigned>(ny));
    switch (out_dtype) {
        case B200Q_F32:
            combine_kernel<YT, float><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<float*>(out));
            break;
        case B200Q_F16:
            combine_kernel<YT, __half><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<__half*>(out));
            break;
        case B200Q_BF16:
            combine_kernel<YT, __nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const YT*>(y), inv_perm, weights, k, F, static_cast<__nv_bfloat16*>(out));
            break;
        default: return set_error(B200Q_EINVAL, "bad out_dtype %d", out_dtype);
    }
    return check_cuda(cudaGetLastError(), "moe_combine launch");
}

}  // namespace

int moe_route_small(const float* logits, int64_t T, int E, int k, int32_t* idx, float* weights, int32_t* counts, int32_t* offsets,
                    int32_t* sorted_slot, int32_t* inv_perm, int32_t* src_token, cudaStream_t st) {
    if (T < 1 || T > ROUTE_SMALL_T || E < 1 || E > 32 * MAX_E_PER_LANE || k < 1 || k > 8 || k > E)
        return set_error(B200Q_EINVAL, "moe_route_small: need 1<=T<=16, 1<=E<=256, 1<=k<=min(8,E) (T=%lld E=%d k=%d)", (long long)T, E, k);
    moe_route_small_kernel<<<1, ROUTE_SMALL_T * 32, 0, st>>>(logits, (int)T, E, k, idx, weights, counts, offsets, sorted_slot, inv_perm, src_token);
    return check_cuda(cudaGetLastError(), "moe_route_small launch");
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_moe_topk(const float* logits, int64_t T, int E, int k, int32_t* idx, float* weights,
                   void* stream) {
    if (T < 0 || E <= 0 || E > 32 * MAX_E_PER_LANE || k <= 0 || k > 8 || k > E)
        return set_error(B200Q_EINVAL, "moe_topk: need T>=0, 1<=E<=256, 1<=k<=min(8,E) (T=%lld E=%d k=%d)", (long long)T, E, k);
    if (T == 0) return 0;
    if (!logits || !idx || !weights) return set_error(B200Q_EINVAL, "moe_topk: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    int64_t blocks = (T + TOPK_WARPS - 1) / TOPK_WARPS;
    moe_topk_kernel<<<static_cast<unsigned>(blocks), TOPK_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(logits, T, E, k, idx, weights);
    return check_cuda(cudaGetLastError(), "moe_topk launch");
}

size_t b200q_moe_permute_ws_bytes(int64_t T, int E, int k) {
    int64_t A = T * k;
    int64_t blocks = (A + PERM_CHUNK - 1) / PERM_CHUNK;
    if (blocks < 1) blocks = 1;
    return static_cast<size_t>(blocks) * E * sizeof(int32_t);
}

int b200q_moe_permute(const int32_t* idx, int64_t T, int E, int k, int32_t* counts,
                      int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws,
                      size_t ws_bytes, void* stream) {
    return b200q_moe_permute_mapped(idx, nullptr, T, E, k, counts, offsets, sorted_slot, inv_perm, ws, ws_bytes, stream);
}

int b200q_moe_route(const float* logits, const int32_t* remap, int64_t T, int E, int k, int32_t* idx, float* weights,
                    int32_t* counts, int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws, size_t ws_bytes,
                    void* stream) {
    if (!remap && T >= 1 && T <= ROUTE_SMALL_T && E >= 1 && E <= 32 * MAX_E_PER_LANE && k >= 1 && k <= 8 && k <= E && logits && idx &&
        weights && counts && offsets && sorted_slot && inv_perm)
        return moe_route_small(logits, T, E, k, idx, weights, counts, offsets, sorted_slot, inv_perm, nullptr, static_cast<cudaStream_t>(stream));
    if (int rc = b200q_moe_topk(logits, T, E, k, idx, weights, stream)) return rc;
    return b200q_moe_permute_mapped(idx, remap, T, E, k, counts, offsets, sorted_slot, inv_perm, ws, ws_bytes, stream);
}

int b200q_moe_permute_mapped(const int32_t* idx, const int32_t* remap, int64_t T, int E, int k, int32_t* counts,
                             int32_t* offsets, int32_t* sorted_slot, int32_t* inv_perm, void* ws,
                             size_t ws_bytes, void* stream) {
    if (T < 0 || E <= 0 || E > 1024 || k <= 0) return set_error(B200Q_EINVAL, "moe_permute: bad T/E/k");
    const int64_t A = T * k;
    if (A > 0x7fffffffLL) return set_error(B200Q_EINVAL, "moe_permute: T*k exceeds int32");
    if (!counts || !offsets) return set_error(B200Q_EINVAL, "moe_permute: null pointer");
    if (ws_bytes < b200q_moe_permute_ws_bytes(T, E, k) || !ws) return set_error(B200Q_EWORKSPACE, "moe_permute: workspace too small");
    if (A > 0 && (!idx || !sorted_slot || !inv_perm)) return set_error(B200Q_EINVAL, "moe_permute: null pointer");
    DeviceInfo d;
    if (int rc = current_device(&d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int nblocks = static_cast<int>((A + PERM_CHUNK - 1) / PERM_CHUNK);
    if (nblocks < 1) nblocks = 1;
    int32_t* w = static_cast<int32_t*>(ws);
    perm_hist_kernel<<<nblocks, PERM_THREADS, E * sizeof(int32_t), st>>>(idx, A, E, w, remap);
    perm_scan_kernel<<<1, PERM_THREADS, (E + 1) * sizeof(int32_t), st>>>(w, nblocks, E, counts, offsets, sorted_slot, A, remap);
    if (A > 0)
        perm_scatter_kernel<<<nblocks, PERM_THREADS, (1 + PERM_THREADS / 32) * E * sizeof(int32_t), st>>>(idx, A, E, w, sorted_slot, inv_perm, remap);
    return check_cuda(cudaGetLastError(), "moe_permute launch");
}

int b200q_moe_gather_rows(const void* x, int dtype, const int32_t* sorted_slot, int64_t rows,
                          int k, int64_t d, void* xs, void* stream) {
    if (rows < 0 || k <= 0 || d < 0) return set_error(B200Q_EINVAL, "moe_gather_rows: bad sizes");
    if (rows == 0 || d == 0) return 0;
    int esz = dtype == B200Q_F32 ? 4 : (dtype == B200Q_F16 || dtype == B200Q_BF16) ? 2 : 0;
    if (!esz) return set_error(B200Q_EINVAL, "moe_gather_rows: bad dtype %d", dtype);
    if ((d * esz) % 16 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(xs) & 15))
        return set_error(B200Q_EALIGN, "moe_gather_rows: rows must be 16-byte multiples and 16-byte aligned");
    if (rows > 0x7fffffffLL) return set_error(B200Q_EINVAL, "moe_gather_rows: too many rows");
    DeviceInfo di;
    if (int rc = current_device(&di)) return rc;
    int64_t vpr = d * esz / 16;
    int threads = vpr >= 256 ? 256 : (vpr >= 128 ? 128 : 64);
    gather_rows_kernel<<<static_cast<unsigned>(rows), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), sorted_slot, rows, k, vpr, static_cast<uint4*>(xs));
    return check_cuda(cudaGetLastError(), "moe_gather_rows launch");
}

int b200q_moe_silu_mul(const void* gu, int dtype, int64_t R, int64_t F, void* h, void* stream) {
    if (R < 0 || F < 0) return set_error(B200Q_EINVAL, "moe_silu_mul: bad sizes");
    if (R == 0 || F == 0) return 0;
    DeviceInfo di;
    if (int rc = current_device(&di)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t total = R * F;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)di.sm_count * 32;
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case B200Q_F32: silu_mul_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const float*>(gu), R, F, static_cast<float*>(h)); break;
        case B200Q_F16: silu_mul_kernel<__half><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const __half*>(gu), R, F, static_cast<__half*>(h)); break;
        case B200Q_BF16: silu_mul_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gu), R, F, static_cast<__nv_bfloat16*>(h)); break;
        default: return set_error(B200Q_EINVAL, "moe_silu_mul: bad dtype %d", dtype);
    }
    return check_cuda(cudaGetLastError(), "moe_silu_mul launch");
}

int b200q_moe_combine(const void* y, int dtype, const int32_t* inv_perm, const float* weights,
                      int64_t T, int k, int64_t F, void* out, int out_dtype, void* stream) {
    if (T < 0 || k <= 0 || F < 0) return set_error(B200Q_EINVAL, "moe_combine: bad sizes");
    if (T == 0 || F == 0) return 0;
    if (T > 0x7fffffffLL) return set_error(B200Q_EINVAL, "moe_combine: too many tokens");
    if (!y || !inv_perm || !weights || !out) return set_error(B200Q_EINVAL, "moe_combine: null pointer");
    DeviceInfo di;
    if (int rc = current_device(&di)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case B200Q_F32: return combine_out<float>(y, inv_perm, weights, T, k, F, out, out_dtype, st);
        case B200Q_F16: return combine_out<__half>(y, inv_perm, weights, T, k, F, out, out_dtype, st);
        case B200Q_BF16: return combine_out<__nv_bfloat16>(y, inv_perm, weights, T, k, F, out, out_dtype, st);
    }
    return set_error(B200Q_EINVAL, "moe_combine: bad dtype %d", dtype);
}

}  // extern "C"
