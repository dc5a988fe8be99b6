// Decode path: y[M,N] = x[M,K] @ dequant(W)^T for M <= 16.  HBM-bound: every packed byte is read
// exactly once, by the TMA engine, into shared memory; everything else is on-chip.
//
// Decomposition (DESIGN.md "GEMV"):
//   * K is cut into `nslab` slabs of whole 128-column granules, N into `nrb` row blocks;
//     CTA (rb, slab) streams rows [r0,r1) x slab bytes.  nslab * nrb ~= SM count, one CTA per SM.
//   * a stage of the shared-memory ring = one tile of 16 weight rows x slab bytes, filled by 16
//     cp.async.bulk row copies (UBLKCP) that complete on the stage's mbarrier.  The ring is deep
//     enough that for the Llama shapes the producer issues the CTA's whole slab up front.
//   * consumer warp w owns granules w, w+NW, ... of the slab for the whole kernel, so its x
//     operand (mma B fragments) is built ONCE in registers; the main loop is LDS.128 -> LOP3
//     nibble->half -> HMMA only.  Weights enter the mma as exact integers (q-8); x enters as an
//     fp16 hi/lo split of x * 2^e (per-warp power-of-two e), so the products are exact and the
//     accumulation is fp32: results match the fp32 reference to ~1e-6 relative.
//   * per tile: cross-warp reduction in shared memory (fixed order), epilogue
//     y = s * (acc - (zp-8) * sum(x)); with nslab > 1 the slab partials go to a workspace and the
//     last CTA of a row block (ticket counter) adds them in slab order: deterministic.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output).
#include <cmath>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

namespace {

constexpr int TILE_ROWS = 16;
constexpr int GRAN_K = 128;          // columns per granule
constexpr int GRAN_B = GRAN_K / 2;   // packed bytes per granule per row
constexpr int MAX_SLABS = 16;
constexpr int MAX_RB = 1024;         // ticket counters at the head of the workspace

struct GemvParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    void* y;
    float* part;            // [nslab][M][N] fp32 slab partials (nslab > 1)
    unsigned int* tickets;  // [nrb]
    int x_dtype, y_dtype;
    int M, N, K;
    int nslab, nrb, G;      // G = K / 128
    int stages;
    int pitch;              // bytes between rows of a stage
    int wait_weights;       // 1: weights may be written by the preceding kernel -> wait first
};

__device__ __forceinline__ void load4f(const void* x, int dtype, int64_t idx, float (&v)[4]) {
    if (dtype == B200Q_F32) {
        float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + idx);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if (dtype == B200Q_F16) {
        uint2 r = *reinterpret_cast<const uint2*>(static_cast<const __half*>(x) + idx);
        float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x));
        float2 b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        uint2 r = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x) + idx);
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}

__device__ __forceinline__ void store_y(void* y, int dtype, int64_t idx, float v) {
    if (dtype == B200Q_F32) static_cast<float*>(y)[idx] = v;
    else if (dtype == B200Q_F16) static_cast<__half*>(y)[idx] = __float2half_rn(v);
    else static_cast<__nv_bfloat16*>(y)[idx] = __float2bfloat16_rn(v);
}

// 8 nibbles of one 32-bit word -> four half2 holding the exact integers (n - 8):
// r[0] = (n0,n4)  r[1] = (n1,n5)  r[2] = (n2,n6)  r[3] = (n3,n7)
__device__ __forceinline__ void nibbles_to_half2x4_s8(uint32_t w, uint32_t (&r)[4]) {
    constexpr uint32_t MAGIC = 0x64006400u;                    // half2(1024, 1024)
    constexpr uint32_t LO = 0x000f000fu, HI = 0x00f000f0u;
    const __half2 bias = __halves2half2(__ushort_as_half(0x6408), __ushort_as_half(0x6408));     // 1032
    const __half2 sixteenth = __halves2half2(__ushort_as_half(0x2C00), __ushort_as_half(0x2C00)); // 1/16
    const __half2 neg72 = __halves2half2(__ushort_as_half(0xD480), __ushort_as_half(0xD480));     // -72
    uint32_t e0 = lop3_and_or(w, LO, MAGIC);
    uint32_t e1 = lop3_and_or(w, HI, MAGIC);
    uint32_t w2 = w >> 8;
    uint32_t e2 = lop3_and_or(w2, LO, MAGIC);
    uint32_t e3 = lop3_and_or(w2, HI, MAGIC);
    __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&e0), bias);
    __half2 h1 = __hfma2(*reinterpret_cast<__half2*>(&e1), sixteenth, neg72);
    __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&e2), bias);
    __half2 h3 = __hfma2(*reinterpret_cast<__half2*>(&e3), sixteenth, neg72);
    r[0] = *reinterpret_cast<uint32_t*>(&h0);
    r[1] = *reinterpret_cast<uint32_t*>(&h1);
    r[2] = *reinterpret_cast<uint32_t*>(&h2);
    r[3] = *reinterpret_cast<uint32_t*>(&h3);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Shared memory carve-up (dynamic):
//   [0, 1024)                      : mbarriers full[S], empty[S] (S <= 32), flags
//   [1024, 1024 + red_bytes)       : red[2][NW][NT*4][16] f32, sx[NW][NT*4] f32, descale[NW] f32
//   [ring_off, ...)                : stages x 16 x pitch bytes
template <int NW, int NT>
struct SmemLayout {
    static constexpr int COLS = NT * 4;
    static constexpr int RED_FLOATS = 2 * NW * COLS * TILE_ROWS;
    static constexpr int SX_FLOATS = NW * COLS;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int RED_OFF = BAR_BYTES;
    static constexpr int SX_OFF = RED_OFF + RED_FLOATS * 4;
    static constexpr int DS_OFF = SX_OFF + SX_FLOATS * 4;
    static constexpr int RING_OFF = ((DS_OFF + NW * 4 + 127) / 128) * 128;
};

template <int NW, int GPW, int NT>
__global__ void __launch_bounds__((NW + 1) * 32, 1) gemv_kernel(const GemvParams p) {
    using L = SmemLayout<NW, NT>;
    constexpr int COLS = NT * 4;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    float* red = reinterpret_cast<float*>(smem + L::RED_OFF);
    float* sx = reinterpret_cast<float*>(smem + L::SX_OFF);
    float* descale = reinterpret_cast<float*>(smem + L::DS_OFF);
    volatile int* flag = reinterpret_cast<volatile int*>(smem + 768);
    const uint32_t ring = smem_base + L::RING_OFF;
    const int S = p.stages;
    auto full_bar = [&](int s) { return smem_base + 8u * s; };
    auto empty_bar = [&](int s) { return smem_base + 256u + 8u * s; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab = blockIdx.x % p.nslab, rb = blockIdx.x / p.nslab;
    const int g0 = (int)((int64_t)p.G * slab / p.nslab), g1 = (int)((int64_t)p.G * (slab + 1) / p.nslab);
    const int ng = g1 - g0;                       // granules in this slab
    const int r0 = (int)((int64_t)p.N * rb / p.nrb), r1 = (int)((int64_t)p.N * (rb + 1) / p.nrb);
    const int ntiles = (r1 - r0 + TILE_ROWS - 1) / TILE_ROWS;
    const int slab_bytes = ng * GRAN_B;
    const uint32_t stage_bytes = TILE_ROWS * p.pitch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();

    if (warp == NW) {
        // ------------------------------------------------------------ producer warp
        if (p.wait_weights) pdl_wait();
        const uint64_t pol = policy_evict_first();
        const int64_t row_bytes = p.K / 2;
        const uint8_t* src0 = p.packed + (int64_t)g0 * GRAN_B;
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % S;
            if (i >= S) mbar_wait(empty_bar(s), ((i / S) - 1) & 1);
            const int row = r0 + i * TILE_ROWS;
            const int rows = min(TILE_ROWS, r1 - row);
            if (lane == 0) mbar_arrive_expect_tx(full_bar(s), (uint32_t)(rows * slab_bytes));
            __syncwarp();
            if (lane < rows)
                bulk_g2s_hint(ring + s * stage_bytes + lane * p.pitch, src0 + (int64_t)(row + lane) * row_bytes,
                              (uint32_t)slab_bytes, full_bar(s), pol);
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps
    const int g = lane >> 2, t = lane & 3;
    pdl_wait();   // x (and the output / workspace) belong to the stream-ordered predecessor

    // ---- build the B fragments (x operand) for this warp's granules, once
    uint32_t bf[GPW][NT][4][4];
    {
        const int part = g & 1;
        float amax = 0.0f;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int gq = warp + q * NW;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int m = nt * 4 + (g >> 1);
                if (gq < ng && m < p.M) {
                    const int64_t base = (int64_t)m * p.K + (int64_t)(g0 + gq) * GRAN_K + t * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v[4];
                        load4f(p.x, p.x_dtype, base + 4 * j, v);
                        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        int ex = 0;
        if (amax > 0.0f && amax < INFINITY) {
            (void)frexpf(amax, &ex);      // amax = f * 2^ex, f in [0.5, 1)
            ex = 14 - ex;                 // amax * 2^ex in [2^13, 2^14)
            ex = max(-100, min(100, ex));
        }
        const float up = ldexpf(1.0f, ex);
        if (lane == 0) descale[warp] = ldexpf(1.0f, -ex);

        float sums[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) sums[nt] = 0.0f;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int gq = warp + q * NW;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int m = nt * 4 + (g >> 1);
                const bool live = gq < ng && m < p.M;
                const int64_t base = (int64_t)m * p.K + (int64_t)(g0 + gq) * GRAN_K + t * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[8];
                    if (live) {
                        float a[4], b[4];
                        load4f(p.x, p.x_dtype, base + 8 * j, a);
                        load4f(p.x, p.x_dtype, base + 8 * j + 4, b);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
                    }
                    __half h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        sums[nt] += v[i];
                        const float sv = v[i] * up;
                        const __half hi = __float2half_rn(sv);
                        h[i] = part ? __float2half_rn(sv - __half2float(hi)) : hi;
                    }
                    bf[q][nt][j][0] = pack_h2(h[0], h[4]);
                    bf[q][nt][j][1] = pack_h2(h[1], h[5]);
                    bf[q][nt][j][2] = pack_h2(h[2], h[6]);
                    bf[q][nt][j][3] = pack_h2(h[3], h[7]);
                }
            }
        }
        // sum(x) over this warp's columns, per batch row (needed for the zero-point term)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float s = sums[nt];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (t == 0 && part == 0) sx[warp * COLS + nt * 4 + (g >> 1)] = s;
        }
    }

    const int ctid = threadIdx.x;                     // 0 .. NW*32-1
    const int em = ctid / TILE_ROWS, er = ctid % TILE_ROWS;   // epilogue: column (batch row), tile row
    float sxm = 0.0f;                                  // sum(x[em, slab]) in fixed warp order
    named_bar_sync(1, NW * 32);
    if (em < COLS) {
#pragma unroll
        for (int w = 0; w < NW; ++w) sxm += sx[w * COLS + em];
    }
    float dsw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) dsw[w] = descale[w];

    // ---- main loop over 16-row tiles
    for (int i = 0; i < ntiles; ++i) {
        const int s = i % S;
        mbar_wait(full_bar(s), (i / S) & 1);
        float acc[NT][2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[nt][c][r] = 0.0f;
        const uint32_t sbase = ring + s * stage_bytes + g * p.pitch + t * 16;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int gq = warp + q * NW;
            if (gq < ng) {
                const uint4 lo = lds128(sbase + gq * GRAN_B);
                const uint4 hi = lds128(sbase + 8 * p.pitch + gq * GRAN_B);
                const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w};
                const uint32_t wh[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t al[4], ah[4];
                    nibbles_to_half2x4_s8(wl[j], al);
                    nibbles_to_half2x4_s8(wh[j], ah);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        mma_m16n8k16_f16(acc[nt][0], al[0], ah[0], al[1], ah[1], bf[q][nt][j][0], bf[q][nt][j][1]);
                        mma_m16n8k16_f16(acc[nt][1], al[2], ah[2], al[3], ah[3], bf[q][nt][j][2], bf[q][nt][j][3]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));

        // cross-warp reduction: red[buf][warp][col][row]
        float* rbuf = red + (i & 1) * (NW * COLS * TILE_ROWS) + warp * (COLS * TILE_ROWS);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 4 + t;   // C columns 2t (hi part) and 2t+1 (lo part) of batch row nt*4+t
            rbuf[col * TILE_ROWS + g] = (acc[nt][0][0] + acc[nt][1][0]) + (acc[nt][0][1] + acc[nt][1][1]);
            rbuf[col * TILE_ROWS + g + 8] = (acc[nt][0][2] + acc[nt][1][2]) + (acc[nt][0][3] + acc[nt][1][3]);
        }
        named_bar_sync(1, NW * 32);
        const int row = r0 + i * TILE_ROWS + er;
        if (em < p.M && row < r1) {
            const float* rr = red + (i & 1) * (NW * COLS * TILE_ROWS) + em * TILE_ROWS + er;
            float a = 0.0f;
#pragma unroll
            for (int w = 0; w < NW; ++w) a += rr[w * (COLS * TILE_ROWS)] * dsw[w];
            const float sc = __ldg(p.scales + row), zp = __ldg(p.zps + row);
            const float v = sc * (a - (zp - 8.0f) * sxm);
            if (p.nslab == 1) store_y(p.y, p.y_dtype, (int64_t)em * p.N + row, v);
            else p.part[((int64_t)slab * p.M + em) * p.N + row] = v;
        }
    }

    // ---- cross-slab reduction by the last CTA of the row block (deterministic slab order)
    if (p.nslab > 1) {
        __threadfence();
        named_bar_sync(1, NW * 32);
        if (ctid == 0) {
            const unsigned int old = atomicAdd(p.tickets + rb, 1u);
            *flag = (old == (unsigned)p.nslab - 1u);
        }
        named_bar_sync(1, NW * 32);
        if (*flag) {
            __threadfence();
            const int nrows = r1 - r0;
            for (int idx = ctid; idx < nrows * p.M; idx += NW * 32) {
                const int m = idx / nrows, row = r0 + idx % nrows;
                float a = 0.0f;
                for (int sl = 0; sl < p.nslab; ++sl) a += __ldcg(p.part + ((int64_t)sl * p.M + m) * p.N + row);
                store_y(p.y, p.y_dtype, (int64_t)m * p.N + row, a);
            }
            if (ctid == 0) p.tickets[rb] = 0u;   // leave the workspace zeroed
        }
    }
}

struct GemvConfig {
    int nw, gpw, nt, nslab, nrb, stages, pitch;
    size_t smem;
};

bool plan(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, GemvConfig* c) {
    if (M < 1 || M > 16 || K % GRAN_K != 0 || K <= 0 || N < 1 || N > 0x7fffffff || K > 0x7fffffff) return false;
    const Tuning& tu = tuning();
    const int G = (int)(K / GRAN_K);
    const int nt = M <= 4 ? 1 : (M <= 8 ? 2 : 4);
    int nw = 8;
    if (tu.gemv_warps == 16) nw = 16;
    const int gpw_max = nw == 8 ? (nt == 1 ? 3 : 2) : (nt == 1 ? 2 : 1);
    int ctas = dev.sm_count;
    if (tu.gemv_ctas > 0 && tu.gemv_ctas < ctas) ctas = tu.gemv_ctas;
    int best_slab = 0;
    double best_cost = 1e30;
    for (int ns = 1; ns <= MAX_SLABS && ns <= G; ++ns) {
        const int ng = (G + ns - 1) / ns;
        const int gpw = (ng + nw - 1) / nw;
        if (gpw > gpw_max) continue;
        int nrb = ctas / ns;
        if (nrb < 1) continue;
        if (nrb > N) nrb = (int)N;
        if (nrb > MAX_RB) nrb = MAX_RB;
        // cost ~ bytes of the busiest CTA (+ a small charge per extra slab for the reduction)
        const double rows = (double)((N + nrb - 1) / nrb);
        const double cost = rows * ng * (1.0 + 0.01 * ns) * (1.0 + 0.1 * ((double)gpw * nw / ng - 1.0));
        if (tu.gemv_slabs == ns) { best_slab = ns; break; }
        if (cost < best_cost) { best_cost = cost; best_slab = ns; }
    }
    if (!best_slab) return false;
    const int ns = best_slab;
    const int ng = (G + ns - 1) / ns;
    c->nw = nw;
    c->nt = nt;
    c->gpw = (ng + nw - 1) / nw;
    c->nslab = ns;
    int nrb = ctas / ns;
    if (nrb > N) nrb = (int)N;
    if (nrb > MAX_RB) nrb = MAX_RB;
    c->nrb = nrb;
    c->pitch = (ng | 1) * GRAN_B;
    const int ring_off = nw == 8 ? (nt == 1 ? SmemLayout<8, 1>::RING_OFF : nt == 2 ? SmemLayout<8, 2>::RING_OFF : SmemLayout<8, 4>::RING_OFF)
                                 : (nt == 1 ? SmemLayout<16, 1>::RING_OFF : nt == 2 ? SmemLayout<16, 2>::RING_OFF : SmemLayout<16, 4>::RING_OFF);
    const int stage_bytes = TILE_ROWS * c->pitch;
    const int avail = dev.max_smem_optin - ring_off;
    int stages = avail / stage_bytes;
    const int rows = (int)((N + nrb - 1) / nrb);
    const int ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (stages > ntiles) stages = ntiles;
    if (stages > 32) stages = 32;
    if (tu.gemv_stages > 0 && tu.gemv_stages < stages) stages = tu.gemv_stages;
    if (stages < 1) return false;
    c->stages = stages;
    c->smem = (size_t)ring_off + (size_t)stages * stage_bytes;
    return true;
}

template <int NW, int GPW, int NT>
int launch_inst(const GemvConfig& c, const GemvParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_kernel<NW, GPW, NT>;
    static thread_local int attr_dev_smem[64] = {0};
    int dev = 0;
    B200Q_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_dev_smem[dev] < (int)c.smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        attr_dev_smem[dev] = (int)c.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(c.nslab * c.nrb));
    cfg.blockDim = dim3((NW + 1) * 32);
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, p), "gemv launch");
}

}  // namespace

bool gemv_supported(int64_t M, int64_t N, int64_t K, int x_dtype) {
    (void)x_dtype;
    DeviceInfo d;
    d.sm_count = 148;
    d.max_smem_optin = 232448;
    GemvConfig c;
    return plan(d, M, N, K, &c);
}

size_t gemv_ws_bytes(int64_t M, int64_t N, int64_t K) {
    if (!gemv_supported(M, N, K, B200Q_F32)) return 0;
    return (size_t)MAX_RB * 4 + (size_t)MAX_SLABS * M * N * 4;
}

int launch_gemv(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                const float* scales, const float* zps, void* y, int y_dtype, int64_t M, int64_t N,
                int64_t K, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st) {
    GemvConfig c;
    if (!plan(dev, M, N, K, &c)) return set_error(B200Q_EINVAL, "gemv: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv: x and packed must be 16-byte aligned");
    GemvParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.y = y;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.nslab = c.nslab; p.nrb = c.nrb; p.G = (int)(K / GRAN_K);
    p.stages = c.stages; p.pitch = c.pitch;
    const bool is_static = (flags & B200Q_FLAG_STATIC_WEIGHTS) != 0;
    const bool pdl = tuning().gemv_pdl != 0;
    p.wait_weights = is_static ? 0 : 1;
    if (c.nslab > 1) {
        const size_t need = (size_t)MAX_RB * 4 + (size_t)c.nslab * M * N * 4;
        if (!ws || ws_bytes < need) return set_error(B200Q_EWORKSPACE, "gemv: workspace too small (%zu < %zu)", ws_bytes, need);
        if (reinterpret_cast<uintptr_t>(ws) & 15) return set_error(B200Q_EALIGN, "gemv: workspace must be 16-byte aligned");
        p.tickets = static_cast<unsigned int*>(ws);
        p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + (size_t)MAX_RB * 4);
    }
#define B200Q_GEMV_CASE(NW_, GPW_, NT_) \
    if (c.nw == NW_ && c.gpw == GPW_ && c.nt == NT_) return launch_inst<NW_, GPW_, NT_>(c, p, pdl, st);
    B200Q_GEMV_CASE(8, 1, 1) B200Q_GEMV_CASE(8, 2, 1) B200Q_GEMV_CASE(8, 3, 1)
    B200Q_GEMV_CASE(8, 1, 2) B200Q_GEMV_CASE(8, 2, 2)
    B200Q_GEMV_CASE(8, 1, 4) B200Q_GEMV_CASE(8, 2, 4)
    B200Q_GEMV_CASE(16, 1, 1) B200Q_GEMV_CASE(16, 2, 1)
    B200Q_GEMV_CASE(16, 1, 2) B200Q_GEMV_CASE(16, 1, 4)
#undef B200Q_GEMV_CASE
    return set_error(B200Q_EINVAL, "gemv: no kernel instance for nw=%d gpw=%d nt=%d", c.nw, c.gpw, c.nt);
}

}  // namespace b200q
