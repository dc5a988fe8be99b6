// Decode path on the 5th-generation tensor cores: y[M,N] = x[M,K] @ dequant(W)^T for M <= 8.
//
// The legacy mma.sync path (gemv.cu) is capped by HMMA issue rate on sm_100 (~9 clk per m16n8k16
// per SM, measured: profiles/r01_gemv_notes.md), well below what 6.5 TB/s of packed INT4 needs.
// Here the weights are the A operand of tcgen05.mma and live in TENSOR MEMORY:
//
//   TMA row copies (cp.async.bulk) -> shared-memory ring -> LDS.128 (one thread per weight row)
//   -> LOP3 nibble->fp16-subnormal (no arithmetic) -> tcgen05.st into a TMEM A slot
//   -> tcgen05.mma.kind::f16  D[128 rows x 16] += A[tmem 128 x 16] * B[smem 16 x 16]
//
//   * B is the activation operand: x as an fp16 hi/lo split of x * 2^e (two columns per batch row),
//     converted ONCE per CTA into the canonical K-major core-matrix layout in shared memory, in the
//     nibble order the A registers come out in (so the dequantisation needs no permutes).
//   * one CTA per SM streams rows [r0,r1) (<= 128 per tile) x its K slab; a tile's accumulator is a
//     single 128x16 fp32 TMEM block, read back with tcgen05.ld for the epilogue
//     y = s * ((D_hi + D_lo) * 2^(24-e) - zp * sum(x)).
//   * 16 dequant warps: warp w serves TMEM lane quarter w & 3 (a hardware rule: a warp can only
//     touch lanes 32*(warp%4)..+31) and granules g == (w >> 2) mod 4; warp 16 issues the MMAs.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279.
#include "internal.h"
#include "ptx.cuh"
#include "tc.cuh"

namespace b200q {

namespace {

constexpr int TC_DQ_WARPS = 16;
constexpr int TC_THREADS = (TC_DQ_WARPS + 1) * 32;
constexpr int GRAN_K = 128;              // columns per granule (= 64 TMEM columns = 8 MMAs)
constexpr int GRAN_B = 64;               // packed bytes per granule per row
constexpr int CHUNK_G = 8;               // granules per weight chunk
constexpr int CHUNK_B = CHUNK_G * GRAN_B;   // 512 bytes per row per chunk
constexpr int PITCH = CHUNK_B + 16;      // row pitch in a chunk slot: conflict-free LDS.128 by row
constexpr int A_SLOTS = 6;               // TMEM A ring: 6 x 64 columns (+ 128 for the accumulators) = 512
constexpr int D_CHAINS = 8;              // independent accumulators (16 columns each), summed in the epilogue
constexpr int A_BASE = 128;              // first TMEM column of the A ring
constexpr int A_COLS = 64;
constexpr int TMEM_COLS = 512;
constexpr int MAX_WSLOTS = 8;
constexpr int TILE_ROWS = 128;
constexpr int MAX_SLABS = 16;
constexpr int MAX_RB = 1024;

// shared memory map
constexpr int OFF_WFULL = 0;      // [MAX_WSLOTS][4] mbarriers
constexpr int OFF_WEMPTY = 256;   // [MAX_WSLOTS][4]
constexpr int OFF_AFULL = 512;    // [A_SLOTS]
constexpr int OFF_AEMPTY = 576;   // [A_SLOTS]
constexpr int OFF_XREADY = 640;
constexpr int OFF_DFULL = 648;
constexpr int OFF_DEMPTY = 656;
constexpr int OFF_TMEMPTR = 672;
constexpr int OFF_FLAG = 676;
constexpr int OFF_RED = 1024;     // float red[16][8][2]
constexpr int OFF_ROWC = 2048;    // float descale[8], sumx[8], up[8]
constexpr int OFF_EPI = 2176;     // float epi[128][8]: per-row accumulator sums for the epilogue loop (4 KB)
constexpr int OFF_XOP = 7168;     // x operand, then the weight ring

struct TcParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    void* y;
    float* part;
    unsigned int* tickets;
    int x_dtype, y_dtype;
    int M, N, K;
    int nslab, nrb, G;
    int xg;            // 8-row groups of the x operand: 1 (M <= 4) or 2 (M <= 8)
    int wslots;        // weight ring depth in chunks
    int slot_bytes;    // bytes per chunk slot = tile_rows * PITCH
    int ring_off;
    int wait_weights;
    int debug;
};

template <typename XT> __device__ __forceinline__ void load4f(const XT* x, int64_t idx, float (&v)[4]);
template <> __device__ __forceinline__ void load4f<float>(const float* x, int64_t idx, float (&v)[4]) {
    float4 a = *reinterpret_cast<const float4*>(x + idx);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4f<__half>(const __half* x, int64_t idx, float (&v)[4]) {
    uint2 r = *reinterpret_cast<const uint2*>(x + idx);
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void load4f<__nv_bfloat16>(const __nv_bfloat16* x, int64_t idx, float (&v)[4]) {
    uint2 r = *reinterpret_cast<const uint2*>(x + idx);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ void store_y(void* y, int dtype, int64_t idx, float v) {
    if (dtype == B200Q_F32) static_cast<float*>(y)[idx] = v;
    else if (dtype == B200Q_F16) static_cast<__half*>(y)[idx] = __float2half_rn(v);
    else static_cast<__nv_bfloat16*>(y)[idx] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <typename XT>
__global__ void __launch_bounds__(TC_THREADS, 1) gemv_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.wslots;
    auto wfull = [&](int s, int q) { return sb + OFF_WFULL + 8u * (s * 4 + q); };
    auto wempty = [&](int s, int q) { return sb + OFF_WEMPTY + 8u * (s * 4 + q); };
    auto afull = [&](int a) { return sb + OFF_AFULL + 8u * a; };
    auto aempty = [&](int a) { return sb + OFF_AEMPTY + 8u * a; };
    const uint32_t xready = sb + OFF_XREADY, dfull = sb + OFF_DFULL, dempty = sb + OFF_DEMPTY;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEMPTR);
    volatile int* flag = reinterpret_cast<volatile int*>(smem + OFF_FLAG);
    float* red = reinterpret_cast<float*>(smem + OFF_RED);
    float* rowc = reinterpret_cast<float*>(smem + OFF_ROWC);
    const uint32_t xop = sb + OFF_XOP;
    const uint32_t ring = sb + p.ring_off;

    const int slab = blockIdx.x % p.nslab, rb = blockIdx.x / p.nslab;
    const int g0 = (int)((int64_t)p.G * slab / p.nslab), g1 = (int)((int64_t)p.G * (slab + 1) / p.nslab);
    const int ng = g1 - g0;                                   // granules in this K slab
    const int nc = (ng + CHUNK_G - 1) / CHUNK_G;              // weight chunks per tile
    const int r0 = (int)((int64_t)p.N * rb / p.nrb), r1 = (int)((int64_t)p.N * (rb + 1) / p.nrb);
    const int ntile = (r1 - r0 + TILE_ROWS - 1) / TILE_ROWS;
    const int tile_rows = (r1 - r0 + ntile - 1) / ntile;     // <= 128, balanced over the tiles
    const int kslab0 = g0 * GRAN_K, kslab = ng * GRAN_K;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s)
            for (int q = 0; q < 4; ++q) {
                mbar_init(wfull(s, q), 1);
                mbar_init(wempty(s, q), 4);
            }
        for (int a = 0; a < A_SLOTS; ++a) {
            mbar_init(afull(a), 4);
            mbar_init(aempty(a), 1);
        }
        mbar_init(xready, TC_DQ_WARPS);
        mbar_init(dfull, 1);
        mbar_init(dempty, 4);
        fence_mbar_init();
    }
    if (warp == TC_DQ_WARPS) {
        tmem_alloc(sb + OFF_TMEMPTR, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_launch_dependents();
    const uint32_t tmem = *tmem_ptr;        // lane 0, column 0 of the allocation
    const uint32_t d_tmem = tmem;           // accumulators: D_CHAINS x 16 columns from column 0
    // (back-to-back tcgen05.mma into ONE 128x16 accumulator serialise on the accumulator round trip;
    //  MMA j of every granule therefore owns its own accumulator -- measured in profiles/r01_gemv_notes.md)

    if (warp == TC_DQ_WARPS) {
        // ================================================================= MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(128, 16, 0);
            const uint32_t lbo = p.xg * 128u, sbo = p.xg == 2 ? 128u : 0u;   // xg == 1: both 8-row groups alias
            mbar_wait(xready, 0);
            tc_fence_after_sync();
            int gs = 0;                                   // running granule sequence number
            for (int t = 0; t < ntile; ++t) {
                if (t > 0) {                              // epilogue of the previous tile has drained D
                    mbar_wait(dempty, (t - 1) & 1);
                    tc_fence_after_sync();
                }
                for (int gi = 0; gi < ng; ++gi, ++gs) {
                    const int a = gs % A_SLOTS;
                    mbar_wait(afull(a), (gs / A_SLOTS) & 1);
                    tc_fence_after_sync();
                    const uint32_t a_tmem = tmem + A_BASE + A_COLS * a;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (p.debug == 3) break;
                        const uint32_t kk = gi * 8 + j;                       // k16 step inside the slab
                        const uint64_t bd = smem_desc(xop + kk * 2u * lbo, lbo, sbo, SWIZZLE_NONE);
                        mma_ts_f16(d_tmem + 16 * (j % D_CHAINS), a_tmem + 8 * j, bd, idesc, gi != 0 ? 1u : 0u);
                    }
                    tc_commit(aempty(a));
                }
                tc_commit(dfull);
            }
        }
    } else {
        // ================================================================= dequant warps
        const int q = warp & 3, kg = warp >> 2;
        const int dtid = threadIdx.x;                                  // 0..511
        const uint64_t pol = policy_evict_first();
        const int64_t row_bytes = p.K / 2;

        // chunk sequence number cs = t * nc + c lives in ring slot cs % S; quarter q has its own barrier
        auto issue_chunk = [&](int cs) {
            const int t = cs / nc, c = cs - t * nc, s = cs % S;
            const int row_lo = r0 + t * tile_rows + 32 * q;
            const int row_hi = min(r0 + (t + 1) * tile_rows, r1);
            const int rows = max(0, min(32, row_hi - row_lo));
            const int cb = min(CHUNK_G, ng - c * CHUNK_G) * GRAN_B;
            if (lane == 0) mbar_arrive_expect_tx(wfull(s, q), (uint32_t)(rows * cb));
            __syncwarp();
            if (lane < rows)
                bulk_g2s_hint(ring + s * p.slot_bytes + (32 * q + lane) * PITCH,
                              p.packed + (int64_t)(row_lo + lane) * row_bytes + (int64_t)g0 * GRAN_B + c * CHUNK_B,
                              (uint32_t)cb, wfull(s, q), pol);
        };
        const int total_chunks = ntile * nc;
        if (p.wait_weights) pdl_wait();
        if (p.debug != 2)
            for (int cs = kg; cs < min(S, total_chunks); cs += 4) issue_chunk(cs);

        pdl_wait();          // x, y and the workspace belong to the stream-ordered predecessor
        if (p.debug == 5) {   // ablation: skip the x operand
            __syncwarp();
            if (lane == 0) mbar_arrive(xready);
        } else {

        // ---- x operand.  pass 1: per batch row amax and sum over the slab (block reduction).
        // Loops over batch rows are deliberately NOT unrolled: this kernel runs for a few
        // microseconds and its instruction footprint has to stay inside the instruction cache.
        const XT* xp = static_cast<const XT*>(p.x);
#pragma unroll 1
        for (int m = 0; m < p.M; ++m) {
            float am = 0.0f, sm = 0.0f;
            for (int k = dtid * 4; k < kslab; k += TC_DQ_WARPS * 32 * 4) {
                float v[4];
                load4f<XT>(xp, (int64_t)m * p.K + kslab0 + k, v);
                am = fmaxf(am, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
                sm += (v[0] + v[1]) + (v[2] + v[3]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
            }
            if (lane == 0) {
                red[(warp * 8 + m) * 2] = am;
                red[(warp * 8 + m) * 2 + 1] = sm;
            }
        }
        named_bar_sync(1, TC_DQ_WARPS * 32);
        if (dtid < p.M) {        // per-row scale 2^e with amax * 2^e in [2^13, 2^14), fixed summation order
            float am = 0.0f, sm = 0.0f;
#pragma unroll 1
            for (int w = 0; w < TC_DQ_WARPS; ++w) {
                am = fmaxf(am, red[(w * 8 + dtid) * 2]);
                sm += red[(w * 8 + dtid) * 2 + 1];
            }
            int ex = 0;
            if (am > 0.0f && am < INFINITY) {
                ex = 140 - (int)((__float_as_uint(am) >> 23) & 0xffu);
                ex = max(-100, min(100, ex));
            }
            rowc[dtid] = __uint_as_float((uint32_t)(127 + 24 - ex) << 23);    // descale 2^(24-e)
            rowc[8 + dtid] = sm;
            rowc[16 + dtid] = __uint_as_float((uint32_t)(127 + ex) << 23);    // 2^e
        }
        named_bar_sync(1, TC_DQ_WARPS * 32);
        // pass 2: one 16-byte core-matrix row per item: item -> (k core kc, operand row n = 2m + part)
        {
            const int rows_op = 8 * p.xg, rshift = p.xg == 2 ? 4 : 3;
            const int items = (kslab / 8) * rows_op;
#pragma unroll 1
            for (int it = dtid; it < items; it += TC_DQ_WARPS * 32) {
                const int kc = it >> rshift, n = it & (rows_op - 1);
                const int m = n >> 1, part = n & 1;
                uint4 out = make_uint4(0u, 0u, 0u, 0u);
                if (m < p.M) {
                    float a[4], b[4];
                    const int64_t base = (int64_t)m * p.K + kslab0 + kc * 8;
                    load4f<XT>(xp, base, a);
                    load4f<XT>(xp, base + 4, b);
                    const float v[8] = {a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]};
                    const float u = rowc[16 + m];
                    const float u_hi = u * 0.0625f;          // columns that meet a high nibble (2^-20 vs 2^-24)
                    __half h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float sv = v[i] * ((i & 1) ? u_hi : u);
                        const __half hi = __float2half_rn(sv);
                        h[i] = part ? __float2half_rn(sv - __half2float(hi)) : hi;
                    }
                    // register order of the A side: (n0,n4) (n1,n5) (n2,n6) (n3,n7)
                    out = make_uint4(pack_h2(h[0], h[4]), pack_h2(h[1], h[5]), pack_h2(h[2], h[6]), pack_h2(h[3], h[7]));
                }
                sts128(xop + it * 16u, out);
            }
            fence_proxy_async_smem();       // generic-proxy writes -> visible to tcgen05.mma's smem reads
            __syncwarp();
            if (lane == 0) mbar_arrive(xready);
        }
        }

        // ---- main loop: every warp visits every chunk; inside a chunk it owns granules kg and kg + 4
        for (int t = 0; t < ntile; ++t) {
            const int row_lo = r0 + t * tile_rows + 32 * q;
            const int row_hi = min(r0 + (t + 1) * tile_rows, r1);
            const bool live = row_lo + lane < row_hi;
            for (int c = 0; c < nc; ++c) {
                const int cs = t * nc + c, s = cs % S;
                if (p.debug != 2) mbar_wait(wfull(s, q), (cs / S) & 1);
                const int gend = min(ng, (c + 1) * CHUNK_G);
                uint4 w[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gi = c * CHUNK_G + kg + 4 * h;
                    const uint32_t src = ring + s * p.slot_bytes + (32 * q + lane) * PITCH + (kg + 4 * h) * GRAN_B;
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        w[h][v4] = make_uint4(0u, 0u, 0u, 0u);
                        if (live && gi < gend && p.debug != 1) w[h][v4] = lds128(src + 16 * v4);
                    }
                }
                // the chunk's bytes are in registers: release the ring slot (and refill it)
                if (cs + S < total_chunks) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(wempty(s, q));
                    if ((cs & 3) == kg && p.debug != 2) {
                        mbar_wait(wempty(s, q), (cs / S) & 1);
                        issue_chunk(cs + S);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gi = c * CHUNK_G + kg + 4 * h;
                    if (gi < gend) {
                        const int gs = t * ng + gi, a = gs % A_SLOTS;
                        if (gs >= A_SLOTS) {
                            mbar_wait(aempty(a), ((gs / A_SLOTS) - 1) & 1);
                            tc_fence_after_sync();
                        }
                        const uint32_t dst = tmem + ((uint32_t)(32 * q) << 16) + A_BASE + A_COLS * a;
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            const uint32_t ws[4] = {w[h][v4].x, w[h][v4].y, w[h][v4].z, w[h][v4].w};
                            uint32_t rr[16];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // nibble n left in the low mantissa bits of a zero-exponent fp16 is the
                                // subnormal n * 2^-24 (bits 0-3) or n * 2^-20 (bits 4-7): exact, no arithmetic
                                const uint32_t w2 = ws[j] >> 8;
                                rr[4 * j + 0] = ws[j] & 0x000f000fu;
                                rr[4 * j + 1] = ws[j] & 0x00f000f0u;
                                rr[4 * j + 2] = w2 & 0x000f000fu;
                                rr[4 * j + 3] = w2 & 0x00f000f0u;
                            }
                            if (p.debug != 4) tmem_st16(dst + 16 * v4, rr);
                        }
                        if (p.debug != 4) tmem_wait_st();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(afull(a));
                    }
                }
            }

            // ---- epilogue of tile t (warps 0-3: one per TMEM lane quarter)
            if (kg == 0) {
                mbar_wait(dfull, t & 1);
                tc_fence_after_sync();
                float dsum[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) dsum[i] = 0.0f;
#pragma unroll
                for (int ch = 0; ch < D_CHAINS; ++ch) {
                    uint32_t d[16];
                    tmem_ld16(d_tmem + ((uint32_t)(32 * q) << 16) + 16 * ch, d);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) dsum[i] += __uint_as_float(d[i]);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(dempty);
                const int row = row_lo + lane;
                float* epi = reinterpret_cast<float*>(smem + OFF_EPI) + (32 * q + lane) * 8;
#pragma unroll
                for (int m = 0; m < 8; ++m) epi[m] = dsum[2 * m] + dsum[2 * m + 1];
                if (live) {
                    const float sc = __ldg(p.scales + row), zp = __ldg(p.zps + row);
#pragma unroll 1
                    for (int m = 0; m < p.M; ++m) {
                        const float v = sc * (epi[m] * rowc[m] - zp * rowc[8 + m]);
                        if (p.nslab == 1) store_y(p.y, p.y_dtype, (int64_t)m * p.N + row, v);
                        else p.part[((int64_t)slab * p.M + m) * p.N + row] = v;
                    }
                }
            }
        }

        // ---- cross-slab reduction by the last CTA of the row block (deterministic slab order)
        if (p.nslab > 1 && kg == 0) {
            __threadfence();
            named_bar_sync(2, 128);
            if (dtid == 0) {
                const unsigned int old = atomicAdd(p.tickets + rb, 1u);
                *flag = (old == (unsigned)p.nslab - 1u);
            }
            named_bar_sync(2, 128);
            if (*flag) {
                __threadfence();
                const int nrows = r1 - r0;
                const int t128 = q * 32 + lane;
                for (int idx = t128; idx < nrows * p.M; idx += 128) {
                    const int m = idx / nrows, row = r0 + idx % nrows;
                    float acc = 0.0f;
                    for (int sl = 0; sl < p.nslab; ++sl) acc += __ldcg(p.part + ((int64_t)sl * p.M + m) * p.N + row);
                    store_y(p.y, p.y_dtype, (int64_t)m * p.N + row, acc);
                }
                if (t128 == 0) p.tickets[rb] = 0u;
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == TC_DQ_WARPS) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

struct TcConfig {
    int nslab, nrb, xg, wslots, slot_bytes, ring_off;
    size_t smem;
};

bool plan(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, TcConfig* c) {
    if (M < 1 || M > 8 || K % GRAN_K != 0 || K <= 0 || N < 1 || N > 0x7fffffff || K > 0x7fffffff) return false;
    const Tuning& tu = tuning();
    const int G = (int)(K / GRAN_K);
    const int xg = M <= 4 ? 1 : 2;
    int ctas = dev.sm_count;
    if (tu.gemv_ctas > 0 && tu.gemv_ctas < ctas) ctas = tu.gemv_ctas;
    double best = 1e30;
    bool found = false;
    for (int ns = 1; ns <= MAX_SLABS && ns <= G; ++ns) {
        if (tu.gemv_slabs > 0 && tu.gemv_slabs != ns) continue;
        const int ng = (G + ns - 1) / ns;
        int nrb = ctas / ns;
        if (nrb < 1) continue;
        if (nrb > N) nrb = (int)N;
        if (nrb > MAX_RB) nrb = MAX_RB;
        const int rows = (int)((N + nrb - 1) / nrb);
        const int ntile = (rows + TILE_ROWS - 1) / TILE_ROWS;
        const int tile_rows = (rows + ntile - 1) / ntile;
        const int nc = (ng + CHUNK_G - 1) / CHUNK_G;
        const int xop_bytes = ng * GRAN_K * 16 * xg;             // K_slab/8 core matrices x 8*xg rows x 16 B
        const int ring_off = ((OFF_XOP + xop_bytes + 127) / 128) * 128;
        const int slot_bytes = ((tile_rows * PITCH + 127) / 128) * 128;
        int wslots = (dev.max_smem_optin - ring_off) / slot_bytes;
        if (wslots > ntile * nc) wslots = ntile * nc;
        if (wslots > MAX_WSLOTS) wslots = MAX_WSLOTS;
        if (tu.gemv_stages > 0 && tu.gemv_stages < wslots) wslots = tu.gemv_stages;
        if (wslots < 2 && wslots < ntile * nc) continue;
        if (wslots < 1) continue;
        double cost = (double)rows * ng * (1.0 + 0.03 * (ns - 1));
        if (wslots < ntile * nc) cost *= 1.0 + 0.1 * (1.0 - (double)wslots / (ntile * nc));
        if (cost < best) {
            best = cost;
            found = true;
            c->nslab = ns; c->nrb = nrb; c->xg = xg; c->wslots = wslots; c->slot_bytes = slot_bytes;
            c->ring_off = ring_off;
            c->smem = (size_t)ring_off + (size_t)wslots * slot_bytes;
        }
    }
    return found;
}

}  // namespace

bool gemv_tc_supported(int64_t M, int64_t N, int64_t K) {
    DeviceInfo d;
    d.sm_count = 148;
    d.max_smem_optin = 232448;
    TcConfig c;
    return plan(d, M, N, K, &c);
}

size_t gemv_tc_ws_bytes(int64_t M, int64_t N, int64_t K) {
    if (!gemv_tc_supported(M, N, K)) return 0;
    return (size_t)MAX_RB * 4 + (size_t)MAX_SLABS * M * N * 4;
}

int launch_gemv_tc(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                   const float* scales, const float* zps, void* y, int y_dtype, int64_t M, int64_t N,
                   int64_t K, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st) {
    TcConfig c;
    if (!plan(dev, M, N, K, &c)) return set_error(B200Q_EINVAL, "gemv_tc: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv_tc: x and packed must be 16-byte aligned");
    TcParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.y = y;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.nslab = c.nslab; p.nrb = c.nrb; p.G = (int)(K / GRAN_K);
    p.xg = c.xg; p.wslots = c.wslots; p.slot_bytes = c.slot_bytes; p.ring_off = c.ring_off;
    const bool is_static = (flags & B200Q_FLAG_STATIC_WEIGHTS) != 0;
    const bool pdl = tuning().gemv_pdl != 0;
    p.wait_weights = is_static ? 0 : 1;
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    if (c.nslab > 1) {
        const size_t need = (size_t)MAX_RB * 4 + (size_t)c.nslab * M * N * 4;
        if (!ws || ws_bytes < need) return set_error(B200Q_EWORKSPACE, "gemv_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        if (reinterpret_cast<uintptr_t>(ws) & 15) return set_error(B200Q_EALIGN, "gemv_tc: workspace must be 16-byte aligned");
        p.tickets = static_cast<unsigned int*>(ws);
        p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + (size_t)MAX_RB * 4);
    }
    void (*kfn)(const TcParams) = x_dtype == B200Q_F32 ? gemv_tc_kernel<float>
                                  : x_dtype == B200Q_F16 ? gemv_tc_kernel<__half> : gemv_tc_kernel<__nv_bfloat16>;
    static thread_local int attr_dev_smem[64][3] = {{0}};
    int devi = 0;
    B200Q_CUDA(cudaGetDevice(&devi));
    if (devi >= 0 && devi < 64 && attr_dev_smem[devi][x_dtype] < (int)c.smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        attr_dev_smem[devi][x_dtype] = (int)c.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(c.nslab * c.nrb));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, p), "gemv_tc launch");
}

}  // namespace b200q
