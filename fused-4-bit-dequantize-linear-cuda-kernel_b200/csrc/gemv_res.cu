// Decode path, "resident slab" variant: y[M,N] = x[M,K] @ dequant(W)^T for M <= 8 when a CTA's whole
// share of W (its rows x K/2 packed bytes) fits in shared memory next to the x operand.
//
// Same arithmetic as gemv.cu (exact integers: u8 nibbles x signed base-256 limbs of round(x * 2^e) on
// IMMA m16n8k32, s32 accumulation, fp64 epilogue) and therefore bit-identical results; what differs is
// the amount of code every warp executes.  gemv.cu spends ~1.1 k - 1.8 k instructions per warp per launch,
// most of them outside the 5-tile main loop (profiles/r01_gemv_notes.md): with 16 warps on 4 schedulers
// that is 2 - 4 us of pure issue time.  This kernel
//   * has no ring: tile i is ONE bulk copy into its own slot with its own single-use mbarrier;
//   * builds the x operand cooperatively: every thread converts 8 consecutive activations once
//     (coalesced loads, 8 F2I) into the four limb planes of a shared-memory image laid out in mma
//     B-fragment order, so a lane fetches the operand of a granule with two 16-byte loads;
//   * reduces the 16 warp partials of a tile with shared-memory integer atomics (exact and order
//     independent) into a 512-byte plane per (tile, n-tile): no partial buffers, no reduction rounds,
//     one thread per output in the epilogue;
//   * needs no workspace and no second kernel: consecutive decode layers are main -> main, which
//     programmatic dependent launch overlaps CTA by CTA.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output).
#include <cmath>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

// per-CTA phase timestamps (globaltimer ns) and first / last CTA start / end of the last 64 launches;
// bench-only (gemv_debug bit 3), read back with b200q_debug_read_prof_res / b200q_debug_wall_res
__device__ long long g_res_prof[256 * 16];
__device__ unsigned long long g_res_wall[64 * 4];

namespace {

constexpr int NW = 16;               // warps per CTA
constexpr int NTHR = NW * 32;
constexpr int TILE_ROWS = 16;
constexpr int GRAN_K = 128;
constexpr int GRAN_B = 64;
constexpr int MAX_TILES = 32;

struct ResParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    void* y;
    const uint8_t* next_packed;      // L2 prefetch hint (weights of the next fused linear), may be null
    unsigned long long next_bytes;
    unsigned int next_chunk;         // next_bytes / gridDim.x
    int x_dtype, y_dtype;
    int M, N, K;
    int rows_q, rows_rem;            // CTA b owns rows_q (+1 if b < rows_rem) rows
    int G;                           // K / 128
    int acc_off, img_off, tile_off;  // byte offsets in dynamic shared memory
    int wait_weights;                // 1: weights may be written by the preceding kernel
    int early_tiles;                 // tiles requested before griddepcontrol.wait
    int pf_mode;
    unsigned int launch_no;
    int debug;
};

__device__ __forceinline__ void load8f(const void* x, int dtype, int64_t idx, float (&v)[8]) {
    if (dtype == B200Q_F32) {
        const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + idx);
        const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + idx + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (dtype == B200Q_F16) {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const __half*>(x) + idx);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    } else {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(x) + idx);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}

__device__ __forceinline__ void store_out(void* y, int dtype, int64_t idx, float v) {
    if (dtype == B200Q_F32) static_cast<float*>(y)[idx] = v;
    else if (dtype == B200Q_F16) static_cast<__half*>(y)[idx] = __float2half_rn(v);
    else static_cast<__nv_bfloat16*>(y)[idx] = __float2bfloat16_rn(v);
}

// D(16x8,s32) += A(16x32,u8,row) * B(32x8,s8,col)      SASS: IMMA.16832.U8.S8
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                     uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void red_add_s32(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Shared memory (dynamic):
//   [0, 256)        mbarriers, one per tile
//   [256, 1280)     s_amax[8][16] f32   (then s_up[8] f32 at +512, s_ex[8] s32 at +544, s_txs[8][2] s32 at +576)
//   [1280, 2304)    s_tx[16 warps][8 rows][2] s32
//   [acc_off, ..)   acc[tile][NT*8 columns][16 rows] s32   (zeroed, RED.ADD target)
//   [img_off, ..)   x image of ONE n-tile (two batch rows): [h][granule][limb][t][word] x {b0, b1}
//   [tile_off, ..)  tiles: ntiles x 16 rows x K/2 bytes
constexpr int OFF_AMAX = 256, OFF_UP = 768, OFF_EX = 800, OFF_TXS = 832, OFF_TX = 1280, OFF_END = 2304;

template <int GPW, int NT>
__global__ void __launch_bounds__(NTHR, 1) gemv_res_kernel(const ResParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const bool prof = (p.debug & 8) && tid == 0 && blockIdx.x < 256;
    auto stamp = [&](int i) {
        if (prof) {
            long long tnow;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
            g_res_prof[blockIdx.x * 16 + i] = tnow;
        }
    };
    stamp(0);
    unsigned long long wall0 = 0;
    if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wall0));

    const int b = blockIdx.x;
    const int r0 = b * p.rows_q + min(b, p.rows_rem);
    const int nrows = p.rows_q + (b < p.rows_rem ? 1 : 0);
    const int ntiles = (nrows + TILE_ROWS - 1) / TILE_ROWS;
    const int row_bytes = p.K >> 1;
    const uint32_t tile_bytes = (uint32_t)(TILE_ROWS * row_bytes);
    float* s_amax = reinterpret_cast<float*>(smem + OFF_AMAX);
    float* s_up = reinterpret_cast<float*>(smem + OFF_UP);
    int* s_ex = reinterpret_cast<int*>(smem + OFF_EX);
    int* s_txs = reinterpret_cast<int*>(smem + OFF_TXS);
    int* s_tx = reinterpret_cast<int*>(smem + OFF_TX);
    int* acc = reinterpret_cast<int*>(smem + p.acc_off);
    uint2* img = reinterpret_cast<uint2*>(smem + p.img_off);

    if (tid == 0) {
        for (int i = 0; i < ntiles; ++i) mbar_init(sbase + 8u * i, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < ntiles * NT * 128; i += NTHR) acc[i] = 0;
    __syncthreads();
    pdl_launch_dependents();
    stamp(1);

    const bool skip_loads = (p.debug & 2) != 0;
    auto issue_tile = [&](int i) {
        const int rows = min(TILE_ROWS, nrows - i * TILE_ROWS);
        const uint32_t bytes = (uint32_t)(rows * row_bytes);
        mbar_arrive_expect_tx(sbase + 8u * i, bytes);
        bulk_g2s_hint(sbase + p.tile_off + i * tile_bytes, p.packed + (int64_t)(r0 + i * TILE_ROWS) * row_bytes, bytes,
                      sbase + 8u * i, policy_evict_first());
    };
    auto prefetch_next = [&]() {
        if (!p.next_bytes || p.pf_mode != 1) return;
        const unsigned long long per = (((unsigned long long)p.next_chunk) + 127ull) & ~127ull;
        const unsigned long long beg = min(per * blockIdx.x, p.next_bytes), end = min(beg + per, p.next_bytes);
        for (unsigned long long off = beg; off < end; off += 32768ull) {
            const unsigned int n = (unsigned int)min(32768ull, end - off) & ~15u;
            if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.next_packed + off), "r"(n) : "memory");
        }
    };
    if (p.wait_weights) pdl_wait();
    const int early = min(p.early_tiles, ntiles);
    if (tid == 0 && !skip_loads) {
        for (int i = 0; i < early; ++i) issue_tile(i);
        if (early == ntiles) prefetch_next();
    }
    stamp(2);
    pdl_wait();          // x (and y) belong to the stream-ordered predecessor
    stamp(3);
    if (tid == 0 && !skip_loads && early < ntiles) {
        for (int i = early; i < ntiles; ++i) issue_tile(i);
        prefetch_next();
    }

    // ---- pass 1: amax of every batch row (coalesced float4 / 8-byte loads, all threads)
    {
        const int K8 = p.K >> 3;
        for (int m = 0; m < p.M; ++m) {
            float am = 0.0f;
            for (int i = tid; i < K8; i += NTHR) {
                float v[8];
                load8f(p.x, p.x_dtype, (int64_t)m * p.K + i * 8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) am = fmaxf(am, fabsf(v[e]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
            if (lane == 0) s_amax[m * NW + warp] = am;
        }
    }
    __syncthreads();
    if (warp < p.M) {            // warp m: exponent of batch row m
        float am = lane < NW ? s_amax[warp * NW + lane] : 0.0f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
        int ex = 0;
        if (am > 0.0f && am < INFINITY) ex = max(-96, min(126, 156 - (int)(__float_as_uint(am) >> 23)));
        if (lane == 0) {
            s_ex[warp] = ex;
            s_up[warp] = __uint_as_float((uint32_t)(127 + ex) << 23);
        }
    }
    __syncthreads();
    stamp(4);

    // ---- pass 2: x image of one n-tile at a time, then every lane pulls its B fragments
    // lane (g, t) holds mma column g = 4h + l: limb l of batch row 2*nt + h
    const int l = g & 3, h = g >> 2;
    const int G = p.G;
    uint32_t bf[GPW][NT][4][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int live_rows = min(2, p.M - 2 * nt);          // batch rows of this n-tile that exist (may be <= 0)
        if (nt > 0) __syncthreads();                          // the image is reused
        int slo0 = 0, shi0 = 0, slo1 = 0, shi1 = 0;           // sum of X as (X & 0xffff), (X >> 16): exact in s32
        const int items = live_rows * G * 16;                 // 8 values each
        for (int it = tid; it < items; it += NTHR) {
            const int hh = it >= G * 16 ? 1 : 0;
            const int r = it - hh * G * 16;                   // = k / 8: (granule, t, word)
            const int em = 2 * nt + hh;
            float v[8];
            load8f(p.x, p.x_dtype, (int64_t)em * p.K + r * 8, v);
            const float up = s_up[em];
            uint32_t D[8];
            int a_lo = 0, a_hi = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int X = __float2int_rn(v[i] * up);
                a_lo += X & 0xffff;
                a_hi += X >> 16;
                D[i] = (uint32_t)(X + 0x00808080) ^ 0x00808080u;       // byte l = signed base-256 digit l
            }
            if (hh) { slo1 += a_lo; shi1 += a_hi; } else { slo0 += a_lo; shi0 += a_hi; }
            // 4x4 byte transposes: digit l of the even values -> b0 (meets the low nibbles), odd -> b1
            const uint32_t e0 = __byte_perm(D[0], D[2], 0x5140), e1 = __byte_perm(D[4], D[6], 0x5140);
            const uint32_t e2 = __byte_perm(D[0], D[2], 0x7362), e3 = __byte_perm(D[4], D[6], 0x7362);
            const uint32_t o0 = __byte_perm(D[1], D[3], 0x5140), o1 = __byte_perm(D[5], D[7], 0x5140);
            const uint32_t o2 = __byte_perm(D[1], D[3], 0x7362), o3 = __byte_perm(D[5], D[7], 0x7362);
            uint2* dst = img + ((size_t)(hh * G + (r >> 4)) * 4) * 16 + (r & 15);
            dst[0 * 16] = make_uint2(__byte_perm(e0, e1, 0x5410), __byte_perm(o0, o1, 0x5410));
            dst[1 * 16] = make_uint2(__byte_perm(e0, e1, 0x7632), __byte_perm(o0, o1, 0x7632));
            dst[2 * 16] = make_uint2(__byte_perm(e2, e3, 0x5410), __byte_perm(o2, o3, 0x5410));
            dst[3 * 16] = make_uint2(__byte_perm(e2, e3, 0x7632), __byte_perm(o2, o3, 0x7632));
        }
        if (live_rows > 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                slo0 += __shfl_xor_sync(0xffffffffu, slo0, o);
                shi0 += __shfl_xor_sync(0xffffffffu, shi0, o);
                slo1 += __shfl_xor_sync(0xffffffffu, slo1, o);
                shi1 += __shfl_xor_sync(0xffffffffu, shi1, o);
            }
            if (lane == 0) {
                *reinterpret_cast<int4*>(s_tx + (warp * 8 + 2 * nt) * 2) = make_int4(slo0, shi0, slo1, shi1);
            }
        }
        __syncthreads();
        const bool live = h < live_rows && (p.M > 1 || g < 4);
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
            const int gq = warp + q * NW;
            if (live && gq < G) {
                const uint4* src = reinterpret_cast<const uint4*>(img + ((size_t)(h * G + gq) * 4 + l) * 16 + t * 4);
                v0 = src[0];
                v1 = src[1];
            }
            bf[q][nt][0][0] = v0.x; bf[q][nt][0][1] = v0.y; bf[q][nt][1][0] = v0.z; bf[q][nt][1][1] = v0.w;
            bf[q][nt][2][0] = v1.x; bf[q][nt][2][1] = v1.y; bf[q][nt][3][0] = v1.z; bf[q][nt][3][1] = v1.w;
        }
        if (tid < 2 && tid < live_rows) {                     // exact sum of X of one batch row over all warps
            int alo = 0, ahi = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { alo += s_tx[(w * 8 + 2 * nt + tid) * 2]; ahi += s_tx[(w * 8 + 2 * nt + tid) * 2 + 1]; }
            s_txs[(2 * nt + tid) * 2] = alo;
            s_txs[(2 * nt + tid) * 2 + 1] = ahi;
        }
    }
    stamp(6);

    // scale / zero point of the first output of this thread, fetched now (latency hidden by the main loop)
    const int outs = ntiles * TILE_ROWS * p.M;
    float pre_sc = 0.0f, pre_zp = 0.0f;
    if (tid < outs) {
        const int row = r0 + (tid >> 4) / p.M * TILE_ROWS + (tid & 15);
        if (row < r0 + nrows) { pre_sc = __ldg(p.scales + row); pre_zp = __ldg(p.zps + row); }
    }

    // ---- main loop: one 16-row tile per iteration, this warp's granules warp, warp + 16, ...
    const uint32_t lane_off = (uint32_t)(g * row_bytes + warp * GRAN_B + t * 16);
    const uint32_t acc_lane = sbase + p.acc_off + (uint32_t)((2 * t * TILE_ROWS + g) * 4);
    const bool cols_live = p.M > 1 || t < 2;                   // M == 1: mma columns 4..7 are empty
    for (int i = 0; i < ntiles; ++i) {
        int c[NT][2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int r = 0; r < 4; ++r) c[nt][ch][r] = 0;
        if (!skip_loads) mbar_wait(sbase + 8u * i, 0);
        const uint32_t tb = sbase + p.tile_off + i * tile_bytes + lane_off;
        uint4 lo[GPW], hi[GPW];
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            lo[q] = make_uint4(0u, 0u, 0u, 0u);
            hi[q] = lo[q];
            if (warp + q * NW < G) {
                lo[q] = lds128(tb + q * NW * GRAN_B);
                hi[q] = lds128(tb + q * NW * GRAN_B + 8 * row_bytes);
            }
        }
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const uint32_t wl[4] = {lo[q].x, lo[q].y, lo[q].z, lo[q].w};
            const uint32_t wh[4] = {hi[q].x, hi[q].y, hi[q].z, hi[q].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a0 = wl[j] & 0x0f0f0f0fu, a2 = (wl[j] >> 4) & 0x0f0f0f0fu;   // row g
                const uint32_t a1 = wh[j] & 0x0f0f0f0fu, a3 = (wh[j] >> 4) & 0x0f0f0f0fu;   // row g + 8
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) imma(c[nt][j & 1], a0, a1, a2, a3, bf[q][nt][j][0], bf[q][nt][j][1]);
            }
        }
        if (i < 5) stamp(10 + i);
        // the 16 warps add their partials of tile i into acc[i][column][row] (exact, order independent)
        if (cols_live) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint32_t a = acc_lane + (uint32_t)(((i * NT + nt) * 8) * TILE_ROWS * 4);
                red_add_s32(a, c[nt][0][0] + c[nt][1][0]);                            // (row g,     col 2t)
                red_add_s32(a + TILE_ROWS * 4, c[nt][0][1] + c[nt][1][1]);            // (row g,     col 2t + 1)
                red_add_s32(a + 32, c[nt][0][2] + c[nt][1][2]);                       // (row g + 8, col 2t)
                red_add_s32(a + TILE_ROWS * 4 + 32, c[nt][0][3] + c[nt][1][3]);       // (row g + 8, col 2t + 1)
            }
        }
    }
    stamp(7);
    __syncthreads();
    stamp(8);

    // ---- epilogue: one thread per output (tile, batch row, row): limbs -> sum_k q*X (exact s64),
    // y = s * 2^-e * (sum_k q*X - zp * sum_k X)
    for (int o = tid; o < outs; o += NTHR) {
        const int er = o & 15, rest = o >> 4;
        const int em = rest % p.M, ti = rest / p.M;
        const int row = r0 + ti * TILE_ROWS + er;
        if (row >= r0 + nrows) continue;
        float sc, zp;
        if (o == tid) { sc = pre_sc; zp = pre_zp; }
        else { sc = __ldg(p.scales + row); zp = __ldg(p.zps + row); }
        const int* a4 = acc + (ti * NT * 8 + em * 4) * TILE_ROWS + er;
        const long long a = (long long)a4[0] + ((long long)a4[TILE_ROWS] << 8) + ((long long)a4[2 * TILE_ROWS] << 16) +
                            ((long long)a4[3 * TILE_ROWS] << 24);
        const double down = __longlong_as_double((long long)(1023 - s_ex[em]) << 52);            // 2^-e
        const double tx = (double)s_txs[em * 2] + 65536.0 * (double)s_txs[em * 2 + 1];           // sum_k X
        const float v = sc * (float)(((double)a - (double)zp * tx) * down);
        store_out(p.y, p.y_dtype, (int64_t)em * p.N + row, v);
    }
    stamp(9);
    if (prof) {
        unsigned long long wall1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wall1));
        const unsigned int ln = p.launch_no & 63u;
        atomicMin(&g_res_wall[ln * 4 + 0], wall0);
        atomicMax(&g_res_wall[ln * 4 + 1], wall0);
        atomicMin(&g_res_wall[ln * 4 + 2], wall1);
        atomicMax(&g_res_wall[ln * 4 + 3], wall1);
    }
}

struct ResPlan {
    int grid, gpw, nt, ntiles;
    int acc_off, img_off, tile_off;
    size_t smem;
};

bool plan_res(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, ResPlan* c) {
    if (M < 1 || M > 8 || K <= 0 || K % GRAN_K != 0 || N < 1 || N > 0x7fffffff) return false;
    const int G = (int)(K / GRAN_K);
    if (G > 3 * NW) return false;                                   // GPW <= 3
    int grid = sm_count;
    const int min_grid = (int)((N + TILE_ROWS * MAX_TILES - 1) / (TILE_ROWS * MAX_TILES));
    if (grid > (N + TILE_ROWS - 1) / TILE_ROWS) grid = (int)((N + TILE_ROWS - 1) / TILE_ROWS);   // >= one tile per CTA
    if (grid < min_grid) return false;
    const int rows = (int)((N + grid - 1) / grid);
    const int ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (ntiles > MAX_TILES) return false;
    const int nt = M <= 2 ? 1 : (M <= 4 ? 2 : 4);
    c->grid = grid; c->gpw = (G + NW - 1) / NW; c->nt = nt; c->ntiles = ntiles;
    if (c->gpw * nt > 8) return false;                              // B fragments must stay in registers
    c->acc_off = OFF_END;
    c->img_off = (c->acc_off + ntiles * nt * 512 + 127) / 128 * 128;
    c->tile_off = (c->img_off + 2 * G * 512 + 127) / 128 * 128;    // image: 2 rows x G granules x 4 limbs x 16 x 8 B
    c->smem = (size_t)c->tile_off + (size_t)ntiles * TILE_ROWS * (K / 2);
    return c->smem <= (size_t)max_smem;
}

template <int GPW, int NT>
int launch_res_inst(const ResPlan& c, const ResParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_res_kernel<GPW, NT>;
    static thread_local int attr_dev_smem[64] = {0};
    int dev = 0;
    B200Q_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_dev_smem[dev] < (int)c.smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        attr_dev_smem[dev] = (int)c.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)c.grid);
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, p), "gemv_res launch");
}

}  // namespace

bool gemv_res_supported(int64_t M, int64_t N, int64_t K) {
    ResPlan c;
    return plan_res(148, 232448, M, N, K, &c);
}

int launch_gemv_res(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                    const float* zps, void* y, int y_dtype, int64_t M, int64_t N, int64_t K, unsigned flags,
                    cudaStream_t st, const uint8_t* next_packed, size_t next_bytes) {
    ResPlan c;
    if (!plan_res(dev.sm_count, dev.max_smem_optin, M, N, K, &c))
        return set_error(B200Q_EINVAL, "gemv_res: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv_res: x and packed must be 16-byte aligned");
    ResParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.y = y;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.rows_q = (int)(N / c.grid); p.rows_rem = (int)(N % c.grid);
    p.G = (int)(K / GRAN_K);
    p.acc_off = c.acc_off; p.img_off = c.img_off; p.tile_off = c.tile_off;
    p.wait_weights = (flags & B200Q_FLAG_STATIC_WEIGHTS) ? 0 : 1;
    p.early_tiles = tuning().gemv_early >= 0 ? tuning().gemv_early : MAX_TILES;
    p.pf_mode = tuning().gemv_pf;
    p.next_packed = next_packed;
    p.next_bytes = next_packed && (reinterpret_cast<uintptr_t>(next_packed) & 15) == 0 ? next_bytes : 0;
    p.next_chunk = (unsigned int)(p.next_bytes / (unsigned long long)c.grid);
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    {
        static thread_local unsigned launch_no = 0;
        p.launch_no = launch_no++;
    }
    const bool pdl = tuning().gemv_pdl != 0;
#define B200Q_RES_CASE(GPW_, NT_) \
    if (c.gpw == GPW_ && c.nt == NT_) return launch_res_inst<GPW_, NT_>(c, p, pdl, st);
    B200Q_RES_CASE(1, 1) B200Q_RES_CASE(2, 1) B200Q_RES_CASE(3, 1)
    B200Q_RES_CASE(1, 2) B200Q_RES_CASE(2, 2) B200Q_RES_CASE(3, 2)
    B200Q_RES_CASE(1, 4) B200Q_RES_CASE(2, 4)
#undef B200Q_RES_CASE
    return set_error(B200Q_EINVAL, "gemv_res: no kernel instance for gpw=%d nt=%d", c.gpw, c.nt);
}

}  // namespace b200q

extern "C" int b200q_debug_wall_res(unsigned long long* h_out, int reset) {
    if (reset) {
        unsigned long long init[64 * 4];
        for (int i = 0; i < 64; ++i) { init[4 * i] = ~0ull; init[4 * i + 1] = 0; init[4 * i + 2] = ~0ull; init[4 * i + 3] = 0; }
        return b200q::check_cuda(cudaMemcpyToSymbol(b200q::g_res_wall, init, sizeof(init)), "reset wall");
    }
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_res_wall, sizeof(unsigned long long) * 64 * 4), "read wall");
}

extern "C" int b200q_debug_read_prof_res(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_res_prof, sizeof(long long) * 256 * 16), "read prof");
}
