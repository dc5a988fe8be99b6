// Decode path: y[M,N] = x[M,K] @ dequant(W)^T for M <= 16.  HBM-bound: every packed byte is read
// exactly once, by the TMA engine, into shared memory; everything else is on-chip.
//
// Decomposition (DESIGN.md "GEMV"):
//   * K is cut into `nslab` slabs of whole 128-column granules, N into `nrb` row blocks;
//     CTA (rb, slab) streams rows [r0,r1) x slab bytes.  nslab * nrb ~= SM count, one CTA per SM.
//   * a stage of the shared-memory ring = 16 weight rows x one chunk of NW granules, filled by 16
//     cp.async.bulk row copies (UBLKCP) that complete on the stage's mbarrier.  The ring is deep
//     enough that for the Llama shapes the CTA's whole slab is requested in the first instructions
//     of the kernel (every warp issues its share; there is no dedicated producer warp).
//   * warp w owns granules w, w+NW, ... of the slab for the whole kernel, so its x operand (mma B
//     fragments) is built ONCE in registers; the main loop is LDS.128 -> LOP3 -> HMMA only.
//     Weights enter the mma as exact fp16 SUBNORMALS (nibble * 2^-24, no arithmetic); x enters as
//     an fp16 hi/lo split of x * 2^e (per-warp power-of-two e), so the products are exact and the
//     accumulation is fp32: results match the fp32 reference to ~1e-6 relative.
//   * per-warp tile partials go to shared memory without a barrier; once per round of `rt` tiles:
//     one named barrier, fixed-order cross-warp sum, epilogue y = s * (acc - zp * sum(x)); with
//     nslab > 1 the slab partials go to a workspace and the last CTA of a row block (ticket counter)
//     adds them in slab order: deterministic.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output).
#include <cmath>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

// phase timestamps (SM clock) of every CTA, written when the bench-only "gemv_debug" value has
// bit 3 (8) set; read back with b200q_debug_read_prof
__device__ long long g_gemv_prof[256 * 16];

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
    int stages;             // ring depth (stage = 16 rows x one chunk of NW granules)
    int pitch;              // bytes between rows of a stage
    int rt;                 // tiles per cross-warp reduction round
    int rg, rg_shift;       // live mma columns per n-tile (2, 4 or 8) and log2 of it
    int red_off;            // byte offset of the reduction buffer
    int ring_off;           // byte offset of the ring in dynamic shared memory
    int wait_weights;       // 1: weights may be written by the preceding kernel -> wait first
    int debug;              // bench-only: 1 = skip the mma work, 2 = skip the weight loads
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

// Nibble -> fp16 with NO arithmetic: a nibble left in the low mantissa bits of an fp16 with a zero
// exponent field is the subnormal n * 2^-24 (bits 0-3) or n * 2^-20 (bits 4-7), exactly.  The tensor
// core consumes fp16 subnormals at full precision, so one LOP3 (AND) per two weights is the whole
// dequantisation; the 2^-24 / 2^-20 factors are folded into x (high-nibble columns are pre-scaled
// by 2^-4) and into the epilogue's power-of-two descale.
//   r[0] = (n0,n4) * 2^-24   r[1] = (n1,n5) * 2^-20   r[2] = (n2,n6) * 2^-24   r[3] = (n3,n7) * 2^-20
__device__ __forceinline__ void nibbles_to_subnormal_half2x4(uint32_t w, uint32_t (&r)[4]) {
    constexpr uint32_t LO = 0x000f000fu, HI = 0x00f000f0u;
    const uint32_t w2 = w >> 8;
    r[0] = w & LO;
    r[1] = w & HI;
    r[2] = w2 & LO;
    r[3] = w2 & HI;
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Shared memory carve-up (dynamic):
//   [0, 512) full mbarriers, [512, 1024) empty mbarriers, [1024, 1088) flags
//   [1088, ..)   rowc[48] f32, pro[NW][16][2] f32 (prologue scratch)
//   [XF_OFF, ..) xf: mma-B vectors, ng * NT * rg * 256 bytes
//   [red_off,..) red[rt][NW][NT*4][16] f32
//   [ring_off, ...)  stages x 16 x pitch bytes
constexpr int BAR_BYTES = 1024;
constexpr int MISC_OFF = 1024;
constexpr int MAX_STAGES = 64;

template <int NW, int NT>
struct SmemLayout {
    static constexpr int COLS = NT * 4;
    static constexpr int ROWC_OFF = 1088;
    static constexpr int PRO_OFF = ROWC_OFF + 48 * 4;
    static constexpr int XF_OFF = ((PRO_OFF + NW * 16 * 2 * 4 + 127) / 128) * 128;
    static constexpr int RED_TILE_BYTES = NW * COLS * TILE_ROWS * 4;
};

template <int NW, int GPW, int NT>
__global__ void __launch_bounds__(NW * 32, 1) gemv_kernel(const GemvParams p) {
    using L = SmemLayout<NW, NT>;
    constexpr int COLS = NT * 4;
    constexpr int CH = NT == 1 ? 4 : 2;            // independent mma accumulator chains per n-tile
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    float* red = reinterpret_cast<float*>(smem + p.red_off);
    volatile int* flag = reinterpret_cast<volatile int*>(smem + MISC_OFF);
    const uint32_t ring = smem_base + p.ring_off;
    const int S = p.stages;
    const int dbg = p.debug & 7;
    auto full_bar = [&](int s) { return smem_base + 8u * s; };
    auto empty_bar = [&](int s) { return smem_base + 512u + 8u * s; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool prof = (p.debug & 8) && threadIdx.x == 0 && blockIdx.x < 256;
    auto stamp = [&](int i) { if (prof) g_gemv_prof[blockIdx.x * 16 + i] = clock64(); };
    stamp(0);
    const int slab = blockIdx.x % p.nslab, rb = blockIdx.x / p.nslab;
    const int g0 = (int)((int64_t)p.G * slab / p.nslab), g1 = (int)((int64_t)p.G * (slab + 1) / p.nslab);
    const int ng = g1 - g0;                       // granules in this slab
    const int r0 = (int)((int64_t)p.N * rb / p.nrb), r1 = (int)((int64_t)p.N * (rb + 1) / p.nrb);
    const int ntiles = (r1 - r0 + TILE_ROWS - 1) / TILE_ROWS;
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

    // ---- weight streaming.  Stage st = (tile st / nq, chunk st % nq): 16 rows x NW granules, one
    // cp.async.bulk per row.  Stage st is issued by warp st % NW: the first min(S, total) stages
    // right here (for the Llama shapes that is the CTA's whole slab), later ones as ring slots
    // free up (see the refill hook in the main loop).
    const int nq = (ng + NW - 1) / NW;             // chunks per tile in this slab
    const int total_stages = ntiles * nq;
    const uint64_t pol = policy_evict_first();
    const uint8_t* src0 = p.packed + (int64_t)g0 * GRAN_B;
    auto issue_stage = [&](int st, int slot) {
        const int ti = st / nq, q = st - ti * nq;
        const int row = r0 + ti * TILE_ROWS;
        const int rows = min(TILE_ROWS, r1 - row);
        const int cb = min(NW, ng - q * NW) * GRAN_B;
        if (lane == 0) mbar_arrive_expect_tx(full_bar(slot), (uint32_t)(rows * cb));
        __syncwarp();
        if (lane < rows)
            bulk_g2s_hint(ring + slot * stage_bytes + lane * p.pitch,
                          src0 + (int64_t)(row + lane) * (p.K / 2) + q * NW * GRAN_B, (uint32_t)cb,
                          full_bar(slot), pol);
    };
    stamp(1);
    if (p.wait_weights) pdl_wait();
    if (dbg != 2)
        for (int st = warp; st < min(S, total_stages); st += NW) issue_stage(st, st);
    stamp(2);

    // ---------------------------------------------------------------- consumer warps
    const int g = lane >> 2, t = lane & 3;
    const int ctid = threadIdx.x;                     // 0 .. NW*32-1
    pdl_wait();   // x (and the output / workspace) belong to the stream-ordered predecessor
    stamp(3);

    // ---- x operand, built cooperatively ONCE per CTA (every CTA needs all of x[:, slab]; doing it
    // per warp made this prologue the longest phase of the kernel, profiles/r01_gemv_notes.md).
    //  pass 1: per batch row amax and sum over the slab           -> rowc (2^e, 2^(24-e), sum x)
    //  pass 2: one 16-byte mma-B vector per item                  -> xf in shared memory
    //  then every lane loads its own B fragments with LDS.128.
    // Lane (g,t) feeds mma column g: batch row nt*4 + g/2, part g&1 (0: fp16(x*2^e), 1: the fp16
    // remainder).  Vector j of a granule covers columns k0 + 32t + 8j + {0..7}, packed in the nibble
    // order of nibbles_to_subnormal_half2x4; the columns that meet high nibbles carry 2^-4.
    float* pro = reinterpret_cast<float*>(smem + L::PRO_OFF);     // [NW][16][2] scratch
    float* rowc = reinterpret_cast<float*>(smem + L::ROWC_OFF);   // descale[16], sumx[16], up[16]
    const uint32_t xf = smem_base + L::XF_OFF;
    const int kslab0 = g0 * GRAN_K, kslab = ng * GRAN_K;
    const int rg = p.rg, rgs = p.rg_shift;                        // live mma columns per n-tile (2, 4 or 8)
#pragma unroll 1
    for (int m = 0; m < p.M; ++m) {
        float am = 0.0f, sm = 0.0f;
        for (int k = ctid * 4; k < kslab; k += NW * 32 * 4) {
            float v[4];
            load4f(p.x, p.x_dtype, (int64_t)m * p.K + kslab0 + k, v);
            am = fmaxf(am, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
            sm += (v[0] + v[1]) + (v[2] + v[3]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
        }
        if (lane == 0) {
            pro[(warp * 16 + m) * 2] = am;
            pro[(warp * 16 + m) * 2 + 1] = sm;
        }
    }
    named_bar_sync(1, NW * 32);
    stamp(4);
    if (ctid < p.M) {       // per-row scale 2^e with amax * 2^e in [2^13, 2^14); fixed summation order
        float am = 0.0f, sm = 0.0f;
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
            am = fmaxf(am, pro[(w * 16 + ctid) * 2]);
            sm += pro[(w * 16 + ctid) * 2 + 1];
        }
        int ex = 0;
        if (am > 0.0f && am < INFINITY) {
            ex = 140 - (int)((__float_as_uint(am) >> 23) & 0xffu);
            ex = max(-100, min(100, ex));
        }
        rowc[ctid] = __uint_as_float((uint32_t)(127 + 24 - ex) << 23);     // descale 2^(24-e)
        rowc[16 + ctid] = sm;
        rowc[32 + ctid] = __uint_as_float((uint32_t)(127 + ex) << 23);     // 2^e
    }
    named_bar_sync(1, NW * 32);
    {
        const int items = ng * NT * 16 * rg;        // (granule, n-tile, j, g, t)
#pragma unroll 1
        for (int it = ctid; it < items; it += NW * 32) {
            const int t_ = it & 3, g_ = (it >> 2) & (rg - 1);
            int rest = it >> (2 + rgs);
            const int j = rest & 3;
            rest >>= 2;
            const int nt = rest % NT, gq = rest / NT;
            const int m = nt * 4 + (g_ >> 1), part = g_ & 1;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (m < p.M) {
                float a[4], b[4];
                const int64_t base = (int64_t)m * p.K + kslab0 + gq * GRAN_K + t_ * 32 + j * 8;
                load4f(p.x, p.x_dtype, base, a);
                load4f(p.x, p.x_dtype, base + 4, b);
                const float u = rowc[32 + m], u_hi = u * 0.0625f;
                // (n0,n4) (n1,n5) (n2,n6) (n3,n7): odd columns meet high nibbles (2^-20 instead of 2^-24)
                const float s0 = a[0] * u, s4 = b[0] * u, s1 = a[1] * u_hi, s5 = b[1] * u_hi;
                const float s2 = a[2] * u, s6 = b[2] * u, s3 = a[3] * u_hi, s7 = b[3] * u_hi;
                __half2 h0 = __floats2half2_rn(s0, s4), h1 = __floats2half2_rn(s1, s5);
                __half2 h2 = __floats2half2_rn(s2, s6), h3 = __floats2half2_rn(s3, s7);
                if (part) {
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const float2 f2 = __half22float2(h2), f3 = __half22float2(h3);
                    h0 = __floats2half2_rn(s0 - f0.x, s4 - f0.y);
                    h1 = __floats2half2_rn(s1 - f1.x, s5 - f1.y);
                    h2 = __floats2half2_rn(s2 - f2.x, s6 - f2.y);
                    h3 = __floats2half2_rn(s3 - f3.x, s7 - f3.y);
                }
                out = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                 *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
            sts128(xf + it * 16u, out);
        }
    }
    named_bar_sync(1, NW * 32);
    stamp(5);
    uint32_t bf[GPW][NT][4][4];
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        const int gq = warp + q * NW;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (gq < ng && g < rg) v = lds128(xf + (((((gq * NT + nt) * 4 + j) << rgs) + g) * 4 + t) * 16u);
                bf[q][nt][j][0] = v.x; bf[q][nt][j][1] = v.y; bf[q][nt][j][2] = v.z; bf[q][nt][j][3] = v.w;
            }
    }

    stamp(6);
    // ---- main loop: tiles of 16 rows; per tile GPW stages (one granule of this warp in each)
    int s = 0, ph = 0, st = 0;
    int round_first = 0;           // first tile of the current reduction round
    for (int i = 0; i < ntiles; ++i) {
        float acc[NT][CH][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[nt][c][r] = 0.0f;
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            if (q * NW < ng) {                                   // stage exists (uniform over the CTA)
                if (dbg != 2) mbar_wait(full_bar(s), ph);
                if (warp + q * NW < ng && dbg != 1) {
                    const uint32_t sbase = ring + s * stage_bytes + g * p.pitch + warp * GRAN_B + t * 16;
                    const uint4 lo = lds128(sbase);
                    const uint4 hi = lds128(sbase + 8 * p.pitch);
                    const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w};
                    const uint32_t wh[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t al[4], ah[4];
                        nibbles_to_subnormal_half2x4(wl[j], al);
                        nibbles_to_subnormal_half2x4(wh[j], ah);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            mma_m16n8k16_f16(acc[nt][(2 * j) % CH], al[0], ah[0], al[1], ah[1], bf[q][nt][j][0], bf[q][nt][j][1]);
                            mma_m16n8k16_f16(acc[nt][(2 * j + 1) % CH], al[2], ah[2], al[3], ah[3], bf[q][nt][j][2], bf[q][nt][j][3]);
                        }
                    }
                }
                __syncwarp();
                if (st + S < total_stages) {                     // ring smaller than the slab: recycle
                    if (lane == 0) mbar_arrive(empty_bar(s));
                    if (st % NW == warp && dbg != 2) {       // this warp refills the slot
                        mbar_wait(empty_bar(s), ph);
                        issue_stage(st + S, s);
                    }
                }
                ++st;
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        // this warp's partial of tile i -> red[i - round_first][warp][col][row]   (no barrier here)
        float* rbuf = red + ((i - round_first) * NW + warp) * (COLS * TILE_ROWS);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 4 + t;   // C columns 2t (hi part) and 2t+1 (lo part) of batch row nt*4+t
            float top = 0.0f, bot = 0.0f;   // rows g and g+8: hi-part column + lo-part column, all chains
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                top += acc[nt][c][0] + acc[nt][c][1];
                bot += acc[nt][c][2] + acc[nt][c][3];
            }
            rbuf[col * TILE_ROWS + g] = top;
            rbuf[col * TILE_ROWS + g + 8] = bot;
        }
        if (i + 1 - round_first == p.rt || i + 1 == ntiles) {
            // ---- end of a round: fixed-order cross-warp sum, zero-point term, scale, store
            if (i + 1 == ntiles) stamp(7);
            named_bar_sync(1, NW * 32);
            if (i + 1 == ntiles) stamp(8);
            const int nt_round = i + 1 - round_first;
            const int total = nt_round * TILE_ROWS * p.M;
            for (int idx = ctid; idx < total; idx += NW * 32) {
                const int er = idx & (TILE_ROWS - 1);
                const int rest = idx >> 4;
                const int em = rest % p.M, j = rest / p.M;
                const int row = r0 + (round_first + j) * TILE_ROWS + er;
                if (row < r1) {
                    const float* rr = red + (j * NW) * (COLS * TILE_ROWS) + em * TILE_ROWS + er;
                    float a = 0.0f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) a += rr[w * (COLS * TILE_ROWS)];
                    const float sc = __ldg(p.scales + row), zp = __ldg(p.zps + row);
                    const float v = sc * (a * rowc[em] - zp * rowc[16 + em]);
                    if (p.nslab == 1) store_y(p.y, p.y_dtype, (int64_t)em * p.N + row, v);
                    else p.part[((int64_t)slab * p.M + em) * p.N + row] = v;
                }
            }
            round_first = i + 1;
            if (i + 1 < ntiles) named_bar_sync(1, NW * 32);    // red is reused by the next round
        }
    }

    stamp(9);
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
    int nw, gpw, nt, nslab, nrb, stages, pitch, rt, ring_off, red_off, rg;
    size_t smem;
};

// kernel instances that exist (NW, GPW, NT); see launch_gemv
bool has_instance(int nw, int gpw, int nt) {
    if (nw == 16) return (nt == 1 && gpw <= 3) || (nt == 2 && gpw <= 2) || (nt == 4 && gpw == 1);
    if (nw == 8) return (nt == 1 && gpw <= 4) || (nt == 2 && gpw <= 2) || (nt == 4 && gpw <= 2);
    return false;
}

int xf_off_for(int nw, int nt) {
    if (nw == 16) return nt == 1 ? SmemLayout<16, 1>::XF_OFF : nt == 2 ? SmemLayout<16, 2>::XF_OFF : SmemLayout<16, 4>::XF_OFF;
    return nt == 1 ? SmemLayout<8, 1>::XF_OFF : nt == 2 ? SmemLayout<8, 2>::XF_OFF : SmemLayout<8, 4>::XF_OFF;
}

bool plan(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, GemvConfig* c) {
    if (M < 1 || M > 16 || K % GRAN_K != 0 || K <= 0 || N < 1 || N > 0x7fffffff || K > 0x7fffffff) return false;
    const Tuning& tu = tuning();
    const int G = (int)(K / GRAN_K);
    const int nt = M <= 4 ? 1 : (M <= 8 ? 2 : 4);
    int ctas = dev.sm_count;
    if (tu.gemv_ctas > 0 && tu.gemv_ctas < ctas) ctas = tu.gemv_ctas;
    double best_cost = 1e30;
    bool found = false;
    for (int nw = 16; nw >= 8; nw -= 8) {
        if (tu.gemv_warps > 0 && tu.gemv_warps != nw) continue;
        for (int ns = 1; ns <= MAX_SLABS && ns <= G; ++ns) {
            if (tu.gemv_slabs > 0 && tu.gemv_slabs != ns) continue;
            const int ng = (G + ns - 1) / ns;
            const int gpw = (ng + nw - 1) / nw;
            if (!has_instance(nw, gpw, nt)) continue;
            int nrb = ctas / ns;
            if (nrb < 1) continue;
            if (nrb > N) nrb = (int)N;
            if (nrb > MAX_RB) nrb = MAX_RB;
            const int rows = (int)((N + nrb - 1) / nrb);
            const int ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
            const int pitch = ((ng < nw ? ng : nw) | 1) * GRAN_B;
            const int stage_bytes = TILE_ROWS * pitch;
            const int red_tile = nw * nt * 4 * TILE_ROWS * 4;
            int rt = 32768 / red_tile;
            if (rt < 1) rt = 1;
            if (rt > ntiles) rt = ntiles;
            const int rg = nt > 1 ? 8 : (M == 1 ? 2 : (M == 2 ? 4 : 8));
            const int red_off = ((xf_off_for(nw, nt) + ng * nt * rg * 256 + 127) / 128) * 128;
            const int ring_off = ((red_off + rt * red_tile + 127) / 128) * 128;
            int stages = (dev.max_smem_optin - ring_off) / stage_bytes;
            if (stages > ntiles * gpw) stages = ntiles * gpw;
            if (stages > MAX_STAGES) stages = MAX_STAGES;
            if (tu.gemv_stages > 0 && tu.gemv_stages < stages) stages = tu.gemv_stages;
            if (stages < 1) continue;
            // cost ~ bytes of the busiest CTA; small charges for the cross-slab reduction, for idle
            // warps in the last chunk, for a ring that cannot hold the whole slab, and for 8 warps
            double cost = (double)rows * ng * (1.0 + 0.03 * (ns - 1)) * (1.0 + 0.1 * ((double)gpw * nw / ng - 1.0));
            if (stages < ntiles * gpw) cost *= 1.05;
            if (nw == 8) cost *= 1.10;
            if (cost < best_cost) {
                best_cost = cost;
                found = true;
                c->nw = nw; c->gpw = gpw; c->nt = nt; c->nslab = ns; c->nrb = nrb; c->stages = stages;
                c->pitch = pitch; c->rt = rt; c->ring_off = ring_off; c->red_off = red_off; c->rg = rg;
                c->smem = (size_t)ring_off + (size_t)stages * stage_bytes;
            }
        }
    }
    return found;
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
    cfg.blockDim = dim3(NW * 32);
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
    p.stages = c.stages; p.pitch = c.pitch; p.rt = c.rt; p.ring_off = c.ring_off; p.red_off = c.red_off;
    p.rg = c.rg; p.rg_shift = c.rg == 2 ? 1 : (c.rg == 4 ? 2 : 3);
    const bool is_static = (flags & B200Q_FLAG_STATIC_WEIGHTS) != 0;
    const bool pdl = tuning().gemv_pdl != 0;
    p.wait_weights = is_static ? 0 : 1;
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    if (c.nslab > 1) {
        const size_t need = (size_t)MAX_RB * 4 + (size_t)c.nslab * M * N * 4;
        if (!ws || ws_bytes < need) return set_error(B200Q_EWORKSPACE, "gemv: workspace too small (%zu < %zu)", ws_bytes, need);
        if (reinterpret_cast<uintptr_t>(ws) & 15) return set_error(B200Q_EALIGN, "gemv: workspace must be 16-byte aligned");
        p.tickets = static_cast<unsigned int*>(ws);
        p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + (size_t)MAX_RB * 4);
    }
#define B200Q_GEMV_CASE(NW_, GPW_, NT_) \
    if (c.nw == NW_ && c.gpw == GPW_ && c.nt == NT_) return launch_inst<NW_, GPW_, NT_>(c, p, pdl, st);
    B200Q_GEMV_CASE(16, 1, 1) B200Q_GEMV_CASE(16, 2, 1) B200Q_GEMV_CASE(16, 3, 1)
    B200Q_GEMV_CASE(16, 1, 2) B200Q_GEMV_CASE(16, 2, 2) B200Q_GEMV_CASE(16, 1, 4)
    B200Q_GEMV_CASE(8, 1, 1) B200Q_GEMV_CASE(8, 2, 1) B200Q_GEMV_CASE(8, 3, 1) B200Q_GEMV_CASE(8, 4, 1)
    B200Q_GEMV_CASE(8, 1, 2) B200Q_GEMV_CASE(8, 2, 2)
    B200Q_GEMV_CASE(8, 1, 4) B200Q_GEMV_CASE(8, 2, 4)
#undef B200Q_GEMV_CASE
    return set_error(B200Q_EINVAL, "gemv: no kernel instance for nw=%d gpw=%d nt=%d", c.nw, c.gpw, c.nt);
}

}  // namespace b200q

/* bench-only: copy the per-CTA phase timestamps of the last profiled GEMV launch (256 x 16 int64) */
extern "C" int b200q_debug_read_prof(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_gemv_prof, sizeof(long long) * 256 * 16), "read prof");
}
