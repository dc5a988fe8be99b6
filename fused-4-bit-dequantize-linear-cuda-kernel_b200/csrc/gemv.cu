// Decode path, ring kernel: y[M,N] = x[M,K] @ dequant(W)^T for M <= 8 on the shapes the resident-slab kernel
// (gemv_res.cu) does not take -- a CTA's share of W does not fit in shared memory even with K split over a cluster
// (e.g. 14336 -> 4096), or M > 2 with K > 6144.  HBM-bound: every packed byte is read exactly once, by the TMA
// engine, into a shared-memory ring; everything else is on-chip.
//
// Decomposition:
//   * K is cut into `nslab` slabs of whole 128-column granules, N into `nrb` row blocks;
//     CTA (rb, slab) streams rows [r0,r1) x slab bytes.  nslab * nrb ~= SM count, one CTA per SM.
//   * a stage of the ring = 16 weight rows x the slab's bytes of a row (one contiguous bulk copy per tile when
//     nslab == 1, else one copy per row with an odd 64-byte pitch); slots are recycled through empty barriers.
//   * arithmetic is EXACT INTEGER, as in gemv_res.cu: x is a per-row fixed-point number (even columns
//     Xe = round(x 2^e), odd columns Xo = round(x 2^(e-4))) cut into four signed base-256 limbs, one IMMA m16n8k32
//     (u8 x s8 -> s32) column per (batch row, limb); the raw packed byte meets Xe, the masked byte 16 q_hi meets
//     Xo - Xe, so the nibbles are never widened.  sum q*X and sum X are exact: bit-reproducible whatever the order.
//   * warp w owns granules w, w+NW, ... of the slab for the whole kernel, so its x operand (mma B fragments) sits in
//     registers; it is built inside the CTA (block amax, then a warp-local conversion through a 512-byte staging
//     slot) -- or, with the bench key gemv_xprep = 1, once per launch by xprep_kernel (an extra kernel in the
//     dependency chain: slower, kept as a measured alternative).
//   * per-warp tile partials (s32) go to shared memory without a barrier; once per round of `rt` tiles: one named
//     barrier, cross-warp sum, epilogue y = s * 2^-e * (sum q*X - zp * sum X); with nslab > 1 the slab partials go
//     to a workspace and the last CTA of a row block (ticket counter) adds them in slab order: deterministic.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output).
#include <cmath>
#include "internal.h"
#include "ptx.cuh"

namespace b200q {

// bench-only (-DB200Q_PROF build, tools/prof_gemv.py): ablation switches (tuning key gemv_debug: 1 = skip the mma work,
// 2 = skip the weight loads), phase timestamps of every CTA (bit 8; 16 = globaltimer instead of the SM clock) and the
// wall-clock of the first CTA start / last CTA end of the last 64 profiled launches.  The product build compiles all of
// it out.
#ifdef B200Q_PROF
__device__ long long g_gemv_prof[256 * 16];
__device__ unsigned long long g_gemv_wall[64 * 4];
#define B200Q_GEMV_DBG(p) ((p).debug)
#else
#define B200Q_GEMV_DBG(p) 0
#endif

namespace {

constexpr int TILE_ROWS = 16;
constexpr int GRAN_K = 128;          // columns per granule
constexpr int GRAN_B = GRAN_K / 2;   // packed bytes per granule per row
constexpr int MAX_SLABS = 16;
constexpr int MAX_RB = 1024;         // ticket counters at the head of the workspace
constexpr int MAX_STAGES = 64;
constexpr int LIMBS = 4;             // signed base-256 digits of round(x * 2^e)
constexpr int XHDR_INTS = 4 + 2 * MAX_SLABS;   // per batch row: {e, 0, 0, 0}, then {sum lo16, sum hi16} per K slab

struct GemvParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    void* y;
    const uint8_t* ximg;    // B-fragment image of x (xprep_kernel)
    const int* xhdr;        // per batch row: XHDR_INTS ints (xprep_kernel)
    float* part;            // [nslab][M][N] fp32 slab partials (nslab > 1)
    unsigned int* tickets;  // [nrb]
    int x_dtype, y_dtype;
    int M, N, K;
    int nslab, nrb;
    int rows_q, rows_rem;   // N = nrb * rows_q + rows_rem: row block rb has rows_q (+1 if rb < rows_rem) rows
    int gran_q, gran_rem;   // K / 128 granules split the same way over the slabs
    int stages;             // ring depth (stage = 16 rows x one chunk of granules)
    int whole_row;          // 1: a stage holds the whole slab row (all GPW granules of every warp), else NW granules
    int contig;             // 1 (nslab == 1, whole_row): a tile is ONE contiguous bulk copy, pitch = K/2
    int pitch;              // bytes between rows of a stage
    int rt;                 // tiles per cross-warp reduction round
    int rg, rg_shift;       // live mma columns per n-tile (4 or 8) and log2 of it
    int red_off, ring_off;  // byte offsets in dynamic shared memory
    int after_x;            // 1: the held-back tiles go out after the amax barrier instead of right after the wait
    int early_tiles;        // contiguous mode: tiles requested before griddepcontrol.wait (the rest after the x loads)
    int fused_x;            // 1: every CTA builds its own x operand (no xprep_kernel in the dependency chain)
    int wait_weights;       // 1: weights may be written by the preceding kernel -> wait first
    const uint8_t* next_packed;        // optional: weights of the NEXT fused linear on this stream ...
    unsigned long long next_bytes;     // ... this launch pulls them into L2 behind its own stream (0 = none)
    unsigned int next_chunk;           // next_bytes / grid size
    int pf_mode;            // how next_packed is prefetched (tuning key gemv_pf)
    unsigned int launch_no; // bench-only: slot of the wall-clock record
    int debug;              // bench-only: low bits 1 = skip the mma work, 2 = skip the weight loads; 8 = timestamps
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

// D(16x8,s32) += A(16x32,u8,row) * B(32x8,s8,col)      SASS: IMMA.16832.U8.S8
__device__ __forceinline__ void mma_m16n8k32_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Shared memory carve-up (dynamic):
//   [0, 512) full mbarriers, [512, 1024) empty mbarriers, [1024, 1088) flags
//   [red_off,..)  red[rt][NW][NT*8 columns][16] s32
//   [ring_off, ...)  stages x 16 x pitch bytes
constexpr int MISC_OFF = 1024;
constexpr int RED_OFF = 1152;

template <int NW, int GPW, int NT>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) gemv_kernel(const GemvParams p) {
    constexpr int COLS = NT * 8;                   // mma columns: (batch row, limb), two batch rows per n-tile
    constexpr int NTHR = NW * 32;
    constexpr int RPW = TILE_ROWS / NW;            // stage rows issued per warp (1 or 2)
    constexpr int CH = NT == 1 ? 4 : 2;            // independent IMMA accumulator chains per n-tile
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    int* red = reinterpret_cast<int*>(smem + p.red_off);
    volatile int* flag = reinterpret_cast<volatile int*>(smem + MISC_OFF);
    const uint32_t ring = smem_base + p.ring_off;
    const int S = p.stages;
    const int dbg = B200Q_GEMV_DBG(p) & 3;
    auto full_bar = [&](int s) { return smem_base + 8u * s; };
    auto empty_bar = [&](int s) { return smem_base + 512u + 8u * s; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctid = threadIdx.x;
#ifdef B200Q_PROF
    const bool prof = (p.debug & 8) && threadIdx.x == 0 && blockIdx.x < 256;
    auto stamp = [&](int i) {
        if (prof) {
            long long tnow;
            if (p.debug & 16) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));   // ns, comparable across SMs
            else tnow = clock64();
            g_gemv_prof[blockIdx.x * 16 + i] = tnow;
        }
    };
    unsigned long long wall0 = 0;
    if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wall0));
#else
    auto stamp = [](int) {};
#endif
    stamp(0);
    const int slab = p.nslab == 1 ? 0 : (int)(blockIdx.x % p.nslab);
    const int rb = p.nslab == 1 ? (int)blockIdx.x : (int)(blockIdx.x / p.nslab);
    const int g0 = slab * p.gran_q + min(slab, p.gran_rem);
    const int ng = p.gran_q + (slab < p.gran_rem ? 1 : 0);        // granules in this slab
    const int r0 = rb * p.rows_q + min(rb, p.rows_rem);
    const int r1 = r0 + p.rows_q + (rb < p.rows_rem ? 1 : 0);
    const int ntiles = (r1 - r0 + TILE_ROWS - 1) / TILE_ROWS;
    const int nq = p.whole_row ? 1 : (ng + NW - 1) / NW;          // chunks (= stages) per tile
    const int total_stages = ntiles * nq;
    const uint32_t stage_bytes = TILE_ROWS * p.pitch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), p.contig ? 1 : NW);
            mbar_init(empty_bar(s), NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    stamp(1);

    // ---- weight streaming.  Stage st = (tile, chunk): 16 rows x up to NW granules.  Warp w owns row
    // w (and w + 8 when NW == 8) of every stage: its lane 0 arms the stage barrier with that row's
    // byte count and issues the row's bulk copy.  The first min(S, total) stages go out right here
    // (for the Llama shapes: the CTA's whole slab); later ones as ring slots free up.
    const uint64_t pol = policy_evict_first();
    const uint8_t* src0 = p.packed + (int64_t)g0 * GRAN_B;
    const int64_t row_bytes = p.K / 2;
    auto issue_stage = [&](int ti, int q, int slot) {         // call with lane 0 only
        const int row = r0 + ti * TILE_ROWS;
        const uint32_t cb = (uint32_t)((p.whole_row ? ng : min(NW, ng - q * NW)) * GRAN_B);
        uint32_t tx = 0;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr)
            if (row + warp + rr * NW < r1) tx += cb;
        if (tx) mbar_arrive_expect_tx(full_bar(slot), tx);
        else mbar_arrive(full_bar(slot));
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp + rr * NW;
            if (row + r < r1)
                bulk_g2s_hint(ring + slot * stage_bytes + r * p.pitch,
                              src0 + (int64_t)(row + r) * row_bytes + q * NW * GRAN_B, cb, full_bar(slot), pol);
        }
    };
    // contiguous mode (one K slab): the rows of a tile are adjacent in memory, so the whole tile is
    // ONE bulk copy of up to 32 KB issued by warp 0 -- the TMA unit accepts a copy only every ~18 clk,
    // which made 16 copies per tile the longest part of the prologue (profiles/r01_gemv_notes.md).
    auto issue_tile = [&](int ti, int slot) {                 // warp 0, lane 0
        const int row = r0 + ti * TILE_ROWS;
        const uint32_t bytes = (uint32_t)(min(TILE_ROWS, r1 - row) * row_bytes);
        mbar_arrive_expect_tx(full_bar(slot), bytes);
        bulk_g2s_hint(ring + slot * stage_bytes, p.packed + (int64_t)row * row_bytes, bytes, full_bar(slot), pol);
    };
    // cross-layer software pipelining: this CTA's share of the NEXT layer's packed weights goes to L2 behind
    // its own tiles, so HBM keeps streaming through this launch's epilogue and the next launch's prologue
    // and the next launch's bulk copies hit L2
    auto prefetch_next_bulk = [&]() {
        const unsigned long long per = (((unsigned long long)p.next_chunk) + 127ull) & ~127ull;
        const unsigned long long beg = min(per * blockIdx.x, p.next_bytes), end = min(beg + per, p.next_bytes);
        for (unsigned long long off = beg; off < end; off += 32768ull) {
            const unsigned int n = (unsigned int)min(32768ull, end - off) & ~15u;
            if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.next_packed + off), "r"(n) : "memory");
        }
    };
    auto prefetch_next_lines = [&]() {
        const unsigned long long per = (((unsigned long long)p.next_chunk) + 127ull) & ~127ull;
        const unsigned long long beg = min(per * blockIdx.x, p.next_bytes), end = min(beg + per, p.next_bytes);
        for (unsigned long long off = beg + (unsigned long long)threadIdx.x * 128ull; off < end; off += (unsigned long long)NTHR * 128ull)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.next_packed + off) : "memory");
    };
    if (p.wait_weights) pdl_wait();
    if (lane == 0 && dbg != 2) {
        if (p.contig) {
            if (warp == 0) {
                for (int ti = 0; ti < ntiles && ti < S && ti < p.early_tiles; ++ti) issue_tile(ti, ti);
                if (p.next_bytes && p.pf_mode == 1 && p.early_tiles >= min(ntiles, S)) prefetch_next_bulk();
            }
        } else {
            int st = 0;
            for (int ti = 0; ti < ntiles && st < S; ++ti)
                for (int q = 0; q < nq && st < S; ++q, ++st) issue_stage(ti, q, st);
        }
    }
    if (p.next_bytes && dbg != 2) {
        if (p.pf_mode == 2) prefetch_next_lines();
        if (p.pf_mode == 3 && warp == NW - 1 && lane == 0) prefetch_next_bulk();
    }
    stamp(2);

    const int g = lane >> 2, t = lane & 3;
    pdl_wait();   // x (and the output / workspace) belong to the stream-ordered predecessor
    stamp(3);
    // tiles held back so that the x loads below (issued by the other warps at the same moment) do not queue
    // behind the whole weight stream in the L2 slices (tuning key gemv_early)
    auto issue_late = [&]() {
        for (int ti = max(p.early_tiles, 0); ti < ntiles && ti < S; ++ti) issue_tile(ti, ti);
        if (p.next_bytes && p.pf_mode == 1) prefetch_next_bulk();
    };
    const bool late = p.contig && warp == 0 && lane == 0 && dbg != 2 && p.early_tiles < min(ntiles, S);
    if (late && !p.after_x) issue_late();

    // ---- x operand (mma B fragments, registers).  Two ways:
    //  * fused (default): every CTA derives it straight from x.  One pass for the per-row amax (block
    //    reduction), then each lane converts exactly the values its own fragments hold and keeps only its
    //    limb; no shared-memory image, no second kernel.  The dependency chain of consecutive layers is
    //    then main -> main, which programmatic dependent launch overlaps CTA by CTA; with a preparation
    //    kernel in between the next launch's CTAs did not start before the LAST CTA of this one had ended
    //    (profiles/r01_gemv_notes.md, wall-clock table).
    //  * image: prepared once per launch by xprep_kernel (below) in global memory; every lane pulls its
    //    registers with two coalesced 16-byte loads per (granule, n-tile).
    const int rg = p.rg, rgs = p.rg_shift;
    uint32_t bf[GPW][NT][4][2];
    __shared__ float s_amax[8 * NW];
    __shared__ int s_tx[NW * 8 * 2];
    __shared__ int s_ex[8];
    __shared__ __align__(16) uint2 s_stage[NW * 64];
    if (p.fused_x) {
        // pass 1: amax of every batch row
        // (in contiguous mode warp 0 is still pushing the tile copies into the memory pipe: its x loads would
        // queue behind 150 KB of weight requests, so the other warps cover x and theirs go out first)
        const int skip = (p.contig && NW > 1) ? 32 : 0;
        float am[2 * NT];
#pragma unroll
        for (int m = 0; m < 2 * NT; ++m) am[m] = 0.0f;
#pragma unroll
        for (int m = 0; m < 2 * NT; ++m) {
            if (m < p.M && ctid >= skip) {
#pragma unroll 2
                for (int k = (ctid - skip) * 4; k < p.K; k += (NTHR - skip) * 4) {
                    float a[4];
                    load4f(p.x, p.x_dtype, (int64_t)m * p.K + k, a);
                    am[m] = fmaxf(am[m], fmaxf(fmaxf(fabsf(a[0]), fabsf(a[1])), fmaxf(fabsf(a[2]), fabsf(a[3]))));
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 2 * NT; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) am[m] = fmaxf(am[m], __shfl_xor_sync(0xffffffffu, am[m], o));
            if (lane == 0) s_amax[m * NW + warp] = am[m];
        }
        __syncthreads();
        if (late && p.after_x) issue_late();      // x is on chip: now the weight requests may flood the L2 slices
        stamp(4);
        // pass 2, warp-local: the 16 (t, word) items of one granule of one batch row are converted once, by
        // lanes 0..15 (8 values each: F2I runs at 16 / clk / SM, so no lane converts a value twice), cut into
        // limbs, and handed to the lanes that hold them as mma columns through a 512-byte staging slot
        const int l = g & 3, h = g >> 2;
        uint2* stg = s_stage + warp * 64;                         // [limb][t][word] -> {b0, b1}
#pragma unroll
        for (int q = 0; q < GPW; ++q)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) { bf[q][nt][j][0] = 0u; bf[q][nt][j][1] = 0u; }
#pragma unroll
        for (int em = 0; em < 2 * NT; ++em) {
            if (em < p.M) {
                const int nt = em >> 1, hh = em & 1;
                float amx = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) amx = fmaxf(amx, s_amax[em * NW + w]);
                int ex = 0;
                if (amx > 0.0f && amx < INFINITY) ex = max(-96, min(126, 155 - (int)(__float_as_uint(amx) >> 23)));
                const float up = __uint_as_float((uint32_t)(127 + ex) << 23);
                int slo = 0, shi = 0;                              // sum of X as (X & 0xffff), (X >> 16): exact in s32
                const float up16 = up * 0.0625f;
#pragma unroll
                for (int q = 0; q < GPW; ++q) {
                    if (warp + q * NW < ng) {                      // warp-uniform
                        if (lane < 16) {
                            const int t_ = lane >> 2, j_ = lane & 3;
                            float a[4], b[4];
                            const int64_t base = (int64_t)em * p.K + (int64_t)(g0 + warp + q * NW) * GRAN_K + t_ * 32 + j_ * 8;
                            load4f(p.x, p.x_dtype, base, a);
                            load4f(p.x, p.x_dtype, base + 4, b);
                            uint32_t D[8];
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                // even column: Xe = round(x 2^e); odd column: Xo = round(x 2^(e-4)), carried as
                                // Z = Xo - Xe: the raw packed byte q_lo + 16 q_hi meets Xe, the masked byte 16 q_hi meets Z
                                const int Xe0 = __float2int_rn(a[2 * i] * up), Xo0 = __float2int_rn(a[2 * i + 1] * up16);
                                const int Xe1 = __float2int_rn(b[2 * i] * up), Xo1 = __float2int_rn(b[2 * i + 1] * up16);
                                slo += (Xe0 & 0xffff) + ((Xo0 << 4) & 0xffff) + (Xe1 & 0xffff) + ((Xo1 << 4) & 0xffff);
                                shi += (Xe0 >> 16) + ((Xo0 << 4) >> 16) + (Xe1 >> 16) + ((Xo1 << 4) >> 16);
                                D[2 * i] = (uint32_t)(Xe0 + 0x00808080) ^ 0x00808080u;      // byte l = signed base-256 digit l
                                D[2 * i + 1] = (uint32_t)(Xo0 - Xe0 + 0x00808080) ^ 0x00808080u;
                                D[4 + 2 * i] = (uint32_t)(Xe1 + 0x00808080) ^ 0x00808080u;
                                D[5 + 2 * i] = (uint32_t)(Xo1 - Xe1 + 0x00808080) ^ 0x00808080u;
                            }
                            // 4x4 byte transposes: digit l of the four Xe -> b0 (meets the raw bytes), of the four Z -> b1
                            const uint32_t e0 = __byte_perm(D[0], D[2], 0x5140), e1 = __byte_perm(D[4], D[6], 0x5140);
                            const uint32_t e2 = __byte_perm(D[0], D[2], 0x7362), e3 = __byte_perm(D[4], D[6], 0x7362);
                            const uint32_t o0 = __byte_perm(D[1], D[3], 0x5140), o1 = __byte_perm(D[5], D[7], 0x5140);
                            const uint32_t o2 = __byte_perm(D[1], D[3], 0x7362), o3 = __byte_perm(D[5], D[7], 0x7362);
                            stg[0 * 16 + lane] = make_uint2(__byte_perm(e0, e1, 0x5410), __byte_perm(o0, o1, 0x5410));
                            stg[1 * 16 + lane] = make_uint2(__byte_perm(e0, e1, 0x7632), __byte_perm(o0, o1, 0x7632));
                            stg[2 * 16 + lane] = make_uint2(__byte_perm(e2, e3, 0x5410), __byte_perm(o2, o3, 0x5410));
                            stg[3 * 16 + lane] = make_uint2(__byte_perm(e2, e3, 0x7632), __byte_perm(o2, o3, 0x7632));
                        }
                        __syncwarp();
                        if (h == hh && g < rg) {
                            const uint4 v0 = *reinterpret_cast<const uint4*>(stg + (l * 4 + t) * 4);
                            const uint4 v1 = *reinterpret_cast<const uint4*>(stg + (l * 4 + t) * 4 + 2);
                            bf[q][nt][0][0] = v0.x; bf[q][nt][0][1] = v0.y; bf[q][nt][1][0] = v0.z; bf[q][nt][1][1] = v0.w;
                            bf[q][nt][2][0] = v1.x; bf[q][nt][2][1] = v1.y; bf[q][nt][3][0] = v1.z; bf[q][nt][3][1] = v1.w;
                        }
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    slo += __shfl_xor_sync(0xffffffffu, slo, o);
                    shi += __shfl_xor_sync(0xffffffffu, shi, o);
                }
                if (lane == 0) {
                    s_tx[(warp * 8 + em) * 2] = slo;
                    s_tx[(warp * 8 + em) * 2 + 1] = shi;
                    if (warp == 0) s_ex[em] = ex;
                }
            }
        }
    } else {
#pragma unroll
    for (int q = 0; q < GPW; ++q) {
        const int gq = g0 + warp + q * NW;                      // granule index in the whole K
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
            if (warp + q * NW < ng && g < rg) {
                const uint8_t* a = p.ximg + (size_t)(((((gq * NT + nt) << rgs) + g) * 4 + t) * 32);
                v0 = ldg_nc_v4(a);
                v1 = ldg_nc_v4(a + 16);
            }
            bf[q][nt][0][0] = v0.x; bf[q][nt][0][1] = v0.y; bf[q][nt][1][0] = v0.z; bf[q][nt][1][1] = v0.w;
            bf[q][nt][2][0] = v1.x; bf[q][nt][2][1] = v1.y; bf[q][nt][3][0] = v1.z; bf[q][nt][3][1] = v1.w;
        }
    }
    }
    stamp(6);

    // ---- main loop: tiles of 16 rows; per tile nq stages (one granule of this warp in each)
    int s = 0, ph = 0, st = 0;
    int rti = 0, rtq = 0;          // (tile, chunk) of the next stage to re-issue in recycle mode
    {   // the first S stages were issued above: advance the refill cursor past them
        const int adv = min(S, total_stages);
        rti = adv / nq;
        rtq = adv - rti * nq;
    }
    int round_first = 0;           // first tile of the current reduction round
    // scale / zero point of the first output this thread will finish (first pass of the first round's
    // epilogue), fetched now so that their latency is hidden behind the main loop
    float pre_sc = 0.0f, pre_zp = 0.0f;
    {
        const int er = (ctid >> 2) & (TILE_ROWS - 1), j = (ctid >> 6) / p.M;
        const int row = r0 + j * TILE_ROWS + er;
        if ((ctid & 3) == 0 && j < min(p.rt, ntiles) && row < r1) { pre_sc = __ldg(p.scales + row); pre_zp = __ldg(p.zps + row); }
    }
    for (int i = 0; i < ntiles; ++i) {
        int acc[NT][CH][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[nt][c][r] = 0;
        if (p.whole_row) {
            // one stage per tile: wait once, pull this warp's GPW granules (rows g and g+8) into registers
            // with back-to-back LDS.128, then run the IMMAs over CH independent accumulator chains
            if (dbg != 2) mbar_wait(full_bar(s), ph);
            uint4 lo[GPW], hi[GPW];
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                lo[q] = make_uint4(0u, 0u, 0u, 0u);
                hi[q] = lo[q];
                if (warp + q * NW < ng && dbg != 1) {
                    const uint32_t sbase = ring + s * stage_bytes + g * p.pitch + (warp + q * NW) * GRAN_B + t * 16;
                    lo[q] = lds128(sbase);
                    hi[q] = lds128(sbase + 8 * p.pitch);
                }
            }
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                const uint32_t wl[4] = {lo[q].x, lo[q].y, lo[q].z, lo[q].w};
                const uint32_t wh[4] = {hi[q].x, hi[q].y, hi[q].z, hi[q].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // no nibble widening (see gemv_res.cu): (q_lo + 16 q_hi) Xe + 16 q_hi (Xo - Xe) = q_lo Xe + q_hi 16 Xo
                    const uint32_t a0 = wl[j], a2 = wl[j] & 0xf0f0f0f0u;   // row g
                    const uint32_t a1 = wh[j], a3 = wh[j] & 0xf0f0f0f0u;   // row g + 8
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        mma_m16n8k32_u8s8(acc[nt][(q * 4 + j) % CH], a0, a1, a2, a3, bf[q][nt][j][0], bf[q][nt][j][1]);
                }
            }
            __syncwarp();
            if (st + S < total_stages) {                         // ring smaller than the slab: recycle
                if (lane == 0) {
                    mbar_arrive(empty_bar(s));
                    if (dbg != 2 && (!p.contig || warp == 0)) {
                        mbar_wait(empty_bar(s), ph);             // every warp has read the slot
                        if (p.contig) issue_tile(rti, s);
                        else issue_stage(rti, rtq, s);
                    }
                }
                if (++rtq == nq) { rtq = 0; ++rti; }
                __syncwarp();
            }
            ++st;
            if (++s == S) { s = 0; ph ^= 1; }
        } else {
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                if (q < nq) {                                    // stage exists (uniform over the CTA)
                    if (dbg != 2) mbar_wait(full_bar(s), ph);
                    if (warp + q * NW < ng && dbg != 1) {
                        const uint32_t sbase = ring + s * stage_bytes + g * p.pitch + warp * GRAN_B + t * 16;
                        const uint4 lo = lds128(sbase);
                        const uint4 hi = lds128(sbase + 8 * p.pitch);
                        const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w};
                        const uint32_t wh[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t a0 = wl[j], a2 = wl[j] & 0xf0f0f0f0u;
                            const uint32_t a1 = wh[j], a3 = wh[j] & 0xf0f0f0f0u;
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt)
                                mma_m16n8k32_u8s8(acc[nt][j % CH], a0, a1, a2, a3, bf[q][nt][j][0], bf[q][nt][j][1]);
                        }
                    }
                    __syncwarp();
                    if (st + S < total_stages) {
                        if (lane == 0) {
                            mbar_arrive(empty_bar(s));
                            if (dbg != 2) {
                                mbar_wait(empty_bar(s), ph);
                                issue_stage(rti, rtq, s);
                            }
                        }
                        if (++rtq == nq) { rtq = 0; ++rti; }
                        __syncwarp();
                    }
                    ++st;
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
        if (i < 5) stamp(10 + i);
        // this warp's partial of tile i -> red[i - round_first][warp][col][row]   (no barrier here)
        int* rbuf = red + ((i - round_first) * NW + warp) * (COLS * TILE_ROWS);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 8 + 2 * t;
            int c4[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) c4[r] += acc[nt][c][r];
            rbuf[col * TILE_ROWS + g] = c4[0];
            rbuf[(col + 1) * TILE_ROWS + g] = c4[1];
            rbuf[col * TILE_ROWS + g + 8] = c4[2];
            rbuf[(col + 1) * TILE_ROWS + g + 8] = c4[3];
        }
        if (i + 1 - round_first == p.rt || i + 1 == ntiles) {
            // ---- end of a round: cross-warp sum (exact), limbs -> value, zero-point term, scale, store
            if (i + 1 == ntiles) stamp(7);
            named_bar_sync(1, NTHR);
            if (i + 1 == ntiles) stamp(8);
            // one thread per (tile, batch row, row, limb): sum the NW warp partials (exact s32), then the
            // four limb lanes of a quad combine through shuffles into sum_k q*X (exact s64)
            const int nt_round = i + 1 - round_first;
            const int total = nt_round * TILE_ROWS * p.M * LIMBS;
            for (int idx0 = 0; idx0 < total; idx0 += NTHR) {
                const int idx = idx0 + ctid;
                const int l = idx & 3, er = (idx >> 2) & (TILE_ROWS - 1);
                const int rest = idx >> 6;
                const int em = rest % p.M, j = rest / p.M;
                const int row = r0 + (round_first + j) * TILE_ROWS + er;
                const bool ok = idx < total && row < r1;
                float sc = 0.0f, zp = 0.0f;
                if (round_first == 0 && idx0 == 0) { sc = pre_sc; zp = pre_zp; }
                else if (ok && l == 0) { sc = __ldg(p.scales + row); zp = __ldg(p.zps + row); }
                long long a = 0;
                if (ok) {
                    const int* rr = red + (j * NW) * (COLS * TILE_ROWS) + (em * LIMBS + l) * TILE_ROWS + er;
                    int a32 = 0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) a32 += rr[w * (COLS * TILE_ROWS)];
                    a = (long long)a32 * (1LL << (8 * l));
                }
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                if (ok && l == 0) {
                    // header of batch row em: {e, -, -, -} then per slab {sum(X & 0xffff), sum(X >> 16)}
                    int ex, tlo, thi;
                    if (p.fused_x) {
                        ex = s_ex[em];
                        tlo = 0; thi = 0;
#pragma unroll
                        for (int w = 0; w < NW; ++w) { tlo += s_tx[(w * 8 + em) * 2]; thi += s_tx[(w * 8 + em) * 2 + 1]; }
                    } else {
                        const int* hdr = p.xhdr + em * XHDR_INTS;
                        ex = hdr[0]; tlo = hdr[4 + 2 * slab]; thi = hdr[5 + 2 * slab];
                    }
                    const double down = __longlong_as_double((long long)(1023 - ex) << 52);   // 2^-e
                    const double tx = (double)tlo + 65536.0 * (double)thi;   // sum_k X over this slab
                    const float v = sc * (float)(((double)a - (double)zp * tx) * down);
                    if (p.nslab == 1) store_y(p.y, p.y_dtype, (int64_t)em * p.N + row, v);
                    else p.part[((int64_t)slab * p.M + em) * p.N + row] = v;
                }
            }
            round_first = i + 1;
            if (i + 1 < ntiles) named_bar_sync(1, NTHR);    // red is reused by the next round
        }
    }
    stamp(9);
#ifdef B200Q_PROF
    if (prof) {
        unsigned long long wall1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wall1));
        const unsigned int ln = p.launch_no & 63u;
        atomicMin(&g_gemv_wall[ln * 4 + 0], wall0);
        atomicMax(&g_gemv_wall[ln * 4 + 1], wall0);
        atomicMin(&g_gemv_wall[ln * 4 + 2], wall1);
        atomicMax(&g_gemv_wall[ln * 4 + 3], wall1);
    }
#endif

    // ---- cross-slab reduction by the last CTA of the row block (deterministic slab order)
    if (p.nslab > 1) {
        __threadfence();
        named_bar_sync(1, NTHR);
        if (ctid == 0) {
            const unsigned int old = atomicAdd(p.tickets + rb, 1u);
            *flag = (old == (unsigned)p.nslab - 1u);
        }
        named_bar_sync(1, NTHR);
        if (*flag) {
            __threadfence();
            const int nrows = r1 - r0;
            for (int idx = ctid; idx < nrows * p.M; idx += NTHR) {
                const int m = idx / nrows, row = r0 + idx % nrows;
                float a = 0.0f;
                for (int sl = 0; sl < p.nslab; ++sl) a += __ldcg(p.part + ((int64_t)sl * p.M + m) * p.N + row);
                store_y(p.y, p.y_dtype, (int64_t)m * p.N + row, a);
            }
            if (ctid == 0) p.tickets[rb] = 0u;   // leave the workspace zeroed
        }
    }
}

// ---------------------------------------------------------------------------------------------
// x operand preparation: one CTA per stored batch row.  X = round(x * 2^e) with the row's own
// power-of-two scale (amax * 2^e in [2^29, 2^30)), cut into four signed base-256 digits; digit l of
// the 8 values k0 + 32t + 8j + {0..7} is stored as the two B registers of IMMA j in mma column
// 4h + l: {X0,X2,X4,X6} (meets the low nibbles) and {X1,X3,X5,X7} (high nibbles).
// Also emits, per batch row, e and the exact sum of X over every K slab (two s32 halves).
struct XprepParams {
    const void* x;
    uint8_t* img;
    int* hdr;
    int x_dtype, M, K, G, NT, rg_shift;
    int nslab, gran_q, gran_rem;
};

constexpr int XP_THREADS = 512;

__global__ void __launch_bounds__(XP_THREADS) xprep_kernel(const XprepParams p) {
    __shared__ float s_amax[XP_THREADS / 32];
    __shared__ int s_ls[2 * 1024];                       // per granule {sum lo16, sum hi16}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hs = p.rg_shift - 2;
    const int nt = blockIdx.x >> hs, h = blockIdx.x & ((1 << hs) - 1);
    const int m = nt * 2 + h;
    pdl_launch_dependents();
    pdl_wait();                                          // x belongs to the stream-ordered predecessor
    const int items = p.G * 16;                          // (granule, word j, t)
    auto slot_of = [&](int it) {
        const int t_ = it & 3, j = (it >> 2) & 3, gq = it >> 4;
        return (size_t)((((((gq * p.NT + nt) << p.rg_shift) + 4 * h) * 4 + t_) * 4 + j) * 8);
    };
    if (m >= p.M) {                                      // unused batch row of the last n-tile: zero columns
        for (int it = threadIdx.x; it < items; it += XP_THREADS)
#pragma unroll
            for (int l = 0; l < 4; ++l) *reinterpret_cast<uint2*>(p.img + slot_of(it) + 128 * l) = make_uint2(0u, 0u);
        return;
    }
    float am = 0.0f;
    for (int k = threadIdx.x * 4; k < p.K; k += XP_THREADS * 4) {
        float a[4];
        load4f(p.x, p.x_dtype, (int64_t)m * p.K + k, a);
        am = fmaxf(am, fmaxf(fmaxf(fabsf(a[0]), fabsf(a[1])), fmaxf(fabsf(a[2]), fabsf(a[3]))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    if (lane == 0) s_amax[warp] = am;
    __syncthreads();
    am = 0.0f;
#pragma unroll
    for (int w = 0; w < XP_THREADS / 32; ++w) am = fmaxf(am, s_amax[w]);
    int ex = 0;
    if (am > 0.0f && am < INFINITY) ex = max(-96, min(126, 155 - (int)(__float_as_uint(am) >> 23)));
    const float up = __uint_as_float((uint32_t)(127 + ex) << 23);
    for (int it0 = 0; it0 < items; it0 += XP_THREADS) {            // items is a multiple of 16, not of 32:
        const int it = it0 + threadIdx.x;                          // all lanes take part in the shuffles
        const bool active = it < items;
        const int t_ = it & 3, j = (it >> 2) & 3, gq = it >> 4;
        float v[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (active) {
            float a[4], b[4];
            const int64_t base = (int64_t)m * p.K + gq * GRAN_K + t_ * 32 + j * 8;
            load4f(p.x, p.x_dtype, base, a);
            load4f(p.x, p.x_dtype, base + 4, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
        }
        uint32_t D[8];
        int slo = 0, shi = 0;                            // sum of X as (X & 0xffff) and (X >> 16): exact in s32
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int Xe = __float2int_rn(v[2 * i] * up), Xo = __float2int_rn(v[2 * i + 1] * (up * 0.0625f));
            slo += (Xe & 0xffff) + ((Xo << 4) & 0xffff);
            shi += (Xe >> 16) + ((Xo << 4) >> 16);
            D[2 * i] = (uint32_t)(Xe + 0x00808080) ^ 0x00808080u;       // byte l = signed base-256 digit l
            D[2 * i + 1] = (uint32_t)(Xo - Xe + 0x00808080) ^ 0x00808080u;
        }
        // 4x4 byte transposes: digit l of the even values -> lo[l], of the odd values -> hi[l]
        const uint32_t e0 = __byte_perm(D[0], D[2], 0x5140), e1 = __byte_perm(D[4], D[6], 0x5140);
        const uint32_t e2 = __byte_perm(D[0], D[2], 0x7362), e3 = __byte_perm(D[4], D[6], 0x7362);
        const uint32_t o0 = __byte_perm(D[1], D[3], 0x5140), o1 = __byte_perm(D[5], D[7], 0x5140);
        const uint32_t o2 = __byte_perm(D[1], D[3], 0x7362), o3 = __byte_perm(D[5], D[7], 0x7362);
        const uint32_t lo[4] = {__byte_perm(e0, e1, 0x5410), __byte_perm(e0, e1, 0x7632),
                                __byte_perm(e2, e3, 0x5410), __byte_perm(e2, e3, 0x7632)};
        const uint32_t hi[4] = {__byte_perm(o0, o1, 0x5410), __byte_perm(o0, o1, 0x7632),
                                __byte_perm(o2, o3, 0x5410), __byte_perm(o2, o3, 0x7632)};
        if (active) {
            const size_t slot = slot_of(it);
#pragma unroll
            for (int l = 0; l < 4; ++l) *reinterpret_cast<uint2*>(p.img + slot + 128 * l) = make_uint2(lo[l], hi[l]);
        }
        // the 16 lanes (t, word) of a granule: their sum of X goes to the granule's slot (plain stores)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            slo += __shfl_xor_sync(0xffffffffu, slo, o);
            shi += __shfl_xor_sync(0xffffffffu, shi, o);
        }
        if (active && (lane & 15) == 0 && gq < 1024) { s_ls[2 * gq] = slo; s_ls[2 * gq + 1] = shi; }
    }
    __syncthreads();
    // header: {e, 0, 0, 0}, then per K slab the exact sum of X (warp `slab` adds its granules)
    int* hdr = p.hdr + m * XHDR_INTS;
    if (threadIdx.x == 0) hdr[0] = ex;
    if (warp < p.nslab) {
        const int g0 = warp * p.gran_q + min(warp, p.gran_rem);
        const int ng = p.gran_q + (warp < p.gran_rem ? 1 : 0);
        int alo = 0, ahi = 0;
        for (int gq = lane; gq < ng; gq += 32) { alo += s_ls[2 * (g0 + gq)]; ahi += s_ls[2 * (g0 + gq) + 1]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            alo += __shfl_xor_sync(0xffffffffu, alo, o);
            ahi += __shfl_xor_sync(0xffffffffu, ahi, o);
        }
        if (lane == 0) { hdr[4 + 2 * warp] = alo; hdr[5 + 2 * warp] = ahi; }
    }
}

struct GemvConfig {
    int nw, gpw, nt, nslab, nrb, stages, whole_row, contig, pitch, rt, ring_off, red_off, rg;
    size_t smem;
};

// kernel instances that exist (NW, GPW, NT); see launch_gemv
bool has_instance(int nw, int gpw, int nt) {
    if (nw == 16) return (nt == 1 && gpw <= 3) || (nt == 2 && gpw <= 3) || (nt == 4 && gpw <= 2);
    if (nw == 8) return (nt == 1 && gpw <= 4) || (nt == 2 && gpw <= 4) || (nt == 4 && gpw <= 2);
    return false;
}

bool plan(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, GemvConfig* c) {
    if (M < 1 || M > 8 || K % GRAN_K != 0 || K <= 0 || N < 1 || N > 0x7fffffff || K > 0x7fffffff) return false;
    const Tuning& tu = tuning();
    const int G = (int)(K / GRAN_K);
    const int nt = M <= 2 ? 1 : (M <= 4 ? 2 : 4);      // two batch rows (x 4 limbs) per n-tile
    const int rg = M == 1 ? 4 : 8;
    int ctas = dev.sm_count;
    if (tu.gemv_ctas > 0 && tu.gemv_ctas < ctas) ctas = tu.gemv_ctas;
    double best_cost = 1e30;
    bool found = false;
    // "two layers per SM" mode (tuning key gemv_occ2, off by default): 8-warp CTAs that use at most half
    // of the shared memory and registers, so the CTAs of the NEXT launch (programmatic dependent launch,
    // static weights) become resident next to the running ones and prefetch their first tiles meanwhile.
    // Measured slower (11.5 us vs 9.1 us): profiles/r01_gemv_notes.md
    const bool occ2 = tu.gemv_occ2 != 0;
    constexpr int STATIC_SMEM = 12 * 1024;   // s_amax, s_tx, s_ex, s_stage of gemv_kernel (10.6 KB at 16 warps)
    const int smem_budget = occ2 ? (dev.max_smem_optin - 2048) / 2 - 1024 - STATIC_SMEM : dev.max_smem_optin - STATIC_SMEM;
    for (int nw = 16; nw >= 8; nw -= 8) {
        if (occ2 && nw != 8) continue;
        if (!occ2 && tu.gemv_warps > 0 && tu.gemv_warps != nw) continue;
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
            const int red_tile = nw * nt * 8 * TILE_ROWS * 4;
            int rt = (occ2 ? 8192 : 32768) / red_tile;
            if (rt < 1) rt = 1;
            if (rt > ntiles) rt = ntiles;
            // one round (one barrier + one epilogue per CTA) whenever the partials of all tiles and the whole
            // slab still fit next to each other
            if (!occ2 && rt < ntiles && ns == 1 &&
                RED_OFF + ntiles * red_tile + 128 + ntiles * TILE_ROWS * ng * GRAN_B <= smem_budget) rt = ntiles;
            const int red_off = RED_OFF;
            const int ring_off = ((red_off + rt * red_tile + 127) / 128) * 128;
            // stage = 16 rows x the whole slab row when at least 3 of those fit (fewest, largest bulk
            // copies: the TMA unit accepts one every ~18 clk), else 16 rows x NW granules
            int whole_row = 1;
            const int contig = ns == 1 ? 1 : 0;          // one K slab: a tile's rows are adjacent in memory
            int pitch = contig ? ng * GRAN_B : (ng | 1) * GRAN_B;
            int stage_bytes = TILE_ROWS * pitch;
            int per_tile = 1;
            int stages = (smem_budget - ring_off) / stage_bytes;
            if (!contig && stages < (ntiles < 3 ? ntiles : 3)) {
                whole_row = 0;
                pitch = ((ng < nw ? ng : nw) | 1) * GRAN_B;
                stage_bytes = TILE_ROWS * pitch;
                per_tile = gpw;
                stages = (smem_budget - ring_off) / stage_bytes;
            }
            if (stages > ntiles * per_tile) stages = ntiles * per_tile;
            if (stages > MAX_STAGES) stages = MAX_STAGES;
            if (tu.gemv_stages > 0 && tu.gemv_stages < stages) stages = tu.gemv_stages;
            if (stages < 1) continue;
            // cost ~ bytes of the busiest CTA; small charges for the cross-slab reduction, for idle
            // warps in the last chunk, for a ring that cannot hold the whole slab, and for 8 warps
            double cost = (double)rows * ng * (1.0 + 0.03 * (ns - 1)) * (1.0 + 0.1 * ((double)gpw * nw / ng - 1.0));
            if (stages < ntiles * per_tile) cost *= 1.05;
            if (nw == 8 && !occ2) cost *= 1.10;
            if (cost < best_cost) {
                best_cost = cost;
                found = true;
                c->nw = nw; c->gpw = gpw; c->nt = nt; c->nslab = ns; c->nrb = nrb; c->stages = stages;
                c->whole_row = whole_row; c->contig = contig;
                c->pitch = pitch; c->rt = rt; c->ring_off = ring_off; c->red_off = red_off;
                c->rg = rg;
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
        if (tuning().gemv_occ2) B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
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

bool gemv_supported(int64_t M, int64_t N, int64_t K, int x_dtype, const DeviceInfo* dev) {
    (void)x_dtype;
    DeviceInfo d;
    if (dev) d = *dev;
    else if (current_device(&d)) {   // no device yet (workspace sizing before the first launch): a whole B200
        d.sm_count = 148;
        d.max_smem_optin = 232448;
    }
    GemvConfig c;
    return plan(d, M, N, K, &c);
}

// workspace: [tickets MAX_RB u32][x headers 8 x XHDR_INTS s32][x image, 2 alternating copies][slab partials]
static size_t ximg_bytes(int64_t K) { return (size_t)(K / GRAN_K) * 4 * 8 * 128; }   // NT <= 4, rg <= 8
constexpr size_t WS_HDR_OFF = (size_t)MAX_RB * 4;
constexpr size_t WS_IMG_OFF = WS_HDR_OFF + 2 * 8 * XHDR_INTS * 4;

size_t gemv_ws_bytes(int64_t M, int64_t N, int64_t K) {
    if (!gemv_supported(M, N, K, B200Q_F32)) return 0;
    return WS_IMG_OFF + 2 * ximg_bytes(K) + (size_t)MAX_SLABS * M * N * 4;
}

int launch_gemv(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                const float* scales, const float* zps, void* y, int y_dtype, int64_t M, int64_t N,
                int64_t K, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st,
                const uint8_t* next_packed, size_t next_bytes) {
    GemvConfig c;
    if (!plan(dev, M, N, K, &c)) return set_error(B200Q_EINVAL, "gemv: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv: x and packed must be 16-byte aligned");
    GemvParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.y = y;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.nslab = c.nslab; p.nrb = c.nrb;
    p.rows_q = (int)(N / c.nrb); p.rows_rem = (int)(N % c.nrb);
    const int G = (int)(K / GRAN_K);
    p.gran_q = G / c.nslab; p.gran_rem = G % c.nslab;
    p.stages = c.stages; p.whole_row = c.whole_row; p.contig = c.contig; p.pitch = c.pitch; p.rt = c.rt;
    p.ring_off = c.ring_off; p.red_off = c.red_off;
    p.rg = c.rg; p.rg_shift = c.rg == 4 ? 2 : 3;
    const bool is_static = (flags & B200Q_FLAG_STATIC_WEIGHTS) != 0;
    const bool pdl = tuning().gemv_pdl != 0;
    p.wait_weights = is_static ? 0 : 1;
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    p.pf_mode = tuning().gemv_pf;
    p.fused_x = tuning().gemv_xprep == 0;
    p.early_tiles = (tuning().gemv_early >= 0 && p.fused_x) ? tuning().gemv_early : 1 << 20;
    p.after_x = 0;
    if (tuning().gemv_early == -2 && p.fused_x) { p.early_tiles = 0; p.after_x = 1; }
    p.next_packed = next_packed;
    p.next_bytes = next_packed && (reinterpret_cast<uintptr_t>(next_packed) & 15) == 0 ? next_bytes : 0;
    p.next_chunk = (unsigned int)(p.next_bytes / (unsigned long long)(c.nslab * c.nrb));
    {
        static thread_local unsigned launch_no = 0;
        p.launch_no = launch_no++;
    }
    {
        const size_t need = WS_IMG_OFF + 2 * ximg_bytes(K) + (c.nslab > 1 ? (size_t)c.nslab * M * N * 4 : 0);
        if (!ws || ws_bytes < need) return set_error(B200Q_EWORKSPACE, "gemv: workspace too small (%zu < %zu)", ws_bytes, need);
        if (reinterpret_cast<uintptr_t>(ws) & 127) return set_error(B200Q_EALIGN, "gemv: workspace must be 128-byte aligned");
        uint8_t* w8 = static_cast<uint8_t*>(ws);
        // the x image / header alternate between two copies: the preparation of launch i+1 may overlap
        // (programmatic dependent launch) CTAs of launch i that still hold pointers into copy i
        static thread_local unsigned flip = 0;
        flip ^= 1u;
        p.tickets = reinterpret_cast<unsigned int*>(w8);
        int* hdr = reinterpret_cast<int*>(w8 + WS_HDR_OFF) + flip * 8 * XHDR_INTS;
        uint8_t* img = w8 + WS_IMG_OFF + flip * ximg_bytes(K);
        p.xhdr = hdr;
        p.ximg = img;
        p.part = reinterpret_cast<float*>(w8 + WS_IMG_OFF + 2 * ximg_bytes(K));
        XprepParams xp{};
        xp.x = x; xp.img = img; xp.hdr = hdr; xp.x_dtype = x_dtype; xp.M = (int)M; xp.K = (int)K; xp.G = G;
        xp.NT = c.nt; xp.rg_shift = p.rg_shift; xp.nslab = c.nslab; xp.gran_q = p.gran_q; xp.gran_rem = p.gran_rem;
        if (G > 1024) return set_error(B200Q_EINVAL, "gemv: K > 131072 not supported");
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(c.nt * (c.rg / 4)));
        cfg.blockDim = dim3(XP_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attrs[1];
        attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = pdl ? 1 : 0;
        if (!p.fused_x) B200Q_CUDA(cudaLaunchKernelEx(&cfg, xprep_kernel, xp));
    }
#define B200Q_GEMV_CASE(NW_, GPW_, NT_) \
    if (c.nw == NW_ && c.gpw == GPW_ && c.nt == NT_) return launch_inst<NW_, GPW_, NT_>(c, p, pdl, st);
    B200Q_GEMV_CASE(16, 1, 1) B200Q_GEMV_CASE(16, 2, 1) B200Q_GEMV_CASE(16, 3, 1)
    B200Q_GEMV_CASE(16, 1, 2) B200Q_GEMV_CASE(16, 2, 2) B200Q_GEMV_CASE(16, 3, 2)
    B200Q_GEMV_CASE(16, 1, 4) B200Q_GEMV_CASE(16, 2, 4)
    B200Q_GEMV_CASE(8, 1, 1) B200Q_GEMV_CASE(8, 2, 1) B200Q_GEMV_CASE(8, 3, 1) B200Q_GEMV_CASE(8, 4, 1)
    B200Q_GEMV_CASE(8, 1, 2) B200Q_GEMV_CASE(8, 2, 2) B200Q_GEMV_CASE(8, 3, 2) B200Q_GEMV_CASE(8, 4, 2)
    B200Q_GEMV_CASE(8, 1, 4) B200Q_GEMV_CASE(8, 2, 4)
#undef B200Q_GEMV_CASE
    return set_error(B200Q_EINVAL, "gemv: no kernel instance for nw=%d gpw=%d nt=%d", c.nw, c.gpw, c.nt);
}

}  // namespace b200q

#ifdef B200Q_PROF
/* bench-only: wall-clock records (64 launches x {min start, max start, min end, max end}); reset = 1 re-arms them */
extern "C" int b200q_debug_wall(unsigned long long* h_out, int reset) {
    if (reset) {
        unsigned long long init[64 * 4];
        for (int i = 0; i < 64; ++i) { init[4 * i] = ~0ull; init[4 * i + 1] = 0; init[4 * i + 2] = ~0ull; init[4 * i + 3] = 0; }
        return b200q::check_cuda(cudaMemcpyToSymbol(b200q::g_gemv_wall, init, sizeof(init)), "reset wall");
    }
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_gemv_wall, sizeof(unsigned long long) * 64 * 4), "read wall");
}

/* bench-only: copy the per-CTA phase timestamps of the last profiled GEMV launch (256 x 16 int64) */
extern "C" int b200q_debug_read_prof(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_gemv_prof, sizeof(long long) * 256 * 16), "read prof");
}
#endif
