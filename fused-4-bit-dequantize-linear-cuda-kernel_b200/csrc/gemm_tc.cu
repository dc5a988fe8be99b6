// Prefill / grouped (MoE) path on the 5th-generation tensor cores:
//     Y[m, n] = sum_k X[m, k] * dequant(W)[n, k]        for M >= 9 rows (and every grouped call)
//
// "Swap-AB" tcgen05 GEMM: the 128 x K slab of INT4 weights is the A operand and lives in TENSOR
// MEMORY; the activations are the B operand in shared memory (TMA, 128-byte swizzle), BN = 32 .. 256
// token rows per tile, so one CTA tile is D[128 weight rows x BN tokens] in TMEM (fp32).  BN <= 192 leaves
// room for two accumulators (the epilogue of tile i overlaps the MMAs of tile i + 1); 32 / 64 serve decode-sized
// batches.  Persistent grid; whole tiles round-robin, or stream-K (k-blocks dealt out evenly, partial
// accumulators through the workspace) when the only wave is less than half full; grouped mode for the MoE
// layer, optionally with the SiLU-gate in the epilogue (rows of w1 / w3 interleaved).
//
//   warp 0      TMA producer: per 64-column k-block one activation tile (256 x 64 fp16, plus the
//               low-part tile for fp32 inputs) and one packed-weight tile (128 rows x 32 bytes)
//   warps 4-11  dequant (two groups on alternate k-blocks): thread = weight row; LDS the row's 32 packed bytes, nibble -> fp16 SUBNORMAL
//               with one LOP3 per two weights (no arithmetic), tcgen05.st into a TMEM A slot
//   warp 1      MMA issuer: 4 (x2 for hi/lo) tcgen05.mma.kind::f16 M=128 N=256 K=16 per k-block,
//               A from TMEM, B from shared memory; tcgen05.commit frees the stage / the A slot
//   warps 12-15 epilogue: tcgen05.ld the accumulator, y = s_n * (D * 2^(24-e_m) - zp_n * sum_k x[m,k]),
//               coalesced stores (a warp writes 32 consecutive n of one token row)
//
// The activations are prepared by xprep_gemm_kernel: per token row a power-of-two scale 2^e
// (amax * 2^e in [2^13, 2^14)), fp16 hi part (and fp16 remainder for fp32 inputs, so that products are
// exact to ~2^-22), columns permuted inside each group of 8 to the order the A registers come out in
// (k0,k4,k1,k5,k2,k6,k3,k7) with the odd columns pre-scaled by 2^-4 (they meet nibbles left at bits
// 4-7 of the fp16 mantissa = q * 2^-20 instead of q * 2^-24), plus the row sum for the zero-point term.
//
// Grouped mode (MoE): token rows [starts[e], ends[e]) use expert e of packed [E, N, K/2]; tiles are
// enumerated on the device from the (device-resident) ranges, no host synchronisation.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (M > 1 re-streams the weights per
// row) and csrc/moe_int4_kernel.cu:17-90 (one CTA per expert).
#include <cuda.h>
#include <mutex>
#include <type_traits>
#include "internal.h"
#include "ptx.cuh"
#include "tc.cuh"

namespace b200q {

// bench-only (-DB200Q_PROF builds of tools/): ablation switches (tuning key gemm_debug) and a per-k-block SM-clock
// trace of CTA 0 (bit 4, tools/gemm_trace.py).  The product build compiles all of it out.
#ifdef B200Q_PROF
#define B200Q_GEMM_DBG(p) ((p).debug)
__device__ long long g_gemm_trace[8 * 96];
#define B200Q_TRACE(row, idx) g_gemm_trace[(row) * 96 + (idx)] = clock64()
#else
#define B200Q_GEMM_DBG(p) 0
#define B200Q_TRACE(row, idx) ((void)0)
#endif

namespace {

constexpr int BM = 128;            // weight rows per tile (UMMA M, TMEM lanes)
// token rows per tile (UMMA N, accumulator columns) is a template parameter BN in {128, 192, 256}: with
// BN <= 192 two accumulators fit next to the A ring (2*BN + 128 <= 512 TMEM columns) and the epilogue of
// tile i overlaps the MMAs of tile i+1; BN = 256 has one accumulator.
constexpr int BK = 64;             // k-block: one 128-byte swizzle row of fp16
// TMEM A ring: A_SLOTS k-blocks x 32 columns.  The ring is a latency loop (dequant -> tcgen05.st -> wait::st ->
// mbarrier -> MMA -> commit -> mbarrier -> dequant, ~3 k clk per lap): with 4 slots a k-block cannot take less than
// ~850 clk whatever the MMAs cost (gemm_debug ablation: removing every TMA load changed nothing), with 8 slots the
// MMAs (716 clk per k-block at 256 tokens) are the limit again.  192 tokens x 2 accumulators leave room for 4 only.
// A pipeline stage holds KSUB (1 or 2) such 64-column sub-blocks: every stage costs ~300 - 550 clk of mbarrier /
// tcgen05.commit hand-shakes whatever it carries (ablation with loads, stores and MMAs removed,
// tools/gemm_ablate.py), so two sub-blocks per stage halve that overhead per byte.
__host__ __device__ constexpr int a_slots(int bn, int ksub) { return ksub == 2 ? 4 : (bn == 192 ? 4 : 8); }
constexpr int TMEM_COLS = 512;
constexpr int GEMM_THREADS = 16 * 32;
constexpr int W_TILE_BYTES = BM * (BK / 2);    // 4 KB
constexpr int MAX_STAGES = 24;     // small token tiles (8 KB stages) need many stages to keep enough bytes in flight

// shared memory map
constexpr int OFF_FULL = 0;        // [MAX_STAGES]
constexpr int OFF_EMPTY = 256;     // [MAX_STAGES]
constexpr int OFF_AFULL = 512;     // [A_SLOTS]
constexpr int OFF_AEMPTY = 576;    // [A_SLOTS]
constexpr int OFF_DFULL = 640;     // [2]
constexpr int OFF_DEMPTY = 656;    // [2]
constexpr int OFF_TMEMPTR = 672;
constexpr int OFF_TOK = 1024;      // float descale[256], rowsum[256]
constexpr int OFF_STAGES = 4096;   // 1024-byte aligned stage buffers

struct GemmParams {
    const float* scales;
    const float* zps;
    void* y;
    const float* descale;    // [R] 2^(24 - e_m)
    const float* rowsum;     // [R] sum_k x[m, k]
    const int32_t* starts;   // grouped: [E] device, else nullptr
    const int32_t* ends;
    const int32_t* emap;     // grouped: weight expert of every range (nullptr: range e uses expert e)
    int E;
    int y_dtype;
    int R, N, K;             // token rows, weight rows (per expert), columns
    int bn;                  // token rows per tile
    int mt_bound;            // upper bound of m-tiles (grouped: ceil(R/BN) + E)
    int n_tiles;             // ceil(N / BM)
    int stages;
    int stage_bytes;
    // stream-K (plain linear with a badly quantised last wave): CTA c owns k-blocks [c*skq + min(c, skr), ...) of the
    // linearised (tile, k-block) space; a tile cut between CTAs is finished by the CTA that holds its last k-block,
    // the others publish their fp32 partial accumulator (workspace slot = CTA index) and raise a flag
    int mt_major;            // tile order: 0 = all token tiles of a weight tile first; 1 = all weight tiles of a token tile first
                             //    (many, mostly empty ranges: the valid tiles are then a PREFIX of the index space, so the
                             //    persistent CTAs stay balanced whatever the bound on the number of token tiles is)
    int gated;               // 1: weight rows are interleaved (2f: gate, 2f + 1: up); the epilogue writes
                             //    h[m, f] = silu(gate) * up into y [R, N / 2] (fused SiLU-gate of the MoE layer)
    int debug;               // bench-only ablations, see the TMA producer
    int sk;                  // 0: whole tiles, round-robin
    int skq, skr;
    float* part;             // [grid][BN][BM] fp32
    unsigned int* flags;     // [grid], zero between launches
};

struct TileInfo {
    int e, m0, mend, n0;     // e: weight expert of the tile's range
};

// tile index -> (expert, first token row, end of the group, first weight row); false = no such tile
__device__ __forceinline__ bool locate(const GemmParams& p, int t, TileInfo& ti) {
    int nt, mt;
    if (p.mt_major) { mt = t / p.n_tiles; nt = t - mt * p.n_tiles; }
    else { nt = t / p.mt_bound; mt = t - nt * p.mt_bound; }
    if (nt >= p.n_tiles || mt >= p.mt_bound) return false;
    ti.n0 = nt * BM;
    if (!p.starts) {
        ti.e = 0;
        ti.m0 = mt * p.bn;
        ti.mend = p.R;
        return ti.m0 < p.R;
    }
    int rem = mt;
    for (int e = 0; e < p.E; ++e) {
        const int lo = max(p.starts[e], 0), hi = min(p.ends[e], p.R);
        const int cnt = hi - lo;
        if (cnt <= 0) continue;
        const int tiles = (cnt + p.bn - 1) / p.bn;
        if (rem < tiles) {
            ti.e = p.emap ? p.emap[e] : e;
            ti.m0 = lo + rem * p.bn;
            ti.mend = hi;
            return true;
        }
        rem -= tiles;
    }
    return false;
}

// One unit of work of a CTA: k-blocks [kb0, kb1) of tile ti.  publish: the tile's last k-block belongs to another
// CTA, hand the partial accumulator over; nadd > 0: add the partials of CTAs [c0, c0 + nadd) before the epilogue.
struct Item {
    TileInfo ti;
    int kb0, kb1, publish, nadd, c0;
};

__device__ __forceinline__ int sk_begin(const GemmParams& p, int c) { return c * p.skq + min(c, p.skr); }
__device__ __forceinline__ int sk_owner(const GemmParams& p, int pos) {
    const int big = p.skr * (p.skq + 1);
    return pos < big ? pos / (p.skq + 1) : p.skr + (pos - big) / p.skq;
}

// calls body(item) for every work item of this CTA, in the same order in every warp role
template <typename F>
__device__ __forceinline__ void for_each_item(const GemmParams& p, int KB, int total_tiles, F&& body) {
    Item it;
    if (!p.sk) {
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            if (!locate(p, t, it.ti)) {
                if (p.mt_major) break;          // the valid tiles are a prefix of the index space in this order
                continue;
            }
            it.kb0 = 0; it.kb1 = KB; it.publish = 0; it.nadd = 0; it.c0 = 0;
            body(it);
        }
        return;
    }
    const int beg = sk_begin(p, blockIdx.x), end = sk_begin(p, blockIdx.x + 1);
    // publish first (the head of the LAST tile of the range), then everything else in order: a CTA never waits for
    // a partial before it has handed out its own, so the flags cannot deadlock
    const int last_t = (end - 1) / KB;
    const bool last_cut = end > beg && end < (last_t + 1) * KB;
    if (last_cut) {
        locate(p, last_t, it.ti);
        it.kb0 = max(beg, last_t * KB) - last_t * KB; it.kb1 = end - last_t * KB; it.publish = 1; it.nadd = 0; it.c0 = 0;
        body(it);
    }
    const int stop = last_cut ? max(beg, last_t * KB) : end;
    for (int pos = beg; pos < stop;) {
        const int t = pos / KB;
        locate(p, t, it.ti);
        it.kb0 = pos - t * KB; it.kb1 = KB; it.publish = 0;
        it.c0 = it.kb0 ? sk_owner(p, t * KB) : 0;
        it.nadd = it.kb0 ? (int)blockIdx.x - it.c0 : 0;
        body(it);
        pos = (t + 1) * KB;
    }
}

template <int PARTS, int BN, int KSUB>   // PARTS 1: fp16 activations (hi only); 2: hi + lo (fp32 activations)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
               const __grid_constant__ CUtensorMap map_w, const GemmParams p) {
    constexpr int X_TILE_BYTES = BN * BK * 2;
    constexpr int A_SLOTS = a_slots(BN, KSUB);
    constexpr int A_COLS = 32 * KSUB;                                    // TMEM columns of one A slot
    constexpr int DBUF = 2 * BN + A_SLOTS * A_COLS <= TMEM_COLS ? 2 : 1; // accumulators in TMEM
    constexpr int A_BASE = DBUF * BN;                                    // first column of the A ring
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    auto full = [&](int s) { return sb + OFF_FULL + 8u * s; };
    auto empty = [&](int s) { return sb + OFF_EMPTY + 8u * s; };
    auto afull = [&](int a) { return sb + OFF_AFULL + 8u * a; };
    auto aempty = [&](int a) { return sb + OFF_AEMPTY + 8u * a; };
    auto dfull = [&](int b) { return sb + OFF_DFULL + 8u * b; };
    auto dempty = [&](int b) { return sb + OFF_DEMPTY + 8u * b; };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEMPTR);
    float* tok = reinterpret_cast<float*>(smem + OFF_TOK);
    auto xh_smem = [&](int s) { return sb + OFF_STAGES + (uint32_t)s * p.stage_bytes; };
    auto xl_smem = [&](int s) { return xh_smem(s) + KSUB * X_TILE_BYTES; };
    auto w_smem = [&](int s) { return xh_smem(s) + PARTS * KSUB * X_TILE_BYTES; };
    const int KB = p.K / (BK * KSUB);                                    // stages per tile
    const int dbg = B200Q_GEMM_DBG(p);
    (void)dbg;
    const int total_tiles = p.n_tiles * p.mt_bound;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 5);          // 4 dequant warps + the MMA commit
        }
        for (int a = 0; a < A_SLOTS; ++a) {
            mbar_init(afull(a), 4);
            mbar_init(aempty(a), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(dfull(b), 1);
            mbar_init(dempty(b), 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(sb + OFF_TMEMPTR, TMEM_COLS);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_xh);
        if (PARTS == 2) tma_prefetch_desc(&map_xl);
        tma_prefetch_desc(&map_w);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    pdl_wait();       // barriers, TMEM and descriptors are set up: now wait for the activation preparation kernel

    if (warp == 0) {
        // =============================================================== TMA producer
        if (lane == 0) {
            int s = 0, ph = 0, it = 0;
            for_each_item(p, KB, total_tiles, [&](const Item& w) {
                const TileInfo& ti = w.ti;
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    if (it >= S) mbar_wait(empty(s), ph ^ 1);
                    if ((dbg & 16) && blockIdx.x == 0 && it < 96) B200Q_TRACE(0, it);
                    // bench-only ablation (tuning key gemm_debug): 1 = no weight loads, 2 = no activation loads
                    const uint32_t xb = (dbg & 2) ? 0u : (uint32_t)(PARTS * KSUB * X_TILE_BYTES);
                    const uint32_t wb = (dbg & 1) ? 0u : (uint32_t)(KSUB * W_TILE_BYTES);
                    mbar_arrive_expect_tx(full(s), xb + wb);
                    if (xb) {
#pragma unroll
                        for (int sub = 0; sub < KSUB; ++sub) {         // one 128-byte-swizzle atom [BN][64] per sub-block
                            tma_load_2d(xh_smem(s) + sub * X_TILE_BYTES, &map_xh, (kb * KSUB + sub) * BK, ti.m0, full(s));
                            if (PARTS == 2) tma_load_2d(xl_smem(s) + sub * X_TILE_BYTES, &map_xl, (kb * KSUB + sub) * BK, ti.m0, full(s));
                        }
                    }
                    if (wb) tma_load_2d(w_smem(s), &map_w, kb * KSUB * (BK / 2), ti.e * p.N + ti.n0, full(s));   // [128][32 KSUB] bytes
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            });
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(BM, BN, 0);
            int s = 0, ph = 0, ait = 0, li = 0;
            for_each_item(p, KB, total_tiles, [&](const Item& w) {
                const int ab = li % DBUF;
                const uint32_t d_tmem = tmem + ab * BN;
                if (li >= DBUF) {                               // the epilogue has drained this accumulator
                    mbar_wait(dempty(ab), ((li / DBUF) - 1) & 1);
                    tc_fence_after_sync();
                }
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++ait) {
                    const int a = ait % A_SLOTS;
                    mbar_wait(full(s), ph);
                    if ((dbg & 16) && blockIdx.x == 0 && ait < 96) B200Q_TRACE(1, ait);
                    mbar_wait(afull(a), (ait / A_SLOTS) & 1);
                    if ((dbg & 16) && blockIdx.x == 0 && ait < 96) B200Q_TRACE(2, ait);
                    tc_fence_after_sync();
                    const uint32_t a_tmem = tmem + A_BASE + A_COLS * a;
#pragma unroll
                    for (int sub = 0; sub < KSUB; ++sub) {
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            if (dbg & 4) break;                     // ablation: no MMAs, only the commits
                            const uint64_t bh = smem_desc(xh_smem(s) + sub * X_TILE_BYTES + kk * 32, 16, 1024, SWIZZLE_128B);
                            mma_ts_f16(d_tmem, a_tmem + 32 * sub + 8 * kk, bh, idesc, (kb > w.kb0 || sub || kk) ? 1u : 0u);
                            if (PARTS == 2) {
                                const uint64_t bl = smem_desc(xl_smem(s) + sub * X_TILE_BYTES + kk * 32, 16, 1024, SWIZZLE_128B);
                                mma_ts_f16(d_tmem, a_tmem + 32 * sub + 8 * kk, bl, idesc, 1u);
                            }
                        }
                    }
                    tc_commit(empty(s));
                    tc_commit(aempty(a));
                    if ((dbg & 16) && blockIdx.x == 0 && ait < 96) B200Q_TRACE(3, ait);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                tc_commit(dfull(ab));
                ++li;
            });
        }
    } else if (warp >= 4 && warp < 12) {
        // =============================================================== dequant warps
        // two groups of four warps (one warp per TMEM lane quarter) take alternate k-blocks: one group
        // alone (LDS -> LOP3 -> tcgen05.st -> wait::st -> arrive, ~300 clk per k-block) is barely faster
        // than the 384..512 clk the tensor core needs for the same k-block
        const int q = warp & 3, grp = (warp - 4) >> 2;
        const int r = 32 * q + lane;                                // weight row of the tile = TMEM lane
        int s = 0, ph = 0, ait = 0;
        for_each_item(p, KB, total_tiles, [&](const Item& w) {
            for (int kb = w.kb0; kb < w.kb1; ++kb, ++ait) {
                if ((ait & 1) != grp) {                             // the other group's k-block
                    // Still observe its fill.  With an ODD number of stages (fp32 activations on 256-token tiles: three 68 KB
                    // stages) a group meets a stage only every other phase, and a parity wait that is two phases behind
                    // the barrier returns at once: the group then read a stage before its data had landed -- stream-K with
                    // tiles cut between ~10 CTAs (short items, the groups run far ahead) gave 0.3 relative error on
                    // M = 4096, K = 8192, N = 128 fp32 (tools/sk_check.py, tests/test_gpu_linear.py::test_stream_k_few_long_tiles)
                    mbar_wait(full(s), ph);
                    if (++s == S) { s = 0; ph ^= 1; }
                    continue;
                }
                mbar_wait(full(s), ph);
                const bool tr = (dbg & 16) && blockIdx.x == 0 && q == 0 && lane == 0 && ait < 96;
                if (tr) B200Q_TRACE(4, ait);
                uint4 wv[2 * KSUB];
#pragma unroll
                for (int i = 0; i < 2 * KSUB; ++i) wv[i] = lds128(w_smem(s) + r * (KSUB * BK / 2) + 16 * i);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(s));               // the packed bytes are in registers
                const int a = ait % A_SLOTS;
                if (ait >= A_SLOTS) {
                    mbar_wait(aempty(a), ((ait / A_SLOTS) - 1) & 1);
                    tc_fence_after_sync();
                }
                if (tr) B200Q_TRACE(5, ait);
                const uint32_t dst = tmem + ((uint32_t)(32 * q) << 16) + A_BASE + A_COLS * a;
#pragma unroll
                for (int half = 0; half < 2 * KSUB; ++half) {
                    const uint32_t ws[4] = {wv[half].x, wv[half].y, wv[half].z, wv[half].w};
                    uint32_t rr[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // nibble n left in the low mantissa bits of a zero-exponent fp16 is the subnormal
                        // n * 2^-24 (bits 0-3) or n * 2^-20 (bits 4-7): exact, no arithmetic
                        const uint32_t w = ws[j], w2 = w >> 8;
                        rr[4 * j + 0] = w & 0x000f000fu;
                        rr[4 * j + 1] = w & 0x00f000f0u;
                        rr[4 * j + 2] = w2 & 0x000f000fu;
                        rr[4 * j + 3] = w2 & 0x00f000f0u;
                    }
                    if (!(dbg & 8)) tmem_st16(dst + 16 * half, rr);     // ablation 8: no tcgen05.st
                }
                if (!(dbg & 8)) tmem_wait_st();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(afull(a));
                if (tr) B200Q_TRACE(6, ait);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        });
    } else if (warp >= 12) {
        // =============================================================== epilogue warps
        const int q = warp & 3;
        const int et = (warp - 12) * 32 + lane;                     // 0..127
        int li = 0;
        for_each_item(p, KB, total_tiles, [&](const Item& w) {
            const TileInfo& ti = w.ti;
            const int ab = li % DBUF;
            if (w.publish) {
                // ---- stream-K: hand the raw partial accumulator to the CTA that finishes this tile
                float* dst = p.part + (size_t)blockIdx.x * (BM * BN) + 32 * q + lane;
                mbar_wait(dfull(ab), (li / DBUF) & 1);
                tc_fence_after_sync();
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t d[32];
                    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + ab * BN + 32 * c, d);
                    tmem_wait_ld();
                    if (c == BN / 32 - 1) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(dempty(ab));
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(dst + (32 * c + j) * BM, __uint_as_float(d[j]));
                }
                __threadfence();
                named_bar_sync(1, 128);
                if (et == 0) {
                    __threadfence();
                    atomicExch(p.flags + blockIdx.x, 1u);
                }
                ++li;
                return;
            }
            // per-token scalars of this tile (written after the previous tile's readers are done)
            named_bar_sync(1, 128);
            for (int j = et; j < BN; j += 128) {
                const int m = ti.m0 + j;
                tok[j] = m < ti.mend ? __ldg(p.descale + m) : 0.0f;
                tok[BN + j] = m < ti.mend ? __ldg(p.rowsum + m) : 0.0f;
            }
            if (w.nadd > 0 && et < w.nadd) {
                // ---- stream-K: wait (bounded) for the partials of the CTAs that hold the head of this tile
                // (all CTAs are resident and every CTA publishes before it waits, so this cannot dead-lock; a partial
                // that does not show up within ~1 s means a broken launch: trap -> sticky CUDA error, never a wrong sum)
                volatile unsigned int* f = p.flags + w.c0 + et;
                unsigned spin = 0;
                while (*f == 0u) {
                    __nanosleep(64);
                    if (++spin > (1u << 24)) __trap();
                }
                __threadfence();
            }
            named_bar_sync(1, 128);
            const int n = ti.n0 + 32 * q + lane;
            const bool n_ok = n < p.N;
            const float sc = n_ok ? __ldg(p.scales + (int64_t)ti.e * p.N + n) : 0.0f;
            const float zp = n_ok ? __ldg(p.zps + (int64_t)ti.e * p.N + n) : 0.0f;
            mbar_wait(dfull(ab), (li / DBUF) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t d[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + ab * BN + 32 * c, d);
                tmem_wait_ld();
                if (c == BN / 32 - 1) {                             // accumulator fully read: release it
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dempty(ab));
                }
                for (int a = 0; a < w.nadd; ++a) {                  // fixed order: deterministic
                    const float* src = p.part + (size_t)(w.c0 + a) * (BM * BN) + 32 * q + lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = __float_as_uint(__uint_as_float(d[j]) + __ldcg(src + (32 * c + j) * BM));
                }
                if (p.gated) {
                    // fused SiLU-gate: lanes 2i (gate row) and 2i + 1 (up row) hold the two projections of h column
                    // ti.n0 / 2 + 16 q + i.  Two tokens per step: the even lane finishes token j (it fetches the up value from
                    // its neighbour), the odd lane token j + 1 (it fetches the gate value) -- one shuffle, one SiLU and one
                    // store per lane and token PAIR (every lane does useful work; the first form spent the exp / divide of
                    // the odd lanes on nothing and made the first GEMM of a MoE layer epilogue-bound: 8.4 vs 7.0 ms)
                    const int64_t F = p.N >> 1;
                    const bool odd = (lane & 1) != 0;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float v0 = sc * (__uint_as_float(d[j]) * tok[32 * c + j] - zp * tok[BN + 32 * c + j]);
                        const float v1 = sc * (__uint_as_float(d[j + 1]) * tok[32 * c + j + 1] - zp * tok[BN + 32 * c + j + 1]);
                        const float other = __shfl_xor_sync(0xffffffffu, odd ? v0 : v1, 1);
                        const float v = odd ? other : v0, u = odd ? v1 : other;       // gate, up of this lane's token
                        const int m = ti.m0 + 32 * c + j + (odd ? 1 : 0);
                        if (n_ok && m < ti.mend) {
                            const float hv = v / (1.0f + __expf(-v)) * u;
                            const int64_t o = (int64_t)m * F + (n >> 1);
                            if (p.y_dtype == B200Q_F32) static_cast<float*>(p.y)[o] = hv;
                            else if (p.y_dtype == B200Q_F16) static_cast<__half*>(p.y)[o] = __float2half_rn(hv);
                            else static_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(hv);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int m = ti.m0 + 32 * c + j;
                        if (n_ok && m < ti.mend) {
                            const float v = sc * (__uint_as_float(d[j]) * tok[32 * c + j] - zp * tok[BN + 32 * c + j]);
                            const int64_t o = (int64_t)m * p.N + n;
                            if (p.y_dtype == B200Q_F32) static_cast<float*>(p.y)[o] = v;
                            else if (p.y_dtype == B200Q_F16) static_cast<__half*>(p.y)[o] = __float2half_rn(v);
                            else static_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(v);
                        }
                    }
                }
            }
            if (w.nadd > 0) {                                       // leave the flags zeroed for the next launch
                named_bar_sync(1, 128);
                if (et < w.nadd) p.flags[w.c0 + et] = 0u;
            }
            ++li;
        });
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// Activation preparation: one warp per token row (see the header comment).
struct XprepGemmParams {
    const void* x;
    __half* xh;
    __half* xl;       // nullptr for 16-bit inputs
    float* descale;
    float* rowsum;
    int* nf;          // [R] 1: the row holds NaN / Inf (recomputed by the fix-up pass)
    int x_dtype, R, K;
};

template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// eight consecutive values of a row as they come out of memory
template <typename T> struct Raw8 {
    uint4 r;                                                 // 16-bit types
    __device__ __forceinline__ void load(const T* p) { r = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
    // (the compiler would otherwise unpack once and keep the floats alive across the block reduction)
    __device__ __forceinline__ void opaque() { asm volatile("" : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f;
            if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
    __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
    __device__ __forceinline__ void opaque() {}
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T>
__global__ void __launch_bounds__(256) xprep_gemm_kernel(const XprepGemmParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= p.R) return;
    const T* xr = static_cast<const T*>(p.x) + m * p.K;
    // amax through the bit patterns (non-negative floats order like unsigned integers; NaN / Inf sort above
    // every finite value instead of being dropped by fmaxf)
    unsigned int ub = 0u;
    float sum = 0.0f;
    for (int k = lane * 8; k < p.K; k += 256) {
        float v[8];
        ld8<T>(xr + k, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { ub = max(ub, __float_as_uint(v[i]) & 0x7fffffffu); sum += v[i]; }
    }
    ub = __reduce_max_sync(0xffffffffu, ub);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    int ex = 0;
    if (ub > 0u && ub < 0x7f800000u) ex = max(-100, min(100, 140 - (int)(ub >> 23)));
    if (lane == 0) p.nf[m] = ub >= 0x7f800000u ? 1 : 0;
    const float up = __uint_as_float((uint32_t)(127 + ex) << 23), up_hi = up * 0.0625f;
    if (lane == 0) {
        p.descale[m] = __uint_as_float((uint32_t)(127 + 24 - ex) << 23);
        p.rowsum[m] = sum;
    }
    __half* hr = p.xh + m * p.K;
    __half* lr = p.xl ? p.xl + m * p.K : nullptr;
    for (int k = lane * 8; k < p.K; k += 256) {
        float v[8];
        ld8<T>(xr + k, v);
        // (k0,k4) (k1,k5) (k2,k6) (k3,k7): the order the A registers hold the nibbles in
        const float s0 = v[0] * up, s4 = v[4] * up, s1 = v[1] * up_hi, s5 = v[5] * up_hi;
        const float s2 = v[2] * up, s6 = v[6] * up, s3 = v[3] * up_hi, s7 = v[7] * up_hi;
        __half2 h0 = __floats2half2_rn(s0, s4), h1 = __floats2half2_rn(s1, s5);
        __half2 h2 = __floats2half2_rn(s2, s6), h3 = __floats2half2_rn(s3, s7);
        *reinterpret_cast<uint4*>(hr + k) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        if (lr) {
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
            __half2 l0 = __floats2half2_rn(s0 - f0.x, s4 - f0.y), l1 = __floats2half2_rn(s1 - f1.x, s5 - f1.y);
            __half2 l2 = __floats2half2_rn(s2 - f2.x, s6 - f2.y), l3 = __floats2half2_rn(s3 - f3.x, s7 - f3.y);
            *reinterpret_cast<uint4*>(lr + k) = make_uint4(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1),
                                                           *reinterpret_cast<uint32_t*>(&l2), *reinterpret_cast<uint32_t*>(&l3));
        }
    }
}

// Same preparation, 128 threads per token row (two rows per block), K <= NCH * 1024: every thread issues all its
// loads up front and keeps the values in registers, so a row costs ONE trip to memory plus a block reduction
// (the warp-per-row kernel above walks the row twice with one load in flight: 8 us for 256 rows).
template <typename T, int NCH>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 2 : 1) xprep_gemm_rows_kernel(const XprepGemmParams p) {
    __shared__ unsigned int s_am[2][4];
    __shared__ float s_sum[2][4];
    pdl_launch_dependents();                        // the GEMM's CTAs may set themselves up meanwhile
    const int rt = threadIdx.x & 127, rw = threadIdx.x >> 7, lane = threadIdx.x & 31, w4 = (threadIdx.x >> 5) & 3;
    const int64_t m = (int64_t)blockIdx.x * 2 + rw;
    const bool row_ok = m < p.R;
    const T* xr = static_cast<const T*>(p.x) + (row_ok ? m : 0) * p.K;
    // the row stays in registers AS LOADED (16-bit inputs: 4 registers per 8 values, unpacked once for the amax / sum and once
    // for the output): K = 14336 as floats was 149 registers per thread = one block per SM, 3.7 TB/s on the 1.8 GB of a MoE step
    Raw8<T> raw[NCH];
    unsigned int ub = 0u;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int k = i * 1024 + rt * 8;
        if (row_ok && k < p.K) raw[i].load(xr + k);
        else raw[i].zero();
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        float v[8];
        raw[i].unpack(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { ub = max(ub, __float_as_uint(v[j]) & 0x7fffffffu); sum += v[j]; }
    }
    ub = __reduce_max_sync(0xffffffffu, ub);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) { s_am[rw][w4] = ub; s_sum[rw][w4] = sum; }
    __syncthreads();
    ub = max(max(s_am[rw][0], s_am[rw][1]), max(s_am[rw][2], s_am[rw][3]));
    sum = (s_sum[rw][0] + s_sum[rw][1]) + (s_sum[rw][2] + s_sum[rw][3]);
    if (!row_ok) return;
    int ex = 0;
    if (ub > 0u && ub < 0x7f800000u) ex = max(-100, min(100, 140 - (int)(ub >> 23)));
    if (rt == 0) p.nf[m] = ub >= 0x7f800000u ? 1 : 0;
    const float up = __uint_as_float((uint32_t)(127 + ex) << 23), up_hi = up * 0.0625f;
    if (rt == 0) {
        p.descale[m] = __uint_as_float((uint32_t)(127 + 24 - ex) << 23);
        p.rowsum[m] = sum;
    }
    __half* hr = p.xh + m * p.K;
    __half* lr = p.xl ? p.xl + m * p.K : nullptr;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int k = i * 1024 + rt * 8;
        if (k >= p.K) continue;
        float v[8];
        raw[i].opaque();
        raw[i].unpack(v);
        // (k0,k4) (k1,k5) (k2,k6) (k3,k7): the order the A registers hold the nibbles in
        const float s0 = v[0] * up, s4 = v[4] * up, s1 = v[1] * up_hi, s5 = v[5] * up_hi;
        const float s2 = v[2] * up, s6 = v[6] * up, s3 = v[3] * up_hi, s7 = v[7] * up_hi;
        __half2 h0 = __floats2half2_rn(s0, s4), h1 = __floats2half2_rn(s1, s5);
        __half2 h2 = __floats2half2_rn(s2, s6), h3 = __floats2half2_rn(s3, s7);
        *reinterpret_cast<uint4*>(hr + k) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        if (lr) {
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
            __half2 l0 = __floats2half2_rn(s0 - f0.x, s4 - f0.y), l1 = __floats2half2_rn(s1 - f1.x, s5 - f1.y);
            __half2 l2 = __floats2half2_rn(s2 - f2.x, s6 - f2.y), l3 = __floats2half2_rn(s3 - f3.x, s7 - f3.y);
            *reinterpret_cast<uint4*>(lr + k) = make_uint4(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1),
                                                           *reinterpret_cast<uint32_t*>(&l2), *reinterpret_cast<uint32_t*>(&l3));
        }
    }
}

template <typename T>
int launch_xprep(const XprepGemmParams& xp, int64_t M, int64_t K, cudaStream_t st) {
    const unsigned blocks2 = (unsigned)((M + 1) / 2);
    if (K <= 4096) xprep_gemm_rows_kernel<T, 4><<<blocks2, 256, 0, st>>>(xp);
    else if (K <= 8192) xprep_gemm_rows_kernel<T, 8><<<blocks2, 256, 0, st>>>(xp);
    else if (K <= 16384) xprep_gemm_rows_kernel<T, 16><<<blocks2, 256, 0, st>>>(xp);
    else xprep_gemm_kernel<T><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(xp);
    return check_cuda(cudaGetLastError(), "xprep launch");
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, int esz, const void* base, uint64_t inner, uint64_t outer,
                uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(B200Q_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * (uint64_t)esz};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200Q_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int x_dtype, int y_dtype) {
    (void)x_dtype; (void)y_dtype;
    return M >= 1 && K >= 128 && K % 128 == 0 && N >= 1 && M < (1ll << 30) && N < (1ll << 30) && K <= (1 << 20);
}

// workspace: [4 KB reserved: the decode kernels keep their zeroed ticket counters there -- callers
// share one workspace between all paths][xh R*K fp16][xl R*K fp16][descale R f32][rowsum R f32]
constexpr size_t WS_RESERVED = 4096;
// stream-K partial accumulators: one [256][128] fp32 slot per CTA (the flags live in the reserved 4 KB)
constexpr int SK_MAX_CTAS = 160;
constexpr size_t SK_PART_BYTES = (size_t)SK_MAX_CTAS * 256 * BM * 4;
size_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t K) {
    if (!gemm_tc_supported(M, N, K, B200Q_F32, B200Q_F32)) return 0;
    return WS_RESERVED + 2 * align_up((size_t)M * K * 2, 1024) + 3 * align_up((size_t)M * 4, 1024) + 1024 + SK_PART_BYTES;
}

int launch_gemm_tc(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed,
                   const float* scales, const float* zps, void* y, int y_dtype, int64_t M,
                   int64_t N, int64_t K, const int32_t* starts, const int32_t* ends, int E,
                   void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st, int gated, const int32_t* emap,
                   int n_wexperts) {
    (void)flags;
    if (!gemm_tc_supported(M, N, K, x_dtype, y_dtype)) return set_error(B200Q_EINVAL, "gemm_tc: unsupported shape");
    if (gated && (N & 1)) return set_error(B200Q_EINVAL, "gemm_tc: the gated epilogue needs an even number of weight rows");
    const size_t need = gemm_tc_ws_bytes(M, N, K);
    if (!ws || ws_bytes < need) return set_error(B200Q_EWORKSPACE, "gemm_tc: workspace too small (%zu < %zu)", ws_bytes, need);
    // the tensor maps want 1024-byte aligned bases: align inside the (1 KB larger) workspace
    uint8_t* w8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + WS_RESERVED + 1023) & ~(uintptr_t)1023);
    const size_t xbytes = align_up((size_t)M * K * 2, 1024), sbytes = align_up((size_t)M * 4, 1024);
    __half* xh = reinterpret_cast<__half*>(w8);
    __half* xl = reinterpret_cast<__half*>(w8 + xbytes);
    float* descale = reinterpret_cast<float*>(w8 + 2 * xbytes);
    float* rowsum = reinterpret_cast<float*>(w8 + 2 * xbytes + sbytes);
    int* nf = reinterpret_cast<int*>(w8 + 2 * xbytes + 2 * sbytes);
    const int parts = x_dtype == B200Q_F32 ? 2 : 1;

    XprepGemmParams xp{x, xh, parts == 2 ? xl : nullptr, descale, rowsum, nf, x_dtype, (int)M, (int)K};
    if (x_dtype == B200Q_F32) { if (int rc = launch_xprep<float>(xp, M, K, st)) return rc; }
    else if (x_dtype == B200Q_F16) { if (int rc = launch_xprep<__half>(xp, M, K, st)) return rc; }
    else { if (int rc = launch_xprep<__nv_bfloat16>(xp, M, K, st)) return rc; }

    CUtensorMap map_xh, map_xl, map_w;
    // Stages of two 64-column sub-blocks (template parameter KSUB = 2) were tried to halve the per-stage
    // hand-shakes: +3 % at 256-token tiles, -5 % at 192 (one accumulator left), and a hang at M = 32768 that was not
    // tracked down -- not instantiated.
    const int ksub = 1;
    // token-tile height: fewest (waves x time per wave).  Measured per-wave time (tools/sweep_gemm_bn.py):
    // BN = 192 (two accumulators, epilogue overlapped, 4 A slots) takes 0.74 of a BN = 256 (8 A slots) wave; BN = 128 is never
    // better (64-clk MMAs run into the per-instruction floor) and is only reachable through gemm_bn.
    const int groups_n = starts ? E : 1;
    const int n_tiles_h = (int)((N + BM - 1) / BM);
    int bn = 256;
    {
        double best = 1e30;
        const int cand[2] = {256, 192};
        for (int i = 0; i < 2; ++i) {
            const long long mt = (M + cand[i] - 1) / cand[i] + (starts ? groups_n / 2 : 0);
            const long long tiles = mt * n_tiles_h;
            const long long waves = (tiles + dev.sm_count - 1) / dev.sm_count;
            const double cost = (double)waves * (cand[i] == 256 ? 1.0 : 0.74);
            if (cost < best) { best = cost; bn = cand[i]; }
        }
        // few rows in total (decode-sized batches, M = 9..64, and MoE layers with a handful of tokens): a 256-token
        // tile would spend 179 clk per MMA on empty columns; with 32 / 64 columns an MMA costs the 64 clk it takes to
        // feed the A operand from tensor memory, i.e. the kernel runs at the weight-streaming rate of that floor
        if (M <= 32) bn = 32;
        else if (M <= 64) bn = 64;
        const int fb = tuning().gemm_bn;
        if (fb == 32 || fb == 64 || fb == 128 || fb == 192 || fb == 256) bn = fb;
    }
    // stream-K instead of whole tiles when the only wave is less than half full (plain linear): 256-token tiles,
    // the k-blocks of all tiles dealt out evenly over the SMs.  Measured (tools/run_gemm_sk.py, bench_gemm.py):
    // 11008 -> 4096 (32 weight-row tiles) M = 256 +28 %, M = 512 +17 %; with fuller waves the two extra
    // un-overlapped accumulator drains per CTA cost more than the balance gains (4096 -> 11008 M = 512: -37 %).
    int sk = 0, skq = 0, skr = 0, sk_grid = 0;
    if (!starts && !gated && tuning().gemm_sk != 0) {
        const int fbs = tuning().gemm_bn;
        const int bnsk = (fbs == 128 || fbs == 192 || fbs == 256) ? fbs : (M <= 64 ? bn : 256);   // small batches keep their small token tile
        const long long tiles256 = ((M + bnsk - 1) / bnsk) * n_tiles_h;
        const long long waves = (tiles256 + dev.sm_count - 1) / dev.sm_count;
        const double fill = (double)tiles256 / (double)(waves * dev.sm_count);
        const long long total_kb = tiles256 * (K / (BK * ksub));
        const int g = dev.sm_count < SK_MAX_CTAS ? dev.sm_count : SK_MAX_CTAS;
        // (small batches, 32 / 64-token tiles: the partial accumulators are 16 - 32 KB, so stream-K pays up to a fill of
        // 0.8 -- 4096 -> 11008 (86 tiles on 148 SMs) M = 17..32: 37.8 -> 29.4 us, M = 64: 39.4 -> 34.2 us)
        const bool want = waves == 1 && (fill < 0.5 || (M <= 64 && fill < 0.8));
        if ((tuning().gemm_sk > 0 || want) && total_kb / g >= 8 && tiles256 <= 100000) {
            sk = 1; bn = bnsk; sk_grid = g;
            skq = (int)(total_kb / g); skr = (int)(total_kb % g);
        }
    }
    if (int rc = make_map_2d(&map_xh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, xh, (uint64_t)K, (uint64_t)M, BK, bn, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_2d(&map_xl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, parts == 2 ? xl : xh, (uint64_t)K, (uint64_t)M, BK, bn, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int groups = starts ? E : 1;
    const int wgroups = (starts && emap) ? n_wexperts : groups;           // experts in the weight tensor
    if (int rc = make_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, packed, (uint64_t)(K / 2), (uint64_t)wgroups * N, ksub * BK / 2, BM, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;

    GemmParams p{};
    p.scales = scales; p.zps = zps; p.y = y; p.descale = descale; p.rowsum = rowsum;
    p.starts = starts; p.ends = ends; p.emap = starts ? emap : nullptr; p.E = groups; p.y_dtype = y_dtype;
    p.R = (int)M; p.N = (int)N; p.K = (int)K;
    p.bn = bn;
    p.mt_bound = (int)((M + bn - 1) / bn) + (starts ? E : 0);
    p.n_tiles = n_tiles_h;
    p.stage_bytes = ksub * (parts * bn * BK * 2 + W_TILE_BYTES);
    int stages = (dev.max_smem_optin - OFF_STAGES) / p.stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return set_error(B200Q_EINVAL, "gemm_tc: not enough shared memory");
    p.stages = stages;
    p.sk = sk; p.skq = skq; p.skr = skr;
    p.debug = tuning().gemm_debug > 0 ? tuning().gemm_debug : 0;
    p.gated = gated;
    // tile order.  Weight-tile-major (0) makes every weight tile stream ALL activation rows again: fine while they stay in L2
    // (dense prefill: M K 2 bytes), ruinous for a MoE layer (Mixtral, 16384 tokens: 268 MB of x 224 times, 940 MB of h 32
    // times -- the layer was DRAM-bound: 17.3 ms).  Token-tile-major (1): the CTAs of a wave share a few activation tiles
    // and stream the expert's weights, which fit in L2 (13.97 ms; tools/tile_order.py).  Dense shapes gain <= 8 % and only
    // when activations outgrow L2 while (wave's activation tiles + all weights) still fit: they keep order 0.
    p.mt_major = tuning().gemm_mt_major >= 0 ? (tuning().gemm_mt_major != 0) : (starts ? 1 : 0);
    p.part = reinterpret_cast<float*>(w8 + 2 * xbytes + 3 * sbytes);
    p.flags = reinterpret_cast<unsigned int*>(ws);
    const size_t smem = (size_t)OFF_STAGES + (size_t)stages * p.stage_bytes;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
    const int bi = bn == 32 ? 0 : (bn == 64 ? 1 : (bn == 128 ? 2 : (bn == 192 ? 3 : 4)));
    const KernelFn table[2][5] = {
        {gemm_tc_kernel<1, 32, 1>, gemm_tc_kernel<1, 64, 1>, gemm_tc_kernel<1, 128, 1>, gemm_tc_kernel<1, 192, 1>, gemm_tc_kernel<1, 256, 1>},
        {gemm_tc_kernel<2, 32, 1>, gemm_tc_kernel<2, 64, 1>, gemm_tc_kernel<2, 128, 1>, gemm_tc_kernel<2, 192, 1>, gemm_tc_kernel<2, 256, 1>}};
    const int ti_ = parts - 1;
    KernelFn kfn = table[ti_][bi];
    static thread_local int attr_done[64][2][5] = {{{0}}};
    int devi = 0;
    B200Q_CUDA(cudaGetDevice(&devi));
    if (devi >= 0 && devi < 64 && attr_done[devi][ti_][bi] < (int)smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[devi][ti_][bi] = (int)smem;
    }
    const long long total_tiles = (long long)p.n_tiles * p.mt_bound;
    int grid = dev.sm_count;
    if (total_tiles < grid) grid = (int)total_tiles;
    if (grid < 1) grid = 1;
    if (sk) grid = sk_grid;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    if (int rc = check_cuda(cudaLaunchKernelEx(&cfg, kfn, map_xh, map_xl, map_w, p), "gemm_tc launch")) return rc;
    // token rows with NaN / Inf (flagged by the preparation kernel): recomputed in the reference's order
    return launch_nonfinite_fixup(x, x_dtype, packed, scales, zps, nf, y, y_dtype, M, N, K, starts, ends, groups, gated, st,
                                  starts ? emap : nullptr);
}

}  // namespace b200q

#ifdef B200Q_PROF
/* bench-only: per-k-block SM-clock trace of CTA 0 (gemm_debug bit 4), 8 x 96 int64 */
extern "C" int b200q_debug_gemm_trace(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_gemm_trace, sizeof(long long) * 8 * 96), "read gemm trace");
}
#endif
