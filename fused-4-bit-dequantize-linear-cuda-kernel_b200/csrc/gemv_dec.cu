// Decode path (M = 1..16): y[M,N] = x[M,K] @ dequant(W)^T (+ bias) with the CTA's whole share of W resident in
// shared memory.  HBM-bound by design: every packed byte is read once, by the TMA engine.
//
// Arithmetic: exact integers (unchanged from round 1).  x is a per-row fixed-point number (even columns
// Xe = round(x 2^e), odd columns Xo = round(x 2^(e-4))) cut into four signed base-256 digits = four columns of IMMA
// m16n8k32 (u8 x s8 -> s32) per batch row.  The nibbles are never widened: the raw packed byte q_lo + 16 q_hi meets
// the digits of Xe, the masked byte 16 q_hi meets the digits of Z = Xo - Xe.  s32 partial sums, s64 when the digits
// are combined, y = s * 2^-e * (sum q X - zp * sum X) with one fp32 rounding: results do not depend on the
// summation order.  A batch row that contains NaN / Inf is recomputed in the reference's order (w = (q - zp) * s,
// fp32 FMA), so non-finite inputs propagate exactly like dequantize + F.linear (python/quantize.py:172, 202).
//
// Structure:
//   * weights: tile i = 16 rows x K/2 bytes arrives as 3-D TMA tensor boxes [16 rows][chunk pairs][128 bytes] (pair =
//     256 columns) with the 128-byte swizzle, one box and one mbarrier per (tile, pair group): shared memory holds
//     [pair][row][128 B], which ldmatrix.x4 turns into IMMA A fragments without bank conflicts whatever K is.  K only
//     has to be a multiple of 256; no K split, no repacked copy of the weights;
//   * warp w owns the pairs w, w + 16, ... of every tile: it loads, converts and keeps the B fragments of exactly those
//     columns (exchange through its own 2 KB of shared memory, __syncwarp only); the one block barrier of a pass is the
//     row amax;
//   * sum X for the zero-point term: a row of bytes 0x11 behind the CTA's last weight row makes the tensor cores
//     deliver it as one more output row;
//   * cross-warp reduction, one pass of <= 2 batch rows: every (tile, warp) parks its 16 x 8 partial tile in its own
//     slot, folded once after the loop (integer adds: exact, order independent);
//     otherwise pipelined: partial tiles go through a double buffer (mbarrier full / empty) and all 512 threads fold
//     tile i - 1 into the accumulator with red.shared.add while the tensor cores work on tile i + 1;
//   * M = 3..16: passes of up to four batch rows (two n-tiles) over the resident tiles -- the weights are read from
//     HBM once whatever M is;
//   * GEN instances (template flag): ring of tile buffers for CTAs whose rows do not fit (a buffer is requested again
//     once all warps are done with it; outputs window by window), grouped mode over experts with device-side row
//     offsets and a token map (MoE decode), gated epilogue silu(gate) * up.  The plain instances carry none of it:
//     code that runs once per launch costs 5 - 12 clk per instruction and warp here (profiles/r02_decode_notes.md);
//   * no integer division in the kernel: geometry comes from the host.
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output, M x weight traffic).
#include <cuda.h>
#include <cmath>
#include <mutex>
#include "internal.h"
#include "ptx.cuh"
#include "tc.cuh"
#include "dec.cuh"

namespace b200q {

// bench-only (-DB200Q_PROF build, tools/prof_dec.py): per-CTA wall-clock stamps of the phases of the last launch
#ifdef B200Q_PROF
__device__ long long g_dec_prof[256 * 16];
#define B200Q_STAMP(i)                                                               \
    do {                                                                             \
        if (p.debug && tid == 0 && blockIdx.x < 256) {                               \
            long long t_;                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                   \
            g_dec_prof[blockIdx.x * 16 + (i)] = t_;                                  \
        }                                                                            \
    } while (0)
#define B200Q_ABL(bit) ((p.debug & (bit)) != 0)     // 2: no IMMA, 4: no weight traffic / waits, 8: no weight LDS, 16: no hand-over / reduce
#else
#define B200Q_STAMP(i) ((void)0)
#define B200Q_ABL(bit) false
#endif

namespace {

constexpr int NW = 16;               // warps per CTA
constexpr int NTHR = NW * 32;
constexpr int TILE_ROWS = 16;
constexpr int PAIR_BYTES = TILE_ROWS * 128;      // one pair (256 columns) of one tile in shared memory
constexpr int MAX_BARS = 64;                     // (tile, pair group) barriers

// shared memory map (bytes)
constexpr int OFF_BARS = 0;          // [MAX_BARS] tile barriers
constexpr int OFF_FULL = 512;        // [2]
constexpr int OFF_EMPTY = 528;       // [2]
constexpr int OFF_FLAG = 544;        // bit m: batch row m holds NaN / Inf
constexpr int OFF_SUMX = 576;        // [16 batch rows][4 digits] s32: sum_k X of every batch row (kept for the later windows)
constexpr int OFF_EX = 832;          // [16] exponent of every batch row
constexpr int OFF_AMAX = 1024;       // [2 (pass parity)][4 rows][16 warps] u32
constexpr int OFF_CONS = 1536;       // [MAX_BARS] "consumed" barriers of the tile buffers (ring mode)
constexpr int OFF_DYN = 2048;        // partial-tile buffers (their owner's slots double as its operand exchange space), accumulator, tiles
constexpr int WARP_RED = 1024;       // bytes of a warp in one partial-tile buffer (pipelined reduction)

struct DecParams {
    const void* x;
    const uint8_t* packed;
    const float* scales;
    const float* zps;
    const float* bias;               // may be null
    void* y;
    const uint8_t* next_packed;      // L2 prefetch hint (weights of the next fused linear), may be null
    unsigned long long next_bytes;
    unsigned int next_chunk;         // next_bytes / gridDim.x
    int x_dtype, y_dtype;
    int M, N, K;
    int rows_q, rows_rem;            // CTA b owns rows_q (+1 if b < rows_rem) row units (host-computed: no division in the kernel)
    int npairs;                      // ceil(K / 256)
    int nbars;                       // pair groups (barriers) per tile
    int chunk;                       // pairs per group
    int ntiles_max;                  // tile buffers W = tiles per window
    int tile_bytes;
    int tile_off;                    // byte offset of tile 0 in dynamic shared memory (1024-aligned)
    int red_off, fin_off;            // partial-tile buffers, accumulator
    int slots;                       // 1 (single pass, M <= 2): one slot per (tile, warp) at red_off, folded after the loop
    int npasses;
    int wait_weights;                // 1: weights may be written by the preceding kernel
    int early_tiles;                 // a + 10 b: a tiles requested before griddepcontrol.wait, b more behind the x loads
    int pf_mode;                     // next-layer L2 prefetch: 1 behind the last own request, 2 before the own requests, 3 after the operand build
    int gated;                       // 1: rows 2f / 2f+1 are the gate / up projection of column f; y is h [M, N/2] = silu(gate) * up
    // work items: blockIdx.x = (expert e, CTA b of the expert).  A CTA whose rows do not fit in its W tile buffers
    // (Mixtral's 14336-wide projections) runs them as a ring: a buffer is requested again as soon as all warps are
    // done with it, and the outputs are produced window by window (W tiles each).
    // (GEN instances only; the plain instances -- one expert, every tile resident -- carry none of this code)
    int win_hi[2], win_lo[2];        // {R windows, T0 tiles of the first one} for CTAs with rows_q + 1 / rows_q row units
    const int32_t* offsets;          // grouped (MoE decode): rows [offsets[e], offsets[e+1]) of x / y belong to expert e of
                                     //   packed [E, N, K/2] (device memory; experts without rows exit at once); else nullptr
    const int32_t* row_map;          // grouped: batch row r of the group reads x[row_map[offsets[e] + r]] (the token of that
                                     //   sorted position: x is read in place, no gathered copy); y rows stay in sorted order
    int n_groups;                    // grouped: experts in the weight tensor; gridDim.y < n_groups: blockIdx.y is the RANK of the
                                     // CTA's expert among the experts that have rows (a decode step hits few experts: the CTAs of
                                     // the others would each take their turn on an SM -- shared memory for one CTA -- only to exit)
    int debug;                       // B200Q_PROF builds: record phase stamps
};

// weight requests [from, to) (request = (tile, pair group), ~32 KB): one 3-D box [16 rows][chunk pairs][128 B] each,
// one elected thread.  Not inlined: three call sites (before griddepcontrol.wait, behind the x loads, after the operand
// build), and code that runs once costs ~10 clk per instruction whatever it does (profiles/r02_decode_notes.md).
__device__ __noinline__ void dec_issue_tiles(const CUtensorMap* tmap, uint32_t bar0, uint32_t dst0, int row0, int from, int to,
                                             int nbars, int chunk, int tile_bytes) {
    const uint64_t pol = policy_evict_first();
    int i = nbars == 1 ? from : from / nbars, grp = from - i * nbars;
    for (int op = from; op < to; ++op) {                      // op = tile * nbars + pair group
        const uint32_t bar = bar0 + 8u * (uint32_t)op;
        mbar_arrive_expect_tx(bar, (uint32_t)(chunk * PAIR_BYTES));
        tma_box_3d(dst0 + (uint32_t)(i * tile_bytes + grp * chunk * PAIR_BYTES), tmap, 0, row0 + i * TILE_ROWS, grp * chunk, bar, pol);
        if (++grp == nbars) { grp = 0; ++i; }
    }
}

// Tile order of the GEN instances.  The CTA's positions (weight rows + one more: the row of 0x11 bytes) form S tiles
// of 16, grouped into R windows of W tiles (W = tile buffers).  The windows are processed from the LAST one (partial,
// T0 tiles; it holds the 0x11 row, whose result every epilogue needs) to the first; s counts tiles in processing order
// and tile s lives in buffer s % W.
__device__ __forceinline__ int ring_row_tile(int s, int W, int T0, int R) {      // tile index in row order, s <= W (windows 0 and 1)
    return s < T0 ? (R - 1) * W + s : (R - 2) * W + (s - T0);
}
// first fill of the ring: tiles [0, min(S, W)) -- buffer = tile
__device__ __noinline__ void dec_issue_ring(const CUtensorMap* tmap, uint32_t bar0, uint32_t dst0, int row0, int from, int to,
                                            int nbars, int chunk, int tile_bytes, int W, int T0, int R) {
    const uint64_t pol = policy_evict_first();
    int s = nbars == 1 ? from : from / nbars, grp = from - s * nbars;
    for (int op = from; op < to; ++op) {
        const uint32_t bar = bar0 + 8u * (uint32_t)op;
        mbar_arrive_expect_tx(bar, (uint32_t)(chunk * PAIR_BYTES));
        tma_box_3d(dst0 + (uint32_t)(s * tile_bytes + grp * chunk * PAIR_BYTES), tmap, 0, row0 + ring_row_tile(s, W, T0, R) * TILE_ROWS,
                   grp * chunk, bar, pol);
        if (++grp == nbars) { grp = 0; ++s; }
    }
}

// GPW2: pairs per warp; NT: n-tiles (two batch rows each) per pass; GEN: grouped (MoE decode: blockIdx.y = expert,
// device-side row offsets) and / or more tiles than tile buffers (ring).  GEN = false is the Llama-shape kernel: every
// GEN feature is compiled out of it.
template <int GPW2, int NT, bool GEN>
__global__ void __launch_bounds__(NTHR, 1) gemv_dec_kernel(const __grid_constant__ CUtensorMap tmap, const DecParams p) {
    constexpr int MB = 2 * NT;       // batch rows per pass
    constexpr int RB = NW * WARP_RED;   // bytes of one partial-tile buffer
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    B200Q_STAMP(0);
    // work item -> (expert, CTA of the expert)
    const int b = (int)blockIdx.x;
    int expert = GEN ? (int)blockIdx.y : 0;
    int Mrows = p.M;                                          // batch rows of this work item
    int64_t xrow0 = 0;                                        // ... and where they start in x / y
    if constexpr (GEN) {
        if (p.offsets) {
            int lo, hi;
            if ((int)gridDim.y < p.n_groups) {                // compact grid (n_groups <= 32): lane e looks at expert e, every warp alike
                const int e = lane < p.n_groups ? lane : 0;
                const int lo_e = p.offsets[e], hi_e = p.offsets[e + 1];
                const bool has = lane < p.n_groups && hi_e > lo_e;
                const unsigned act = __ballot_sync(0xffffffffu, has);
                const unsigned sel = __ballot_sync(0xffffffffu, has && __popc(act & ((1u << lane) - 1u)) == (int)blockIdx.y);
                if (!sel) return;                             // uniform: fewer experts with rows than grid rows
                expert = __ffs(sel) - 1;
                lo = __shfl_sync(0xffffffffu, lo_e, expert); hi = __shfl_sync(0xffffffffu, hi_e, expert);
            } else {
                lo = p.offsets[expert]; hi = p.offsets[expert + 1];
            }
            Mrows = min(hi - lo, 16);
            xrow0 = lo;
            if (Mrows <= 0) return;                           // uniform: this expert has no tokens
        }
    }
    auto x_row = [&](int m) -> int64_t {                      // row of x that batch row m of this work item reads
        if constexpr (GEN) { if (p.row_map) return p.row_map[xrow0 + m]; }
        return xrow0 + m;
    };
    // (gated: rows are dealt out in gate / up pairs, so both projections of an output column meet in one CTA)
    const int unit = p.gated ? 2 : 1;
    const int orow0 = unit * (b * p.rows_q + min(b, p.rows_rem));    // first row of this CTA inside the expert's N rows
    const int nrows = unit * (p.rows_q + (b < p.rows_rem ? 1 : 0));
    if constexpr (GEN) { if (nrows <= 0) return; }            // uniform
    const int r0 = GEN ? expert * p.N + orow0 : orow0;        // ... and in the stacked weight tensor [E N, K/2]
    const int npasses = GEN ? (Mrows + MB - 1) / MB : p.npasses;
    // One more row: a row of bytes 0x11 (q_lo = q_hi = 1) behind the last weight row makes the tensor cores deliver
    // sum_k X (the zero-point term) as one more output row -- no extra arithmetic in the operand build.
    const int S = (nrows + 1 + TILE_ROWS - 1) / TILE_ROWS;   // tiles of this CTA
    const int W = p.ntiles_max;                               // tile buffers = tiles per window (plain instances: >= S)
    const int R = GEN ? (b < p.rows_rem ? p.win_hi[0] : p.win_lo[0]) : 1;       // windows
    const int T0 = GEN ? (b < p.rows_rem ? p.win_hi[1] : p.win_lo[1]) : S;      // tiles of the window processed first (the last rows)
    const bool ring = GEN && S > W;
    const int tf = (nrows >> 4) - (R - 1) * W, rf = nrows & 15;      // tile (of the first window) / row of the 0x11 row
    int* s_ex = reinterpret_cast<int*>(smem + OFF_EX);
    unsigned int* s_flag = reinterpret_cast<unsigned int*>(smem + OFF_FLAG);
    unsigned int* s_amax = reinterpret_cast<unsigned int*>(smem + OFF_AMAX);
    int* fin = reinterpret_cast<int*>(smem + p.fin_off);
    int* s_sumx = reinterpret_cast<int*>(smem + OFF_SUMX);
    auto tile_bar = [&](int buf, int grp) { return sbase + OFF_BARS + 8u * (uint32_t)(buf * p.nbars + grp); };
    auto cons_bar = [&](int buf, int grp) { return sbase + OFF_CONS + 8u * (uint32_t)(buf * p.nbars + grp); };
    auto full_bar = [&](int bb) { return sbase + OFF_FULL + 8u * (uint32_t)bb; };
    auto empty_bar = [&](int bb) { return sbase + OFF_EMPTY + 8u * (uint32_t)bb; };

    if (tid == 0) {
        const int nb = min(S, W) * p.nbars;
        for (int i = 0; i < nb; ++i) mbar_init(sbase + OFF_BARS + 8u * i, 1);
        if constexpr (GEN) {
            if (ring)                                        // one arrival per pair of the group (by the warp that owns the pair)
                for (int i = 0; i < nb; ++i) mbar_init(sbase + OFF_CONS + 8u * i, (uint32_t)min(p.chunk, p.npairs - (i % p.nbars) * p.chunk));
        }
        for (int i = 0; i < 2; ++i) { mbar_init(full_bar(i), NW); mbar_init(empty_bar(i), NW); }
        fence_mbar_init();
        *s_flag = 0u;
    }
    if (!p.slots)                                            // pipelined reduction: integer atomics into a zeroed accumulator
        for (int i = tid; i < npasses * W * NT * 128; i += NTHR) fin[i] = 0;
    __syncthreads();
    pdl_launch_dependents();
    B200Q_STAMP(1);

    // ---- weight requests (lane 0 of the last warp).  Tile i, pair P lands at tile_off + i * tile_bytes + P * 2 KB.
    // Staged (tuning key gemv_early = a + 10 * b): a tiles before griddepcontrol.wait, b more once the x loads are
    // in flight, the rest when the x operand is built -- bounds what queues ahead of the x loads.
    const bool issuer = warp == NW - 1 && lane == 0 && !B200Q_ABL(4);
    const int nops = min(S, W) * p.nbars;                    // first fill: requests of ~32 KB, (tile, pair group)
    auto issue = [&](int from, int to) {
        if constexpr (GEN) dec_issue_ring(&tmap, sbase + OFF_BARS, sbase + p.tile_off, r0, from, to, p.nbars, p.chunk, p.tile_bytes, W, T0, R);
        else dec_issue_tiles(&tmap, sbase + OFF_BARS, sbase + p.tile_off, r0, from, to, p.nbars, p.chunk, p.tile_bytes);
    };
    // (GEN, issuer) next refill: request (tile nx_s = processing order, pair group nx_grp) into buffer nx_buf once the
    // tile that is there now has been consumed (parity nx_par of its barrier); nx_rt: its tile index in row order
    int nx_s = W, nx_grp = 0, nx_buf = 0, nx_rt = 0;
    uint32_t nx_par = 0u;
    if constexpr (GEN) { if (ring) nx_rt = ring_row_tile(W, W, T0, R); }
    const int early = min(p.early_tiles % 10, nops);
    const int mid = min(early + p.early_tiles / 10, nops);
    auto prefetch_next = [&]() {
        if (!p.next_bytes) return;
        const unsigned long long per = (((unsigned long long)p.next_chunk) + 127ull) & ~127ull;
        const unsigned long long beg = min(per * blockIdx.x, p.next_bytes), end = min(beg + per, p.next_bytes);
        for (unsigned long long off = beg; off < end; off += 32768ull) {
            const unsigned int n = (unsigned int)min(32768ull, end - off) & ~15u;
            if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.next_packed + off), "r"(n) : "memory");
        }
    };
    if (p.wait_weights) pdl_wait();
    if (issuer) {
        tma_prefetch_desc(&tmap);
        if (p.pf_mode == 2) prefetch_next();
        issue(0, early);
        if (early == nops && p.pf_mode == 1) prefetch_next();
    }
    B200Q_STAMP(2);

    // per-lane constants of the main loop: ldmatrix row address of this lane for the four 32-byte steps of a pair
    // (lane i supplies row (i & 7) + 8 ((i >> 3) & 1) of the 16-byte column 2 c + (i >> 4); 128-byte swizzle: column ^ row)
    uint32_t offc[4];
    {
        const int ri = lane & 7, mi = lane >> 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) offc[c] = (uint32_t)((ri + 8 * (mi & 1)) * 128 + (((2 * c + (mi >> 1)) ^ ri) << 4));
    }
    const uint32_t red = sbase + p.red_off;
    // Operand exchange space of this warp: four 512-byte parts = its OWN partial-tile slots (slots of tiles 0..3, or its
    // kilobyte in both buffers of the pipelined reduction) -- nobody else touches them before the main loop.
    const uint32_t xbase = red + (uint32_t)(warp * (p.slots ? 512 : WARP_RED));
    const uint32_t xs1 = p.slots ? (uint32_t)(NW * 512) : 512u, xs2 = p.slots ? (uint32_t)(2 * NW * 512) : (uint32_t)RB;
    auto xpart = [&](int k) { return xbase + (uint32_t)(k & 1) * xs1 + (uint32_t)(k >> 1) * xs2; };
    const int hsel = g >> 2, lsel = g & 3;                   // lane (g, t) holds mma column g: digit lsel of batch row 2 nt + hsel
    // ring refills (issuer only): request tile s + W as soon as every warp is done with tile s; lim: first tile (processing
    // order) that must not be requested yet
    auto advance = [&](int lim_s, bool block) {
        while (nx_s < lim_s) {
            const uint32_t cb = cons_bar(nx_buf, nx_grp);
            if (block) mbar_wait(cb, nx_par);
            else if (!mbar_try_wait(cb, nx_par)) break;
            const uint32_t bar = tile_bar(nx_buf, nx_grp);
            mbar_arrive_expect_tx(bar, (uint32_t)(p.chunk * PAIR_BYTES));
            tma_box_3d(sbase + p.tile_off + (uint32_t)(nx_buf * p.tile_bytes + nx_grp * p.chunk * PAIR_BYTES), &tmap, 0,
                       r0 + nx_rt * TILE_ROWS, nx_grp * p.chunk, bar, policy_evict_first());
            if (++nx_grp == p.nbars) {
                nx_grp = 0;
                ++nx_s;
                if (++nx_buf == W) { nx_buf = 0; nx_par ^= 1u; }
                nx_rt += (nx_buf == (T0 == W ? 0 : T0)) ? 1 - 2 * W : 1;     // a new window starts one window further up
            }
        }
    };

    // pair group (= barrier) of each of this warp's pairs: computed here, not in the loop (an integer division costs
    // ~150 clk on the critical path of every tile)
    int grp_of[GPW2];
#pragma unroll
    for (int q = 0; q < GPW2; ++q) grp_of[q] = GPW2 == 1 ? 0 : (warp + NW * q) / p.chunk;
    int it = 0;                                              // partial tiles handed over so far (all passes)
    auto reduce_tile = [&](int itx, int pass, int tile) {
        const int bb = itx & 1;
        mbar_wait(full_bar(bb), (uint32_t)((itx >> 1) & 1));
        constexpr int WORDS = NT * 128, SRC = NTHR / WORDS, PER = NW / SRC;
        const int rho = tid % WORDS, sq = tid / WORDS;
        const uint32_t src = red + (uint32_t)(bb * RB + sq * PER * WARP_RED + rho * 4);
        int s = 0;
#pragma unroll
        for (int w = 0; w < PER; ++w) s += lds32(src + (uint32_t)(w * WARP_RED));
        red_add_s32(sbase + p.fin_off + (uint32_t)((((pass * W + tile) * NT) * 128 + rho) * 4), s);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(bb));
    };

    pdl_wait();              // x (and y) belong to the stream-ordered predecessor
    B200Q_STAMP(3);
    uint32_t bf[GPW2][4][NT][4];                              // per 32-byte step: {e-word b0, e-word b1, z-word b0, z-word b1}
#pragma unroll 1
  for (int k = 0; k < R; ++k) {                               // windows of W tiles, last rows first (plain instances: one)
    const int Tk = k == 0 ? T0 : W;                           // tiles of this window
    const int pos0 = (R - 1 - k) * W * TILE_ROWS;             // ... and its first row (relative to the CTA's rows)
    // scale / zero point of this thread's output row (one row per thread and window), fetched early (latency hidden
    // by the main loop)
    const bool mine = tid < Tk * TILE_ROWS && pos0 + tid < nrows;
    float pre_sc = 0.0f, pre_zp = 0.0f;
    if (mine) { pre_sc = __ldg(p.scales + r0 + pos0 + tid); pre_zp = __ldg(p.zps + r0 + pos0 + tid); }
#pragma unroll 1
    for (int pass = 0; pass < npasses; ++pass) {
        const int m0 = pass * MB;
        const bool first = k == 0 && pass == 0;
      if (k == 0 || npasses > 1) {                            // (one pass: the operand stays in registers for every window)
        float xv[GPW2][MB][8];
        // ---- x of this pass: this warp's columns only, all loads in flight at once
#pragma unroll
        for (int q = 0; q < GPW2; ++q) {
            const int col = (warp + NW * q) * 256 + lane * 8;
#pragma unroll
            for (int hr = 0; hr < MB; ++hr) {
#pragma unroll
                for (int e = 0; e < 8; ++e) xv[q][hr][e] = 0.0f;
                if (col < p.K && m0 + hr < Mrows) load8f(p.x, p.x_dtype, x_row(m0 + hr) * p.K + col, xv[q][hr]);
            }
        }
        if (first && issuer && early < mid) {
            issue(early, mid);
            if (mid == nops && p.pf_mode == 1) prefetch_next();
        }
        int ex[MB];
#pragma unroll
        for (int hr = 0; hr < MB; ++hr) ex[hr] = 0;
        // ---- row amax (non-negative floats order like their bit patterns: one REDUX per row); NaN / Inf show up
        // as an exponent field of 0xff
#pragma unroll
        for (int hr = 0; hr < MB; ++hr) {
            if (m0 + hr < Mrows) {                            // uniform
                unsigned int u = 0u;
#pragma unroll
                for (int q = 0; q < GPW2; ++q)
#pragma unroll
                    for (int e = 0; e < 8; ++e) u = max(u, __float_as_uint(xv[q][hr][e]) & 0x7fffffffu);
                u = __reduce_max_sync(0xffffffffu, u);
                if (lane == 0) s_amax[((pass & 1) * 4 + hr) * NW + warp] = u;
            }
        }
        __syncthreads();
        if (first) B200Q_STAMP(4);
#pragma unroll
        for (int hr = 0; hr < MB; ++hr) {
            if (m0 + hr < Mrows) {                            // uniform
                const unsigned int u = __reduce_max_sync(0xffffffffu, lane < NW ? s_amax[((pass & 1) * 4 + hr) * NW + lane] : 0u);
                if (u >= 0x7f800000u) {
                    if (tid == 0) atomicOr(s_flag, 1u << (m0 + hr));
                } else if (u > 0u) {
                    ex[hr] = max(-96, min(126, 155 - (int)(u >> 23)));
                }
                if (tid == 0) s_ex[m0 + hr] = ex[hr];
            }
        }

        // ---- operand: digits of Xe / Z, exchanged inside the warp.  Source lane i (0..31) of a pair holds columns
        // 8 i .. 8 i + 7 = packed bytes 4 i .. 4 i + 3: word (i & 3) of the 16-byte column i >> 2, i.e. the bytes lane quad
        // t = i & 3 multiplies in step c = i >> 3, first (b0) or second (b1) half (i >> 2) & 1.
        const uint32_t soff = (uint32_t)((((lane >> 3) * 4 + (lane & 3)) * 16) + ((lane >> 2) & 1) * 4);
        const uint32_t ssrc = xpart(2 * hsel + (lsel >> 1)) + (uint32_t)((lsel & 1) * 256 + t * 16);
#pragma unroll
        for (int q = 0; q < GPW2; ++q) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                  // the two batch rows of this n-tile: independent chains
                    const int hr = 2 * nt + hh;
                    if (m0 + hr >= Mrows) continue;               // uniform: no such batch row (odd M)
                    const float up = __uint_as_float((uint32_t)(127 + ex[hr]) << 23), up16 = up * 0.0625f;
                    uint32_t D[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // even column: Xe = round(x 2^e); odd column: Xo = round(x 2^(e-4)), carried as Z = Xo - Xe
                        const int Xe = __float2int_rn(xv[q][hr][2 * i] * up);
                        const int Xo = __float2int_rn(xv[q][hr][2 * i + 1] * up16);
                        D[2 * i] = (uint32_t)(Xe + 0x00808080) ^ 0x00808080u;      // byte l = signed base-256 digit l
                        D[2 * i + 1] = (uint32_t)(Xo - Xe + 0x00808080) ^ 0x00808080u;
                    }
                    // 4x4 byte transposes: digit l of the four Xe -> e-word l (meets the raw bytes), of the four Z -> z-word l
                    const uint32_t e0 = __byte_perm(D[0], D[2], 0x5140), e1 = __byte_perm(D[4], D[6], 0x5140);
                    const uint32_t e2 = __byte_perm(D[0], D[2], 0x7362), e3 = __byte_perm(D[4], D[6], 0x7362);
                    const uint32_t o0 = __byte_perm(D[1], D[3], 0x5140), o1 = __byte_perm(D[5], D[7], 0x5140);
                    const uint32_t o2 = __byte_perm(D[1], D[3], 0x7362), o3 = __byte_perm(D[5], D[7], 0x7362);
                    const uint32_t d01 = xpart(2 * hh) + soff, d23 = xpart(2 * hh + 1) + soff;      // digits 0, 1 / 2, 3 of row hh
                    sts32(d01, __byte_perm(e0, e1, 0x5410));       sts32(d01 + 8, __byte_perm(o0, o1, 0x5410));
                    sts32(d01 + 256, __byte_perm(e0, e1, 0x7632)); sts32(d01 + 264, __byte_perm(o0, o1, 0x7632));
                    sts32(d23, __byte_perm(e2, e3, 0x5410));       sts32(d23 + 8, __byte_perm(o2, o3, 0x5410));
                    sts32(d23 + 256, __byte_perm(e2, e3, 0x7632)); sts32(d23 + 264, __byte_perm(o2, o3, 0x7632));
                }
                __syncwarp();
                if (m0 + 2 * nt + hsel < Mrows) {                 // this lane's mma column belongs to a batch row that exists
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 v = lds128(ssrc + (uint32_t)(c * 64));
                        bf[q][c][nt][0] = v.x; bf[q][c][nt][1] = v.y; bf[q][c][nt][2] = v.z; bf[q][c][nt][3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) bf[q][c][nt][j] = 0u;
                }
                __syncwarp();
            }
        }
      }
        if (first && issuer) {
            if (mid < nops) issue(mid, nops);
            if (p.pf_mode == 3 || (p.pf_mode == 1 && mid < nops)) prefetch_next();
        }
        if (first) B200Q_STAMP(5);

        // ---- main loop: one 16-row tile per iteration, this warp's pairs of it
        for (int i = 0; i < Tk; ++i) {
            // (GEN) buffer of the tile and parity of its barriers: tile s = T0 + (k - 1) W + i of the processing order
            int buf = i;
            uint32_t par = 0u;
            if constexpr (GEN) {
                if (k > 0) {
                    const int tq = T0 + i, wrap = tq >= W ? 1 : 0;
                    buf = tq - wrap * W;
                    par = (uint32_t)((k - 1 + wrap) & 1);
                }
            }
            // raw bytes x digits of Xe, masked bytes x digits of Z (two chains of dependent IMMAs; four were not faster)
            int c0[NT][4], c1[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) { c0[nt][r] = 0; c1[nt][r] = 0; }
            const uint32_t tb = sbase + p.tile_off + (uint32_t)(buf * p.tile_bytes);
#pragma unroll
            for (int q = 0; q < GPW2; ++q) {
                const int P = warp + NW * q;
                if (P < p.npairs) {                           // uniform
                    if (pass == 0 && !B200Q_ABL(4)) mbar_wait(tile_bar(buf, grp_of[q]), par);
                    const uint32_t pb = tb + (uint32_t)(P * PAIR_BYTES);
                    if (first && i == tf) {                   // the row of 0x11 bytes behind the last weight row
                        if (lane < 8) sts128(pb + (uint32_t)(rf * 128 + lane * 16), make_uint4(0x11111111u, 0x11111111u, 0x11111111u, 0x11111111u));
                        __syncwarp();
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        // no nibble extraction: sum_k (q_lo + 16 q_hi) Xe + (16 q_hi) (Xo - Xe) = sum_k q_lo Xe + q_hi 16 Xo
                        uint32_t a[4];
                        ldsm_x4(a, pb + offc[c]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) imma(c0[nt], a[0], a[1], a[2], a[3], bf[q][c][nt][0], bf[q][c][nt][1]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) a[r] &= 0xf0f0f0f0u;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) imma(c1[nt], a[0], a[1], a[2], a[3], bf[q][c][nt][2], bf[q][c][nt][3]);
                    }
                    if constexpr (GEN) {
                        if (ring && pass == npasses - 1) {    // uniform: hand the buffer back, refill what is free
                            if (k + 1 < R) {                  // (a tile of the last window is never refilled... its successor is)
                                __syncwarp();
                                if (lane == 0) mbar_arrive(cons_bar(buf, grp_of[q]));
                            }
                            if (issuer) advance(S, false);
                        }
                    }
                }
            }
            if (p.slots) {
                // single pass, two batch rows: every (tile, warp) has its own slot, no synchronisation in the loop
                sts128(red + (uint32_t)((i * NW + warp) * 512 + lane * 16),
                       make_uint4((uint32_t)(c0[0][0] + c1[0][0]), (uint32_t)(c0[0][1] + c1[0][1]),
                                  (uint32_t)(c0[0][2] + c1[0][2]), (uint32_t)(c0[0][3] + c1[0][3])));
                if (first && i < 6) B200Q_STAMP(6 + i);
                continue;
            }
            // hand the partial tile over (double buffer; the buffer was last read for tile it - 2)
            const int bb = it & 1;
            if (it >= 2) mbar_wait(empty_bar(bb), (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                sts128(red + (uint32_t)(bb * RB + warp * WARP_RED + nt * 512 + lane * 16),
                       make_uint4((uint32_t)(c0[nt][0] + c1[nt][0]), (uint32_t)(c0[nt][1] + c1[nt][1]),
                                  (uint32_t)(c0[nt][2] + c1[nt][2]), (uint32_t)(c0[nt][3] + c1[nt][3])));
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(bb));
            if (i >= 1) reduce_tile(it - 1, pass, i - 1);
            if (first && i < 6) B200Q_STAMP(6 + i);
            ++it;
        }
        if (!p.slots) reduce_tile(it - 1, pass, Tk - 1);
        if (first) B200Q_STAMP(12);
        // (the exchange space of the next pass is the warp's own partial-tile slots: the amax barrier of that pass comes
        // after every warp has folded the last tile of this one, so they are free again)
    }

    // ======== after the last pass of the window: fold, epilogue
    __syncthreads();
    if constexpr (GEN) {
        if (ring && issuer) advance(min(S, T0 + k * W + W), true);      // every buffer of this window is free
    }
    if (p.slots) {
        // fold the 16 warp slots of every tile (integer adds: exact, order independent)
        // (M = 1: only mma columns 0..3 are live, i.e. the words of lane quads t = 0, 1)
        const int sh = Mrows == 1 ? 6 : 7;
        for (int v = tid; v < (Tk << sh); v += NTHR) {
            const int tile = v >> sh, rem = v & ((1 << sh) - 1);
            const int word = Mrows == 1 ? ((rem >> 3) * 16 + (rem & 7)) : rem;
            const uint32_t src = red + (uint32_t)((tile * NW) * 512 + word * 4);
            int s = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) s += lds32(src + (uint32_t)(k * 512));
            fin[tile * 128 + word] = s;
        }
        __syncthreads();
    }
    if (k == 0) B200Q_STAMP(13);

    // ---- epilogue: one thread per output (tile, batch row, row): digits -> sum_k q*X (exact s64),
    // y = s * 2^-e * (sum_k q*X - zp * sum_k X) (+ bias)
    const unsigned int flagged = *s_flag;
    // word of (mma row r, column col) in a 16 x 8 tile stored as [lane = 4 (r & 7) + col / 2][reg = 2 (r >> 3) + (col & 1)]
    auto word_of = [](int r, int col) { return (((r & 7) * 4 + (col >> 1)) * 4) + ((r >> 3) * 2 + (col & 1)); };
    {
        if (k == 0 && R > 1 && tid < 4 * Mrows) {            // sum_k X of every batch row, for the epilogues of the later windows
            const int m = tid >> 2, pass = m / MB, mm = m - pass * MB;
            s_sumx[tid] = fin[((pass * W + tf) * NT + (mm >> 1)) * 128 + word_of(rf, 4 * (mm & 1) + (tid & 3))];
        }
        const int ti = tid >> 4, r = tid & 15, row = orow0 + pos0 + tid;      // output column (row of the expert's weight matrix)
        const float sc = pre_sc, zp = pre_zp;
        const float bias = (mine && p.bias) ? __ldg(p.bias + row) : 0.0f;
        const int zi = __float2int_rn(zp);
        const bool zint = (float)zi == zp && zi >= -32768 && zi <= 32767;
        for (int m = 0; m < Mrows; ++m) {
            if ((flagged >> m) & 1u) continue;               // uniform
            float v = 0.0f;
            if (mine) {
                const int pass = m / MB, mm = m - pass * MB, nt = mm >> 1, h = mm & 1;
                const int* f = fin + ((pass * W + ti) * NT + nt) * 128;
                const int* ft = fin + ((pass * W + tf) * NT + nt) * 128;     // the 0x11 row: sum_k X
                long long a = 0, txl = 0;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    a += (long long)f[word_of(r, 4 * h + l)] << (8 * l);
                    txl += (long long)(k == 0 ? ft[word_of(rf, 4 * h + l)] : s_sumx[4 * m + l]) << (8 * l);
                }
                const int ex = s_ex[m];
                if (zint && ex >= -126) {
                    // quantiser-made zero points are integers: a - zp * sum X exactly in s64, ONE rounding to fp32 (the
                    // same value the fp64 expression below rounds to), then the exact power of two and the scale
                    v = sc * (__ll2float_rn(a - (long long)zi * txl) * __uint_as_float((uint32_t)(127 - ex) << 23));
                } else {
                    const double down = __longlong_as_double((long long)(1023 - ex) << 52);                // 2^-e
                    v = sc * (float)(((double)a - (double)zp * (double)txl) * down);
                }
            }
            if (p.gated) {
                // fused gate + up pair: the even row (gate) fetches its neighbour's value (up) and writes silu(gate) * up
                const float u = __shfl_down_sync(0xffffffffu, v, 1);
                if (mine && !(tid & 1)) store_out(p.y, p.y_dtype, (xrow0 + m) * (p.N >> 1) + (row >> 1), v / (1.0f + __expf(-v)) * u);
            } else if (mine) {
                store_out(p.y, p.y_dtype, (xrow0 + m) * p.N + row, v + bias);
            }
        }
    }
    if (k + 1 < R && !p.slots) {                              // the accumulator of the next window starts from zero
        __syncthreads();
        for (int i = tid; i < npasses * W * NT * 128; i += NTHR) fin[i] = 0;
        __syncthreads();
    }
  }
    // ---- batch rows with NaN / Inf: the reference's arithmetic (dequantise, then fp32 multiply-add), so that
    // non-finite values propagate as in F.linear; one warp per output, weights re-read from global memory
    if (const unsigned int flagged = *s_flag) {
        const int64_t row_bytes = p.K >> 1;
        for (int m = 0; m < Mrows; ++m) {
            if (!((flagged >> m) & 1u)) continue;
            auto ref_row = [&](int row) {                     // row: index in the stacked weight tensor
                const float sc = __ldg(p.scales + row), zp = __ldg(p.zps + row);
                const uint8_t* wr = p.packed + (int64_t)row * row_bytes;
                float acc = 0.0f;
                for (int kb = lane; kb < row_bytes; kb += 32) {
                    const unsigned int byte = wr[kb];
                    const float w0 = ((float)(byte & 15u) - zp) * sc, w1 = ((float)(byte >> 4) - zp) * sc;
                    acc = fmaf(w0, load1f(p.x, p.x_dtype, x_row(m) * p.K + 2 * kb), acc);
                    acc = fmaf(w1, load1f(p.x, p.x_dtype, x_row(m) * p.K + 2 * kb + 1), acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                return acc;
            };
            for (int rc = unit * warp; rc < nrows; rc += unit * NW) {
                const int row = orow0 + rc;
                float acc = ref_row(r0 + rc);
                if (p.gated) {
                    const float up = ref_row(r0 + rc + 1);
                    if (lane == 0) store_out(p.y, p.y_dtype, (xrow0 + m) * (p.N >> 1) + (row >> 1), acc / (1.0f + __expf(-acc)) * up);
                } else if (lane == 0) {
                    if (p.bias) acc += __ldg(p.bias + row);
                    store_out(p.y, p.y_dtype, (xrow0 + m) * p.N + row, acc);
                }
            }
        }
    }
    B200Q_STAMP(14);
}

struct DecPlan {
    int grid_g, rows_q, rows_rem, s_max, gpw2, nt, npasses, ntiles, npairs, nbars, chunk, tile_bytes, tile_off;
    int red_off, fin_off, slots;
    size_t smem;
};

// M: (largest) number of batch rows of a work item; N: weight rows (per expert)
bool plan_dec(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, DecPlan* c, int gated = 0) {
    if (M < 1 || M > 16 || K < 256 || K % 256 != 0 || K > 16384 || N < 1 || N > 0x7fffffff) return false;
    if (gated && (N & 1)) return false;
    const int unit = gated ? 2 : 1;                                // rows are dealt out in gate / up pairs
    c->npairs = (int)(K / 256);
    c->gpw2 = (c->npairs + NW - 1) / NW;                          // <= 4
    c->nt = (M >= 3 && c->gpw2 <= 2) ? 2 : 1;                     // B fragments: 16 gpw2 nt registers
    const int mb = 2 * c->nt;
    c->npasses = (int)((M + mb - 1) / mb);
    c->nbars = c->gpw2;
    c->chunk = (c->npairs + c->nbars - 1) / c->nbars;
    c->tile_bytes = c->nbars * c->chunk * PAIR_BYTES;             // >= npairs * 2 KB: a 3-D box always has room
    int cap = tuning().gemv_ctas > 0 ? tuning().gemv_ctas : sm_count;
    if (cap > sm_count) cap = sm_count;
    const int64_t units = N / unit;
    int64_t g = (N + TILE_ROWS - 2 - (unit - 1)) / (TILE_ROWS - unit);       // few rows: at most 15 (14) per CTA (+ the 0x11 row = one tile)
    if (g > cap) g = cap;
    if (g > units) g = units;
    if (g < 1) g = 1;
    c->grid_g = (int)g;
    c->rows_q = (int)(units / g); c->rows_rem = (int)(units % g);
    const int64_t br = unit * ((units + g - 1) / g);              // most rows of a CTA
    if (br > 0x3fffffff) return false;
    const int S = (int)((br + 1 + TILE_ROWS - 1) / TILE_ROWS);    // one more row: the 0x11 row that yields sum_k X
    c->s_max = S;
    // as many tile buffers as fit (all S tiles for the Llama shapes: no refills); 16 rows x 16 tiles = one output row
    // per thread in the epilogue of a window
    const int wmax = tuning().gemv_bufs > 0 ? tuning().gemv_bufs : 16;
    for (int w = S < wmax ? S : wmax; w >= 1; --w) {
        if (w * c->nbars > MAX_BARS) continue;
        c->ntiles = w;
        const int fin_bytes = c->npasses * w * c->nt * 512;
        const int red_bytes = 2 * NW * WARP_RED;
        const int slot_bytes = (w > 4 ? w : 4) * NW * 512;        // >= 4 slots per warp: its exchange space
        // (1) single pass of <= 2 batch rows: one slot per (tile, warp); (2) double-buffered pipelined reduction
        for (int form = (c->nt == 1 && c->npasses == 1 && tuning().gemv_slots != 0) ? 0 : 1; form < 2; ++form) {
            c->slots = form == 0;
            c->red_off = OFF_DYN;
            c->fin_off = c->red_off + (c->slots ? slot_bytes : red_bytes);
            c->tile_off = (c->fin_off + fin_bytes + 1023) / 1024 * 1024;
            c->smem = (size_t)c->tile_off + (size_t)w * c->tile_bytes;
            if (c->smem <= (size_t)max_smem) return true;
        }
    }
    return false;
}

}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn dec_encode_fn() {
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

// tensor maps only depend on (address, N, K, form): a decode loop calls the same layers again and again
struct MapKey { const void* ptr; int64_t N, K; int form; };
struct MapSlot { MapKey key; CUtensorMap map; bool ok; };

int dec_weight_map(const uint8_t* packed, int64_t N, int64_t K, int chunk, CUtensorMap_st* out) {
    constexpr int SLOTS = 64;
    static thread_local MapSlot cache[SLOTS];
    static thread_local bool init = false;
    if (!init) { for (auto& s : cache) s.ok = false; init = true; }
    const size_t h = ((reinterpret_cast<uintptr_t>(packed) >> 8) * 0x9E3779B97F4A7C15ull >> 40) % SLOTS;
    MapSlot& s = cache[h];
    if (s.ok && s.key.ptr == packed && s.key.N == N && s.key.K == K && s.key.form == chunk) { *out = s.map; return 0; }
    EncodeTiledFn fn = dec_encode_fn();
    if (!fn) return set_error(B200Q_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // packed [N][K/2] viewed as [row][column of 128 bytes][byte]: the box [16 rows][chunk columns][128 B] lands in shared
    // memory as [column][row][128 B] with the 128-byte swizzle keyed by the row -- the layout ldmatrix wants
    cuuint64_t dims[3] = {128, (cuuint64_t)N, (cuuint64_t)(K / 256)};
    cuuint64_t strides[2] = {(cuuint64_t)(K / 2), 128};
    cuuint32_t box[3] = {128, TILE_ROWS, (cuuint32_t)chunk};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(&s.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(packed), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { s.ok = false; return set_error(B200Q_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r); }
    s.key = MapKey{packed, N, K, chunk};
    s.ok = true;
    *out = s.map;
    return 0;
}

namespace {

template <int GPW2, int NT, bool GEN>
int launch_dec_inst(const DecPlan& c, dim3 grid, const CUtensorMap& map, const DecParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_dec_kernel<GPW2, NT, GEN>;
    static thread_local int attr_dev_smem[64] = {0};
    int dev = 0;
    B200Q_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_dev_smem[dev] < (int)c.smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        attr_dev_smem[dev] = (int)c.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    int na = 0;
    if (pdl) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, map, p), "gemv_dec launch");
}

}  // namespace

bool gemv_dec_supported(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated) {
    DecPlan c;
    return plan_dec(dev.sm_count, dev.max_smem_optin, M, N, K, &c, gated);
}

bool gemv_dec_resident(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated) {
    DecPlan c;
    return plan_dec(dev.sm_count, dev.max_smem_optin, M, N, K, &c, gated) && c.s_max <= c.ntiles;
}

int launch_gemv_dec(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                    const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                    unsigned flags, cudaStream_t st, const uint8_t* next_packed, size_t next_bytes, int gated,
                    const int32_t* offsets, int n_experts, const int32_t* row_map, int max_groups) {
    DecPlan c;
    if (n_experts < 1) n_experts = 1;
    // grid rows: one per expert, or -- when the caller knows that at most max_groups experts can have rows -- one per possible group
    const int grid_y = (offsets && max_groups >= 1 && max_groups < n_experts && n_experts <= 32) ? max_groups : n_experts;
    if (!plan_dec(dev.sm_count, dev.max_smem_optin, M, N, K, &c, gated))
        return set_error(B200Q_EINVAL, "gemv_dec: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv_dec: x and packed must be 16-byte aligned");
    DecParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.bias = gated ? nullptr : bias; p.y = y;
    p.gated = gated;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.rows_q = c.rows_q; p.rows_rem = c.rows_rem;
    const int unit = gated ? 2 : 1;
    auto windows = [&](int units, int* out) {                 // {R, T0} of a CTA with that many row units
        const int S = (units * unit + 1 + TILE_ROWS - 1) / TILE_ROWS;
        out[0] = (S + c.ntiles - 1) / c.ntiles;
        out[1] = S - (out[0] - 1) * c.ntiles;
    };
    windows(c.rows_q + 1, p.win_hi);
    windows(c.rows_q, p.win_lo);
    const bool gen = offsets != nullptr || n_experts > 1 || c.s_max > c.ntiles;
    p.offsets = offsets;
    p.row_map = offsets ? row_map : nullptr;
    p.n_groups = n_experts;
    p.npairs = c.npairs; p.nbars = c.nbars; p.chunk = c.chunk;
    p.ntiles_max = c.ntiles; p.tile_bytes = c.tile_bytes; p.tile_off = c.tile_off; p.npasses = c.npasses;
    p.red_off = c.red_off; p.fin_off = c.fin_off; p.slots = c.slots;
    p.wait_weights = (flags & B200Q_FLAG_STATIC_WEIGHTS) ? 0 : 1;
    p.early_tiles = tuning().gemv_early >= 0 ? tuning().gemv_early : 92;     // two tiles up front, the rest behind the x loads
    p.next_packed = next_packed;
    p.pf_mode = tuning().gemv_pf;
    p.next_bytes = (tuning().gemv_pf != 0 && next_packed && (reinterpret_cast<uintptr_t>(next_packed) & 15) == 0) ? next_bytes : 0;
    const int grid_total = c.grid_g * n_experts;
    if (gen && n_experts > 1) p.next_bytes = 0;
    p.next_chunk = (unsigned int)(p.next_bytes / (unsigned long long)grid_total);
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    CUtensorMap map;
    if (int rc = dec_weight_map(packed, N * n_experts, K, c.chunk, &map)) return rc;
    const bool pdl = tuning().gemv_pdl != 0;
    const dim3 grid((unsigned)c.grid_g, (unsigned)grid_y);
#define B200Q_DEC_CASE(GPW2_, NT_)                                                                                  \
    if (c.gpw2 == GPW2_ && c.nt == NT_)                                                                             \
        return gen ? launch_dec_inst<GPW2_, NT_, true>(c, grid, map, p, pdl, st) : launch_dec_inst<GPW2_, NT_, false>(c, grid, map, p, pdl, st);
    B200Q_DEC_CASE(1, 1) B200Q_DEC_CASE(2, 1) B200Q_DEC_CASE(3, 1) B200Q_DEC_CASE(4, 1)
    B200Q_DEC_CASE(1, 2) B200Q_DEC_CASE(2, 2)
#undef B200Q_DEC_CASE
    return set_error(B200Q_EINVAL, "gemv_dec: no kernel instance for gpw2=%d nt=%d", c.gpw2, c.nt);
}

}  // namespace b200q

#ifdef B200Q_PROF
extern "C" int b200q_debug_read_prof_dec(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_dec_prof, sizeof(long long) * 256 * 16), "read prof");
}
#endif
