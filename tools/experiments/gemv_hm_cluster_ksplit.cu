// EXPERIMENT (not built, not shipped): gemv_hm.cu with K slices over a thread-block cluster.  Row groups of 2 / 4 CTAs, each
// CTA takes K / ck columns of the group's rows (x traffic per SM / ck, one pair per warp even for K = 11008), partial
// tiles pushed to their owner through distributed shared memory (st.shared::cluster + one cluster barrier per pass).
// Parity-green (tests of the product kernel with hm_ck = 1 / 2 / 4 forced), measured on B200 (us per launch, 24-layer
// pool, CUDA graph), 4096 -> 11008:   M = 4: ck 1 7.82 | ck 2 9.05 | ck 4 17.0     M = 8: 9.40 | 10.39 | 19.3
//                                     M = 16: 16.8 | 18.6 | 36.2                  11008 -> 4096 M = 8: 12.9 | - | 19.2
// Why it loses: (a) a cluster lives inside one GPC and the GPCs of this part do not hold a multiple of 4 SMs each: with one
// CTA per SM, 37 clusters of 4 do not fit at once (two waves); sized to the resident count, the rows of a group no longer
// fit in shared memory; (b) clusters of 2 fit, but the cluster barrier + remote stores + the cluster launch cost more
// than the halved activation traffic saves.
// Decode path for batches of 3..16 tokens: y[M,N] = x[M,K] @ dequant(W)^T (+ bias), the CTA's whole share of W resident
// in shared memory, nibbles turned into fp16 with one LOP3 per two weights and fed to mma.sync m16n8k16 (HMMA).
//
// Why a second decode kernel: the exact-integer kernel (gemv_dec.cu) spends four IMMA columns per batch row, so its
// tensor and operand work grows linearly in M (M = 8: 15 us, M = 16: 28 us at 4096 -> 11008).  Here a weight fragment
// of 16 rows x 16 columns meets EIGHT tokens in one HMMA: per 256 weights one instruction for 16-bit activations, two
// for fp32 ones (fp16 hi + lo parts) -- 16 times fewer tensor instructions per token, HBM-bound up to M = 8.
//
// Arithmetic:
//   * weights: a 32-bit word of a row = nibbles n0..n7 = columns 8j..8j+7 (python/quantize.py:120-122).  w & 0x000f000f
//     is the fp16 pair (n0, n4) * 2^-24 (subnormals), w & 0x00f000f0 = (n1, n5) * 2^-20, the same of w >> 8 = (n2, n6),
//     (n3, n7): four A registers from one SHF + four LOP3, no arithmetic.  The K order inside an MMA is whatever this
//     yields; the B operand (x) is built in the same order, odd columns pre-multiplied by 2^-4;
//   * x: every (warp, pair, token) has its own power-of-two scale 2^e (amax of its 256 columns -> [2^14, 2^15)), fp16
//     hi = rn(x 2^e), and for fp32 inputs lo = rn(x 2^e - hi) in a second n-tile: q * hi and q * lo are exact in the
//     fp32 accumulator, x is represented to 2^-22 relative (2^-39 of the amax for the smallest elements);
//   * y = s * (sum_k q x - zp * sum_k x) + bias, sum_k x in fp32 from the loaded values; partial sums of the 16 warps
//     are folded in a fixed order: results are deterministic (not bit-identical to gemv_dec.cu: fp32 accumulation,
//     max error ~1e-6 relative vs the float64 oracle);
//   * a token whose amax is NaN / Inf is recomputed in the reference's order (w = (q - zp) * s, fp32 FMA), so
//     non-finite inputs propagate exactly like dequantize + F.linear (python/quantize.py:172, 202).
//
// Structure (one CTA per SM, 16 warps):
//   * weights: tile i = 16 rows x K/2 bytes, 3-D TMA boxes [16 rows][chunk pairs][128 B] (pair = 256 columns), 128-byte
//     swizzle, one single-use mbarrier per (pair group, tile), all requested up front (staged around the x loads);
//     ldmatrix.x4 delivers the words of rows g / g + 8 every lane needs;
//   * warp w owns the pairs w, w + 16, ...: lane (g, t) loads the 64 values of token g its B fragments are made of
//     straight from global memory (no exchange through shared memory), the amax / sum of the token's 256 columns is a
//     two-step quad reduction: NO block barrier before the main loop;
//   * per (pair, tile): 4 x (ldmatrix.x4, 4 SHF, 16 LOP3, 4 NT HMMA); the partial tile 16 rows x 8 tokens is descaled
//     and accumulated into the warp's OWN slot of the tile; one barrier, then all threads fold the 16 slots of every
//     output and write y;
//   * M > 8: a second pass over the resident tiles (weights come from HBM once).
//
// Reference being replaced: csrc/quantized_linear_kernel.cu:90-279 (one thread per output, M x weight traffic).
#include <cuda.h>
#include <cmath>
#include "internal.h"
#include "ptx.cuh"
#include "tc.cuh"
#include "dec.cuh"

namespace b200q {

// bench-only (-DB200Q_PROF build, tools/prof_hm.py): per-CTA wall-clock stamps of the phases of the last launch
#ifdef B200Q_PROF
__device__ long long g_hm_prof[256 * 16];
#define HM_STAMP(i)                                                                  \
    do {                                                                             \
        if (p.debug && tid == 0 && blockIdx.x < 256) {                               \
            long long t_;                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                   \
            g_hm_prof[blockIdx.x * 16 + (i)] = t_;                                   \
        }                                                                            \
    } while (0)
#else
#define HM_STAMP(i) ((void)0)
#endif
#ifdef B200Q_PROF
#define HM_ABL(bit) ((p.debug & (bit)) != 0)     // 2: no HMMA, 8: no LDSM
#else
#define HM_ABL(bit) false
#endif

namespace {

constexpr int NW = 16;               // warps per CTA
constexpr int NTHR = NW * 32;
constexpr int TILE_ROWS = 16;
constexpr int PAIR_BYTES = TILE_ROWS * 128;      // one pair (256 columns) of one tile in shared memory
constexpr int MAX_BARS = 64;                     // (pair group, tile) barriers
constexpr int MB = 8;                            // tokens per pass (one n-tile)
constexpr int MAX_TILES = 20;                    // tiles of a CTA (row group of a cluster): <= 320 rows
constexpr int MAX_CK = 4;                        // CTAs of a cluster (K slices)

// shared memory map (bytes)
constexpr int OFF_BARS = 0;          // [MAX_BARS]
constexpr int OFF_FLAG = 512;        // bit m: token m holds NaN / Inf (this CTA's columns)
constexpr int OFF_RFLAG = 528;       // [MAX_CK] the flags of every rank of the cluster (pushed by the ranks)
constexpr int OFF_RSX = 576;         // [2 (pass parity)][MAX_CK][MB] f32: sum_k x of every rank's columns (pushed by the ranks)
constexpr int OFF_SX = 1024;         // [16][MB] f32: sum_k x of a pair owner's columns, per token of the pass
constexpr int OFF_PAR = 2048;        // [3][320] f32: scale, zero point, bias of the CTA's rows (for the fold)
constexpr int PAR_STRIDE = MAX_TILES * TILE_ROWS;
constexpr int OFF_SLOTS = 6144;      // [tile][slot][row][token] f32 partial tiles (512 B each); then the cluster receive
                                     // buffers [2][rank][owned tile][row][token]; then the tiles (1024-aligned)
constexpr int SLOT_BYTES = MB * TILE_ROWS * 4;

struct HmParams {
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
    int ck;                          // CTAs per cluster = K slices (1, 2, 4); consecutive CTAs form a row group
    int rows_q, rows_rem;            // row group b owns rows_q (+1 if b < rows_rem) row units (host-computed: no division in the kernel)
    int npairs;                      // K / 256
    int pp;                          // pairs per K slice (ceil(npairs / ck))
    int gpw;                         // pairs per warp (ceil(pp / 16))
    int wpp_shift;                   // log2 of the warps that share a pair (gpw == 1: they split its tiles)
    int ns;                          // partial-tile slots per tile: min(pp, 16)
    int nbars;                       // pair groups (barriers) per tile
    int chunk;                       // pairs per group
    int tile_bytes;
    int tile_off;                    // byte offset of tile 0 (1024-aligned)
    int recv_off;                    // cluster receive buffers
    int jt;                          // owned tiles per rank (ceil(max tiles / ck))
    int npasses;
    int wait_weights;                // 1: weights may be written by the preceding kernel
    int early_tiles;                 // a + 10 b: a fifths of the requests before griddepcontrol.wait, b more behind the x loads
    int pf_mode;                     // next-layer L2 prefetch: 0 off, 2 before the own requests, else after the first operand build
    int gated;                       // 1: rows 2f / 2f+1 are the gate / up projection of column f; y is h [M, N/2] = silu(gate) * up
    int debug;                       // B200Q_PROF builds: record phase stamps
};

// weight requests [from, to): request op = (pair group, tile) in that order (the main loop walks pairs outside, tiles
// inside), one 3-D box [16 rows][chunk pairs][128 B] each, barrier index = op.  One elected thread; not inlined (three
// call sites).
__device__ __noinline__ void hm_issue(const CUtensorMap* tmap, uint32_t bar0, uint32_t dst0, int row0, int pair0, int from, int to, int S,
                                      int nbars, int chunk, int tile_bytes) {
    const uint64_t pol = policy_evict_first();
    int grp = nbars == 1 ? 0 : from / S, i = from - grp * S;
    for (int op = from; op < to; ++op) {
        const uint32_t bar = bar0 + 8u * (uint32_t)op;
        mbar_arrive_expect_tx(bar, (uint32_t)(chunk * PAIR_BYTES));
        tma_box_3d(dst0 + (uint32_t)(i * tile_bytes + grp * chunk * PAIR_BYTES), tmap, 0, row0 + i * TILE_ROWS, pair0 + grp * chunk, bar, pol);
        if (++i == S) { i = 0; ++grp; }
    }
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// eight consecutive activations as floats; the dtype is a template parameter so that the 8 (16) loads of an operand
// build are straight-line code, all in flight at once (a run-time dtype switch serialises them: one L2 round trip each)
template <int XT>
__device__ __forceinline__ void hm_load8(const void* x, int64_t idx, float2 (&v)[4]) {
    if constexpr (XT == B200Q_F32) {                          // 32-byte aligned (the launcher checks x)
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y)
                     : "l"(static_cast<const float*>(x) + idx));
    } else {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(x) + idx);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (XT == B200Q_F16) v[i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            else v[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        }
    }
}

// XT: activation dtype.  fp32 -> two n-tiles per pass (hi parts, lo parts); else one.
// PK (fp32, M <= 4): ONE n-tile, column 2j = hi part, 2j + 1 = lo part of token j -- half the tensor work of the
// eight-token form.
template <int XT, bool PK>
__global__ void __launch_bounds__(NTHR, 1) gemv_hm_kernel(const __grid_constant__ CUtensorMap tmap, const HmParams p) {
    constexpr bool F32 = XT == B200Q_F32;
    constexpr int NT = (F32 && !PK) ? 2 : 1;
    static_assert(!PK || F32, "the packed form is for fp32 activations");
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    HM_STAMP(0);
    // row group b (all columns of its rows, split over the ck CTAs of a cluster), K slice rank
    const int CK = p.ck;
    const int b = CK == 1 ? (int)blockIdx.x : (CK == 2 ? (int)(blockIdx.x >> 1) : (int)(blockIdx.x >> 2));
    const int rank = (int)blockIdx.x - b * CK;
    // (gated: rows are dealt out in gate / up pairs, so both projections of an output column meet in one CTA)
    const int unit = p.gated ? 2 : 1;
    const int r0 = unit * (b * p.rows_q + min(b, p.rows_rem));           // first weight row of this row group
    const int nrows = unit * (p.rows_q + (b < p.rows_rem ? 1 : 0));
    const int S = (nrows + TILE_ROWS - 1) / TILE_ROWS;                   // tiles of this CTA (all resident)
    const int pair0 = rank * p.pp;                                       // this CTA's pairs: [pair0, pair0 + PP)
    const int PP = min(p.pp, p.npairs - pair0);
    unsigned int* s_flag = reinterpret_cast<unsigned int*>(smem + OFF_FLAG);
    float* s_sx = reinterpret_cast<float*>(smem + OFF_SX);
    float* s_par = reinterpret_cast<float*>(smem + OFF_PAR);
    const int nops = S * p.nbars;

    if (tid == 0) {
        for (int i = 0; i < nops; ++i) mbar_init(sbase + OFF_BARS + 8u * i, 1);
        fence_mbar_init();
        *s_flag = 0u;
    }
    __syncthreads();
    pdl_launch_dependents();
    HM_STAMP(1);

    // ---- this warp's pairs and tiles.  One round (gpw == 1): 2^wpp_shift warps share a pair and take every
    // 2^wpp_shift-th tile of it; several rounds (wide K, no cluster): warp w owns the pairs w, w + 16, ... of every tile.
    const int WPP = 1 << p.wpp_shift;
    const int lp0 = warp >> p.wpp_shift, sub = warp & (WPP - 1);
    const int nwa = min(NW, PP << p.wpp_shift);               // warps that own at least one pair
    const bool issuer = warp == nwa - 1 && lane == 0;
    auto issue = [&](int from, int to) {
        hm_issue(&tmap, sbase + OFF_BARS, sbase + p.tile_off, r0, pair0, from, to, S, p.nbars, p.chunk, p.tile_bytes);
    };
    // weight requests (one elected thread), staged: a fifths before griddepcontrol.wait, b more once the x loads are
    // in flight, the rest when the first operand is built
    const int early = min(((p.early_tiles % 10) * nops + 4) / 5, nops);
    const int mid = min(early + ((p.early_tiles / 10) * nops + 4) / 5, nops);
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
    }

    // ldmatrix row address of this lane for the four 32-byte steps of a pair (lane i supplies row (i & 7) + 8 ((i >> 3) & 1)
    // of the 16-byte column 2 c + (i >> 4); 128-byte swizzle: column ^ row)
    uint32_t offc[4];
    {
        const int ri = lane & 7, mi = lane >> 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) offc[c] = (uint32_t)((ri + 8 * (mi & 1)) * 128 + (((2 * c + (mi >> 1)) ^ ri) << 4));
    }
    // scale / zero point / bias of the CTA's rows (<= 320): fetched early, parked in shared memory for the fold
    float sc = 0.0f, zp = 0.0f, bias = 0.0f;
    if (tid < nrows) {
        sc = __ldg(p.scales + r0 + tid);
        zp = __ldg(p.zps + r0 + tid);
        if (p.bias) bias = __ldg(p.bias + r0 + tid);
    }
    pdl_wait();              // x (and y) belong to the stream-ordered predecessor
    HM_STAMP(2);
    constexpr uint32_t M0 = 0x000f000fu, M1 = 0x00f000f0u;
#pragma unroll 1
    for (int pass = 0; pass < p.npasses; ++pass) {
        const int m0 = pass * MB;
        if (pass > 0) __syncthreads();                        // the fold of the previous pass has read every slot
        const int tk = PK ? (g >> 1) : g;                     // token (of the pass) of this lane's B column
        const bool tok = m0 + tk < p.M;                       // ... exists
        float sxacc = 0.0f;                                   // sum_k x of token tk over this warp's pairs (quad-uniform)
#pragma unroll 1
        for (int q = 0; q < p.gpw; ++q) {
            const int lp = lp0 + NW * q;                      // pair of this round (local index in the K slice)
            if (lp >= PP) break;                              // uniform
            const int P = pair0 + lp;
            const bool first = pass == 0 && q == 0;
            // ---- x of token tk, columns 256 P + 64 c + 32 h + 8 t + (0..7): exactly what this lane's B fragments hold
            // (fp32: one 32-byte load per step -- the four lanes of a quad read one whole 128-byte line per instruction)
            float2 xv[8][4];
            {
                const int64_t base = (int64_t)(m0 + tk) * p.K + P * 256 + t * 8;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) xv[ch][e] = make_float2(0.0f, 0.0f);
                    if (tok) hm_load8<XT>(p.x, base + ch * 32, xv[ch]);
                }
            }
            if (first && issuer && early < mid) issue(early, mid);
            // ---- amax and sum of the token's 256 columns: the four lanes of a quad hold them all.  (fmaxf drops NaN:
            // a NaN shows up in the sum, Inf in the amax)
            float am = 0.0f;
            float2 s2 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    am = fmaxf(am, fmaxf(fabsf(xv[ch][e].x), fabsf(xv[ch][e].y)));
                    s2 = __fadd2_rn(s2, xv[ch][e]);
                }
            float s = s2.x + s2.y;
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            sxacc += s;
            if (first) HM_STAMP(3);
            const int E = (int)(__float_as_uint(am) >> 23);
            if ((E == 255 || s != s) && t == 0) atomicOr(s_flag, 1u << (m0 + tk));
            const int ex = min(126, 141 - E);                 // amax * 2^ex in [2^14, 2^15)
            const float up = __uint_as_float((uint32_t)(127 + ex) << 23);
            const float2 up2 = make_float2(up, up * 0.0625f); // (even column, odd column: its nibble arrives 2^4 too large)
            // descale of the accumulator columns 2t, 2t + 1 (tokens 2t, 2t + 1 of the pass): 2^24 (subnormal nibbles) * 2^-ex
            // (PK: columns 2t, 2t + 1 = hi / lo of token t, whose lanes are g = 2t, 2t + 1: the same scale)
            const float dn = __uint_as_float((uint32_t)(127 - ex) << 23);
            const float d0 = __shfl_sync(0xffffffffu, dn, 8 * t), d1 = __shfl_sync(0xffffffffu, dn, 8 * t + 4);

            // ---- B fragments: per 32-byte step c, half h: MMA alpha = nibbles (0,4 | 1,5), beta = (2,6 | 3,7) of every word
            uint32_t bf[8][2][NT][2];                         // [2 c + h][alpha / beta][hi / lo][b0, b1]
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float2 a[4];                                  // (x0, x1 / 16) (x2, x3 / 16) (x4, x5 / 16) (x6, x7 / 16), scaled
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e] = __fmul2_rn(xv[ch][e], up2);
                bf[ch][0][0][0] = pack_h2(a[0].x, a[2].x); bf[ch][0][0][1] = pack_h2(a[0].y, a[2].y);
                bf[ch][1][0][0] = pack_h2(a[1].x, a[3].x); bf[ch][1][0][1] = pack_h2(a[1].y, a[3].y);
                if constexpr (F32) {
                    const float2 h04 = unpack_h2(bf[ch][0][0][0]), h15 = unpack_h2(bf[ch][0][0][1]);
                    const float2 h26 = unpack_h2(bf[ch][1][0][0]), h37 = unpack_h2(bf[ch][1][0][1]);
                    const uint32_t l0 = pack_h2(a[0].x - h04.x, a[2].x - h04.y), l1 = pack_h2(a[0].y - h15.x, a[2].y - h15.y);
                    const uint32_t l2 = pack_h2(a[1].x - h26.x, a[3].x - h26.y), l3 = pack_h2(a[1].y - h37.x, a[3].y - h37.y);
                    if constexpr (PK) {
                        if (g & 1) { bf[ch][0][0][0] = l0; bf[ch][0][0][1] = l1; bf[ch][1][0][0] = l2; bf[ch][1][0][1] = l3; }
                    } else {
                        bf[ch][0][1][0] = l0; bf[ch][0][1][1] = l1; bf[ch][1][1][0] = l2; bf[ch][1][1][1] = l3;
                    }
                }
            }
            if (first && issuer) {
                if (mid < nops) issue(mid, nops);
                if (p.pf_mode != 0 && p.pf_mode != 2) prefetch_next();
            }
            if (first) HM_STAMP(4);
            int grp = 0;
            if (p.nbars > 1) grp = lp / p.chunk;
            const int sl = p.gpw == 1 ? lp : warp;            // partial-tile slot of this warp's contribution

            // ---- this pair of this warp's tiles
#pragma unroll 1
            for (int i = sub; i < S; i += WPP) {
                if (pass == 0) mbar_wait(sbase + OFF_BARS + 8u * (uint32_t)(grp * S + i), 0u);
                const uint32_t pb = sbase + p.tile_off + (uint32_t)(i * p.tile_bytes + lp * PAIR_BYTES);
                float acc[2][NT][4];                          // two chains (alpha, beta) per n-tile
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[j][nt][r] = 0.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t a[4];
                    if (!HM_ABL(8)) ldsm_x4(a, pb + offc[c]); else { a[0] = c; a[1] = lane; a[2] = i; a[3] = 7; }                // a0 / a1: rows g / g + 8, bytes 32 c + 4 t ..; a2 / a3: + 16 bytes
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w0 = a[2 * h], w1 = a[2 * h + 1];
                        const uint32_t v0 = w0 >> 8, v1 = w1 >> 8;
                        if (!HM_ABL(2)) for (int nt = 0; nt < NT; ++nt) {
                            mma_m16n8k16_f16(acc[0][nt], w0 & M0, w1 & M0, w0 & M1, w1 & M1, bf[2 * c + h][0][nt][0], bf[2 * c + h][0][nt][1]);
                            mma_m16n8k16_f16(acc[1][nt], v0 & M0, v1 & M0, v0 & M1, v1 & M1, bf[2 * c + h][1][nt][0], bf[2 * c + h][1][nt][1]);
                        }
                        else { acc[0][0][0] += __uint_as_float((w0 & M0) ^ (w1 & M1) ^ (v0 & M0) ^ (v1 & M1) ^ (w0 & M1) ^ (w1 & M0) ^ (v0 & M1) ^ (v1 & M0)); }
                    }
                }
                // partial tile -> this warp's own slot of the tile, descaled: q was 2^-24 too small
                float v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float a = acc[0][0][r] + acc[1][0][r];
                    if constexpr (NT == 2) a += acc[0][1][r] + acc[1][1][r];
                    v[r] = (a * 16777216.0f) * ((r & 1) ? d1 : d0);
                }
                if constexpr (PK) { v[0] += v[1]; v[2] += v[3]; }      // hi + lo column of token t
                // slot layout [row][token]: lane (g, t) owns tokens 2t, 2t + 1 of rows g and g + 8 -- two conflict-free 8-byte stores
                float2* slot = reinterpret_cast<float2*>(smem + OFF_SLOTS + (i * p.ns + sl) * SLOT_BYTES);
                if constexpr (PK) {                           // token t of rows g, g + 8 (tokens 4..7 of the slot are never read for a result)
                    float* s1 = reinterpret_cast<float*>(slot);
                    if (q > 0) { v[0] += s1[g * 8 + t]; v[2] += s1[(g + 8) * 8 + t]; }
                    s1[g * 8 + t] = v[0];
                    s1[(g + 8) * 8 + t] = v[2];
                } else {
                    float2 lo2 = make_float2(v[0], v[1]), hi2 = make_float2(v[2], v[3]);
                    if (q > 0) {
                        const float2 o0 = slot[g * 4 + t], o1 = slot[(g + 8) * 4 + t];
                        lo2.x += o0.x; lo2.y += o0.y; hi2.x += o1.x; hi2.y += o1.y;
                    }
                    slot[g * 4 + t] = lo2;
                    slot[(g + 8) * 4 + t] = hi2;
                }
                if (first && i < 5) HM_STAMP(5 + i);
            }
        }
        if (pass == 0) HM_STAMP(10);
        if (t == 0 && sub == 0 && lp0 < PP) s_sx[(p.gpw == 1 ? lp0 : warp) * MB + (PK ? 4 * (g & 1) + tk : g)] = (PK && (g & 1)) ? 0.0f : sxacc;
        if (pass == 0 && tid < PAR_STRIDE) { s_par[tid] = sc; s_par[PAR_STRIDE + tid] = zp; s_par[2 * PAR_STRIDE + tid] = bias; }
        __syncthreads();
        if (pass == 0) HM_STAMP(11);

        // ---- fold the partial tiles of this CTA.  Thread = (tile, row, token quad tq, half of the slots): 16-byte loads
        // of 4 tokens of its row + of their sum_k x, the two halves meet in one shuffle step -- a fixed order:
        // deterministic.  Lanes 0..7 of a quarter warp read 128 consecutive bytes: no bank conflicts.
        const int NSL = p.gpw == 1 ? PP : NW;                 // slots that hold a contribution
        const int tq = tid & 1, half = (tid >> 3) & 1;
        const int row = ((tid >> 1) & 3) + 4 * ((tid >> 4) & 3);
        const int par = pass & 1;
        auto emit = [&](int tile, float4 a, float4 s, unsigned int flagged) {      // final values of (tile, row, tokens 4 tq ..): scale, store
            const int rowpos = tile * TILE_ROWS + row;
            const bool mine = rowpos < nrows && half == 0;
            const float rsc = s_par[rowpos], rzp = s_par[PAR_STRIDE + rowpos], rbias = s_par[2 * PAR_STRIDE + rowpos];
            float vv[4] = {rsc * fmaf(-rzp, s.x, a.x), rsc * fmaf(-rzp, s.y, a.y), rsc * fmaf(-rzp, s.z, a.z), rsc * fmaf(-rzp, s.w, a.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + tq * 4 + j;
                if (p.gated) {
                    // fused gate + up pair: the even row (gate) fetches its neighbour's value (up, lane ^ 2) and writes silu(gate) * up
                    const float uu = __shfl_xor_sync(0xffffffffu, vv[j], 2);
                    if (mine && !(row & 1) && m < p.M && !((flagged >> m) & 1u))
                        store_out(p.y, p.y_dtype, (int64_t)m * (p.N >> 1) + ((r0 + rowpos) >> 1), vv[j] / (1.0f + __expf(-vv[j])) * uu);
                } else if (mine && m < p.M && !((flagged >> m) & 1u)) {
                    store_out(p.y, p.y_dtype, (int64_t)m * p.N + r0 + rowpos, vv[j] + rbias);
                }
            }
        };
        // sum_k x of the 4 tokens over this CTA's pairs
        float4 sloc = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const float4* ssx = reinterpret_cast<const float4*>(s_sx) + tq;
#pragma unroll 4
            for (int w = half; w < NSL; w += 2) {
                const float4 sv = ssx[w * (MB / 4)];
                sloc.x += sv.x; sloc.y += sv.y; sloc.z += sv.z; sloc.w += sv.w;
            }
            sloc.x += __shfl_xor_sync(0xffffffffu, sloc.x, 8); sloc.y += __shfl_xor_sync(0xffffffffu, sloc.y, 8);
            sloc.z += __shfl_xor_sync(0xffffffffu, sloc.z, 8); sloc.w += __shfl_xor_sync(0xffffffffu, sloc.w, 8);
        }
        const uint32_t recv = sbase + (uint32_t)p.recv_off + (uint32_t)(par * CK * p.jt) * SLOT_BYTES;
#pragma unroll 1
        for (int tile = tid >> 6; tile < S; tile += 8) {      // uniform per warp
            const float4* src = reinterpret_cast<const float4*>(smem + OFF_SLOTS + (tile * p.ns) * SLOT_BYTES) + row * 2 + tq;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int w = half; w < NSL; w += 2) {
                const float4 av = src[w * (SLOT_BYTES / 16)];
                a.x += av.x; a.y += av.y; a.z += av.z; a.w += av.w;
            }
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 8); a.y += __shfl_xor_sync(0xffffffffu, a.y, 8);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, 8); a.w += __shfl_xor_sync(0xffffffffu, a.w, 8);
            if (CK == 1) {
                emit(tile, a, sloc, *s_flag);
            } else if (half == 0) {
                // K slices: hand the tile to its owner (rank tile % ck, its tile tile / ck) -- asynchronous remote stores
                const int owner = tile & (CK - 1), j = CK == 2 ? (tile >> 1) : (tile >> 2);
                st_cluster_v4(mapa(recv + (uint32_t)((rank * p.jt + j) * SLOT_BYTES + row * 32 + tq * 16), (uint32_t)owner), a.x, a.y, a.z, a.w);
            }
        }
        if (pass == 0) HM_STAMP(13);
        if (CK > 1) {
            if (tid < 2 * CK) {                               // this rank's sum_k x (tokens 4 tq ..) and flags to every rank
                const int dst = tid >> 1;
                st_cluster_v4(mapa(sbase + OFF_RSX + (uint32_t)(((par * MAX_CK + rank) * MB + tq * 4) * 4), (uint32_t)dst), sloc.x, sloc.y, sloc.z, sloc.w);
                if (tq == 0) st_cluster_u32(mapa(sbase + OFF_RFLAG + 4u * (uint32_t)rank, (uint32_t)dst), *s_flag);
            }
            cluster_sync_all();
            unsigned int flagged = 0u;
            float4 stot = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < CK; ++r) {                    // fixed order: every rank computes the same sums
                flagged |= *reinterpret_cast<const unsigned int*>(smem + OFF_RFLAG + 4 * r);
                const float4 sv = *reinterpret_cast<const float4*>(smem + OFF_RSX + ((par * MAX_CK + r) * MB + tq * 4) * 4);
                stot.x += sv.x; stot.y += sv.y; stot.z += sv.z; stot.w += sv.w;
            }
#pragma unroll 1
            for (int j = tid >> 6; j * CK + rank < S; j += 8) {       // owned tiles, uniform per warp
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < CK; ++r) {
                    const float4 av = lds128f(recv + (uint32_t)((r * p.jt + j) * SLOT_BYTES + row * 32 + tq * 16));
                    a.x += av.x; a.y += av.y; a.z += av.z; a.w += av.w;
                }
                emit(j * CK + rank, a, stot, flagged);
            }
        }
        if (pass < 1) HM_STAMP(12 + pass);
    }

    HM_STAMP(14);
    // ---- tokens with NaN / Inf: the reference's arithmetic (dequantise, then fp32 multiply-add), so that non-finite
    // values propagate as in F.linear; one warp per output, weights re-read from global memory.  (K slices: every
    // rank recomputes the rows of the tiles it owns, over all of K)
    unsigned int flagged = *s_flag;
    if (CK > 1) {
        flagged = 0u;
        for (int r = 0; r < CK; ++r) flagged |= *reinterpret_cast<const unsigned int*>(smem + OFF_RFLAG + 4 * r);
    }
    if (flagged) {
        const int64_t row_bytes = p.K >> 1;
        for (int m = 0; m < p.M; ++m) {
            if (!((flagged >> m) & 1u)) continue;
            auto ref_row = [&](int rw) {
                const float rs = __ldg(p.scales + rw), rz = __ldg(p.zps + rw);
                const uint8_t* wr = p.packed + (int64_t)rw * row_bytes;
                float acc = 0.0f;
                for (int kb = lane; kb < row_bytes; kb += 32) {
                    const unsigned int byte = wr[kb];
                    const float w0 = ((float)(byte & 15u) - rz) * rs, w1 = ((float)(byte >> 4) - rz) * rs;
                    acc = fmaf(w0, load1f(p.x, p.x_dtype, (int64_t)m * p.K + 2 * kb), acc);
                    acc = fmaf(w1, load1f(p.x, p.x_dtype, (int64_t)m * p.K + 2 * kb + 1), acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                return acc;
            };
            for (int rc = unit * warp; rc < nrows; rc += unit * NW) {
                if (((rc >> 4) & (CK - 1)) != rank) continue;        // uniform: a tile of another rank
                const int rw = r0 + rc;
                float acc = ref_row(rw);
                if (p.gated) {
                    const float upv = ref_row(rw + 1);
                    if (lane == 0) store_out(p.y, p.y_dtype, (int64_t)m * (p.N >> 1) + (rw >> 1), acc / (1.0f + __expf(-acc)) * upv);
                } else if (lane == 0) {
                    if (p.bias) acc += __ldg(p.bias + rw);
                    store_out(p.y, p.y_dtype, (int64_t)m * p.N + rw, acc);
                }
            }
        }
    }
}

struct HmPlan {
    int grid, ck, rows_q, rows_rem, s_max, npairs, pp, gpw, wpp_shift, ns, nbars, chunk, tile_bytes, tile_off, recv_off, jt, npasses;
    size_t smem;
};

// Clusters that can be resident at once (one CTA per SM, all shared memory): a cluster lives inside one GPC, and the
// GPCs of a B200 do not hold a multiple of 4 SMs each -- 148 CTAs in clusters of 4 run in two waves.  Asked once.
template <int XT, bool PK>
__global__ void gemv_hm_kernel(const __grid_constant__ CUtensorMap tmap, const HmParams p);
int hm_max_clusters(int sm_count, int max_smem, int ck) {
    static int cache[MAX_CK + 1] = {0, 0, 0, 0, 0};
    if (ck <= 1) return sm_count;
    if (cache[ck] > 0) return cache[ck];
    auto kfn = gemv_hm_kernel<B200Q_F32, false>;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sm_count / ck * ck));
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = (size_t)max_smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)ck; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    cache[ck] = n;
    return n;
}

// one candidate: ck K slices per row group
bool plan_hm_ck(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, int gated, int ck, HmPlan* c) {
    const int unit = gated ? 2 : 1;
    c->ck = ck;
    c->npairs = (int)(K / 256);
    c->pp = (c->npairs + ck - 1) / ck;
    if (ck > 1 && (c->pp > NW || c->pp * (ck - 1) >= c->npairs)) return false;      // one round per K slice, no empty slice
    c->gpw = (c->pp + NW - 1) / NW;                              // <= 4
    c->wpp_shift = 0;
    if (c->gpw == 1) while ((c->pp << (c->wpp_shift + 1)) <= NW) ++c->wpp_shift;
    c->ns = c->gpw == 1 ? c->pp : NW;
    c->nbars = c->gpw;
    c->chunk = (c->pp + c->nbars - 1) / c->nbars;
    c->tile_bytes = c->nbars * c->chunk * PAIR_BYTES;            // >= pp * 2 KB: a 3-D box always has room
    c->npasses = (int)((M + MB - 1) / MB);
    int cap = tuning().gemv_ctas > 0 ? tuning().gemv_ctas : sm_count;
    if (cap > sm_count) cap = sm_count;
    const int64_t units = N / unit;
    int64_t g = (N + TILE_ROWS - 1) / TILE_ROWS;                 // few rows: one tile per CTA
    if (g > cap) g = cap;
    if (g > units) g = units;
    if (g < 1) g = 1;
    if (ck > 1 && g < cap) return false;                         // K slices only when every SM has work
    int64_t groups = g / ck;
    if (ck > 1) {                                                // ... and only as many clusters as are resident at once
        const int mc = hm_max_clusters(sm_count, max_smem, ck);
        if (mc < 1) return false;
        if (groups > mc) groups = mc;
        g = groups * ck;
    }
    c->grid = (int)g;
    c->rows_q = (int)(units / groups); c->rows_rem = (int)(units % groups);
    const int64_t br = unit * ((units + groups - 1) / groups);   // most rows of a row group
    const int64_t S = (br + TILE_ROWS - 1) / TILE_ROWS;
    if (S > MAX_TILES || S * c->nbars > MAX_BARS) return false;
    c->s_max = (int)S;
    c->jt = (int)((S + ck - 1) / ck);
    c->recv_off = OFF_SLOTS + (int)S * c->ns * SLOT_BYTES;
    const int recv_bytes = ck > 1 ? 2 * ck * c->jt * SLOT_BYTES : 0;
    c->tile_off = (c->recv_off + recv_bytes + 1023) / 1024 * 1024;
    c->smem = (size_t)c->tile_off + (size_t)S * c->tile_bytes;
    return c->smem <= (size_t)max_smem;
}

bool plan_hm(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, int gated, HmPlan* c) {
    if (M < 1 || M > 16 || K < 256 || K % 256 != 0 || K > 16384 || N < 1 || N > 0x3fffffff) return false;
    if (gated && (N & 1)) return false;
    // K slices over a cluster cut the activation traffic per SM (every CTA of an N split needs all of x: M K 4 bytes
    // against N K / (2 SMs) bytes of weights) and give every warp one pair: measured faster from ... (tools/dec_tune.py)
    const int want = tuning().hm_ck;
    if (want > 0) return plan_hm_ck(sm_count, max_smem, M, N, K, gated, want, c);
    const int64_t xbytes = M * K * 4, wbytes = N * K / 2 / (sm_count > 0 ? sm_count : 1);
    if (2 * xbytes >= wbytes && plan_hm_ck(sm_count, max_smem, M, N, K, gated, 4, c)) return true;
    return plan_hm_ck(sm_count, max_smem, M, N, K, gated, 1, c) || plan_hm_ck(sm_count, max_smem, M, N, K, gated, 4, c) ||
           plan_hm_ck(sm_count, max_smem, M, N, K, gated, 2, c);
}

template <int XT, bool PK>
int launch_hm_inst(const HmPlan& c, const CUtensorMap& map, const HmParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_hm_kernel<XT, PK>;
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
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (pdl) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (c.ck > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = (unsigned)c.ck;
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, map, p), "gemv_hm launch");
}

}  // namespace

bool gemv_hm_supported(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated) {
    HmPlan c;
    return plan_hm(dev.sm_count, dev.max_smem_optin, M, N, K, gated, &c);
}

int launch_gemv_hm(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                   const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                   unsigned flags, cudaStream_t st, const uint8_t* next_packed, size_t next_bytes, int gated) {
    HmPlan c;
    if (!plan_hm(dev.sm_count, dev.max_smem_optin, M, N, K, gated, &c))
        return set_error(B200Q_EINVAL, "gemv_hm: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & (x_dtype == B200Q_F32 ? 31 : 15)) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv_hm: packed must be 16-byte aligned, x 32-byte (fp32) / 16-byte aligned");
    HmParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.bias = gated ? nullptr : bias; p.y = y;
    p.gated = gated;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.ck = c.ck;
    p.rows_q = c.rows_q; p.rows_rem = c.rows_rem;
    p.npairs = c.npairs; p.pp = c.pp; p.gpw = c.gpw; p.wpp_shift = c.wpp_shift; p.ns = c.ns; p.nbars = c.nbars; p.chunk = c.chunk;
    p.tile_bytes = c.tile_bytes; p.tile_off = c.tile_off; p.recv_off = c.recv_off; p.jt = c.jt; p.npasses = c.npasses;
    p.wait_weights = (flags & B200Q_FLAG_STATIC_WEIGHTS) ? 0 : 1;
    p.early_tiles = tuning().gemv_early >= 0 ? tuning().gemv_early : 92;
    p.next_packed = next_packed;
    p.pf_mode = tuning().gemv_pf;
    p.next_bytes = (tuning().gemv_pf != 0 && next_packed && (reinterpret_cast<uintptr_t>(next_packed) & 15) == 0) ? next_bytes : 0;
    p.next_chunk = (unsigned int)(p.next_bytes / (unsigned long long)c.grid);
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    CUtensorMap map;
    if (int rc = dec_weight_map(packed, N, K, c.chunk, &map)) return rc;
    const bool pdl = tuning().gemv_pdl != 0;
    if (x_dtype == B200Q_F32)
        return (M <= 4 && tuning().hm_packed != 0) ? launch_hm_inst<B200Q_F32, true>(c, map, p, pdl, st) : launch_hm_inst<B200Q_F32, false>(c, map, p, pdl, st);
    if (x_dtype == B200Q_F16) return launch_hm_inst<B200Q_F16, false>(c, map, p, pdl, st);
    return launch_hm_inst<B200Q_BF16, false>(c, map, p, pdl, st);
}

}  // namespace b200q

#ifdef B200Q_PROF
extern "C" int b200q_debug_read_prof_hm(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_hm_prof, sizeof(long long) * 256 * 16), "read prof");
}
#endif
