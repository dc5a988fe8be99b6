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
constexpr int HM_MAX_M = 32;                     // four passes (the non-finite flags are one 32-bit word); grouped: 16 rows per expert
constexpr int MAX_TILES = 8;                     // 128 row positions: one per thread of a quarter of the CTA in the fold

// shared memory map (bytes)
constexpr int OFF_BARS = 0;          // [MAX_BARS]
constexpr int OFF_FLAG = 512;        // bit m: token m holds NaN / Inf
constexpr int OFF_SX = 1024;         // [NW][MB] f32: sum_k x of the warp's columns, per token of the pass
constexpr int OFF_PAR = 2048;        // [3][128] f32: scale, zero point, bias of the CTA's rows (for the fold)
constexpr int OFF_SLOTS = 4096;      // [tile][warp][row][token] f32 partial tiles (512 B each), then the tiles (1024-aligned)
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
    int rows_q, rows_rem;            // CTA b owns rows_q (+1 if b < rows_rem) row units (host-computed: no division in the kernel)
    int npairs;                      // K / 256
    int gpw;                         // pairs per warp (ceil(npairs / 16))
    int nbars;                       // pair groups (barriers) per tile
    int chunk;                       // pairs per group
    int tile_bytes;
    int tile_off;                    // byte offset of tile 0 (1024-aligned)
    int npasses;
    int wait_weights;                // 1: weights may be written by the preceding kernel
    int early_tiles;                 // a + 10 b: a requests before griddepcontrol.wait, b more behind the x loads
    int pf_mode;                     // next-layer L2 prefetch: 0 off, 2 before the own requests, else after the first operand build
    int gated;                       // 1: rows 2f / 2f+1 are the gate / up projection of column f; y is h [M, N/2] = silu(gate) * up
    // grouped (MoE decode, GRP instances): blockIdx.y = expert of packed [E, N, K/2]; rows [offsets[e], offsets[e+1]) of y
    // belong to it (device memory, <= 16 rows; experts without rows exit at once); row r of the group reads
    // x[row_map[offsets[e] + r]] (x is read in place, no gathered copy) or x[offsets[e] + r] when row_map is null
    const int32_t* offsets;
    const int32_t* row_map;
    int kgroup, ngroups;             // group-wise scales (GS instances): scales / zps are [N][K / kgroup], kgroup % 128 == 0
    int debug;                       // B200Q_PROF builds: record phase stamps
};

// weight requests [from, to): request op = (pair group, tile) in that order (the main loop walks pairs outside, tiles
// inside), one 3-D box [16 rows][chunk pairs][128 B] each, barrier index = op.  One elected thread; not inlined (three
// call sites).
__device__ __noinline__ void hm_issue(const CUtensorMap* tmap, uint32_t bar0, uint32_t dst0, int row0, int from, int to, int S,
                                      int nbars, int chunk, int tile_bytes) {
    const uint64_t pol = policy_evict_first();
    int grp = nbars == 1 ? 0 : from / S, i = from - grp * S;
    for (int op = from; op < to; ++op) {
        const uint32_t bar = bar0 + 8u * (uint32_t)op;
        mbar_arrive_expect_tx(bar, (uint32_t)(chunk * PAIR_BYTES));
        tma_box_3d(dst0 + (uint32_t)(i * tile_bytes + grp * chunk * PAIR_BYTES), tmap, 0, row0 + i * TILE_ROWS, grp * chunk, bar, pol);
        if (++i == S) { i = 0; ++grp; }
    }
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// c0 + 256 c1 + 65536 c2 of three s32 digit sums (|c| < 2^22) as a float: every digit sum is converted exactly by the
// magic-number add (int -> float on the integer and FMA pipes; I2F runs on the quarter-rate XU pipe, 12 per warp and tile)
__device__ __forceinline__ float digits_to_float(int c0, int c1, int c2) {
    const float f0 = __int_as_float(c0 + 0x4B400000) - 12582912.0f;
    const float f1 = __int_as_float(c1 + 0x4B400000) - 12582912.0f;
    const float f2 = __int_as_float(c2 + 0x4B400000) - 12582912.0f;
    return fmaf(f2, 65536.0f, fmaf(f1, 256.0f, f0));
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

// XT: activation dtype.
// FORM 0 (HMMA): nibbles -> fp16 subnormals, x in fp16: one n-tile per pass for 16-bit activations, two (hi parts, lo
// parts) for fp32 ones.
// FORM 2 (I3, fp32; the default for fp32): x as a 22-bit fixed-point number per (warp, pair, token), cut into three
// base-256 digits = three n-tiles of IMMA m16n8k32 (u8 x u8, top digit u8 x s8); the nibbles are widened to bytes
// (w & 0x0f0f0f0f, (w >> 4) & 0x0f0f0f0f).  Three tensor + six ALU instructions per 512 weights against four + ten of
// the hi / lo HMMA form: the main loop is bound by exactly those (M = 8: 8.8 -> 8.0 us, M = 16: 15.7 -> 13.6 us).
// GS: group-wise scales.  The two 128-column halves of a pair are accumulated separately and meet their own scale / zero
// point of every row in the main-loop epilogue; the slots then hold finished partial outputs and the fold only adds.
template <int XT, int FORM, bool GRP, bool GS>
__global__ void __launch_bounds__(NTHR, 1) gemv_hm_kernel(const __grid_constant__ CUtensorMap tmap, const HmParams p) {
    constexpr bool F32 = XT == B200Q_F32;
    constexpr bool I3 = FORM == 2;
    constexpr int NT = F32 ? 2 : 1;
    static_assert(FORM == 0 || (FORM == 2 && F32), "the integer form is for fp32 activations");
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    HM_STAMP(0);
    const int b = (int)blockIdx.x;
    // (gated: rows are dealt out in gate / up pairs, so both projections of an output column meet in one CTA)
    const int unit = p.gated ? 2 : 1;
    const int r0 = unit * (b * p.rows_q + min(b, p.rows_rem));           // first weight row of this CTA
    const int nrows = unit * (p.rows_q + (b < p.rows_rem ? 1 : 0));
    const int S = (nrows + TILE_ROWS - 1) / TILE_ROWS;                   // tiles of this CTA (all resident)
    // grouped: this CTA's expert, its token rows and where they start in y (and in x, through the row map)
    int Mrows = p.M, npasses = p.npasses, xrow0 = 0, wrow0 = r0;         // wrow0: first row in the stacked weight tensor
    if constexpr (GRP) {
        const int expert = (int)blockIdx.y;
        const int lo = p.offsets[expert], hi = p.offsets[expert + 1];
        Mrows = min(hi - lo, 2 * MB);
        if (Mrows <= 0) return;                                          // uniform: this expert has no tokens
        npasses = (Mrows + MB - 1) / MB;
        xrow0 = lo;
        wrow0 = expert * p.N + r0;
    }
    auto x_row = [&](int m) -> int64_t {                                 // row of x that token row m of this CTA reads
        if constexpr (GRP) { return p.row_map ? (int64_t)p.row_map[xrow0 + m] : (int64_t)(xrow0 + m); }
        return (int64_t)m;
    };
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

    // ---- weight requests (lane 0 of the last warp), staged: a requests before griddepcontrol.wait, b more once the x
    // loads are in flight, the rest when the first operand is built
    const int nwa = min(NW, p.npairs);                        // warps that own at least one pair
    const bool issuer = warp == nwa - 1 && lane == 0;
    auto issue = [&](int from, int to) {
        hm_issue(&tmap, sbase + OFF_BARS, sbase + p.tile_off, wrow0, from, to, S, p.nbars, p.chunk, p.tile_bytes);
    };
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
    if constexpr (GS) {
        // the scales / zero points of this CTA's rows (one contiguous block of nrows * ngroups floats each) -> L2: the main
        // loop asks for eight of them per (pair, tile) and needs them ~0.2 us later -- from DRAM, under the weight
        // stream, every tile would wait for them
        const size_t bytes = (size_t)nrows * (size_t)p.ngroups * 4;
        const char* base[2] = {reinterpret_cast<const char*>(p.scales + (int64_t)wrow0 * p.ngroups),
                               reinterpret_cast<const char*>(p.zps + (int64_t)wrow0 * p.ngroups)};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const size_t lead = reinterpret_cast<uintptr_t>(base[a]) & 127;          // first line starts `lead` bytes before the block
            for (size_t off = (size_t)tid * 128; off < bytes + lead; off += (size_t)NTHR * 128) {
                const char* q = base[a] + (off > lead ? off - lead : 0);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q) : "memory");
            }
        }
    }
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
    // scale / zero point / bias of the CTA's rows: fetched early by the first 128 threads, parked in shared memory for the fold
    float sc = 0.0f, zp = 0.0f, bias = 0.0f;
    if (tid < nrows) {                                        // nrows <= 128
        if constexpr (GS) { sc = 1.0f; }                      // (applied per group in the main loop)
        else {
            sc = __ldg(p.scales + wrow0 + tid);
            zp = __ldg(p.zps + wrow0 + tid);
        }
        if (p.bias) bias = __ldg(p.bias + wrow0 + tid);
    }
    if (warp >= nwa) {                                        // K < 4096: a warp without pairs contributes zeros to the fold
        for (int i = 0; i < S; ++i) sts128(sbase + OFF_SLOTS + (uint32_t)((i * NW + warp) * SLOT_BYTES + lane * 16), make_uint4(0u, 0u, 0u, 0u));
    }
    pdl_wait();              // x (and y) belong to the stream-ordered predecessor
    HM_STAMP(2);
    constexpr uint32_t M0 = 0x000f000fu, M1 = 0x00f000f0u;
#pragma unroll 1
    for (int pass = 0; pass < npasses; ++pass) {
        const int m0 = pass * MB;
        if (pass > 0) __syncthreads();                        // the fold of the previous pass has read every slot
        const int tk = g;                                     // token (of the pass) of this lane's B column
        const bool tok = m0 + tk < Mrows;                     // ... exists
        float sxacc = 0.0f;                                   // sum_k x of token g over this warp's pairs (quad-uniform)
#pragma unroll 1
        for (int q = 0; q < p.gpw; ++q) {
            const int P = warp + NW * q;
            if (P >= p.npairs) break;                         // uniform
            const bool first = pass == 0 && q == 0;
            // ---- x of token g, columns 256 P + 64 c + 32 h + 8 t + (0..7): exactly what this lane's B fragments hold
            // (fp32: one 32-byte load per step -- the four lanes of a quad read one whole 128-byte line per instruction)
            float2 xv[8][4];
            {
                const int64_t base = (tok ? x_row(m0 + tk) : 0) * p.K + P * 256 + t * 8;
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
            float2 s2 = make_float2(0.0f, 0.0f), s2b = make_float2(0.0f, 0.0f);      // (GS: columns 0..127 / 128..255 apart)
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    am = fmaxf(am, fmaxf(fabsf(xv[ch][e].x), fabsf(xv[ch][e].y)));
                    if (GS && ch >= 4) s2b = __fadd2_rn(s2b, xv[ch][e]);
                    else s2 = __fadd2_rn(s2, xv[ch][e]);
                }
            float s = s2.x + s2.y, sb = s2b.x + s2b.y;
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            float sxg[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};   // GS: sum_k x of tokens 2t, 2t + 1 over the two halves of the pair
            if constexpr (GS) {
                sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                sb += __shfl_xor_sync(0xffffffffu, sb, 2);
                sxg[0][0] = __shfl_sync(0xffffffffu, s, 8 * t);  sxg[0][1] = __shfl_sync(0xffffffffu, s, 8 * t + 4);
                sxg[1][0] = __shfl_sync(0xffffffffu, sb, 8 * t); sxg[1][1] = __shfl_sync(0xffffffffu, sb, 8 * t + 4);
                s += sb;                                      // (NaN detection below)
            } else {
                sxacc += s;
            }
            if (first) HM_STAMP(3);
            const int E = (int)(__float_as_uint(am) >> 23);
            if ((E == 255 || s != s) && t == 0) atomicOr(s_flag, 1u << (m0 + tk));
            const int ex = min(126, (I3 ? 148 : 141) - E);    // amax * 2^ex in [2^14, 2^15)   (I3: [2^21, 2^22))
            const float up = __uint_as_float((uint32_t)(127 + ex) << 23);
            const float2 up2 = make_float2(up, up * 0.0625f); // (even column, odd column: its nibble arrives 2^4 too large)
            // descale of the accumulator columns 2t, 2t + 1 (tokens 2t, 2t + 1 of the pass): 2^24 (subnormal nibbles) * 2^-ex
            const float dn = __uint_as_float((uint32_t)(127 - ex) << 23);
            const float d0 = __shfl_sync(0xffffffffu, dn, 8 * t), d1 = __shfl_sync(0xffffffffu, dn, 8 * t + 4);

            // ---- B fragments
            uint32_t bi[I3 ? 8 : 1][3][2];                    // I3: [2 c + h][digit][even columns, odd columns]
            uint32_t bf[I3 ? 1 : 8][2][NT][2];                // HMMA: [2 c + h][alpha / beta][hi / lo][b0, b1]
            if constexpr (I3) {
                // X = rn(x 2^ex) + 2^22 sits in the low 24 bits of fma(x, 2^ex, 1.5 * 2^23): bytes 0, 1 = unsigned digits,
                // byte 2 - 64 = signed top digit.  Columns 2j (even) meet the low nibbles, 2j + 1 the high nibbles.
                const float2 upi = make_float2(up, up), magic = make_float2(12582912.0f, 12582912.0f);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t ev[4], od[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = __ffma2_rn(xv[ch][e], upi, magic);
                        ev[e] = __float_as_uint(a.x); od[e] = __float_as_uint(a.y);
                    }
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        const uint32_t* v = par ? od : ev;
                        const uint32_t t01 = __byte_perm(v[0], v[1], 0x5140), t23 = __byte_perm(v[2], v[3], 0x5140);
                        const uint32_t u01 = __byte_perm(v[0], v[1], 0x0062), u23 = __byte_perm(v[2], v[3], 0x0062);
                        bi[ch][0][par] = __byte_perm(t01, t23, 0x5410);
                        bi[ch][1][par] = __byte_perm(t01, t23, 0x7632);
                        bi[ch][2][par] = (__byte_perm(u01, u23, 0x5410) + 0x40404040u) ^ 0x80808080u;       // byte - 64 as s8
                    }
                }
            } else {
            // per 32-byte step c, half h: MMA alpha = nibbles (0,4 | 1,5), beta = (2,6 | 3,7) of every word
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
            if (p.nbars > 1) grp = P / p.chunk;
            // GS: lane (g, t) fetches the scale / zero point of ONE (half of the pair, row g / g + 8) combination -- half = t >> 1,
            // row = g + 8 (t & 1) -- one tile ahead (the values come out of L2 while the previous tile's MMAs run) and the
            // quad exchanges them by shuffle in the tile epilogue: 4 live registers instead of 8, no exposed load latency
            float csc = 0.0f, czp = 0.0f;
            int64_t gs_off = 0;                               // column of this lane's group in the [N][ngroups] arrays
            auto gs_load = [&](int i, float& s_, float& z_) {
                const int64_t r = (int64_t)min(wrow0 + i * TILE_ROWS + g + 8 * (t & 1), p.N - 1) * p.ngroups + gs_off;
                s_ = __ldg(p.scales + r); z_ = __ldg(p.zps + r);
            };
            if constexpr (GS) {
                gs_off = (P * 256 + 128 * (t >> 1)) / p.kgroup;
                gs_load(0, csc, czp);
            }

            // ---- this pair of every tile
#pragma unroll 1
            for (int i = 0; i < S; ++i) {
                if (pass == 0) mbar_wait(sbase + OFF_BARS + 8u * (uint32_t)(grp * S + i), 0u);
                const uint32_t pb = sbase + p.tile_off + (uint32_t)(i * p.tile_bytes + P * PAIR_BYTES);
                float v[4];
                float va[4];                                  // GS: the first half's values
                float nsc = 0.0f, nzp = 0.0f;                 // GS: this lane's scale / zero point of the next tile (in flight during the MMAs)
                if constexpr (GS) {
                    if (i + 1 < S) gs_load(i + 1, nsc, nzp);
                }
                if constexpr (I3) {
                    int ac[3][4];                             // one chain per digit
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int r = 0; r < 4; ++r) ac[d][r] = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (GS && c == 2) {                   // the first 128 columns are done: keep their values, start over
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                va[r] = digits_to_float(ac[0][r], ac[1][r], ac[2][r]) * ((r & 1) ? d1 : d0);
                                ac[0][r] = 0; ac[1][r] = 0; ac[2][r] = 0;
                            }
                        }
                        uint32_t a[4];
                        ldsm_x4(a, pb + offc[c]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t w0 = a[2 * h], w1 = a[2 * h + 1];
                            const uint32_t e0 = w0 & 0x0f0f0f0fu, e1 = w1 & 0x0f0f0f0fu, o0 = (w0 >> 4) & 0x0f0f0f0fu, o1 = (w1 >> 4) & 0x0f0f0f0fu;
                            imma_uu(ac[0], e0, e1, o0, o1, bi[2 * c + h][0][0], bi[2 * c + h][0][1]);
                            imma_uu(ac[1], e0, e1, o0, o1, bi[2 * c + h][1][0], bi[2 * c + h][1][1]);
                            imma(ac[2], e0, e1, o0, o1, bi[2 * c + h][2][0], bi[2 * c + h][2][1]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)               // digits -> value
                        v[r] = digits_to_float(ac[0][r], ac[1][r], ac[2][r]) * ((r & 1) ? d1 : d0);
                } else {
                float acc[2][NT][4];                          // two chains (alpha, beta) per n-tile
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[j][nt][r] = 0.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (GS && c == 2) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            float a = acc[0][0][r] + acc[1][0][r];
                            if constexpr (NT == 2) a += acc[0][1][r] + acc[1][1][r];
                            va[r] = (a * 16777216.0f) * ((r & 1) ? d1 : d0);
#pragma unroll
                            for (int j = 0; j < 2; ++j)
#pragma unroll
                                for (int nt = 0; nt < NT; ++nt) acc[j][nt][r] = 0.0f;
                        }
                    }
                    uint32_t a[4];
                    if (!HM_ABL(8)) ldsm_x4(a, pb + offc[c]); else { a[0] = c; a[1] = lane; a[2] = i; a[3] = 7; }                // a0 / a1: rows g / g + 8, bytes 32 c + 4 t ..; a2 / a3: + 16 bytes
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w0 = a[2 * h], w1 = a[2 * h + 1];
                        const uint32_t v0 = w0 >> 8, v1 = w1 >> 8;
#pragma unroll
                        if (!HM_ABL(2)) for (int nt = 0; nt < NT; ++nt) {
                            mma_m16n8k16_f16(acc[0][nt], w0 & M0, w1 & M0, w0 & M1, w1 & M1, bf[2 * c + h][0][nt][0], bf[2 * c + h][0][nt][1]);
                            mma_m16n8k16_f16(acc[1][nt], v0 & M0, v1 & M0, v0 & M1, v1 & M1, bf[2 * c + h][1][nt][0], bf[2 * c + h][1][nt][1]);
                        }
                        else { acc[0][0][0] += __uint_as_float((w0 & M0) ^ (w1 & M1) ^ (v0 & M0) ^ (v1 & M1) ^ (w0 & M1) ^ (w1 & M0) ^ (v0 & M1) ^ (v1 & M0)); }
                    }
                }
                // partial tile -> this warp's own slot of the tile, descaled: q was 2^-24 too small
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float a = acc[0][0][r] + acc[1][0][r];
                    if constexpr (NT == 2) a += acc[0][1][r] + acc[1][1][r];
                    v[r] = (a * 16777216.0f) * ((r & 1) ? d1 : d0);
                }
                }
                if constexpr (GS) {                           // v = second half so far: scale both halves, subtract the zero-point terms
                    float gsc[2][2], gzp[2][2];               // [half][row g / g + 8] from the quad's lanes 2 half + row
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            gsc[hh][rr] = __shfl_sync(0xffffffffu, csc, (lane & ~3) | (2 * hh + rr));
                            gzp[hh][rr] = __shfl_sync(0xffffffffu, czp, (lane & ~3) | (2 * hh + rr));
                        }
                    csc = nsc; czp = nzp;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        v[r] = gsc[0][r >> 1] * fmaf(-gzp[0][r >> 1], sxg[0][r & 1], va[r]) + gsc[1][r >> 1] * fmaf(-gzp[1][r >> 1], sxg[1][r & 1], v[r]);
                }
                // slot layout [row][token]: lane (g, t) owns tokens 2t, 2t + 1 of rows g and g + 8 -- two conflict-free 8-byte stores
                float2* slot = reinterpret_cast<float2*>(smem + OFF_SLOTS + (i * NW + warp) * SLOT_BYTES);
                float2 lo2 = make_float2(v[0], v[1]), hi2 = make_float2(v[2], v[3]);
                if (q > 0) {
                    const float2 o0 = slot[g * 4 + t], o1 = slot[(g + 8) * 4 + t];
                    lo2.x += o0.x; lo2.y += o0.y; hi2.x += o1.x; hi2.y += o1.y;
                }
                slot[g * 4 + t] = lo2;
                slot[(g + 8) * 4 + t] = hi2;
                if (first && i < 5) HM_STAMP(5 + i);
            }
        }
        if (pass == 0) HM_STAMP(10);
        if (t == 0) s_sx[warp * MB + g] = sxacc;
        if (pass == 0 && tid < 128) { s_par[tid] = sc; s_par[128 + tid] = zp; s_par[256 + tid] = bias; }
        __syncthreads();
        if (pass == 0) HM_STAMP(11);

        // ---- fold the 16 warps and write y.  Thread = (tile, row, token quad tq, half of the warps): 8 x (16-byte load
        // of 4 tokens of its row + 16-byte load of their sum_k x), the two halves meet in one shuffle step -- a fixed
        // order: deterministic.  Lanes 0..7 of a quarter warp read 128 consecutive bytes: no bank conflicts.
        // (the whole fold is 80 + 80 vector loads per SM; one 4-byte load per (output, warp) was 1 us of LSU time)
        {
            const unsigned int flagged = *s_flag;
            const int tq = tid & 1, half = (tid >> 3) & 1, tile = tid >> 6;
            const int row = ((tid >> 1) & 3) + 4 * ((tid >> 4) & 3);
            const int rowpos = tile * TILE_ROWS + row;
            if (tile < S) {                                   // uniform per warp
                const float4* src = reinterpret_cast<const float4*>(smem + OFF_SLOTS + ((tile * NW + half * 8) * SLOT_BYTES)) + row * 2 + tq;
                const float4* ssx = reinterpret_cast<const float4*>(s_sx + half * 8 * MB) + tq;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < 8; ++w) {                 // (idle warps hold zeros)
                    const float4 av = src[w * (SLOT_BYTES / 16)], sv = ssx[w * (MB / 4)];
                    a.x += av.x; a.y += av.y; a.z += av.z; a.w += av.w;
                    s.x += sv.x; s.y += sv.y; s.z += sv.z; s.w += sv.w;
                }
                a.x += __shfl_xor_sync(0xffffffffu, a.x, 8); a.y += __shfl_xor_sync(0xffffffffu, a.y, 8);
                a.z += __shfl_xor_sync(0xffffffffu, a.z, 8); a.w += __shfl_xor_sync(0xffffffffu, a.w, 8);
                s.x += __shfl_xor_sync(0xffffffffu, s.x, 8); s.y += __shfl_xor_sync(0xffffffffu, s.y, 8);
                s.z += __shfl_xor_sync(0xffffffffu, s.z, 8); s.w += __shfl_xor_sync(0xffffffffu, s.w, 8);
                if (pass == 0) HM_STAMP(13);
                const bool mine = rowpos < nrows;
                const float rsc = s_par[rowpos], rzp = s_par[128 + rowpos], rbias = s_par[256 + rowpos];
                float vv[4] = {rsc * fmaf(-rzp, s.x, a.x), rsc * fmaf(-rzp, s.y, a.y), rsc * fmaf(-rzp, s.z, a.z), rsc * fmaf(-rzp, s.w, a.w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int m = m0 + tq * 4 + j;
                    if (p.gated) {
                        // fused gate + up pair: the even row (gate) fetches its neighbour's value (up, lane ^ 2) and writes silu(gate) * up
                        const float uu = __shfl_xor_sync(0xffffffffu, vv[j], 2);
                        if (mine && half == 0 && !(row & 1) && m < Mrows && !((flagged >> m) & 1u))
                            store_out(p.y, p.y_dtype, (int64_t)(xrow0 + m) * (p.N >> 1) + ((r0 + rowpos) >> 1), vv[j] / (1.0f + __expf(-vv[j])) * uu);
                    } else if (mine && half == 0 && m < Mrows && !((flagged >> m) & 1u)) {
                        store_out(p.y, p.y_dtype, (int64_t)(xrow0 + m) * p.N + r0 + rowpos, vv[j] + rbias);
                    }
                }
            }
        }
        if (pass < 1) HM_STAMP(12 + pass);
    }

    HM_STAMP(14);
    // ---- tokens with NaN / Inf: the reference's arithmetic (dequantise, then fp32 multiply-add), so that non-finite
    // values propagate as in F.linear; one warp per output, weights re-read from global memory
    if (const unsigned int flagged = *s_flag) {
        const int64_t row_bytes = p.K >> 1;
        for (int m = 0; m < Mrows; ++m) {
            if (!((flagged >> m) & 1u)) continue;
            auto ref_row = [&](int row) {
                float rs = 0.0f, rz = 0.0f;
                if constexpr (!GS) { rs = __ldg(p.scales + row); rz = __ldg(p.zps + row); }
                const uint8_t* wr = p.packed + (int64_t)row * row_bytes;
                float acc = 0.0f;
                for (int kb = lane; kb < row_bytes; kb += 32) {
                    if constexpr (GS) {
                        const int64_t gi = (int64_t)row * p.ngroups + (2 * kb) / p.kgroup;
                        rs = __ldg(p.scales + gi); rz = __ldg(p.zps + gi);
                    }
                    const unsigned int byte = wr[kb];
                    const float w0 = ((float)(byte & 15u) - rz) * rs, w1 = ((float)(byte >> 4) - rz) * rs;
                    acc = fmaf(w0, load1f(p.x, p.x_dtype, x_row(m) * p.K + 2 * kb), acc);
                    acc = fmaf(w1, load1f(p.x, p.x_dtype, x_row(m) * p.K + 2 * kb + 1), acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                return acc;
            };
            for (int rc = unit * warp; rc < nrows; rc += unit * NW) {
                const int row = r0 + rc, wrow = wrow0 + rc;   // output column / row of the stacked weight tensor
                float acc = ref_row(wrow);
                if (p.gated) {
                    const float upv = ref_row(wrow + 1);
                    if (lane == 0) store_out(p.y, p.y_dtype, (int64_t)(xrow0 + m) * (p.N >> 1) + (row >> 1), acc / (1.0f + __expf(-acc)) * upv);
                } else if (lane == 0) {
                    if (p.bias) acc += __ldg(p.bias + wrow);
                    store_out(p.y, p.y_dtype, (int64_t)(xrow0 + m) * p.N + row, acc);
                }
            }
        }
    }
}

struct HmPlan {
    int grid, rows_q, rows_rem, s_max, npairs, gpw, nbars, chunk, tile_bytes, tile_off, npasses;
    size_t smem;
};

bool plan_hm(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, int gated, HmPlan* c) {
    if (M < 1 || M > HM_MAX_M || K < 256 || K % 256 != 0 || K > 16384 || N < 1 || N > 0x3fffffff) return false;
    if (gated && (N & 1)) return false;
    const int unit = gated ? 2 : 1;
    c->npairs = (int)(K / 256);
    c->gpw = (c->npairs + NW - 1) / NW;                          // <= 4
    c->nbars = c->gpw;
    c->chunk = (c->npairs + c->nbars - 1) / c->nbars;
    c->tile_bytes = c->nbars * c->chunk * PAIR_BYTES;            // >= npairs * 2 KB: a 3-D box always has room
    c->npasses = (int)((M + MB - 1) / MB);
    int cap = tuning().gemv_ctas > 0 ? tuning().gemv_ctas : sm_count;
    if (cap > sm_count) cap = sm_count;
    const int64_t units = N / unit;
    // One CTA per SM when its rows fit in shared memory; else 2, 3, ... waves of smaller CTAs (Llama's fused gate + up
    // pair, Mixtral's 14336-wide projections): each wave pays the prologue again, the weights are still read once.
    const int max_waves = tuning().hm_waves > 0 ? tuning().hm_waves : 4;
    for (int waves = 1; waves <= max_waves; ++waves) {
        int64_t g = (N + TILE_ROWS - 1) / TILE_ROWS;             // few rows: one tile per CTA
        if (g > (int64_t)cap * waves) g = (int64_t)cap * waves;
        if (g > units) g = units;
        if (g < 1) g = 1;
        c->grid = (int)g;
        c->rows_q = (int)(units / g); c->rows_rem = (int)(units % g);
        const int64_t br = unit * ((units + g - 1) / g);         // most rows of a CTA
        const int64_t S = (br + TILE_ROWS - 1) / TILE_ROWS;
        if (S > MAX_TILES || S * c->nbars > MAX_BARS) continue;
        c->s_max = (int)S;
        c->tile_off = (OFF_SLOTS + (int)S * NW * SLOT_BYTES + 1023) / 1024 * 1024;
        c->smem = (size_t)c->tile_off + (size_t)S * c->tile_bytes;
        if (c->smem <= (size_t)max_smem) return true;
        if (g < (int64_t)cap * waves) break;                     // more waves would not make the CTAs smaller
    }
    return false;
}

template <int XT, int FORM, bool GRP, bool GS>
int launch_hm_inst(const HmPlan& c, int n_experts, const CUtensorMap& map, const HmParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_hm_kernel<XT, FORM, GRP, GS>;
    static thread_local int attr_dev_smem[64] = {0};
    int dev = 0;
    B200Q_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && attr_dev_smem[dev] < (int)c.smem) {
        B200Q_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        attr_dev_smem[dev] = (int)c.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)c.grid, (unsigned)n_experts);
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
    return check_cuda(cudaLaunchKernelEx(&cfg, kfn, map, p), "gemv_hm launch");
}

}  // namespace

bool gemv_hm_supported(const DeviceInfo& dev, int64_t M, int64_t N, int64_t K, int gated) {
    HmPlan c;
    if (!plan_hm(dev.sm_count, dev.max_smem_optin, M, N, K, gated, &c)) return false;
    // measured crossover with the tcgen05 GEMM (tools/dec_tune.py): K = 14336 needs several waves of one-tile CTAs with four
    // pairs per warp; two passes over them (M > 8) are no faster than the GEMM (14336 -> 4096 M = 16: 39.2 vs 37.7 us)
    if (tuning().force_path != 7 && M > MB && c.gpw >= 4 && c.grid > dev.sm_count) return false;
    // 25 .. 32 tokens (four passes): ahead of the 32-token tiles of the GEMM only for one wave of CTAs and K <= 8192
    // (4096 -> 11008 M = 32: 24.1 vs 30.8 us; 11008 -> 4096: 36.0 vs 33.7; 4096 -> 14336: 37.2 vs 35.9; up to M = 24 always:
    // 18.5 vs 30.7, 27.3 vs 33.4, 28.6 vs 35.7 us)
    if (tuning().force_path != 7 && M > 3 * MB && (c.gpw > 2 || c.grid > dev.sm_count)) return false;
    return true;
}

int launch_gemv_hm(const DeviceInfo& dev, const void* x, int x_dtype, const uint8_t* packed, const float* scales,
                   const float* zps, const float* bias, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                   unsigned flags, cudaStream_t st, const uint8_t* next_packed, size_t next_bytes, int gated,
                   const int32_t* offsets, int n_experts, const int32_t* row_map, int kgroup) {
    HmPlan c;
    if (n_experts < 1) n_experts = 1;
    const bool grp = offsets != nullptr;
    if (!plan_hm(dev.sm_count, dev.max_smem_optin, M, N, K, gated, &c))
        return set_error(B200Q_EINVAL, "gemv_hm: unsupported shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if ((reinterpret_cast<uintptr_t>(x) & (x_dtype == B200Q_F32 ? 31 : 15)) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return set_error(B200Q_EALIGN, "gemv_hm: packed must be 16-byte aligned, x 32-byte (fp32) / 16-byte aligned");
    HmParams p{};
    p.x = x; p.packed = packed; p.scales = scales; p.zps = zps; p.bias = gated ? nullptr : bias; p.y = y;
    p.gated = gated;
    p.x_dtype = x_dtype; p.y_dtype = y_dtype;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.rows_q = c.rows_q; p.rows_rem = c.rows_rem;
    p.npairs = c.npairs; p.gpw = c.gpw; p.nbars = c.nbars; p.chunk = c.chunk;
    p.tile_bytes = c.tile_bytes; p.tile_off = c.tile_off; p.npasses = c.npasses;
    p.wait_weights = (flags & B200Q_FLAG_STATIC_WEIGHTS) ? 0 : 1;
    p.early_tiles = tuning().gemv_early >= 0 ? tuning().gemv_early : 92;
    p.next_packed = next_packed;
    p.pf_mode = tuning().gemv_pf;
    p.next_bytes = (tuning().gemv_pf != 0 && next_packed && (reinterpret_cast<uintptr_t>(next_packed) & 15) == 0) ? next_bytes : 0;
    p.next_chunk = (unsigned int)(p.next_bytes / (unsigned long long)c.grid);
    p.debug = tuning().gemv_debug > 0 ? tuning().gemv_debug : 0;
    p.offsets = offsets; p.row_map = grp ? row_map : nullptr;
    if (grp) p.next_bytes = 0;
    p.kgroup = kgroup; p.ngroups = kgroup > 0 ? (int)(K / kgroup) : 0;
    CUtensorMap map;
    if (int rc = dec_weight_map(packed, N * n_experts, K, c.chunk, &map)) return rc;
    const bool pdl = tuning().gemv_pdl != 0;
    if (kgroup > 0) {                                       // group-wise scales (plain linear only)
        if (grp || gated || kgroup % 128 != 0 || K % kgroup != 0) return set_error(B200Q_EINVAL, "gemv_hm: group-wise scales need kgroup %% 128 == 0, K %% kgroup == 0, no grouping / gate");
        if (x_dtype == B200Q_F32) return launch_hm_inst<B200Q_F32, 2, false, true>(c, 1, map, p, pdl, st);
        if (x_dtype == B200Q_F16) return launch_hm_inst<B200Q_F16, 0, false, true>(c, 1, map, p, pdl, st);
        return launch_hm_inst<B200Q_BF16, 0, false, true>(c, 1, map, p, pdl, st);
    }
    if (grp) {
        if (x_dtype == B200Q_F32) return launch_hm_inst<B200Q_F32, 2, true, false>(c, n_experts, map, p, pdl, st);
        if (x_dtype == B200Q_F16) return launch_hm_inst<B200Q_F16, 0, true, false>(c, n_experts, map, p, pdl, st);
        return launch_hm_inst<B200Q_BF16, 0, true, false>(c, n_experts, map, p, pdl, st);
    }
    if (x_dtype == B200Q_F32)
        return tuning().hm_i3 != 0 ? launch_hm_inst<B200Q_F32, 2, false, false>(c, 1, map, p, pdl, st) : launch_hm_inst<B200Q_F32, 0, false, false>(c, 1, map, p, pdl, st);
    if (x_dtype == B200Q_F16) return launch_hm_inst<B200Q_F16, 0, false, false>(c, 1, map, p, pdl, st);
    return launch_hm_inst<B200Q_BF16, 0, false, false>(c, 1, map, p, pdl, st);
}

}  // namespace b200q

#ifdef B200Q_PROF
extern "C" int b200q_debug_read_prof_hm(long long* h_out) {
    return b200q::check_cuda(cudaMemcpyFromSymbol(h_out, b200q::g_hm_prof, sizeof(long long) * 256 * 16), "read prof");
}
#endif
