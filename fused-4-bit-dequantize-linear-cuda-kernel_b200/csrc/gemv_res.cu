// Decode path, "resident slab" kernel (the default for M <= 8): y[M,N] = x[M,K] @ dequant(W)^T when a CTA's whole
// share of W (its rows x K/2 packed bytes, or half of that with K split over a 2-CTA cluster) fits in shared memory
// next to the x operand.  HBM-bound by design: every packed byte is read once, by the TMA engine.
//
// Arithmetic: exact integers.  x is a per-row fixed-point number (even columns Xe = round(x 2^e), odd columns
// Xo = round(x 2^(e-4))) cut into four signed base-256 limbs = four columns of IMMA m16n8k32 (u8 x s8 -> s32) per batch
// row.  The nibbles are never widened: the raw packed byte q_lo + 16 q_hi meets the limbs of Xe, the masked byte
// 16 q_hi meets the limbs of Z = Xo - Xe.  s32 partial sums, s64 when the limbs are combined,
// y = s * 2^-e * (sum q X - zp * sum X) with one fp32 rounding: results do not depend on the summation order.
//
// Structure (profiles/r01_gemv_notes.md has the measurements behind each choice):
//   * no ring: tile i (16 rows) lands in its own slot with its own single-use mbarrier.  Two bulk copies per
//     tile, the second half skewed by 64 bytes, make the 8-row LDS.128 of a warp conflict-free; with K split
//     over a cluster: one copy per row, rows 64 bytes off a 128-byte multiple apart;
//   * requests are staged (two tiles before griddepcontrol.wait, the rest behind the x loads) so that the
//     16 KB of x do not queue behind 22 MB of weight requests; optionally the NEXT layer's weights are
//     prefetched into L2 (cp.async.bulk.prefetch.L2) behind the own ones;
//   * the x operand is built inside the CTA: one coalesced 32-byte load per thread (kept in registers for
//     M <= 2), block amax with REDUX + one barrier, 8 F2I per thread, byte transposes (PRMT) into a shared-memory
//     image in mma B-fragment order; a lane then fetches the operand of a granule with two 16-byte loads;
//   * cross-warp reduction: per-warp partial slots (plain stores) and a 4-threads-per-output epilogue for
//     M <= 2; replicated shared-memory integer atomics for M = 3..8; with K split, the exact s64 partials of
//     the second CTA reach the leader through distributed shared memory;
//   * no workspace and no second kernel: consecutive decode layers are main -> main, which programmatic
//     dependent launch overlaps CTA by CTA.
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
constexpr int ACC_CS = 20;           // ints between the columns of an acc plane: 16 rows + 4 (RED.ADD of a warp hits 32 banks)
constexpr int ACC_PLANE = 8 * ACC_CS;
// acc is replicated R = 8 / NT times (warp w adds into replica w % R): shared-memory atomics to ONE address from
// 16 warps serialise at ~60 clk each, which used to show up as 0.5 us between the last tile and the epilogue
constexpr int ACC_REPS_X_NT = 8;

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
    int nslab;                       // CTAs of one cluster that split K (1: no cluster)
    int gran_q, gran_rem;            // K / 128 granules dealt out over the slabs: slab s has gran_q (+1 if s < gran_rem)
    int pitch;                       // nslab > 1: bytes between the rows of a tile in shared memory (= 64 mod 128)
    int xbuf_off, outs_cap;          // nslab > 1: exchange buffer [nslab][outs_cap] s64 + [nslab][8][2] s32 in the leader
    int acc_off, img_off, tile_off;  // byte offsets in dynamic shared memory
    int wait_weights;                // 1: weights may be written by the preceding kernel
    int early_tiles;                 // tiles requested before griddepcontrol.wait
    int pf_mode;
    int slots;                       // 1 (M <= 2): per-warp partial slots (plain stores) instead of shared-memory atomics
    int slots_alias;                 // the slots reach into the x image: barrier before the first store
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

__device__ __forceinline__ void st_cluster_b64(uint32_t cluster_addr, long long v) {
    asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(cluster_addr), "l"(v) : "memory");
}

__device__ __forceinline__ void red_add_s32(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Shared memory (dynamic):
//   [0, 256)        mbarriers, one per tile
//   [256, 1280)     s_amax[8][16] f32   (then s_up[8] f32 at +512, s_ex[8] s32 at +544, s_txs[8][2] s32 at +576)
//   [1280, 2304)    s_tx[16 warps][8 rows][2] s32
//   [acc_off, ..)   acc[8/NT replicas][tile][NT*8 columns][ACC_CS] s32   (zeroed, RED.ADD target; 16 rows + 4 pad per column)
//   [img_off, ..)   x image of ONE n-tile (two batch rows): [h][granule][limb][t][word] x {b0, b1}
//   [tile_off, ..)  tiles: ntiles x 16 rows x K/2 bytes
constexpr int OFF_AMAX = 256, OFF_EX = 800, OFF_TXS = 832, OFF_TX = 1280, OFF_END = 2304;

template <int GPW, int NT, bool SPLIT>   // SPLIT: K cut over the CTAs of a cluster (compile-time, so that the common case pays nothing)
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

    // K split: the nslab CTAs of a cluster own consecutive granule ranges of the same rows; their exact integer
    // partials meet in the leader's shared memory (DSMEM), so the result is still rounded once
    const int nslab = SPLIT ? p.nslab : 1;
    const int slab = SPLIT ? (int)(blockIdx.x % nslab) : 0;
    const int b = SPLIT ? (int)(blockIdx.x / nslab) : (int)blockIdx.x;
    const int g0 = slab * p.gran_q + min(slab, p.gran_rem);
    const int G = p.gran_q + (slab < p.gran_rem ? 1 : 0);
    const int r0 = b * p.rows_q + min(b, p.rows_rem);
    const int nrows = p.rows_q + (b < p.rows_rem ? 1 : 0);
    const int ntiles = (nrows + TILE_ROWS - 1) / TILE_ROWS;
    const int row_bytes = p.K >> 1;
    // A tile is TWO bulk copies: rows 0..7 at offset 0 and rows 8..15 at offset 8 * row_bytes + 64.  An LDS.128 of
    // a warp fetches four rows of each half (mma row g <-> tile row perm(g), below): with row_bytes a multiple of
    // 128 the two halves then sit in complementary 64-byte windows and the 8 rows x 64 B touch every bank exactly
    // four times (no conflicts) -- without one copy per padded row.
    // With a K split a row's slab is not contiguous with the next row's: one copy per row, rows p.pitch apart.
    const uint32_t tile_bytes = SPLIT ? (uint32_t)(TILE_ROWS * p.pitch) : (uint32_t)(TILE_ROWS * row_bytes + 128);
    const uint32_t half_off = (uint32_t)(8 * row_bytes + 64);
    float* s_amax = reinterpret_cast<float*>(smem + OFF_AMAX);
    int* s_ex = reinterpret_cast<int*>(smem + OFF_EX);
    int* s_txs = reinterpret_cast<int*>(smem + OFF_TXS);
    int* s_tx = reinterpret_cast<int*>(smem + OFF_TX);
    int* acc = reinterpret_cast<int*>(smem + p.acc_off);
    uint2* img = reinterpret_cast<uint2*>(smem + p.img_off);

    if (tid == 0) {
        for (int i = 0; i < ntiles; ++i) mbar_init(sbase + 8u * i, 1);
        fence_mbar_init();
    }
    constexpr int R = ACC_REPS_X_NT / NT;
    const int rep_ints = ntiles * NT * ACC_PLANE;
    if (!p.slots)
        for (int i = tid; i < R * rep_ints; i += NTHR) acc[i] = 0;
    __syncthreads();
    pdl_launch_dependents();
    stamp(1);

    // mma row (0..15) -> row of the tile, see the tile layout above
    auto tile_row = [&](int er) { return SPLIT ? er : ((er & 1) << 3) + ((er & 8) >> 1) + ((er & 7) >> 1); };
    const bool skip_loads = (p.debug & 2) != 0;
    // the two copies of tile i are issued by lane 0 of warps 2i and 2i + 1 (mod 16); the first also arms the tile's
    // barrier with the byte count of both (complete_tx may run ahead of expect_tx: the phase cannot complete before
    // that arrival)
    const uint64_t pol = policy_evict_first();
    auto issue_tile = [&](int i) {                                // lane 0 of every warp calls this
        const int rows = min(TILE_ROWS, nrows - i * TILE_ROWS);
        const uint32_t bar = sbase + 8u * i, dst = sbase + p.tile_off + i * tile_bytes;
        const uint8_t* src = p.packed + (int64_t)(r0 + i * TILE_ROWS) * row_bytes;
        if (SPLIT) {                                          // row `warp` of the tile: this slab's bytes of it
            if (warp == 0) mbar_arrive_expect_tx(bar, (uint32_t)(rows * G * GRAN_B));
            if (warp < rows)
                bulk_g2s_hint(dst + warp * p.pitch, src + (int64_t)warp * row_bytes + g0 * GRAN_B, (uint32_t)(G * GRAN_B), bar, pol);
        } else if (warp == ((2 * i) & (NW - 1))) {
            mbar_arrive_expect_tx(bar, (uint32_t)(rows * row_bytes));
            bulk_g2s_hint(dst, src, (uint32_t)(min(rows, 8) * row_bytes), bar, pol);
        } else if (warp == ((2 * i + 1) & (NW - 1)) && rows > 8) {
            bulk_g2s_hint(dst + half_off, src + 8 * row_bytes, (uint32_t)((rows - 8) * row_bytes), bar, pol);
        }
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
    // staged requests (tuning key gemv_early = a + 10 * b): a tiles before griddepcontrol.wait, b more once the
    // x loads are in flight, the rest when the x operand is built -- bounds what queues ahead of the x loads
    const int early = min(p.early_tiles % 10, ntiles);
    const int mid = min(early + p.early_tiles / 10, ntiles);
    if (lane == 0 && !skip_loads) {
        for (int i = 0; i < early; ++i) issue_tile(i);
        if (early == ntiles && warp == NW - 1) prefetch_next();
    }
    stamp(2);
    pdl_wait();          // x (and y) belong to the stream-ordered predecessor
    stamp(3);
    // tiles held back (tuning key gemv_early) are requested only after this warp's x loads are in flight, so
    // that those do not queue behind the weight stream
    auto issue_range = [&](int from, int to) {
        if (lane == 0 && !skip_loads && from < to) {
            for (int i = from; i < to; ++i) issue_tile(i);
            if (to == ntiles && warp == NW - 1) prefetch_next();
        }
    };
    auto issue_late = [&]() { issue_range(early, mid); };

    // ---- pass 1: amax of every batch row (coalesced 32-byte loads, all threads).  With one n-tile and at most
    // two items (8 activations each) per thread, the values stay in registers for pass 2.
    const int K8 = p.K >> 3;
    const bool keep = NT == 1 && !SPLIT && p.M * K8 <= 2 * NTHR;
    float xv[2][8];
    if (keep) {
        float am0 = 0.0f, am1 = 0.0f;
        const bool two = p.M * K8 > NTHR;                      // uniform: a second item per thread exists
#pragma unroll
        for (int e = 0; e < 8; ++e) { xv[0][e] = 0.0f; xv[1][e] = 0.0f; }
        if (tid < p.M * K8) load8f(p.x, p.x_dtype, (int64_t)tid * 8, xv[0]);
        if (two && tid + NTHR < p.M * K8) load8f(p.x, p.x_dtype, (int64_t)(tid + NTHR) * 8, xv[1]);
        issue_late();
        {
            float a = 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e) a = fmaxf(a, fabsf(xv[0][e]));
            if (tid >= K8) am1 = a; else am0 = a;
        }
        if (two) {
            float a = 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e) a = fmaxf(a, fabsf(xv[1][e]));
            if (tid + NTHR >= K8) am1 = fmaxf(am1, a); else am0 = fmaxf(am0, a);
        }
        // non-negative floats order like their bit patterns: one REDUX per row instead of five shuffles
        const unsigned int u0 = __reduce_max_sync(0xffffffffu, __float_as_uint(am0));
        if (lane == 0) s_amax[warp] = __uint_as_float(u0);
        if (p.M > 1) {
            const unsigned int u1 = __reduce_max_sync(0xffffffffu, __float_as_uint(am1));
            if (lane == 0) s_amax[NW + warp] = __uint_as_float(u1);
        }
    } else {
        issue_late();
        if (K8 <= NTHR) {
            // a row is at most one item per thread: four ROWS in flight per thread (row after row cost one trip to L2
            // each: 4.5 us for 8 rows while the weight requests are queued in front)
            for (int m0 = 0; m0 < p.M; m0 += 4) {
                float v[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[u][e] = 0.0f;
                    if (m0 + u < p.M && tid < K8) load8f(p.x, p.x_dtype, (int64_t)(m0 + u) * p.K + (int64_t)tid * 8, v[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float am = 0.0f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) am = fmaxf(am, fabsf(v[u][e]));
                    const unsigned int ua = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
                    if (lane == 0 && m0 + u < p.M) s_amax[(m0 + u) * NW + warp] = __uint_as_float(ua);
                }
            }
        } else {
            for (int m = 0; m < p.M; ++m) {
                float am = 0.0f;
                // four loads in flight per thread (a plain strided loop made one trip to L2 per iteration: 2.7 us for
                // the 43 KB of a K = 11008 row)
                for (int i0 = tid; i0 < K8; i0 += 4 * NTHR) {
                    float v[4][8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[u][e] = 0.0f;
                        if (i0 + u * NTHR < K8) load8f(p.x, p.x_dtype, (int64_t)m * p.K + (int64_t)(i0 + u * NTHR) * 8, v[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int e = 0; e < 8; ++e) am = fmaxf(am, fabsf(v[u][e]));
                }
                const unsigned int ua = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
                if (lane == 0) s_amax[m * NW + warp] = __uint_as_float(ua);
            }
        }
    }
    __syncthreads();
    // every warp derives the rows' exponents itself (one LDS + one REDUX per row): no second barrier
    auto row_exp = [&](int m) {
        const unsigned int u = __reduce_max_sync(0xffffffffu, lane < NW ? __float_as_uint(s_amax[m * NW + lane]) : 0u);
        const float am = __uint_as_float(u);
        int ex = 0;
        if (am > 0.0f && am < INFINITY) ex = max(-96, min(126, 155 - (int)(u >> 23)));
        return ex;
    };
    stamp(4);

    // ---- pass 2: x image of one n-tile at a time, then every lane pulls its B fragments
    // lane (g, t) holds mma column g = 4h + l: limb l of batch row 2*nt + h
    const int l = g & 3, h = g >> 2;
    uint32_t bf[GPW][NT][4][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int live_rows = min(2, p.M - 2 * nt);          // batch rows of this n-tile that exist (may be <= 0)
        if (nt > 0) __syncthreads();                          // the image is reused
        int slo0 = 0, shi0 = 0, slo1 = 0, shi1 = 0;           // sum of X as (X & 0xffff), (X >> 16): exact in s32
        const int ex0 = live_rows > 0 ? row_exp(2 * nt) : 0, ex1 = live_rows > 1 ? row_exp(2 * nt + 1) : 0;
        const float up0 = __uint_as_float((uint32_t)(127 + ex0) << 23), up1 = __uint_as_float((uint32_t)(127 + ex1) << 23);
        if (tid == 0) { s_ex[2 * nt] = ex0; s_ex[2 * nt + 1] = ex1; }
        const int items = live_rows * G * 16;                 // 8 values each
        auto convert = [&](int it, const float (&v)[8]) {
            const int hh = it >= G * 16 ? 1 : 0;
            const int r = it - hh * G * 16;                   // = k / 8: (granule, t, word)
            const float up = hh ? up1 : up0, up16 = up * 0.0625f;
            uint32_t D[8];
            int a_lo = 0, a_hi = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // even column: Xe = round(x 2^e); odd column: Xo = round(x 2^(e-4)), carried as Z = Xo - Xe (see
                // the main loop: the raw packed byte q_lo + 16 q_hi meets Xe, the masked byte 16 q_hi meets Z)
                const int Xe = __float2int_rn(v[2 * i] * up);
                const int Xo = __float2int_rn(v[2 * i + 1] * up16);
                const int Z = Xo - Xe, Xo16 = Xo << 4;
                a_lo += (Xe & 0xffff) + (Xo16 & 0xffff);
                a_hi += (Xe >> 16) + (Xo16 >> 16);
                D[2 * i] = (uint32_t)(Xe + 0x00808080) ^ 0x00808080u;      // byte l = signed base-256 digit l
                D[2 * i + 1] = (uint32_t)(Z + 0x00808080) ^ 0x00808080u;
            }
            if (hh) { slo1 += a_lo; shi1 += a_hi; } else { slo0 += a_lo; shi0 += a_hi; }
            // 4x4 byte transposes: digit l of the four Xe -> b0 (meets the raw bytes), of the four Z -> b1 (high nibbles x 16)
            const uint32_t e0 = __byte_perm(D[0], D[2], 0x5140), e1 = __byte_perm(D[4], D[6], 0x5140);
            const uint32_t e2 = __byte_perm(D[0], D[2], 0x7362), e3 = __byte_perm(D[4], D[6], 0x7362);
            const uint32_t o0 = __byte_perm(D[1], D[3], 0x5140), o1 = __byte_perm(D[5], D[7], 0x5140);
            const uint32_t o2 = __byte_perm(D[1], D[3], 0x7362), o3 = __byte_perm(D[5], D[7], 0x7362);
            uint2* dst = img + ((size_t)(hh * G + (r >> 4)) * 4) * 16 + (r & 15);
            dst[0 * 16] = make_uint2(__byte_perm(e0, e1, 0x5410), __byte_perm(o0, o1, 0x5410));
            dst[1 * 16] = make_uint2(__byte_perm(e0, e1, 0x7632), __byte_perm(o0, o1, 0x7632));
            dst[2 * 16] = make_uint2(__byte_perm(e2, e3, 0x5410), __byte_perm(o2, o3, 0x5410));
            dst[3 * 16] = make_uint2(__byte_perm(e2, e3, 0x7632), __byte_perm(o2, o3, 0x7632));
        };
        if (keep) {
            if (tid < items) convert(tid, xv[0]);
            if (tid + NTHR < items) convert(tid + NTHR, xv[1]);
        } else {
            for (int it0 = tid; it0 < items; it0 += 2 * NTHR) {       // two loads in flight per thread
                float v[2][8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int it = it0 + u * NTHR;
                    if (it < items) {
                        const int hh = it >= G * 16 ? 1 : 0;
                        load8f(p.x, p.x_dtype, (int64_t)(2 * nt + hh) * p.K + (int64_t)(g0 * 16 + it - hh * G * 16) * 8, v[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (it0 + u * NTHR < items) convert(it0 + u * NTHR, v[u]);
            }
        }
        if (live_rows > 0) {
            slo0 = __reduce_add_sync(0xffffffffu, slo0);
            shi0 = __reduce_add_sync(0xffffffffu, shi0);
            slo1 = __reduce_add_sync(0xffffffffu, slo1);
            shi1 = __reduce_add_sync(0xffffffffu, shi1);
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
        if (warp >= NW - 2 && NW - 1 - warp < live_rows) {    // warps 15, 14: exact sum of X of one batch row over all warps
            const int em = 2 * nt + (NW - 1 - warp);
            int alo = lane < NW ? s_tx[(lane * 8 + em) * 2] : 0, ahi = lane < NW ? s_tx[(lane * 8 + em) * 2 + 1] : 0;
            alo = __reduce_add_sync(0xffffffffu, alo);
            ahi = __reduce_add_sync(0xffffffffu, ahi);
            if (lane == 0) { s_txs[em * 2] = alo; s_txs[em * 2 + 1] = ahi; }
        }
    }
    issue_range(mid, ntiles);
    stamp(6);

    // scale / zero point of the first output of this thread, fetched now (latency hidden by the main loop)
    const int outs = ntiles * TILE_ROWS * p.M;
    float pre_sc = 0.0f, pre_zp = 0.0f;
    if (NT == 1 && p.slots) {
        const int rest = tid >> 6, row = r0 + (p.M == 1 ? rest : (rest >> 1)) * TILE_ROWS + tile_row((tid >> 2) & 15);
        if ((tid & 3) == 0 && tid < outs * 4 && row < r0 + nrows) { pre_sc = __ldg(p.scales + row); pre_zp = __ldg(p.zps + row); }
    } else if (tid < outs) {
        const int row = r0 + (tid >> 4) / p.M * TILE_ROWS + tile_row(tid & 15);
        if (row < r0 + nrows) { pre_sc = __ldg(p.scales + row); pre_zp = __ldg(p.zps + row); }
    }

    // ---- main loop: one 16-row tile per iteration, this warp's granules warp, warp + 16, ...
    // mma rows g (a0 / a2) and g + 8 (a1 / a3) of this lane <-> tile rows perm(g) and perm(g) + 4:
    // even g -> first half, row g / 2; odd g -> second half, row 8 + g / 2
    const uint32_t lane_off = SPLIT ? (uint32_t)(g * p.pitch + warp * GRAN_B + t * 16)
                                        : (uint32_t)((g & 1) * half_off + (g >> 1) * row_bytes + warp * GRAN_B + t * 16);
    const uint32_t hi_off = SPLIT ? (uint32_t)(8 * p.pitch) : (uint32_t)(4 * row_bytes);
    const uint32_t acc_lane = sbase + p.acc_off + (uint32_t)(((warp % R) * rep_ints + 2 * t * ACC_CS + g) * 4);
    const bool cols_live = p.M > 1 || t < 2;                   // M == 1: mma columns 4..7 are empty
    const int slot_ints = (p.M == 1 ? 4 : 8) * ACC_CS;         // slots mode: ints per (tile, warp)
    if (p.slots_alias) __syncthreads();                        // every lane has read its fragments from the image
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
                hi[q] = lds128(tb + q * NW * GRAN_B + hi_off);
            }
        }
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const uint32_t wl[4] = {lo[q].x, lo[q].y, lo[q].z, lo[q].w};
            const uint32_t wh[4] = {hi[q].x, hi[q].y, hi[q].z, hi[q].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // no nibble extraction: sum_k (q_lo + 16 q_hi) Xe + (16 q_hi) (Xo - Xe) = sum_k q_lo Xe + q_hi 16 Xo.
                // LOP3 / SHF run at half rate on the INT pipe; the 48 of them per tile that a plain
                // nibble -> byte widening costs were what bounded this loop (profiles/r01_gemv_notes.md)
                const uint32_t a0 = wl[j], a2 = wl[j] & 0xf0f0f0f0u;   // row g
                const uint32_t a1 = wh[j], a3 = wh[j] & 0xf0f0f0f0u;   // row g + 8
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) imma(c[nt][j & 1], a0, a1, a2, a3, bf[q][nt][j][0], bf[q][nt][j][1]);
            }
        }
        if (i < 5) stamp(10 + i);
        // the 16 warps add their partials of tile i into acc[i][column][row] (exact, order independent)
        if (NT == 1 && p.slots) {
            // own slot of (tile, warp): plain stores.  (A shared-memory atomic costs the LSU ~10 clk per warp
            // instruction; 320 of them per launch drained for 0.5 us after the last tile.)
            if (cols_live) {
                int* sl = acc + (i * NW + warp) * slot_ints + 2 * t * ACC_CS + g;
                sl[0] = c[0][0][0] + c[0][1][0];
                sl[ACC_CS] = c[0][0][1] + c[0][1][1];
                sl[8] = c[0][0][2] + c[0][1][2];
                sl[ACC_CS + 8] = c[0][0][3] + c[0][1][3];
            }
        } else if (cols_live) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint32_t a = acc_lane + (uint32_t)((i * NT + nt) * ACC_PLANE * 4);
                red_add_s32(a, c[nt][0][0] + c[nt][1][0]);                            // (row g,     col 2t)
                red_add_s32(a + ACC_CS * 4, c[nt][0][1] + c[nt][1][1]);               // (row g,     col 2t + 1)
                red_add_s32(a + 32, c[nt][0][2] + c[nt][1][2]);                       // (row g + 8, col 2t)
                red_add_s32(a + ACC_CS * 4 + 32, c[nt][0][3] + c[nt][1][3]);          // (row g + 8, col 2t + 1)
            }
        }
    }
    stamp(7);
    __syncthreads();
    stamp(8);

    auto finish = [&](long long a, int em, int row, float sc, float zp) {
        const int ex = s_ex[em];
        const long long txl = (long long)s_txs[em * 2] + ((long long)s_txs[em * 2 + 1] << 16);     // sum_k X (exact)
        const int zi = __float2int_rn(zp);
        float v;
        if ((float)zi == zp && zi >= -32768 && zi <= 32767 && ex >= -126) {
            // quantiser-made zero points are integers: a - zp * sum X exactly in s64, ONE rounding to fp32 (the
            // same value the fp64 expression below rounds to), then the exact power of two and the scale
            v = sc * (__ll2float_rn(a - (long long)zi * txl) * __uint_as_float((uint32_t)(127 - ex) << 23));
        } else {
            const double down = __longlong_as_double((long long)(1023 - ex) << 52);                // 2^-e
            v = sc * (float)(((double)a - (double)zp * (double)txl) * down);
        }
        store_out(p.y, p.y_dtype, (int64_t)em * p.N + row, v);
    };
    // with a K split the exact s64 partial of output o goes to slot [slab][o] of the leader's exchange buffer
    auto emit = [&](long long a, int o, int em, int row, float sc, float zp) {
        if (!SPLIT) finish(a, em, row, sc, zp);
        else st_cluster_b64(mapa(sbase + p.xbuf_off + (uint32_t)((slab * p.outs_cap + o) * 8), 0), a);
    };
    if (NT == 1 && p.slots) {
        // four threads per output (one per limb) add the 16 warp slots; the quad combines through shuffles
        const int total = outs * 4;
        for (int idx0 = 0; idx0 < total; idx0 += NTHR) {
            const int idx = idx0 + tid;
            const int lq = idx & 3, er = (idx >> 2) & 15, rest = idx >> 6;
            const int em = p.M == 1 ? 0 : (rest & 1), ti = p.M == 1 ? rest : (rest >> 1);
            const int row = r0 + ti * TILE_ROWS + tile_row(er);
            const bool ok = idx < total && row < r0 + nrows;
            int a32 = 0;
            if (ok) {
                const int* sl = acc + (ti * NW) * slot_ints + (em * 4 + lq) * ACC_CS + er;
#pragma unroll
                for (int w = 0; w < NW; ++w) a32 += sl[w * slot_ints];
            }
            const int l1 = __shfl_down_sync(0xffffffffu, a32, 1), l2 = __shfl_down_sync(0xffffffffu, a32, 2),
                      l3 = __shfl_down_sync(0xffffffffu, a32, 3);
            const long long a = (long long)a32 + ((long long)l1 << 8) + ((long long)l2 << 16) + ((long long)l3 << 24);
            if (ok && lq == 0) {
                float sc, zp;
                if (idx0 == 0) { sc = pre_sc; zp = pre_zp; }
                else { sc = __ldg(p.scales + row); zp = __ldg(p.zps + row); }
                emit(a, (ti * p.M + em) * TILE_ROWS + er, em, row, sc, zp);
            }
        }
    } else
    // ---- epilogue: one thread per output (tile, batch row, row): limbs -> sum_k q*X (exact s64),
    // y = s * 2^-e * (sum_k q*X - zp * sum_k X)
    for (int o = tid; o < outs; o += NTHR) {
        const int er = o & 15, rest = o >> 4;
        const int em = rest % p.M, ti = rest / p.M;
        const int row = r0 + ti * TILE_ROWS + tile_row(er);
        if (row >= r0 + nrows) continue;
        float sc, zp;
        if (o == tid) { sc = pre_sc; zp = pre_zp; }
        else { sc = __ldg(p.scales + row); zp = __ldg(p.zps + row); }
        const int* a4 = acc + (ti * NT * 8 + em * 4) * ACC_CS + er;
        int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            l0 += a4[r * rep_ints]; l1 += a4[r * rep_ints + ACC_CS]; l2 += a4[r * rep_ints + 2 * ACC_CS]; l3 += a4[r * rep_ints + 3 * ACC_CS];
        }
        const long long a = (long long)l0 + ((long long)l1 << 8) + ((long long)l2 << 16) + ((long long)l3 << 24);
        emit(a, o, em, row, sc, zp);
    }
    if (SPLIT) {
        // ---- K split: sum_k X of every slab and the partials above meet in the leader, which finishes the rows
        const uint32_t xtx = sbase + p.xbuf_off + (uint32_t)(nslab * p.outs_cap * 8);
        if (tid < p.M)
            st_cluster_b64(mapa(xtx + (uint32_t)((slab * 8 + tid) * 8), 0),
                           (long long)s_txs[tid * 2] + ((long long)s_txs[tid * 2 + 1] << 16));
        cluster_sync_all();
        if (slab == 0) {
            const long long* xb = reinterpret_cast<const long long*>(smem + p.xbuf_off);
            const long long* xt = xb + nslab * p.outs_cap;
            if (tid < p.M) {
                long long tsum = 0;
                for (int sl = 0; sl < nslab; ++sl) tsum += xt[sl * 8 + tid];
                s_txs[tid * 2] = (int)(tsum & 0xffff);
                s_txs[tid * 2 + 1] = (int)(tsum >> 16);
            }
            named_bar_sync(1, NTHR);
            for (int o = tid; o < outs; o += NTHR) {
                const int er = o & 15, rest = o >> 4;
                const int em = rest % p.M, ti = rest / p.M;
                const int row = r0 + ti * TILE_ROWS + tile_row(er);
                if (row >= r0 + nrows) continue;
                long long a = 0;
                for (int sl = 0; sl < nslab; ++sl) a += xb[sl * p.outs_cap + o];
                finish(a, em, row, __ldg(p.scales + row), __ldg(p.zps + row));
            }
        }
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
    int acc_off, img_off, tile_off, slots, slots_alias;
    int nslab, nrb, pitch, xbuf_off, outs_cap;
    size_t smem;
};

// one candidate: K cut into nslab slabs (cluster of nslab CTAs), the SMs dealt out over the row blocks
bool plan_res_slabs(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, int nslab, ResPlan* c) {
    const int G = (int)(K / GRAN_K);
    const int Gmax = (G + nslab - 1) / nslab;                      // granules of the largest slab
    if (Gmax > 3 * NW || nslab > G) return false;                   // GPW <= 3
    int nrb = sm_count / nslab;
    if (nrb < 1) return false;
    const int min_rb = (int)((N + TILE_ROWS * MAX_TILES - 1) / (TILE_ROWS * MAX_TILES));
    if (nrb > (N + TILE_ROWS - 1) / TILE_ROWS) nrb = (int)((N + TILE_ROWS - 1) / TILE_ROWS);   // >= one tile per CTA
    if (nrb < min_rb) return false;
    const int rows = (int)((N + nrb - 1) / nrb);
    const int ntiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (ntiles > MAX_TILES) return false;
    const int nt = M <= 2 ? 1 : (M <= 4 ? 2 : 4);
    c->nslab = nslab; c->nrb = nrb; c->grid = nrb * nslab; c->gpw = (Gmax + NW - 1) / NW; c->nt = nt; c->ntiles = ntiles;
    if (c->gpw * nt > 8) return false;                              // B fragments must stay in registers
    c->pitch = ((Gmax * GRAN_B + 63) / 128) * 128 + 64;             // >= the slab's bytes of a row, = 64 mod 128
    const int tiles_bytes = nslab > 1 ? ntiles * TILE_ROWS * c->pitch : ntiles * (TILE_ROWS * (int)(K / 2) + 128);
    const int img_rows = M == 1 ? 1 : 2;
    const int img_bytes = img_rows * Gmax * 512;                    // rows x granules x 4 limbs x 16 (t, word) x 8 B
    const int acc_bytes = ACC_REPS_X_NT * ntiles * ACC_PLANE * 4;
    const int slot_bytes = NW * ntiles * (M == 1 ? 4 : 8) * ACC_CS * 4;
    c->outs_cap = ntiles * TILE_ROWS * (int)M;
    const int xbuf_bytes = nslab > 1 ? nslab * (c->outs_cap * 8 + 64) : 0;
    c->acc_off = OFF_END;
    // layouts, roomiest first: (1) acc / slots, then the image; (2) M <= 2: the slots run on into the image (dead by
    // then, one extra barrier); (3) M <= 2: slots and image share one region
    for (int layout = 1; layout <= 3; ++layout) {
        if (layout > 1 && nt != 1) break;
        int region;
        if (layout == 1) { c->img_off = (c->acc_off + (nt == 1 ? slot_bytes : acc_bytes) + 127) / 128 * 128; region = c->img_off + img_bytes; c->slots = nt == 1; c->slots_alias = 0; }
        else if (layout == 2) { c->img_off = (c->acc_off + acc_bytes + 127) / 128 * 128; region = c->img_off + img_bytes; if (region < c->acc_off + slot_bytes) region = c->acc_off + slot_bytes; c->slots = 1; c->slots_alias = 1; }
        else { c->img_off = c->acc_off; region = c->acc_off + (slot_bytes > img_bytes ? slot_bytes : img_bytes); c->slots = 1; c->slots_alias = 1; }
        c->xbuf_off = (region + 127) / 128 * 128;
        c->tile_off = (c->xbuf_off + xbuf_bytes + 127) / 128 * 128;
        c->smem = (size_t)c->tile_off + (size_t)tiles_bytes;
        if (c->smem <= (size_t)max_smem) return true;
    }
    return false;
}

bool plan_res(int sm_count, int max_smem, int64_t M, int64_t N, int64_t K, ResPlan* c) {
    if (M < 1 || M > 8 || K <= 0 || K % GRAN_K != 0 || N < 1 || N > 0x7fffffff) return false;
    const int only = tuning().gemv_slabs;                           // bench hook: force the number of K slabs
    for (int nslab = 1; nslab <= 4; ++nslab)
        if ((only <= 0 || only == nslab) && plan_res_slabs(sm_count, max_smem, M, N, K, nslab, c)) return true;
    return false;
}

template <int GPW, int NT, bool SPLIT>
int launch_res_inst(const ResPlan& c, const ResParams& p, bool pdl, cudaStream_t st) {
    auto kfn = gemv_res_kernel<GPW, NT, SPLIT>;
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
    if (c.nslab > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = (unsigned)c.nslab;
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
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
    p.rows_q = (int)(N / c.nrb); p.rows_rem = (int)(N % c.nrb);
    p.nslab = c.nslab;
    p.gran_q = (int)(K / GRAN_K) / c.nslab; p.gran_rem = (int)(K / GRAN_K) % c.nslab;
    p.pitch = c.pitch; p.xbuf_off = c.xbuf_off; p.outs_cap = c.outs_cap;
    p.acc_off = c.acc_off; p.img_off = c.img_off; p.tile_off = c.tile_off;
    p.slots = c.slots; p.slots_alias = c.slots_alias;
    p.wait_weights = (flags & B200Q_FLAG_STATIC_WEIGHTS) ? 0 : 1;
    p.early_tiles = tuning().gemv_early >= 0 ? tuning().gemv_early : 92;     // default: two tiles up front, the rest behind the x loads (tools/tune_gemv.py sweep)
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
    if (c.gpw == GPW_ && c.nt == NT_) return c.nslab > 1 ? launch_res_inst<GPW_, NT_, true>(c, p, pdl, st) : launch_res_inst<GPW_, NT_, false>(c, p, pdl, st);
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
