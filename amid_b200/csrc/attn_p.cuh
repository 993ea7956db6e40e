// Round-2 causal attention kernels (head_dim 16) on tcgen05 at fp32-level accuracy (FP16-pair split of x3.cuh: every operand
// x = s^-1 (h0 + h1), products h0 h0 + h1 h0 + h0 h1 accumulated in fp32 in tensor memory).  Three kernels:
//   k_attn_fwd_p   forward, one CTA per (sample, head), two CTAs per SM, 64 <= L <= 224
//   k_attn_bwd_t2  backward, two passes (rows = queries for dQ, rows = keys for dV / dK), one CTA per (sample, head), two CTAs
//                  per SM, L <= 224 -- the kernel the x3 train step runs
//   k_attn_bwd_p   backward, ONE pass, 148 persistent warp-specialised CTAs, 64 <= L <= 256 -- runs for 224 < L <= 256 and side
//                  by side with the two-pass kernel in tests/test_gpu_attn.py and tools/prof_attn.py
// What they share (and what changed against attn_tc.cuh, the first tcgen05 version):
//   * operands live in two K-major SWIZZLE_128B ROW tiles [position][q0 | q1 | g0 | g1] and [position][k0 | k1 | v0 | v1]
//     (128-byte rows of FP16 pair pieces).  There are no transposed copies: wherever a GEMM needs an operand the other way
//     round (v in P V, k in dS K, q in dS^T Q, dO in Pd^T dO) the row tile is addressed as an MN-major B operand with a
//     32 / 64-byte offset inside the swizzled row;
//   * piece products by stacking: neighbouring pieces [b0 | b1] are one N = 32 operand, so a k-step costs two MMAs instead
//     of three (the epilogue adds the two 16-column halves); in the single-pass kernel the two A pieces of the staged
//     Pd / dS tiles are the two 64-row halves of one M = 128 MN-major operand (leading-dimension offset = the distance
//     between the piece tiles);
//   * every MMA is issued by one elect.sync lane of a converged warp with descriptors built from a low word + constant high
//     word: ptxas emits back-to-back UTCHMMA (a `threadIdx.x == 0` branch costs a ~12-instruction election loop per MMA);
//   * the exp / dropout / split loops are specialised on training mode and on diagonal chunks; the keep test of a dropout
//     byte is one shift + one unsigned compare.
// The single-pass kernel (k_attn_bwd_p): a unit = [128 queries x 64 keys]; S = Q K^T and dP = dO V^T land in one of two TMEM
// buffers two units ahead; each of the 16 element-wise warps turns 32 rows x 16 keys into Pd = dropout(P) and dS, stores dS
// in place (A operand of dQ += dS [k0 | k1]) and Pd, dS into one of two staging tile sets in shared memory -- the SAME bytes
// are the MN-major A operand of dV += Pd^T [g0 | g1] and dK += dS^T q.  Three issuing warps (S / dP + dQ, dK, dV) and the
// element-wise warps hand buffers over through mbarriers; the next head is prefetched into L2 and converted while the last
// MMAs drain.  Measured at the C3 shape: forward 313 us, two-pass backward 741 us, single-pass backward 805 us.
// mbarrier rule followed everywhere: two commits on one barrier are separated by a CTA barrier (or a dependency chain) that
// every waiter of the first has passed -- otherwise a late waker can find the barrier two phases ahead and wait forever.
// Arithmetic follows torch/nn/functional.py:6630-6647 (q pre-scaled by 0.25, -inf above the diagonal, softmax, dropout
// without renormalisation, P v); the dropout bits are the same counter hash as every other attention kernel.
#pragma once
#include "attn_tc.cuh"

namespace amid {
namespace attn_p {
using namespace tc;
using namespace attn_tc;

constexpr int PMAXL = 256, PMINL = 64;
constexpr int NTH = 608;                            // 16 element-wise warps + 3 MMA-issuing warps
constexpr int ROWT_BYTES = 256 * 128;               // row tile: 256 positions x [4 pieces of 16 fp16]
constexpr int STG_TILE = 128 * 128;                 // [128 query rows][64 keys] fp16
constexpr int STG_SET = 4 * STG_TILE;               // Pd piece 0, Pd piece 1, dS piece 0, dS piece 1
constexpr size_t PBWD_SMEM = 2 * (size_t)ROWT_BYTES + 2 * (size_t)STG_SET + 1024;
constexpr uint32_t C_DP = 64, C_DQ = 256, C_DK = 320, C_DV = 384;   // buffer b: S at 128 b, dP at 128 b + 64; dQ[qt] 32 cols; dK[kh] 16; dV[kh] 32

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// descriptor halves: low word = (address >> 4) | (leading-dimension offset >> 4) << 16, high word constant
// (stride offset 1024 B, version 1, SWIZZLE_128B); an address offset of x bytes is + (x >> 4) on the low word
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t dlo_k(uint32_t addr) { return (addr >> 4) | (1u << 16); }                       // lbo = 16
__device__ __forceinline__ uint32_t dlo_m(uint32_t addr) { return (addr >> 4) | ((uint32_t)(STG_TILE >> 4) << 16); } // lbo = one staging tile
__device__ __forceinline__ void mma_lo(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(idesc), "r"(accum), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void mma_lo_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(idesc), "r"(accum), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// S = Q K^T (3 piece products) and dP = dO V^T into [tS, tS + 64) and [tS + 64, tS + 128); al / bl = low descriptor
// words of the first query row / first key row
__device__ __forceinline__ void issue_sdp(uint32_t tS, uint32_t al, uint32_t bl, int N) {
    const uint32_t id = idesc_f16(N, false, false);
    mma_lo(tS, al, bl, id, 0u);                      // q0 k0
    mma_lo(tS, al + 2, bl, id, 1u);                  // q1 k0
    mma_lo(tS, al, bl + 2, id, 1u);                  // q0 k1
    mma_lo(tS + C_DP, al + 4, bl + 4, id, 0u);       // g0 v0
    mma_lo(tS + C_DP, al + 6, bl + 4, id, 1u);       // g1 v0
    mma_lo(tS + C_DP, al + 4, bl + 6, id, 1u);       // g0 v1
}
// unit u of a head: key blocks of 64 outer, query tiles of 128 from the diagonal on
__device__ __forceinline__ void unit_at(int u, int nqt, int& kh, int& qt) {
    kh = 0;
    for (;;) {
        const int c = nqt - (kh >> 1);
        if (u < c) break;
        u -= c;
        ++kh;
    }
    qt = (kh >> 1) + u;
}

#ifdef AMID_ATTN_DBG
__device__ long long g_dbg[20][256];
#define DBG(tag) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && dbg_i < 254) { g_dbg[threadIdx.x >> 5][dbg_i++] = (long long)(tag); g_dbg[threadIdx.x >> 5][dbg_i++] = clock64(); } } while (0)
#else
#define DBG(tag) do { } while (0)
#endif

struct ShPB {
    uint64_t bar_tiles;       // 16 arrivals: the operand tiles (and ls / dl) of the current head are written
    uint64_t bar_sready[2];   // tcgen05.commit: S and dP of set s are in tensor memory
    uint64_t bar_full[2];     // 16 arrivals: every warp stored its part of the staging tiles and is done reading S / dP
    uint64_t bar_free[2];     // 2 x tcgen05.commit (dK and dV issuers): the MMAs that read the staging tiles of set s are complete
    uint64_t bar_done;        // 3 x tcgen05.commit: every MMA of the head is complete
    uint64_t bar_kvgo[2];     // 1 arrival per unit (indexed by unit parity): the S / dP issuer has queued its MMAs, dK / dV go behind
                              // them.  Two barriers: with one, a dK / dV issuer that wakes up late could find it two phases ahead
    uint32_t tmem;
    float red[4][16];
    float ls[256];            // lse * log2(e) per query (+inf for rows >= L)
    float dl[256];            // delta_i = <dO_i, O_i>
};

struct HeadScal { float f, fdp, sc_q, sc_k, sc_v; };
struct HeadRegs {             // one head's rows as fetched from HBM: thread = (row r0 + 128 i, features 4 c4 .. 4 c4 + 3)
    float4 q[2], k[2], v[2], g[2], o[2];
    float lse[2];
};
__device__ __forceinline__ void fetch_head(HeadRegs& h, const float* __restrict__ q, const float* __restrict__ k,
                                           const float* __restrict__ v, const float* __restrict__ o, const float* __restrict__ dO,
                                           const float* __restrict__ lse, int bh, int L, int tid) {
    const int b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    const int c4 = tid & 3, r0 = tid >> 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int R = r0 + 128 * i;
        h.q[i] = h.k[i] = h.v[i] = h.g[i] = h.o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        h.lse[i] = INFINITY;
        if (R < L) {
            const size_t off = base + (size_t)R * D + 4 * c4;
            h.q[i] = __ldg(reinterpret_cast<const float4*>(q + off));
            h.k[i] = __ldg(reinterpret_cast<const float4*>(k + off));
            h.v[i] = __ldg(reinterpret_cast<const float4*>(v + off));
            h.g[i] = __ldg(reinterpret_cast<const float4*>(dO + off));
            h.o[i] = __ldg(reinterpret_cast<const float4*>(o + off));
            if (c4 == 0) h.lse[i] = __ldg(lse + (size_t)bh * L + R);
        }
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
    return pred != 0;
}
struct Geo {                  // per-launch geometry (depends on L only)
    int L, nqt, NKP, nkh, nunits;
};
__device__ __forceinline__ Geo make_geo(int L) {
    Geo g;
    g.L = L;
    g.nqt = (L + 127) >> 7;
    g.NKP = (L + 15) & ~15;
    g.nkh = (g.NKP + 63) >> 6;
    g.nunits = 0;
    for (int kh = 0; kh < g.nkh; ++kh) g.nunits += g.nqt - (kh >> 1);
    return g;
}

// ---- MMA-issuing warps.  The whole warp runs the control flow (warp-uniform values stay in uniform registers and
// ptxas emits back-to-back UTCHMMA); elect.sync picks the lane that issues.
// role 0: S / dP of both sets and dQ.
__device__ __forceinline__ void mma_role_sq(ShPB& sh, const Geo g, uint32_t tmem, uint32_t qg, uint32_t kv, uint32_t stg, int nbh) {
    constexpr uint32_t idq = idesc_f16(32, false, true);
    const uint32_t qgl = dlo_k(qg), kvl = dlo_k(kv);
    uint32_t cnt_full[2] = {0u, 0u}, nhead = 0;
    int dbg_i = 0;
    for (int bh = blockIdx.x; bh < nbh; bh += gridDim.x, ++nhead) {
        DBG(100);
        mbar_wait(&sh.bar_tiles, nhead & 1);
        fence_after();
        DBG(101);
        for (int u = 0; u < min(2, g.nunits); ++u) {
            int kh, qt;
            unit_at(u, g.nqt, kh, qt);
            if (elect_one()) {
                issue_sdp(tmem + 128 * u, qgl + (uint32_t)(128 * qt) * 8, kvl + (uint32_t)(64 * kh) * 8, min(64, g.NKP - 64 * kh));
                mma_commit(&sh.bar_sready[u]);
            }
            __syncwarp();
        }
        for (int u = 0; u < g.nunits; ++u) {
            const int s = u & 1;
            int kh, qt, kh2 = 0, qt2 = 0;
            unit_at(u, g.nqt, kh, qt);
            const bool more = u + 2 < g.nunits;
            if (more) unit_at(u + 2, g.nqt, kh2, qt2);
            const uint32_t a2 = qgl + (uint32_t)(128 * qt2) * 8, b2 = kvl + (uint32_t)(64 * kh2) * 8;
            const int N2 = min(64, g.NKP - 64 * kh2);
            const int nkq = min(64, g.NKP - 64 * kh) >> 4;
            const uint32_t tq = tmem + C_DQ + 32 * qt;
            const uint32_t ta = tmem + 128 * s, bl = kvl + (uint32_t)(64 * kh) * 8;
            mbar_wait(&sh.bar_full[s], cnt_full[s] & 1);
            ++cnt_full[s];
            fence_after();
            DBG(110 + u);
            if (elect_one()) {
                // dQ[qt] += dS K : A = dS pieces in tensor memory (in place over the consumed score columns: chunk c holds
                // piece 0 in columns [16c, 16c+8) and piece 1 in [16c+8, 16c+16)), B = [k0 | k1] MN-major, N = 32
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks < nkq) {
                        mma_lo_ts(tq, ta + 16 * ks, bl + ks * 128, idq, (kh > 0 || ks > 0) ? 1u : 0u);
                        mma_lo_ts(tq, ta + 16 * ks + 8, bl + ks * 128, idq, 1u);
                    }
                }
                // the same thread issues the next S / dP into this buffer: tcgen05.mma of one thread execute in order
                if (more) {
                    issue_sdp(tmem + 128 * s, a2, b2, N2);
                    mma_commit(&sh.bar_sready[s]);
                }
                if (u == g.nunits - 1) mma_commit(&sh.bar_done);
                mbar_arrive(&sh.bar_kvgo[s]);
            }
            __syncwarp();
            DBG(120 + u);
        }
    }
}
// role 1: dK[kh] += dS^T Q;  role 2: dV[kh] += Pd^T dO.  A = the two pieces stacked along M (MN-major, leading-dimension
// offset = one staging tile), B = b0 then b1 (MN-major, N = 16) into the same 16 accumulator columns.
template <int ROLE>
__device__ __forceinline__ void mma_role_kv(ShPB& sh, const Geo g, uint32_t tmem, uint32_t qg, uint32_t stg, int nbh) {
    constexpr uint32_t idk = idesc_f16(16, true, true), idv = idesc_f16(32, true, true);
    const uint32_t qgl = dlo_k(qg) + (ROLE == 2 ? 4u : 0u);
    uint32_t cnt_full[2] = {0u, 0u};
    for (int bh = blockIdx.x; bh < nbh; bh += gridDim.x) {
        for (int u = 0; u < g.nunits; ++u) {
            const int s = u & 1;
            int kh, qt;
            unit_at(u, g.nqt, kh, qt);
            const int nks = min(8, (min(g.L, 128 * (qt + 1)) - 128 * qt + 15) >> 4);
            const int ks0 = max(0, (64 * kh - 128 * qt) >> 4);        // queries before the first key of the block are masked
            const uint32_t td = ROLE == 1 ? tmem + C_DK + 16 * kh : tmem + C_DV + 32 * kh;
            const uint32_t al = dlo_m(stg + (uint32_t)s * STG_SET + (ROLE == 1 ? 2 * STG_TILE : 0));
            const uint32_t bl = qgl + (uint32_t)(128 * qt) * 8;
            const bool first = qt == (kh >> 1);
            mbar_wait(&sh.bar_full[s], cnt_full[s] & 1);
            mbar_wait(&sh.bar_kvgo[s], cnt_full[s] & 1);
            ++cnt_full[s];
            fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (ks >= ks0 && ks < nks) {
                        const uint32_t acc = (first && ks == ks0) ? 0u : 1u;
                        if (ROLE == 1) {
                            mma_lo(td, al + ks * 128, bl + ks * 128, idk, acc);
                            mma_lo(td, al + ks * 128, bl + ks * 128 + 2, idk, 1u);
                        } else {
                            mma_lo(td, al + ks * 128, bl + ks * 128, idv, acc);        // [a0 ; a1] x [g0 | g1]
                        }
                    }
                }
                mma_commit(&sh.bar_free[s]);
                if (u == g.nunits - 1) mma_commit(&sh.bar_done);
            }
            __syncwarp();
        }
    }
}

// one chunk of a unit: 16 keys of this thread's query row.  S, dP (raw accumulators) -> Pd = dropout(P), dS as FP16 pairs
template <bool TRAIN, bool DIAG>
__device__ __forceinline__ void bwd_chunk(uint32_t ts, uint32_t tp, float f, float li, float Di, float fdp, float dsc, int nvalid,
                                          uint32_t seed, uint32_t site, uint32_t g0, uint32_t thr24,
                                          uint32_t (&pd0)[8], uint32_t (&pd1)[8], uint32_t (&ds0)[8], uint32_t (&ds1)[8]) {
    float sx[16], dp[16];
    tmem_ld16(ts, sx);
    tmem_ld16(tp, dp);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t r = 0u;
        if (TRAIN) r = rng4(seed, site, (uint64_t)(g0 + g));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cc = 4 * g + e;
            float p = ex2(fmaf(sx[cc], f, -li));
            if (DIAG) p = cc < nvalid ? p : 0.f;
            float k2 = dsc;
            if (TRAIN) k2 = (e == 3 ? r : (r << (24 - 8 * e))) >= thr24 ? dsc : 0.f;
            sx[cc] = p * k2;                                   // Pd
            dp[cc] = p * fmaf(dp[cc] * k2, fdp, -Di);          // dS (scaled by sds)
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        split_f16x2(sx[2 * e], sx[2 * e + 1], pd0[e], pd1[e]);
        split_f16x2(dp[2 * e], dp[2 * e + 1], ds0[e], ds1[e]);
    }
}

template <bool TRAIN>
__global__ void __launch_bounds__(NTH, 1)
k_attn_bwd_p(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
             const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
             float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, int nbh, DropCfg dc, uint32_t site) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ ShPB sh;
    uint8_t* QG = align1k(smem_raw);                // [position][q0 | q1 | g0 | g1]
    uint8_t* KV = QG + ROWT_BYTES;                  // [position][k0 | k1 | v0 | v1]
    uint8_t* STG = KV + ROWT_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Geo g = make_geo(L);
    const int nqt = g.nqt, NKP = g.NKP, nkh = g.nkh, Lp4 = ((L + 3) & ~3) >> 2;
    if (warp == 16) tmem_alloc(&sh.tmem, 512);
    if (tid == 0) {
        mbar_init(&sh.bar_tiles, 16);
        for (int s = 0; s < 2; ++s) { mbar_init(&sh.bar_sready[s], 1); mbar_init(&sh.bar_full[s], 16); mbar_init(&sh.bar_free[s], 2); }
        mbar_init(&sh.bar_done, 3);
        mbar_init(&sh.bar_kvgo[0], 1);
        mbar_init(&sh.bar_kvgo[1], 1);
        fence_barrier_init();
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = sh.tmem;
    const uint32_t qg = smem_u32(QG), kv = smem_u32(KV), stg = smem_u32(STG);

    if (warp == 16) {
        mma_role_sq(sh, g, tmem, qg, kv, stg, nbh);
    } else if (warp == 17) {
        mma_role_kv<1>(sh, g, tmem, qg, stg, nbh);
    } else if (warp == 18) {
        mma_role_kv<2>(sh, g, tmem, qg, stg, nbh);
    } else {
        // =============================================================== element-wise warps
        const int lq = warp & 3, cq = warp >> 2;          // lane quarter (query rows), 16-key chunk of a unit
        const uint32_t tl = tmem + ((uint32_t)(32 * lq) << 16);
        const int c4 = tid & 3, r0 = tid >> 2;
        const int rr = 32 * lq + lane;                                         // row inside a query tile
        const uint32_t rowb = (uint32_t)((rr >> 3) * 1024 + (rr & 7) * 128);
        const uint32_t x0 = (((uint32_t)(2 * cq) ^ (uint32_t)rr) & 7u) << 4, x1 = x0 ^ 16u;   // 16-byte units of this chunk
        const uint32_t thr24 = dc.thr16 << 24;
        const float dsc = TRAIN ? dc.scale : 1.0f;
        uint32_t cnt_buf[2] = {0u, 0u}, nhead = 0;       // units started on each buffer (all heads); heads started
        int dbg_i = 0;
        // one head's rows, converted: FP16 pair pieces of this thread's 2 x 4 features of q, dO, k, v; delta; lse
        uint2 cv[2][8];
        float cdl[2], cls[2];
        HeadScal hs;                                     // of the converted head
        HeadRegs hr;
        // stage A: rows from L2 / HBM into registers, the head after it into L2, per-warp maxima into shared memory
        auto prep_a = [&](int hb) {
            fetch_head(hr, q, k, v, o, dO, lse, hb, L, tid);
            const int nb = hb + gridDim.x;
            if (nb < nbh && tid < L) {               // 64-byte row slices: one prefetch per (tensor, row)
                const size_t noff = (size_t)(nb / H) * L * D + (nb % H) * DH + (size_t)tid * D;
                prefetch_l2(q + noff); prefetch_l2(k + noff); prefetch_l2(v + noff); prefetch_l2(dO + noff); prefetch_l2(o + noff);
            }
            float mx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                mx[0] = amax4(hr.q[i], mx[0]); mx[1] = amax4(hr.k[i], mx[1]); mx[2] = amax4(hr.v[i], mx[2]); mx[3] = amax4(hr.g[i], mx[3]);
            }
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const float w = warp_max(mx[kx]);
                if (lane == 0) sh.red[kx][warp] = w;
            }
        };
        // stage B (after a barrier of the 512 element-wise threads): power-of-two scales, FP16 pair pieces, delta
        auto prep_b = [&]() {
            float mx[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                float r = sh.red[kx][lane & 15];
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, off));
                mx[kx] = r;
            }
            float sq, iq, sk, ik, sv, iv, sg, ig, sds, ids;
            pow2_scale(mx[0], sq, iq); pow2_scale(mx[1], sk, ik); pow2_scale(mx[2], sv, iv); pow2_scale(mx[3], sg, ig);
            pow2_scale(64.0f * mx[3] * mx[2] * dsc, sds, ids);        // |dS| <= |dPd| + |delta| <= 2 * 16 gmax vmax scale
            hs.f = iq * ik * LOG2E;                  // raw S -> log2 domain
            hs.fdp = ig * iv * sds;                  // raw dP (times the keep scale) -> dPd * sds
            hs.sc_q = 0.25f * ids * ik; hs.sc_k = ids * iq; hs.sc_v = ig;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float dsum = hr.g[i].x * hr.o[i].x + hr.g[i].y * hr.o[i].y + hr.g[i].z * hr.o[i].z + hr.g[i].w * hr.o[i].w;
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                cdl[i] = dsum * sds;
                cls[i] = hr.lse[i] * LOG2E;
                split4(hr.q[i], sq, cv[i][0], cv[i][1]);
                split4(hr.g[i], sg, cv[i][2], cv[i][3]);
                split4(hr.k[i], sk, cv[i][4], cv[i][5]);
                split4(hr.v[i], sv, cv[i][6], cv[i][7]);
            }
        };
        if ((int)blockIdx.x < nbh) {
            prep_a(blockIdx.x);
            named_sync(1, 512);
            prep_b();
        }
        for (int bh = blockIdx.x; bh < nbh; bh += gridDim.x, ++nhead) {
            DBG(1);
            const int b = bh / H, hd = bh % H;
            const size_t base = (size_t)b * L * D + hd * DH;
            const HeadScal cur = hs;
            // ---- this head's converted rows into the operand tiles (every MMA of the previous head is complete)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int R = r0 + 128 * i;
                if (c4 == 0) { sh.dl[R] = cdl[i]; sh.ls[R] = cls[i]; }
                // 16-byte unit u of row R sits at ((u ^ R) & 7) << 4: pieces at units {0,1}, {2,3}, {4,5}, {6,7}
                const uint32_t ob = (uint32_t)((R >> 3) * 1024 + (R & 7) * 128) + (((((uint32_t)c4 >> 1) ^ (uint32_t)R) & 7u) << 4) + (c4 & 1) * 8;
                *reinterpret_cast<uint2*>(QG + ob) = cv[i][0];
                *reinterpret_cast<uint2*>(QG + (ob ^ 0x20u)) = cv[i][1];
                *reinterpret_cast<uint2*>(QG + (ob ^ 0x40u)) = cv[i][2];
                *reinterpret_cast<uint2*>(QG + (ob ^ 0x60u)) = cv[i][3];
                *reinterpret_cast<uint2*>(KV + ob) = cv[i][4];
                *reinterpret_cast<uint2*>(KV + (ob ^ 0x20u)) = cv[i][5];
                *reinterpret_cast<uint2*>(KV + (ob ^ 0x40u)) = cv[i][6];
                *reinterpret_cast<uint2*>(KV + (ob ^ 0x60u)) = cv[i][7];
            }
            fence_async_smem();
            named_sync(1, 512);                       // ls / dl visible to every element-wise thread
            if (lane == 0) mbar_arrive(&sh.bar_tiles);
            DBG(3);
            const uint32_t rbase = ((uint32_t)bh + dc.bh_off) * (uint32_t)L;

            // ---- every unit of the head: this warp owns one 16-key chunk of its 32 query rows
#pragma unroll 1
            for (int u = 0; u < g.nunits; ++u) {
                const int bf = u & 1;
                int kh, qt;
                unit_at(u, nqt, kh, qt);
                const int Rw = 128 * qt + 32 * lq, i = Rw + lane;
                const int rowlim = min(128 * (qt + 1), (L + 15) & ~15);      // query rows the MMAs of this unit read
                const int j0 = 64 * kh + 16 * cq;                            // first key of the chunk
                const bool store = Rw < rowlim && j0 < NKP;
                const bool work = store && j0 <= Rw + 31 && Rw < L;
                // the dK / dV MMAs skip the query k-steps that lie entirely before the key block: those staging rows are never read
                const bool sstore = store && 32 * lq + 32 > max(0, 64 * kh - 128 * qt);
                uint32_t pd0[8], pd1[8], ds0[8], ds1[8];
                DBG(90 + u);
                mbar_wait(&sh.bar_sready[bf], cnt_buf[bf] & 1);
                fence_after();
                DBG(10 + u);
                if (work) {
                    const float li = sh.ls[i], Di = sh.dl[i];
                    const uint32_t g0 = (rbase + (uint32_t)min(i, L - 1)) * (uint32_t)Lp4 + (uint32_t)(j0 >> 2);
                    const uint32_t ts = tl + 128 * bf + 16 * cq;
                    if (j0 + 15 > Rw)
                        bwd_chunk<TRAIN, true>(ts, ts + C_DP, cur.f, li, Di, cur.fdp, dsc, i - j0 + 1, dc.seed, site, g0, thr24, pd0, pd1, ds0, ds1);
                    else
                        bwd_chunk<TRAIN, false>(ts, ts + C_DP, cur.f, li, Di, cur.fdp, dsc, 16, dc.seed, site, g0, thr24, pd0, pd1, ds0, ds1);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) pd0[e] = pd1[e] = ds0[e] = ds1[e] = 0u;
                }
                // the MMAs that read this buffer's previous staging tiles (unit u - 2)
                DBG(20 + u);
                if (cnt_buf[bf] > 0) mbar_wait(&sh.bar_free[bf], (cnt_buf[bf] - 1) & 1);
                ++cnt_buf[bf];
                DBG(30 + u);
                if (store) {
                    tmem_st8(tl + 128 * bf + 16 * cq, ds0);            // dS pieces in place: the A operand of dQ += dS K
                    tmem_st8(tl + 128 * bf + 16 * cq + 8, ds1);
                }
                if (sstore) {
                    uint8_t* sb = STG + bf * STG_SET + rowb;
                    *reinterpret_cast<uint4*>(sb + x0) = make_uint4(pd0[0], pd0[1], pd0[2], pd0[3]);
                    *reinterpret_cast<uint4*>(sb + x1) = make_uint4(pd0[4], pd0[5], pd0[6], pd0[7]);
                    *reinterpret_cast<uint4*>(sb + STG_TILE + x0) = make_uint4(pd1[0], pd1[1], pd1[2], pd1[3]);
                    *reinterpret_cast<uint4*>(sb + STG_TILE + x1) = make_uint4(pd1[4], pd1[5], pd1[6], pd1[7]);
                    *reinterpret_cast<uint4*>(sb + 2 * STG_TILE + x0) = make_uint4(ds0[0], ds0[1], ds0[2], ds0[3]);
                    *reinterpret_cast<uint4*>(sb + 2 * STG_TILE + x1) = make_uint4(ds0[4], ds0[5], ds0[6], ds0[7]);
                    *reinterpret_cast<uint4*>(sb + 3 * STG_TILE + x0) = make_uint4(ds1[0], ds1[1], ds1[2], ds1[3]);
                    *reinterpret_cast<uint4*>(sb + 3 * STG_TILE + x1) = make_uint4(ds1[4], ds1[5], ds1[6], ds1[7]);
                }
                DBG(60 + u);
                fence_async_smem();
                DBG(70 + u);
                tmem_st_wait();
                fence_before();
                DBG(80 + u);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.bar_full[bf]);
                DBG(40 + u);
            }
            // ---- the next head: fetch (from L2), scales and conversion while this head's last MMAs drain
            DBG(50);
            if (bh + (int)gridDim.x < nbh) {
                prep_a(bh + gridDim.x);
                named_sync(1, 512);
                prep_b();
            }
            DBG(51);
            mbar_wait(&sh.bar_done, nhead & 1);
            fence_after();
            // The exchange buffer below aliases staging set 0.  Its stores are ordered after every warp's last staging stores
            // through bar_full -> tcgen05.commit -> bar_done; this CTA barrier states the same order in a form that
            // compute-sanitizer's racecheck can see (it does not follow commits that arrive on an mbarrier).
            named_sync(1, 512);
            DBG(52);
            // ---- accumulators -> HBM.  warp = (lane quarter lq, c): dQ of query tile c, dK / dV of key block c
            float* xch = reinterpret_cast<float*>(STG);          // [key block][tensor][64 keys][20] (staging tiles are free now)
            if (cq < nqt && 128 * cq + 32 * lq < L) {
                const int i = 128 * cq + 32 * lq + lane;
                float a[16], bq[16];
                tmem_ld16(tl + C_DQ + 32 * cq, a);
                tmem_ld16(tl + C_DQ + 32 * cq + 16, bq);
                if (i < L) {
                    float4* dst = reinterpret_cast<float4*>(dq + base + (size_t)i * D);
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        dst[x] = make_float4((a[4 * x] + bq[4 * x]) * cur.sc_q, (a[4 * x + 1] + bq[4 * x + 1]) * cur.sc_q,
                                             (a[4 * x + 2] + bq[4 * x + 2]) * cur.sc_q, (a[4 * x + 3] + bq[4 * x + 3]) * cur.sc_q);
                }
            }
            // lanes [64,128) hold the products of the second A piece: through shared memory to the lanes of the first
            if (cq < nkh && lq >= 2) {
                float a[16], a2[16];
                float4* d0 = reinterpret_cast<float4*>(xch + ((size_t)(cq * 2 + 0) * 64 + 32 * (lq - 2) + lane) * 20);
                float4* d1 = reinterpret_cast<float4*>(xch + ((size_t)(cq * 2 + 1) * 64 + 32 * (lq - 2) + lane) * 20);
                tmem_ld16(tl + C_DK + 16 * cq, a);
#pragma unroll
                for (int x = 0; x < 4; ++x) d0[x] = make_float4(a[4 * x], a[4 * x + 1], a[4 * x + 2], a[4 * x + 3]);
                tmem_ld16(tl + C_DV + 32 * cq, a);
                tmem_ld16(tl + C_DV + 32 * cq + 16, a2);
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    d1[x] = make_float4(a[4 * x] + a2[4 * x], a[4 * x + 1] + a2[4 * x + 1], a[4 * x + 2] + a2[4 * x + 2], a[4 * x + 3] + a2[4 * x + 3]);
            }
            fence_before();
            named_sync(1, 512);
            if (cq < nkh && lq < 2 && 64 * cq + 32 * lq < L) {
                const int j = 64 * cq + 32 * lq + lane;
                const float4* s0 = reinterpret_cast<const float4*>(xch + ((size_t)(cq * 2 + 0) * 64 + 32 * lq + lane) * 20);
                const float4* s1 = reinterpret_cast<const float4*>(xch + ((size_t)(cq * 2 + 1) * 64 + 32 * lq + lane) * 20);
                float4* dK = reinterpret_cast<float4*>(dk + base + (size_t)j * D);
                float4* dV = reinterpret_cast<float4*>(dv + base + (size_t)j * D);
                float a[16], a2[16];
                tmem_ld16(tl + C_DK + 16 * cq, a);
                if (j < L) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float4 w0 = s0[x];
                        dK[x] = make_float4((a[4 * x] + w0.x) * cur.sc_k, (a[4 * x + 1] + w0.y) * cur.sc_k, (a[4 * x + 2] + w0.z) * cur.sc_k,
                                            (a[4 * x + 3] + w0.w) * cur.sc_k);
                    }
                }
                tmem_ld16(tl + C_DV + 32 * cq, a);
                tmem_ld16(tl + C_DV + 32 * cq + 16, a2);
                if (j < L) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float4 w1 = s1[x];
                        dV[x] = make_float4((a[4 * x] + a2[4 * x] + w1.x) * cur.sc_v, (a[4 * x + 1] + a2[4 * x + 1] + w1.y) * cur.sc_v,
                                            (a[4 * x + 2] + a2[4 * x + 2] + w1.z) * cur.sc_v, (a[4 * x + 3] + a2[4 * x + 3] + w1.w) * cur.sc_v);
                    }
                }
            }
            fence_before();
            DBG(53);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}


// ================================================================================================
// forward: one CTA per (sample, head), 256 threads, two CTAs per SM (tensor memory: 224 score columns + 32 for O).
// Against attn_tc::k_attn_fwd_tc: v enters P V as an MN-major B operand straight from the row tile [k0 | k1 | v0 | v1]
// (no transposed copies, no 2-byte scatter stores in the prologue), [v0 | v1] is ONE N = 32 operand (two MMAs per
// k-step, O = columns [0,16) + [16,32)), warp 0 issues the MMAs through elect.sync (back-to-back UTCHMMA), the
// exp / dropout / split loop is specialised on training mode and on diagonal chunks.
// ================================================================================================
constexpr size_t PFWD_SMEM = 2 * (size_t)ROWT_BYTES + 1024;
constexpr uint32_t CF_O = 224;

struct ShPF {
    uint64_t bar;
    uint32_t tmem;
    float red[3][8];
    float xm[2][128];
    float xl[2][128];
};

// P = 2^(S f - m) of 32 consecutive keys, row sum, dropout, FP16 pairs in place
template <bool TRAIN, bool DIAG>
__device__ __forceinline__ float fwd_chunk(uint32_t taddr, float f, float m, int lane, uint32_t seed, uint32_t site, uint32_t g0,
                                           uint32_t thr24) {
    float a[32];
    uint32_t p0[16], p1[16];
    tmem_ld32(taddr, a);
    float l = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint32_t r = 0u;
        if (TRAIN) r = rng4(seed, site, (uint64_t)(g0 + g));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cc = 4 * g + e;
            float p = ex2(fmaf(a[cc], f, -m));
            if (DIAG) p = cc <= lane ? p : 0.f;
            l += p;
            if (TRAIN) p = (e == 3 ? r : (r << (24 - 8 * e))) >= thr24 ? p : 0.f;
            a[cc] = p;
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) split_f16x2(a[2 * e], a[2 * e + 1], p0[e], p1[e]);
    tmem_st16(taddr, p0);
    tmem_st16(taddr + 16, p1);
    return l;
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 2)
k_attn_fwd_p(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, float* __restrict__ o,
             float* __restrict__ lse, int L, DropCfg dc, uint32_t site) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ ShPF sh;
    uint8_t* Qp = align1k(smem_raw);                // [position][q0 | q1 | - | -]
    uint8_t* KV = Qp + ROWT_BYTES;                  // [position][k0 | k1 | v0 | v1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    const int ntile = (L + 127) >> 7, NKP = (L + 15) & ~15, Lp4 = ((L + 3) & ~3) >> 2;
    if (warp == 0) tmem_alloc(&sh.tmem, 256);
    if (tid == 0) { mbar_init(&sh.bar, 1); fence_barrier_init(); }
    // ---- head slices -> registers, per-slice maxima, FP16 pair row tiles
    const int c4 = tid & 3, r0 = tid >> 2;
    float4 vq[4], vk[4], vv[4];
    float mx[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        vq[i] = vk[i] = vv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (R < L) {
            const size_t off = base + (size_t)R * D + 4 * c4;
            vq[i] = __ldg(reinterpret_cast<const float4*>(q + off));
            vk[i] = __ldg(reinterpret_cast<const float4*>(k + off));
            vv[i] = __ldg(reinterpret_cast<const float4*>(v + off));
        }
        mx[0] = amax4(vq[i], mx[0]); mx[1] = amax4(vk[i], mx[1]); mx[2] = amax4(vv[i], mx[2]);
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
        const float w = warp_max(mx[kx]);
        if (lane == 0) sh.red[kx][warp] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
        float r = sh.red[kx][lane & 7];
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, off));
        mx[kx] = r;
    }
    float sq, iq, sk, ik, sv, iv;
    pow2_scale(mx[0], sq, iq); pow2_scale(mx[1], sk, ik); pow2_scale(mx[2], sv, iv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        if (R >= NKP) continue;                     // rows the MMAs never read (rows [L, NKP) are stored as zeros)
        const uint32_t ob = (uint32_t)((R >> 3) * 1024 + (R & 7) * 128) + (((((uint32_t)c4 >> 1) ^ (uint32_t)R) & 7u) << 4) + (c4 & 1) * 8;
        uint2 p0, p1;
        split4(vq[i], sq, p0, p1);
        *reinterpret_cast<uint2*>(Qp + ob) = p0;
        *reinterpret_cast<uint2*>(Qp + (ob ^ 0x20u)) = p1;
        split4(vk[i], sk, p0, p1);
        *reinterpret_cast<uint2*>(KV + ob) = p0;
        *reinterpret_cast<uint2*>(KV + (ob ^ 0x20u)) = p1;
        split4(vv[i], sv, p0, p1);
        *reinterpret_cast<uint2*>(KV + (ob ^ 0x40u)) = p0;
        *reinterpret_cast<uint2*>(KV + (ob ^ 0x60u)) = p1;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    {   // head slices of the CTA that will follow on this SM slot (CTAs are dispatched in index order, two per SM) into L2
        const int nb = bh + 2 * 148;
        if (nb < (int)gridDim.x && tid < L) {
            const size_t noff = (size_t)(nb / H) * L * D + (nb % H) * DH + (size_t)tid * D;
            prefetch_l2(q + noff); prefetch_l2(k + noff); prefetch_l2(v + noff);
        }
    }
    const uint32_t tmem = sh.tmem;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int half = warp >> 2, row = 32 * (warp & 3) + lane;
    const float f = iq * ik * LOG2E;               // raw accumulator -> scores in the log2 domain
    const float osc = (TRAIN ? dc.scale : 1.0f) * iv;
    const uint32_t thr24 = dc.thr16 << 24;
    const uint32_t ql = dlo_k(smem_u32(Qp)), kl = dlo_k(smem_u32(KV));
    constexpr uint32_t idv = idesc_f16(32, false, true);
    uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        const int Nk = min(NKP, 128 * (t + 1));
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t id = idesc_f16(Nk, false, false);
                const uint32_t qa = ql + (uint32_t)(128 * t) * 8;
                mma_lo(tmem, qa, kl, id, 0u);            // q0 k0
                mma_lo(tmem, qa + 2, kl, id, 1u);        // q1 k0
                mma_lo(tmem, qa, kl + 2, id, 1u);        // q0 k1
                mma_commit(&sh.bar);
            }
            __syncwarp();
        }
        mbar_wait(&sh.bar, phase);
        phase ^= 1;
        fence_after();
        const int Rw = 128 * t + 32 * (warp & 3);           // first query row of this warp
        const int i = Rw + lane;
        const bool wvalid = Rw < L;
        const int nch = (Nk + 31) >> 5, cdiag = Rw >> 5;
        // ---- row maximum (raw accumulator units)
        float mraw = -INFINITY;
        if (wvalid) {
#pragma unroll 1
            for (int c = half; c <= cdiag; c += 2) {
                float a[32];
                tmem_ld32(tl + 32 * c, a);
                if (c == cdiag) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (e > lane) a[e] = -INFINITY;
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) mraw = fmaxf(mraw, a[e]);
            }
        }
        sh.xm[half][row] = mraw;
        __syncthreads();
        const float m = fmaxf(sh.xm[0][row], sh.xm[1][row]) * f;
        // ---- P = 2^(S f - m), row sum, dropout, FP16 pairs in place
        float l = 0.f;
        if (wvalid) {
            const uint32_t rb4 = (((uint32_t)bh + dc.bh_off) * (uint32_t)L + (uint32_t)min(i, L - 1)) * (uint32_t)Lp4;
#pragma unroll 1
            for (int c = half; c < nch; c += 2) {
                if (c < cdiag) {
                    l += fwd_chunk<TRAIN, false>(tl + 32 * c, f, m, lane, dc.seed, site, rb4 + 8 * c, thr24);
                } else if (c == cdiag) {
                    l += fwd_chunk<TRAIN, true>(tl + 32 * c, f, m, lane, dc.seed, site, rb4 + 8 * c, thr24);
                } else {
                    uint32_t z[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) z[e] = 0u;
                    tmem_st16(tl + 32 * c, z);
                    tmem_st16(tl + 32 * c + 16, z);
                }
            }
        }
        sh.xl[half][row] = l;
        tmem_st_wait();
        fence_before();
        __syncthreads();
        if (warp == 0) {
            fence_after();
            if (elect_one()) {
                // O[128 x 32] = P [v0 | v1]: A = P pieces in tensor memory, B = MN-major rows of the key / value tile
                for (int kk = 0; kk < (Nk >> 4); ++kk) {
                    const uint32_t a0 = tmem + 32 * (kk >> 1) + 8 * (kk & 1), bl = kl + (uint32_t)kk * 128 + 4;
                    mma_lo_ts(tmem + CF_O, a0, bl, idv, kk ? 1u : 0u);
                    mma_lo_ts(tmem + CF_O, a0 + 16, bl, idv, 1u);
                }
                mma_commit(&sh.bar);
            }
            __syncwarp();
        }
        mbar_wait(&sh.bar, phase);
        phase ^= 1;
        fence_after();
        if (half == 0 && wvalid) {               // warp-uniform: tcgen05.ld is a warp-collective instruction
            float a[32];
            tmem_ld32(tl + CF_O, a);
            if (i < L) {
                const float lt = sh.xl[0][row] + sh.xl[1][row];
                const float sc = osc / lt;
                float4* dst = reinterpret_cast<float4*>(o + base + (size_t)i * D);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[u] = make_float4((a[4 * u] + a[16 + 4 * u]) * sc, (a[4 * u + 1] + a[17 + 4 * u]) * sc,
                                         (a[4 * u + 2] + a[18 + 4 * u]) * sc, (a[4 * u + 3] + a[19 + 4 * u]) * sc);
                lse[(size_t)bh * L + i] = m * LN2 + logf(lt);
            }
        }
        fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}


// ================================================================================================
// backward, two-pass variant: attn_tc::k_attn_bwd_tc (one CTA per (sample, head), 256 threads, two CTAs per SM, S / dP and
// the in-place Pd / dS pieces in tensor memory, pass A by query rows for dQ, pass B by key rows for dV / dK) with the
// operand handling of this file: no transposed copies (k, dO, q enter the second GEMMs as MN-major B operands from the row
// tiles), [b0 | b1] as one N = 32 operand (two MMAs per k-step instead of three), elect.sync issue with low-word descriptors.
// Score blocks are 96 columns wide so that S, dP (2 x 96) and two 32-column accumulators fit 256 TMEM columns.
// It computes every score twice but keeps two independent CTAs per SM, which the single-pass kernel cannot.
// ================================================================================================
constexpr int T2_BW = 96;
constexpr int T2_ROWS = 224;                       // positions per row tile (L <= 224)
constexpr size_t T2_SMEM = ((size_t)T2_ROWS + 256) * 128 + 1024;   // the k | v tile is read 128 rows at a time from row 128
constexpr uint32_t T2_X = 0, T2_Y = T2_BW, T2_A1 = 192, T2_A2 = 224;

struct ShT2 {
    // Two barriers, not one: after the wait for a block's second GEMMs there is no CTA barrier before warp 0 commits the next
    // block's S / dP.  With a single mbarrier a warp that wakes up late from that wait could find the barrier TWO phases ahead,
    // read its parity as "not complete" and wait forever.  With one barrier per GEMM kind, two commits on the same barrier are
    // always separated by a __syncthreads that every thread reaches only after its previous wait on that barrier.
    uint64_t bar_s, bar_g;
    uint32_t tmem;
    float red[4][8];
    __align__(16) float ls[256];          // lse * log2(e) per query (+inf for rows >= L)
    __align__(16) float dl[256];          // delta_i * sds
};

// pass A chunk (16 keys of this thread's query row): dS scaled by sds into sx
template <bool TRAIN, bool DIAG>
__device__ __forceinline__ void t2_chunk_a(uint32_t ts, uint32_t tp, float f, float li, float Di, float fdp, float dsc, int nvalid,
                                           uint32_t seed, uint32_t site, uint32_t g0, uint32_t thr24, float (&sx)[16]) {
    float dp[16];
    tmem_ld16(ts, sx);
    tmem_ld16(tp, dp);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t r = 0u;
        if (TRAIN) r = rng4(seed, site, (uint64_t)(g0 + g));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * g + e;
            float p = ex2(fmaf(sx[c], f, -li));
            if (DIAG) p = c < nvalid ? p : 0.f;
            float k2 = dsc;
            if (TRAIN) k2 = (e == 3 ? r : (r << (24 - 8 * e))) >= thr24 ? dsc : 0.f;
            sx[c] = p * fmaf(dp[c] * k2, fdp, -Di);
        }
    }
}
// pass B chunk (16 queries i0.. of this thread's key row j): Pd into sx, dS (scaled) into dp.  jrel = j - i0: query e is
// masked when e < jrel.  The four keys of a dropout hash group sit in four neighbouring lanes: lane x of the quad hashes the
// queries e = x, x+4, x+8, x+12 of the chunk and the quad exchanges them.
template <bool TRAIN, bool DIAG>
__device__ __forceinline__ void t2_chunk_b(uint32_t ts, uint32_t tp, float f, float fdp, float dsc, const float* __restrict__ lsp,
                                           const float* __restrict__ dlp, int jrel, int i0, int L, int Lp4, int lane, int jb,
                                           uint32_t seed, uint32_t site, uint32_t hbase, uint32_t thr24, float (&sx)[16], float (&dp)[16]) {
    tmem_ld16(ts, sx);
    tmem_ld16(tp, dp);
    uint32_t hv[4] = {0u, 0u, 0u, 0u};
    if (TRAIN) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int qi = min(i0 + (lane & 3) + 4 * u, L - 1);
            hv[u] = rng4(seed, site, (uint64_t)(hbase + (uint32_t)qi * (uint32_t)Lp4));
        }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 l4 = *reinterpret_cast<const float4*>(lsp + 4 * g), d4 = *reinterpret_cast<const float4*>(dlp + 4 * g);
        const float la[4] = {l4.x, l4.y, l4.z, l4.w}, da[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int e = 4 * g + x;
            float p = ex2(fmaf(sx[e], f, -la[x]));
            if (DIAG) p = e < jrel ? 0.f : p;
            float k2 = dsc;
            if (TRAIN) {
                const uint32_t r = __shfl_sync(0xffffffffu, hv[g], (lane & ~3) | x);
                k2 = (r << jb) >= thr24 ? dsc : 0.f;
            }
            sx[e] = p * k2;
            dp[e] = p * fmaf(dp[e] * k2, fdp, -da[x]);
        }
    }
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 2)
k_attn_bwd_t2(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
              const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
              float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, DropCfg dc, uint32_t site) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ ShT2 sh;
    uint8_t* QG = align1k(smem_raw);               // [position][q0 | q1 | g0 | g1]
    uint8_t* KV = QG + T2_ROWS * 128;              // [position][k0 | k1 | v0 | v1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    const int ntile = (L + 127) >> 7, NKP = (L + 15) & ~15, nblk = (NKP + T2_BW - 1) / T2_BW, Lp4 = ((L + 3) & ~3) >> 2;
    if (warp == 0) tmem_alloc(&sh.tmem, 256);
    if (tid == 0) { mbar_init(&sh.bar_s, 1); mbar_init(&sh.bar_g, 1); fence_barrier_init(); }
    const int c4 = tid & 3, r0 = tid >> 2;
    float4 vq[4], vk[4], vv[4], vg[4];
    float dsum[4];
    float mx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        vq[i] = vk[i] = vv[i] = vg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float ds = 0.f;
        if (R < L) {
            const size_t off = base + (size_t)R * D + 4 * c4;
            vq[i] = __ldg(reinterpret_cast<const float4*>(q + off));
            vk[i] = __ldg(reinterpret_cast<const float4*>(k + off));
            vv[i] = __ldg(reinterpret_cast<const float4*>(v + off));
            vg[i] = __ldg(reinterpret_cast<const float4*>(dO + off));
            const float4 oo = __ldg(reinterpret_cast<const float4*>(o + off));
            ds = vg[i].x * oo.x + vg[i].y * oo.y + vg[i].z * oo.z + vg[i].w * oo.w;
        }
        ds += __shfl_xor_sync(0xffffffffu, ds, 1);
        ds += __shfl_xor_sync(0xffffffffu, ds, 2);
        dsum[i] = ds;
        if (c4 == 0) sh.ls[R] = R < L ? lse[(size_t)bh * L + R] * LOG2E : INFINITY;
        mx[0] = amax4(vq[i], mx[0]); mx[1] = amax4(vk[i], mx[1]); mx[2] = amax4(vv[i], mx[2]); mx[3] = amax4(vg[i], mx[3]);
    }
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
        const float w = warp_max(mx[kx]);
        if (lane == 0) sh.red[kx][warp] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
        float r = sh.red[kx][lane & 7];
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, off));
        mx[kx] = r;
    }
    float sq, iq, sk, ik, sv, iv, sg, ig, sds, ids;
    pow2_scale(mx[0], sq, iq); pow2_scale(mx[1], sk, ik); pow2_scale(mx[2], sv, iv); pow2_scale(mx[3], sg, ig);
    const float dsc = TRAIN ? dc.scale : 1.0f;
    pow2_scale(64.0f * mx[3] * mx[2] * dsc, sds, ids);        // |dS| <= |dPd| + |delta| <= 2 * 16 gmax vmax scale
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        if (c4 == 0) sh.dl[R] = dsum[i] * sds;
        if (R < NKP) {                              // rows the MMAs read (rows [L, NKP) are stored as zeros)
            const uint32_t ob = (uint32_t)((R >> 3) * 1024 + (R & 7) * 128) + (((((uint32_t)c4 >> 1) ^ (uint32_t)R) & 7u) << 4) + (c4 & 1) * 8;
            uint2 a0, a1, b0, b1;
            split4(vq[i], sq, a0, a1);
            split4(vg[i], sg, b0, b1);
            *reinterpret_cast<uint2*>(QG + ob) = a0;
            *reinterpret_cast<uint2*>(QG + (ob ^ 0x20u)) = a1;
            *reinterpret_cast<uint2*>(QG + (ob ^ 0x40u)) = b0;
            *reinterpret_cast<uint2*>(QG + (ob ^ 0x60u)) = b1;
            split4(vk[i], sk, a0, a1);
            split4(vv[i], sv, b0, b1);
            *reinterpret_cast<uint2*>(KV + ob) = a0;
            *reinterpret_cast<uint2*>(KV + (ob ^ 0x20u)) = a1;
            *reinterpret_cast<uint2*>(KV + (ob ^ 0x40u)) = b0;
            *reinterpret_cast<uint2*>(KV + (ob ^ 0x60u)) = b1;
        }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    {   // head slices of the CTA that will follow on this SM slot (CTAs are dispatched in index order, two per SM) into L2
        const int nb = bh + 2 * 148;
        if (nb < (int)gridDim.x && tid < L) {
            const size_t noff = (size_t)(nb / H) * L * D + (nb % H) * DH + (size_t)tid * D;
            prefetch_l2(q + noff); prefetch_l2(k + noff); prefetch_l2(v + noff); prefetch_l2(dO + noff); prefetch_l2(o + noff);
        }
    }
    const uint32_t tmem = sh.tmem;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int half = warp >> 2;
    const float f = iq * ik * LOG2E;               // raw S -> log2 domain
    const float fdp = ig * iv * sds;               // raw dP (times the keep scale) -> dPd * sds
    const uint32_t thr24 = dc.thr16 << 24;
    const uint32_t qgl = dlo_k(smem_u32(QG)), kvl = dlo_k(smem_u32(KV));
    constexpr uint32_t id32 = idesc_f16(32, false, true);
    uint32_t ph_s = 0, ph_g = 0;

    // ---------------- pass A: rows = queries; dQ[128 x 32] += dS [k0 | k1]
    // (Issuing the consuming MMAs of a block together with the S / dP MMAs of the next block under one commit -- one MMA round
    // trip per block instead of two -- was measured: 809 us against 751 us.  The two CTAs of an SM interleave better when a
    // block's two waits are short than when its one wait is long.)
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        const int Rw = 128 * t + 32 * (warp & 3), i = Rw + lane;
        const bool wvalid = Rw < L;
        const int last_q = min(L, 128 * (t + 1)) - 1;
        const float li = sh.ls[i], Di = sh.dl[i];
        const uint32_t rb4 = (((uint32_t)bh + dc.bh_off) * (uint32_t)L + (uint32_t)min(i, L - 1)) * (uint32_t)Lp4;
        bool first = true;
#pragma unroll 1
        for (int kb = 0; kb < nblk; ++kb) {
            const int key0 = T2_BW * kb;
            if (key0 > last_q) break;
            const int Nb = min(min(T2_BW, NKP - key0), ((last_q - key0) & ~15) + 16);      // no key beyond the tile's last query
            if (warp == 0) {
                if (elect_one()) {
                    const uint32_t id = idesc_f16(Nb, false, false);
                    const uint32_t a = qgl + (uint32_t)(128 * t) * 8, bb = kvl + (uint32_t)key0 * 8;
                    mma_lo(tmem + T2_X, a, bb, id, 0u);              // q0 k0
                    mma_lo(tmem + T2_X, a + 2, bb, id, 1u);          // q1 k0
                    mma_lo(tmem + T2_X, a, bb + 2, id, 1u);          // q0 k1
                    mma_lo(tmem + T2_Y, a + 4, bb + 4, id, 0u);      // g0 v0
                    mma_lo(tmem + T2_Y, a + 6, bb + 4, id, 1u);      // g1 v0
                    mma_lo(tmem + T2_Y, a + 4, bb + 6, id, 1u);      // g0 v1
                    mma_commit(&sh.bar_s);
                }
                __syncwarp();
            }
            mbar_wait(&sh.bar_s, ph_s);
            ph_s ^= 1;
            fence_after();
            if (wvalid) {
#pragma unroll 1
                for (int ch = half; ch < Nb / 16; ch += 2) {
                    const int j0 = key0 + 16 * ch;
                    if (j0 > Rw + 31) { zero16(tl + T2_X + 16 * ch); continue; }
                    float sx[16];
                    if (j0 + 15 > Rw)
                        t2_chunk_a<TRAIN, true>(tl + T2_X + 16 * ch, tl + T2_Y + 16 * ch, f, li, Di, fdp, dsc, i - j0 + 1, dc.seed, site,
                                                rb4 + (uint32_t)(j0 >> 2), thr24, sx);
                    else
                        t2_chunk_a<TRAIN, false>(tl + T2_X + 16 * ch, tl + T2_Y + 16 * ch, f, li, Di, fdp, dsc, 16, dc.seed, site,
                                                 rb4 + (uint32_t)(j0 >> 2), thr24, sx);
                    put16(tl + T2_X + 16 * ch, sx);
                }
            }
            tmem_st_wait();
            fence_before();
            __syncthreads();
            if (warp == 0) {
                fence_after();
                if (elect_one()) {
                    for (int kk = 0; kk < Nb / 16; ++kk) {
                        const uint32_t a0 = tmem + T2_X + 16 * kk, bl = kvl + (uint32_t)(key0 + 16 * kk) * 8;
                        mma_lo_ts(tmem + T2_A1, a0, bl, id32, (first && kk == 0) ? 0u : 1u);
                        mma_lo_ts(tmem + T2_A1, a0 + 8, bl, id32, 1u);
                    }
                    mma_commit(&sh.bar_g);
                }
                __syncwarp();
            }
            first = false;
            mbar_wait(&sh.bar_g, ph_g);
            ph_g ^= 1;
            fence_after();
        }
        if (half == 0 && wvalid) {
            float a[32];
            tmem_ld32(tl + T2_A1, a);
            if (i < L) {
                const float sc = 0.25f * ids * ik;
                float4* dst = reinterpret_cast<float4*>(dq + base + (size_t)i * D);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[u] = make_float4((a[4 * u] + a[16 + 4 * u]) * sc, (a[4 * u + 1] + a[17 + 4 * u]) * sc,
                                         (a[4 * u + 2] + a[18 + 4 * u]) * sc, (a[4 * u + 3] + a[19 + 4 * u]) * sc);
            }
        }
        fence_before();
        __syncthreads();
    }

    // ---------------- pass B: rows = keys; dV[128 x 32] += Pd^T [g0 | g1], dK[128 x 32] += dS^T [q0 | q1]
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        const int Jw = 128 * t + 32 * (warp & 3), j = Jw + lane;
        const bool wvalid = Jw < L;
        const uint32_t jg = (uint32_t)(j >> 2);
        const int jb = 24 - 8 * (j & 3);
        const uint32_t hbase = (((uint32_t)bh + dc.bh_off) * (uint32_t)L) * (uint32_t)Lp4 + jg;
        bool first = true;
#pragma unroll 1
        for (int qb = 0; qb < nblk; ++qb) {
            int q0 = T2_BW * qb;
            int Nb = min(T2_BW, NKP - q0);
            if (q0 + Nb <= 128 * t) continue;              // every query of the block precedes every key of the tile
            if (q0 < 128 * t) { Nb -= 128 * t - q0; q0 = 128 * t; }      // no query before the tile's first key
            if (warp == 0) {
                if (elect_one()) {
                    const uint32_t id = idesc_f16(Nb, false, false);
                    const uint32_t a = kvl + (uint32_t)(128 * t) * 8, bb = qgl + (uint32_t)q0 * 8;
                    mma_lo(tmem + T2_X, a, bb, id, 0u);              // k0 q0
                    mma_lo(tmem + T2_X, a + 2, bb, id, 1u);          // k1 q0
                    mma_lo(tmem + T2_X, a, bb + 2, id, 1u);          // k0 q1
                    mma_lo(tmem + T2_Y, a + 4, bb + 4, id, 0u);      // v0 g0
                    mma_lo(tmem + T2_Y, a + 6, bb + 4, id, 1u);      // v1 g0
                    mma_lo(tmem + T2_Y, a + 4, bb + 6, id, 1u);      // v0 g1
                    mma_commit(&sh.bar_s);
                }
                __syncwarp();
            }
            mbar_wait(&sh.bar_s, ph_s);
            ph_s ^= 1;
            fence_after();
            if (wvalid) {
#pragma unroll 1
                for (int ch = half; ch < Nb / 16; ch += 2) {
                    const int i0 = q0 + 16 * ch;
                    if (i0 + 15 < Jw) { zero16(tl + T2_X + 16 * ch); zero16(tl + T2_Y + 16 * ch); continue; }
                    float sx[16], dp[16];
                    if (Jw + 31 > i0)
                        t2_chunk_b<TRAIN, true>(tl + T2_X + 16 * ch, tl + T2_Y + 16 * ch, f, fdp, dsc, sh.ls + i0, sh.dl + i0, j - i0, i0, L, Lp4,
                                                lane, jb, dc.seed, site, hbase, thr24, sx, dp);
                    else
                        t2_chunk_b<TRAIN, false>(tl + T2_X + 16 * ch, tl + T2_Y + 16 * ch, f, fdp, dsc, sh.ls + i0, sh.dl + i0, 0, i0, L, Lp4,
                                                 lane, jb, dc.seed, site, hbase, thr24, sx, dp);
                    put16(tl + T2_X + 16 * ch, sx);
                    put16(tl + T2_Y + 16 * ch, dp);
                }
            }
            tmem_st_wait();
            fence_before();
            __syncthreads();
            if (warp == 0) {
                fence_after();
                if (elect_one()) {
                    for (int kk = 0; kk < Nb / 16; ++kk) {
                        const uint32_t ax = tmem + T2_X + 16 * kk, ay = tmem + T2_Y + 16 * kk;
                        const uint32_t bl = qgl + (uint32_t)(q0 + 16 * kk) * 8;
                        const uint32_t acc = (first && kk == 0) ? 0u : 1u;
                        mma_lo_ts(tmem + T2_A1, ax, bl + 4, id32, acc);         // Pd^T [g0 | g1]
                        mma_lo_ts(tmem + T2_A1, ax + 8, bl + 4, id32, 1u);
                        mma_lo_ts(tmem + T2_A2, ay, bl, id32, acc);             // dS^T [q0 | q1]
                        mma_lo_ts(tmem + T2_A2, ay + 8, bl, id32, 1u);
                    }
                    mma_commit(&sh.bar_g);
                }
                __syncwarp();
            }
            first = false;
            mbar_wait(&sh.bar_g, ph_g);
            ph_g ^= 1;
            fence_after();
        }
        if (wvalid) {                                    // half 0 stores dv, half 1 stores dk
            float a[32];
            tmem_ld32(tl + (half == 0 ? T2_A1 : T2_A2), a);
            if (j < L) {
                const float sc = half == 0 ? ig : ids * iq;
                float4* dst = reinterpret_cast<float4*>((half == 0 ? dv : dk) + base + (size_t)j * D);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[u] = make_float4((a[4 * u] + a[16 + 4 * u]) * sc, (a[4 * u + 1] + a[17 + 4 * u]) * sc,
                                         (a[4 * u + 2] + a[18 + 4 * u]) * sc, (a[4 * u + 3] + a[19 + 4 * u]) * sc);
            }
        }
        fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace attn_p
}  // namespace amid
