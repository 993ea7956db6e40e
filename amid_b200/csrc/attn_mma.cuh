// Causal multi-head attention for short sequences (L <= 512, head_dim 16) on the warp-level
// tensor-core path (mma.sync m16n8k8, TF32 operands, fp32 accumulate).  One CTA per (sample, head);
// k/v (and q/dO in the backward) of the head live in shared memory with a 20-float row stride, which
// makes every fragment load bank-conflict free; rows are padded with zeros to a multiple of 16 so no
// index clamps are needed.  Softmax statistics, dropout and the causal mask are applied on the
// accumulator fragments (only diagonal blocks evaluate the mask); P (or dS) goes straight back into
// the next MMA as the A operand by relabelling the accumulator columns (2t -> k index t, 2t+1 -> k
// index t+4) and reading the B operand rows in the same permuted order.  Row tiles are handed to warps
// through a per-CTA queue, heaviest first.
//
// Arithmetic follows torch/nn/functional.py:6630-6647: S = (0.25 q) k^T with -inf above the
// diagonal, P = softmax(S), dropout(P) (no renormalisation), O = P v.
#pragma once
#include "common.cuh"

namespace amid {
namespace attn {

constexpr int LDS = 20;          // smem row stride (floats)
constexpr int NW = 4;            // warps per CTA (forward)
constexpr int NWB = 8;           // warps per CTA (backward: 2 x ntile work items)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fbits(float x) { return __float_as_uint(x); }
// 2^x for x <= 0 (softmax numerators): one MUFU, inputs below -126 flush to 0
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 3xTF32 split (precision "x3"): x = hi + lo with both halves rounded to TF32; products hi*hi + lo*hi + hi*lo
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
// A fragments of a 16 x 16 block (2 k-steps); lo is used by the 3xTF32 variants only
struct AFrag {
    uint32_t hi[2][4], lo[2][4];
};
template <bool X3>
__device__ __forceinline__ void finish_a(AFrag& a) {
    if (X3) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(__uint_as_float(a.hi[ks][i]), a.hi[ks][i], a.lo[ks][i]);
    }
}
// dynamic work queue of a CTA: warps take items in order (callers order items by decreasing cost)
__device__ __forceinline__ int next_item(int* counter, int lane) {
    int it = 0;
    if (lane == 0) it = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
}

// head slice [L,16] of a [M,128] tensor -> smem [Lpad][LDS], rows >= L zero, optionally scaled
__device__ __forceinline__ void stage(float* s, const float* __restrict__ g, int L, int Lpad, float scale = 1.0f) {
    for (int idx = threadIdx.x; idx < Lpad * 4; idx += blockDim.x) {
        const int r = idx >> 2, c4 = idx & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < L) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * D) + c4);
        *reinterpret_cast<float4*>(s + r * LDS + c4 * 4) = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    }
}
// A fragments (16 rows x 16 cols = 2 k-steps) of rows r0.. from a staged (padded) matrix
__device__ __forceinline__ void load_a16(uint32_t (&a)[2][4], const float* s, int r0, int g, int t) {
    const float* pa = s + (r0 + g) * LDS;
    const float* pb = pa + 8 * LDS;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = fbits(pa[8 * ks + t]);
        a[ks][1] = fbits(pb[8 * ks + t]);
        a[ks][2] = fbits(pa[8 * ks + t + 4]);
        a[ks][3] = fbits(pb[8 * ks + t + 4]);
    }
}
// the same from global memory (row stride D, rows clamped to L-1), scaled
__device__ __forceinline__ void load_a16_g(uint32_t (&a)[2][4], const float* __restrict__ p, int r0, int L, int g, int t, float scale) {
    const float* pa = p + (size_t)min(r0 + g, L - 1) * D;
    const float* pb = p + (size_t)min(r0 + g + 8, L - 1) * D;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = fbits(__ldg(pa + 8 * ks + t) * scale);
        a[ks][1] = fbits(__ldg(pb + 8 * ks + t) * scale);
        a[ks][2] = fbits(__ldg(pa + 8 * ks + t + 4) * scale);
        a[ks][3] = fbits(__ldg(pb + 8 * ks + t + 4) * scale);
    }
}
// D[16 x 8] = A[16 x 16] * X[n0..n0+8][16]^T   (X rows are the n index; k = feature)
template <bool X3>
__device__ __forceinline__ void mma_xt(float (&d)[4], const AFrag& a, const float* x, int n0, int g, int t) {
    const float* p = x + (n0 + g) * LDS + t;
    if (X3) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(p[8 * ks], bh0, bl0);
            split_tf32(p[8 * ks + 4], bh1, bl1);
            mma_tf32(d, a.lo[ks], bh0, bh1);
            mma_tf32(d, a.hi[ks], bl0, bl1);
            mma_tf32(d, a.hi[ks], bh0, bh1);
        }
    } else {
        mma_tf32(d, a.hi[0], fbits(p[0]), fbits(p[4]));
        mma_tf32(d, a.hi[1], fbits(p[8]), fbits(p[12]));
    }
}
// acc[dt][..] += P[16 x 8 (relabelled)] * X[n0..n0+8][16]   (X rows are the k index, permuted 2t / 2t+1).
// The output features are permuted too -- n index g of n-tile dt is feature 2g+dt -- so the B operands of both
// n-tiles come from one 64-bit load per key row and a thread ends up owning the four consecutive features
// 4t..4t+3 of its rows: {acc[0][0], acc[1][0], acc[0][1], acc[1][1]} (row g) and the [..][2], [..][3] set (row g+8).
template <bool X3>
__device__ __forceinline__ void mma_px(float (&acc)[2][4], const float (&p)[4], const float* x, int n0, int g, int t) {
    const float* pa = x + (n0 + 2 * t) * LDS + 2 * g;
    const float2 u = *reinterpret_cast<const float2*>(pa), w = *reinterpret_cast<const float2*>(pa + LDS);
    if (X3) {
        uint32_t ah[4], al[4];
        split_tf32(p[0], ah[0], al[0]); split_tf32(p[2], ah[1], al[1]);
        split_tf32(p[1], ah[2], al[2]); split_tf32(p[3], ah[3], al[3]);
        uint32_t uh, ul, wh, wl;
        split_tf32(u.x, uh, ul); split_tf32(w.x, wh, wl);
        mma_tf32(acc[0], al, uh, wh);
        mma_tf32(acc[0], ah, ul, wl);
        mma_tf32(acc[0], ah, uh, wh);
        split_tf32(u.y, uh, ul); split_tf32(w.y, wh, wl);
        mma_tf32(acc[1], al, uh, wh);
        mma_tf32(acc[1], ah, ul, wl);
        mma_tf32(acc[1], ah, uh, wh);
    } else {
        const uint32_t a[4] = {fbits(p[0]), fbits(p[2]), fbits(p[1]), fbits(p[3])};
        mma_tf32(acc[0], a, fbits(u.x), fbits(w.x));
        mma_tf32(acc[1], a, fbits(u.y), fbits(w.y));
    }
}
// features 4t..4t+3 of row g (hi = 0) or g+8 (hi = 1) out of a mma_px accumulator, scaled
__device__ __forceinline__ float4 px_row(const float (&acc)[2][4], int hi, float sc) {
    return make_float4(acc[0][2 * hi] * sc, acc[1][2 * hi] * sc, acc[0][2 * hi + 1] * sc, acc[1][2 * hi + 1] * sc);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Dropout keep bits of an accumulator fragment whose ROWS are query rows (a = row g, b = row g+8) and
// whose columns are the keys n0+2t, n0+2t+1.  The two threads of a pair (t, t^1) share one 4-key hash
// group, so each computes one of the two rows and they exchange.  rb4_* = (bh*L + row) * Lp / 4.
// The index space of an attention site is B*8*L*ceil(L/4) 4-key groups; the launchers require it to fit 32 bits,
// so the hash input is a 32-bit index (rng4 sees a zero high word: identical bits to the 64-bit form the mask
// export and the fp32 kernels use).
struct RowKeep {
    uint32_t rb4_a, rb4_b;
};
// byte-lane test without the shift: (r & (0xFF << 8*lane)) >= (thr << 8*lane)
struct ByteLane {
    uint32_t mask, thr;
    __device__ __forceinline__ ByteLane(int lane4, uint32_t thr8) : mask(0xFFu << (8 * lane4)), thr(thr8 << (8 * lane4)) {}
    __device__ __forceinline__ bool keep(uint32_t r) const { return (r & mask) >= thr; }
};
__device__ __forceinline__ void keep_rows(const DropCfg& dc, uint32_t site, const RowKeep& rk, const ByteLane& b0,
                                          const ByteLane& b1, int n0, int t, bool (&kp)[4]) {
    const uint32_t idx4 = ((t & 1) ? rk.rb4_b : rk.rb4_a) + (uint32_t)((n0 >> 2) + (t >> 1));
    const uint32_t mine = rng4(dc.seed, site, (uint64_t)idx4);
    const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
    const uint32_t ra = (t & 1) ? other : mine, rb = (t & 1) ? mine : other;
    kp[0] = b0.keep(ra); kp[1] = b1.keep(ra);
    kp[2] = b0.keep(rb); kp[3] = b1.keep(rb);
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
template <bool X3>
__global__ void __launch_bounds__(NW * 32)
k_attn_fwd_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               float* __restrict__ o, float* __restrict__ lse, int L, DropCfg dc, uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int queue;
    const int ntile = (L + 15) / 16, Lpad = ntile * 16, Lp = (L + 3) & ~3;
    float* ks = smem;
    float* vs = ks + Lpad * LDS;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    if (threadIdx.x == 0) queue = 0;
    stage(ks, k + base, L, Lpad);
    stage(vs, v + base, L, Lpad);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const ByteLane bl0(2 * (t & 1), dc.thr16), bl1(2 * (t & 1) + 1, dc.thr16);   // this thread's two key columns
    const uint64_t bhL = (uint64_t)bh * L;
    for (;;) {
        const int item = next_item(&queue, lane);
        if (item >= ntile) break;
        const int rt = ntile - 1 - item;          // heaviest (last) row tile first
        const int r0 = rt * 16;
        AFrag aq;
        load_a16_g(aq.hi, q + base, r0, L, g, t, LOG2E);       // scores in the log2 domain
        finish_a<X3>(aq);
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int row_a = r0 + g, row_b = r0 + g + 8;
        RowKeep rk;
        rk.rb4_a = ((uint32_t)bhL + dc.bh_off * (uint32_t)L + (uint32_t)min(row_a, L - 1)) * (uint32_t)(Lp >> 2);
        rk.rb4_b = ((uint32_t)bhL + dc.bh_off * (uint32_t)L + (uint32_t)min(row_b, L - 1)) * (uint32_t)(Lp >> 2);
        const int kend = min(r0 + 16, L);         // keys [0, kend) can be visible to this tile
        for (int kb = 0; kb < kend; kb += 32) {
            float s[4][4];
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n0 = kb + 8 * nt;
                if (n0 < kend) {                                  // warp-uniform
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                    mma_xt<X3>(s[nt], aq, ks, n0, g, t);
                    if (n0 + 8 > r0 + 1 || n0 + 8 > L) {          // diagonal / ragged block: apply the mask
                        const int c = n0 + 2 * t;
                        if (!(c <= row_a && c < L)) s[nt][0] = -INFINITY;
                        if (!(c + 1 <= row_a && c + 1 < L)) s[nt][1] = -INFINITY;
                        if (!(c <= row_b && c < L)) s[nt][2] = -INFINITY;
                        if (!(c + 1 <= row_b && c + 1 < L)) s[nt][3] = -INFINITY;
                    }
                } else {
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = -INFINITY;
                }
                mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
            }
            const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
            const float sub0 = mn0 == -INFINITY ? 0.f : mn0, sub1 = mn1 == -INFINITY ? 0.f : mn1;
            const float c0 = ex2(m0 - sub0), c1 = ex2(m1 - sub1);
            l0 *= c0; l1 *= c1;
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) { acc[dt][0] *= c0; acc[dt][1] *= c0; acc[dt][2] *= c1; acc[dt][3] *= c1; }
            m0 = mn0; m1 = mn1;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n0 = kb + 8 * nt;
                if (n0 >= kend) continue;                         // warp-uniform
                float p[4];
                p[0] = ex2(s[nt][0] - sub0); p[1] = ex2(s[nt][1] - sub0);
                p[2] = ex2(s[nt][2] - sub1); p[3] = ex2(s[nt][3] - sub1);
                l0 += p[0] + p[1]; l1 += p[2] + p[3];
                if (dc.train) {
                    bool kp[4];
                    keep_rows(dc, site, rk, bl0, bl1, n0, t, kp);
#pragma unroll
                    for (int e = 0; e < 4; ++e) p[e] = kp[e] ? p[e] : 0.f;      // 1/(1-p) folded into the final scale
                }
                mma_px<X3>(acc, p, vs, n0, g, t);
            }
        }
        l0 = quad_sum(l0); l1 = quad_sum(l1);
        const float osc = dc.train ? dc.scale : 1.0f;
        const float i0 = osc / l0, i1 = osc / l1;
        if (row_a < L) {
            *reinterpret_cast<float4*>(o + base + (size_t)row_a * D + 4 * t) = px_row(acc, 0, i0);
            if (t == 0) lse[bhL + row_a] = m0 * LN2 + logf(l0);
        }
        if (row_b < L) {
            *reinterpret_cast<float4*>(o + base + (size_t)row_b * D + 4 * t) = px_row(acc, 1, i1);
            if (t == 0) lse[bhL + row_b] = m1 * LN2 + logf(l1);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// backward: pass A (query tiles -> dq), pass B (key tiles -> dk, dv); P recomputed from lse.
// q is staged pre-multiplied by log2(e) (scores in the log2 domain); dk is rescaled at the store.
// ----------------------------------------------------------------------------------------------
template <bool X3>
__global__ void __launch_bounds__(NWB * 32)
k_attn_bwd_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
               float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, DropCfg dc, uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int queue;
    if (threadIdx.x == 0) queue = 0;
    const int ntile = (L + 15) / 16, Lpad = ntile * 16, Lp = (L + 3) & ~3;
    float* qs = smem;                 // q * log2(e)
    float* ks = qs + Lpad * LDS;
    float* vs = ks + Lpad * LDS;
    float* gs = vs + Lpad * LDS;      // dO
    float* Dv = gs + Lpad * LDS;      // D_i = <dO_i, O_i>
    float* ls = Dv + Lpad;            // lse * log2(e)
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    stage(qs, q + base, L, Lpad, LOG2E);
    stage(ks, k + base, L, Lpad);
    const float sc = dc.train ? dc.scale : 1.0f;      // dropout scale rides on the staged v (dP) and the dv store
    stage(vs, v + base, L, Lpad, sc);
    stage(gs, dO + base, L, Lpad);
    for (int i = threadIdx.x; i < Lpad; i += blockDim.x) {
        float s = 0.f, le = 0.f;
        if (i < L) {
            const float4* po = reinterpret_cast<const float4*>(o + base + (size_t)i * D);
            const float4* pg = reinterpret_cast<const float4*>(dO + base + (size_t)i * D);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 a = __ldg(po + c), bb = __ldg(pg + c);
                s += a.x * bb.x + a.y * bb.y + a.z * bb.z + a.w * bb.w;
            }
            le = lse[(size_t)bh * L + i] * LOG2E;
        }
        Dv[i] = s;
        ls[i] = le;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const ByteLane bl0(2 * (t & 1), dc.thr16), bl1(2 * (t & 1) + 1, dc.thr16);   // this thread's two key columns
    const uint64_t bhL = (uint64_t)bh * L;
    // Work items, heaviest first: item 2n = pass A on query tile ntile-1-n, item 2n+1 = pass B on key tile n.
    for (;;) {
        const int item = next_item(&queue, lane);
        if (item >= 2 * ntile) break;
        if ((item & 1) == 0) {
            // ---------------- pass A: dq[i] = 0.25 * sum_j dS_ij k_j
            const int r0 = (ntile - 1 - (item >> 1)) * 16;
            AFrag aq, ag;
            load_a16(aq.hi, qs, r0, g, t);
            load_a16(ag.hi, gs, r0, g, t);
            finish_a<X3>(aq);
            finish_a<X3>(ag);
            const int row_a = r0 + g, row_b = r0 + g + 8;
            const float la = ls[row_a], lb = ls[row_b], Da = Dv[row_a], Db = Dv[row_b];
            RowKeep rk;
            rk.rb4_a = ((uint32_t)bhL + dc.bh_off * (uint32_t)L + (uint32_t)min(row_a, L - 1)) * (uint32_t)(Lp >> 2);
            rk.rb4_b = ((uint32_t)bhL + dc.bh_off * (uint32_t)L + (uint32_t)min(row_b, L - 1)) * (uint32_t)(Lp >> 2);
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const int kend = min(r0 + 16, L);
            for (int n0 = 0; n0 < kend; n0 += 8) {
                float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                mma_xt<X3>(s, aq, ks, n0, g, t);
                mma_xt<X3>(dp, ag, vs, n0, g, t);
                if (dc.train) {
                    bool kp[4];
                    keep_rows(dc, site, rk, bl0, bl1, n0, t, kp);
#pragma unroll
                    for (int e = 0; e < 4; ++e) dp[e] = kp[e] ? dp[e] : 0.f;
                }
                float ds[4];
                ds[0] = ex2(s[0] - la) * (dp[0] - Da); ds[1] = ex2(s[1] - la) * (dp[1] - Da);
                ds[2] = ex2(s[2] - lb) * (dp[2] - Db); ds[3] = ex2(s[3] - lb) * (dp[3] - Db);
                if (n0 + 8 > r0 + 1 || n0 + 8 > L) {              // diagonal / ragged block
                    const int c = n0 + 2 * t;
                    if (!(c <= row_a && c < L)) ds[0] = 0.f;
                    if (!(c + 1 <= row_a && c + 1 < L)) ds[1] = 0.f;
                    if (!(c <= row_b && c < L)) ds[2] = 0.f;
                    if (!(c + 1 <= row_b && c + 1 < L)) ds[3] = 0.f;
                }
                mma_px<X3>(acc, ds, ks, n0, g, t);
            }
            if (row_a < L) *reinterpret_cast<float4*>(dq + base + (size_t)row_a * D + 4 * t) = px_row(acc, 0, 0.25f);
            if (row_b < L) *reinterpret_cast<float4*>(dq + base + (size_t)row_b * D + 4 * t) = px_row(acc, 1, 0.25f);
        } else {
            // ---------------- pass B: key tile; S^T = K Q^T so that P^T / dS^T land in accumulator layout
            const int j0 = (item >> 1) * 16;
            AFrag ak, av;
            load_a16(ak.hi, ks, j0, g, t);
            load_a16(av.hi, vs, j0, g, t);
            finish_a<X3>(ak);
            finish_a<X3>(av);
            const int key_a = j0 + g, key_b = j0 + g + 8;
            float dka[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            float dva[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            // hash sharing: for a fixed query, the four keys of a hash group sit in the four lanes that differ in
            // g & 3 (lane bits 2,3); lane (g & 3) == h computes hash #h of {(qa,key_a),(qb,key_a),(qa,key_b),(qb,key_b)}
            const int hsel = g & 3;
            const uint32_t grp_a = (uint32_t)(key_a >> 2), grp_b = (uint32_t)(key_b >> 2);
            const int src_base = lane & ~12;
            const ByteLane bka(key_a & 3, dc.thr16), bkb(key_b & 3, dc.thr16);
            for (int i0 = j0; i0 < L; i0 += 8) {        // queries i >= key, 8 at a time
                float st[4] = {0.f, 0.f, 0.f, 0.f}, dpt[4] = {0.f, 0.f, 0.f, 0.f};
                mma_xt<X3>(st, ak, qs, i0, g, t);       // st[key][query] = k_key . q_query (log2 domain)
                mma_xt<X3>(dpt, av, gs, i0, g, t);      // dpt[key][query] = v_key . dO_query
                const int qa = i0 + 2 * t, qb = qa + 1;
                const float lqa = ls[qa], lqb = ls[qb], Dqa = Dv[qa], Dqb = Dv[qb];
                float p[4];
                p[0] = ex2(st[0] - lqa); p[1] = ex2(st[1] - lqb); p[2] = ex2(st[2] - lqa); p[3] = ex2(st[3] - lqb);
                if (i0 < j0 + 16 || i0 + 8 > L) {       // diagonal / ragged block: key <= query < L
                    if (!(key_a <= qa && qa < L)) p[0] = 0.f;
                    if (!(key_a <= qb && qb < L)) p[1] = 0.f;
                    if (!(key_b <= qa && qa < L)) p[2] = 0.f;
                    if (!(key_b <= qb && qb < L)) p[3] = 0.f;
                }
                float pd[4] = {p[0], p[1], p[2], p[3]};
                if (dc.train) {
                    const int qsel = min((hsel & 1) ? qb : qa, L - 1);
                    const uint32_t idx4 = ((uint32_t)bhL + dc.bh_off * (uint32_t)L + (uint32_t)qsel) * (uint32_t)(Lp >> 2) + ((hsel & 2) ? grp_b : grp_a);
                    const uint32_t mine = rng4(dc.seed, site, (uint64_t)idx4);
                    uint32_t r[4];
#pragma unroll
                    for (int h4 = 0; h4 < 4; ++h4) r[h4] = __shfl_sync(0xffffffffu, mine, src_base | (h4 << 2));
                    // r[0]=(qa,key_a) r[1]=(qb,key_a) r[2]=(qa,key_b) r[3]=(qb,key_b); lane inside the group = key & 3
                    const bool k0 = bka.keep(r[0]), k1 = bka.keep(r[1]);
                    const bool k2 = bkb.keep(r[2]), k3 = bkb.keep(r[3]);
                    pd[0] = k0 ? p[0] : 0.f; pd[1] = k1 ? p[1] : 0.f;
                    pd[2] = k2 ? p[2] : 0.f; pd[3] = k3 ? p[3] : 0.f;
                    dpt[0] = k0 ? dpt[0] : 0.f; dpt[1] = k1 ? dpt[1] : 0.f;
                    dpt[2] = k2 ? dpt[2] : 0.f; dpt[3] = k3 ? dpt[3] : 0.f;
                }
                float ds[4];
                ds[0] = p[0] * (dpt[0] - Dqa); ds[1] = p[1] * (dpt[1] - Dqb);
                ds[2] = p[2] * (dpt[2] - Dqa); ds[3] = p[3] * (dpt[3] - Dqb);
                mma_px<X3>(dva, pd, gs, i0, g, t);      // dv[key] += Pd^T[key][query] dO[query]
                mma_px<X3>(dka, ds, qs, i0, g, t);      // dk[key] += dS^T[key][query] q[query] (q carries log2e)
            }
            if (key_a < L) {
                *reinterpret_cast<float4*>(dk + base + (size_t)key_a * D + 4 * t) = px_row(dka, 0, LN2);
                *reinterpret_cast<float4*>(dv + base + (size_t)key_a * D + 4 * t) = px_row(dva, 0, sc);
            }
            if (key_b < L) {
                *reinterpret_cast<float4*>(dk + base + (size_t)key_b * D + 4 * t) = px_row(dka, 1, LN2);
                *reinterpret_cast<float4*>(dv + base + (size_t)key_b * D + 4 * t) = px_row(dva, 1, sc);
            }
        }
    }
}

}  // namespace attn
}  // namespace amid
