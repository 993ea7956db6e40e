// Causal multi-head attention for short sequences (L <= 512, head_dim 16) on the warp-level
// tensor-core path (mma.sync m16n8k8, TF32 operands, fp32 accumulate).  One CTA per (sample, head);
// q/k/v (and dO in the backward) of the head live in shared memory with a 20-float row stride, which
// makes every fragment load bank-conflict free.  Softmax statistics, dropout and the causal mask are
// applied on the accumulator fragments; P (or dS) goes straight back into the next MMA as the A
// operand by relabelling the accumulator columns (2t -> k index t, 2t+1 -> k index t+4) and reading
// the B operand rows in the same permuted order.
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
// dynamic work queue of a CTA: warps take items in order (callers order items by decreasing cost)
__device__ __forceinline__ int next_item(int* counter, int lane) {
    int it = 0;
    if (lane == 0) it = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
}

// head slice [L,16] of a [M,128] tensor -> smem [L][LDS]
__device__ __forceinline__ void stage(float* s, const float* __restrict__ g, int L) {
    for (int idx = threadIdx.x; idx < L * 4; idx += blockDim.x) {
        const int r = idx >> 2, c4 = idx & 3;
        *reinterpret_cast<float4*>(s + r * LDS + c4 * 4) = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * D) + c4);
    }
}
// A fragments (16 rows x 16 cols = 2 k-steps) of rows r0.. from a staged matrix; rows clamped to L-1
__device__ __forceinline__ void load_a16(uint32_t (&a)[2][4], const float* s, int r0, int L, int g, int t) {
    const int ra = min(r0 + g, L - 1), rb = min(r0 + g + 8, L - 1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = fbits(s[ra * LDS + 8 * ks + t]);
        a[ks][1] = fbits(s[rb * LDS + 8 * ks + t]);
        a[ks][2] = fbits(s[ra * LDS + 8 * ks + t + 4]);
        a[ks][3] = fbits(s[rb * LDS + 8 * ks + t + 4]);
    }
}
// the same from global memory (row stride D): used where the matrix is only ever an A operand
__device__ __forceinline__ void load_a16_g(uint32_t (&a)[2][4], const float* __restrict__ p, int r0, int L, int g, int t) {
    const float* pa = p + (size_t)min(r0 + g, L - 1) * D;
    const float* pb = p + (size_t)min(r0 + g + 8, L - 1) * D;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = fbits(__ldg(pa + 8 * ks + t));
        a[ks][1] = fbits(__ldg(pb + 8 * ks + t));
        a[ks][2] = fbits(__ldg(pa + 8 * ks + t + 4));
        a[ks][3] = fbits(__ldg(pb + 8 * ks + t + 4));
    }
}
// D[16 x 8] = A[16 x 16] * X[n0..n0+8][16]^T   (X rows are the n index; k = feature)
__device__ __forceinline__ void mma_xt(float (&d)[4], const uint32_t (&a)[2][4], const float* x, int n0, int L, int g, int t) {
    const int n = min(n0 + g, L - 1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) mma_tf32(d, a[ks], fbits(x[n * LDS + 8 * ks + t]), fbits(x[n * LDS + 8 * ks + t + 4]));
}
// acc[dt][..] += P[16 x 8 (relabelled)] * X[n0..n0+8][16]   (X rows are the k index, permuted 2t / 2t+1)
__device__ __forceinline__ void mma_px(float (&acc)[2][4], const float (&p)[4], const float* x, int n0, int L, int g, int t) {
    const uint32_t a[4] = {fbits(p[0]), fbits(p[2]), fbits(p[1]), fbits(p[3])};
    const int ka = min(n0 + 2 * t, L - 1), kb = min(n0 + 2 * t + 1, L - 1);
#pragma unroll
    for (int dt = 0; dt < 2; ++dt) mma_tf32(acc[dt], a, fbits(x[ka * LDS + 8 * dt + g]), fbits(x[kb * LDS + 8 * dt + g]));
}
// keep bits of the two elements (row, col), (row, col+1) with col even
__device__ __forceinline__ void keep2(const DropCfg& dc, uint32_t site, uint64_t bhL, int row, int col, int Lp, bool& k0, bool& k1) {
    const uint32_t r = rng4(dc.seed, site, ((bhL + row) * Lp + col) >> 2);
    k0 = rng_keep(r, col & 3, dc.thr16);
    k1 = rng_keep(r, (col & 3) + 1, dc.thr16);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NW * 32)
k_attn_fwd_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               float* __restrict__ o, float* __restrict__ lse, int L, DropCfg dc, uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int queue;
    float* ks = smem;
    float* vs = ks + L * LDS;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    if (threadIdx.x == 0) queue = 0;
    stage(ks, k + base, L);
    stage(vs, v + base, L);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntile = (L + 15) / 16, Lp = (L + 3) & ~3;
    const uint64_t bhL = (uint64_t)bh * L;
    // causal work per row tile grows with its index: hand tiles out from the last (heaviest) to the first
    for (;;) {
        {
            const int item = next_item(&queue, lane);
            if (item >= ntile) break;
            const int rt = ntile - 1 - item;
            const int r0 = rt * 16;
            uint32_t aq[2][4];
            load_a16_g(aq, q + base, r0, L, g, t);
            float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const int row_a = r0 + g, row_b = r0 + g + 8;
            const int kend = min(r0 + 16, L);              // keys [0, kend) can be visible to this tile
            for (int kb = 0; kb < kend; kb += 32) {
                float s[4][4];
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n0 = kb + 8 * nt;
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                    if (n0 < kend) mma_xt(s[nt], aq, ks, n0, L, g, t);     // warp-uniform
                    const int c = n0 + 2 * t;
                    s[nt][0] = (n0 < kend && c <= row_a && c < L) ? s[nt][0] : -INFINITY;
                    s[nt][1] = (n0 < kend && c + 1 <= row_a && c + 1 < L) ? s[nt][1] : -INFINITY;
                    s[nt][2] = (n0 < kend && c <= row_b && c < L) ? s[nt][2] : -INFINITY;
                    s[nt][3] = (n0 < kend && c + 1 <= row_b && c + 1 < L) ? s[nt][3] : -INFINITY;
                    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
                }
                const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
                // rows beyond the sequence (clamped loads) can stay at -inf for a while: guard the subtraction
                const float sub0 = mn0 == -INFINITY ? 0.f : mn0, sub1 = mn1 == -INFINITY ? 0.f : mn1;
                const float c0 = ex2((m0 - sub0) * LOG2E), c1 = ex2((m1 - sub1) * LOG2E);
                l0 *= c0; l1 *= c1;
#pragma unroll
                for (int dt = 0; dt < 2; ++dt) { acc[dt][0] *= c0; acc[dt][1] *= c0; acc[dt][2] *= c1; acc[dt][3] *= c1; }
                m0 = mn0; m1 = mn1;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n0 = kb + 8 * nt;
                    if (n0 >= kend) continue;                               // warp-uniform
                    float p[4];
                    p[0] = ex2((s[nt][0] - sub0) * LOG2E); p[1] = ex2((s[nt][1] - sub0) * LOG2E);
                    p[2] = ex2((s[nt][2] - sub1) * LOG2E); p[3] = ex2((s[nt][3] - sub1) * LOG2E);
                    l0 += p[0] + p[1]; l1 += p[2] + p[3];
                    if (dc.train) {
                        bool ka, kb_, kc, kd;
                        keep2(dc, site, bhL, min(row_a, L - 1), n0 + 2 * t, Lp, ka, kb_);
                        keep2(dc, site, bhL, min(row_b, L - 1), n0 + 2 * t, Lp, kc, kd);
                        p[0] = ka ? p[0] * dc.scale : 0.f; p[1] = kb_ ? p[1] * dc.scale : 0.f;
                        p[2] = kc ? p[2] * dc.scale : 0.f; p[3] = kd ? p[3] * dc.scale : 0.f;
                    }
                    mma_px(acc, p, vs, n0, L, g, t);
                }
            }
            l0 = quad_sum(l0); l1 = quad_sum(l1);
            const float i0 = 1.0f / l0, i1 = 1.0f / l1;
            if (row_a < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt)
                    *reinterpret_cast<float2*>(o + base + (size_t)row_a * D + 8 * dt + 2 * t) = make_float2(acc[dt][0] * i0, acc[dt][1] * i0);
                if (t == 0) lse[bhL + row_a] = m0 + logf(l0);
            }
            if (row_b < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt)
                    *reinterpret_cast<float2*>(o + base + (size_t)row_b * D + 8 * dt + 2 * t) = make_float2(acc[dt][2] * i1, acc[dt][3] * i1);
                if (t == 0) lse[bhL + row_b] = m1 + logf(l1);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// backward: pass A (query tiles -> dq), pass B (key tiles -> dk, dv); P recomputed from lse
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NWB * 32)
k_attn_bwd_mma(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
               float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, DropCfg dc, uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int queue;
    if (threadIdx.x == 0) queue = 0;
    float* qs = smem;
    float* ks = qs + L * LDS;
    float* vs = ks + L * LDS;
    float* gs = vs + L * LDS;     // dO
    float* Dv = gs + L * LDS;     // D_i = <dO_i, O_i>
    float* ls = Dv + L;           // lse * log2(e) is NOT folded: plain lse
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    stage(qs, q + base, L);
    stage(ks, k + base, L);
    stage(vs, v + base, L);
    stage(gs, dO + base, L);
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float4* po = reinterpret_cast<const float4*>(o + base + (size_t)i * D);
        const float4* pg = reinterpret_cast<const float4*>(dO + base + (size_t)i * D);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 a = __ldg(po + c), bb = __ldg(pg + c);
            s += a.x * bb.x + a.y * bb.y + a.z * bb.z + a.w * bb.w;
        }
        Dv[i] = s;
        ls[i] = lse[(size_t)bh * L + i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntile = (L + 15) / 16, Lp = (L + 3) & ~3;
    const uint64_t bhL = (uint64_t)bh * L;
    // Work items, heaviest first: item 2n = pass A on query tile ntile-1-n, item 2n+1 = pass B on key tile n.
    for (;;) {
        const int item = next_item(&queue, lane);
        if (item >= 2 * ntile) break;
        if ((item & 1) == 0) {
            // ---------------- pass A: dq[i] = 0.25 * sum_j dS_ij k_j
            const int rt = ntile - 1 - (item >> 1);
            const int r0 = rt * 16;
            uint32_t aq[2][4], ag[2][4];
            load_a16(aq, qs, r0, L, g, t);
            load_a16(ag, gs, r0, L, g, t);
            const int row_a = r0 + g, row_b = r0 + g + 8;
            const int ra = min(row_a, L - 1), rb = min(row_b, L - 1);
            const float la = ls[ra], lb = ls[rb], Da = Dv[ra], Db = Dv[rb];
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const int kend = min(r0 + 16, L);
            for (int n0 = 0; n0 < kend; n0 += 8) {
                float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                mma_xt(s, aq, ks, n0, L, g, t);
                mma_xt(dp, ag, vs, n0, L, g, t);
                const int c = n0 + 2 * t;
                bool k0 = true, k1 = true, k2 = true, k3 = true;
                if (dc.train) { keep2(dc, site, bhL, ra, c, Lp, k0, k1); keep2(dc, site, bhL, rb, c, Lp, k2, k3); }
                const float sc = dc.train ? dc.scale : 1.0f;
                float ds[4];
                ds[0] = (c <= row_a && c < L) ? ex2((s[0] - la) * LOG2E) * ((k0 ? dp[0] * sc : 0.f) - Da) : 0.f;
                ds[1] = (c + 1 <= row_a && c + 1 < L) ? ex2((s[1] - la) * LOG2E) * ((k1 ? dp[1] * sc : 0.f) - Da) : 0.f;
                ds[2] = (c <= row_b && c < L) ? ex2((s[2] - lb) * LOG2E) * ((k2 ? dp[2] * sc : 0.f) - Db) : 0.f;
                ds[3] = (c + 1 <= row_b && c + 1 < L) ? ex2((s[3] - lb) * LOG2E) * ((k3 ? dp[3] * sc : 0.f) - Db) : 0.f;
                mma_px(acc, ds, ks, n0, L, g, t);
            }
            if (row_a < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt)
                    *reinterpret_cast<float2*>(dq + base + (size_t)row_a * D + 8 * dt + 2 * t) = make_float2(acc[dt][0] * 0.25f, acc[dt][1] * 0.25f);
            }
            if (row_b < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt)
                    *reinterpret_cast<float2*>(dq + base + (size_t)row_b * D + 8 * dt + 2 * t) = make_float2(acc[dt][2] * 0.25f, acc[dt][3] * 0.25f);
            }
        } else {
            // ---------------- pass B: key tile; S^T = K Q^T so that P^T / dS^T land in accumulator layout
            const int kt = item >> 1;
            const int j0 = kt * 16;
            uint32_t ak[2][4], av[2][4];
            load_a16(ak, ks, j0, L, g, t);
            load_a16(av, vs, j0, L, g, t);
            const int key_a = j0 + g, key_b = j0 + g + 8;
            float dka[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            float dva[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            for (int i0 = j0 & ~7; i0 < L; i0 += 8) {      // queries i >= key; 8-query blocks from the tile's first key
                float st[4] = {0.f, 0.f, 0.f, 0.f}, dpt[4] = {0.f, 0.f, 0.f, 0.f};
                mma_xt(st, ak, qs, i0, L, g, t);           // st[key][query] = k_key . q_query
                mma_xt(dpt, av, gs, i0, L, g, t);          // dpt[key][query] = v_key . dO_query
                const int qa = i0 + 2 * t, qb = qa + 1;    // the two query columns of this thread
                const int qca = min(qa, L - 1), qcb = min(qb, L - 1);
                const float lqa = ls[qca], lqb = ls[qcb], Dqa = Dv[qca], Dqb = Dv[qcb];
                const float sc = dc.train ? dc.scale : 1.0f;
                float pd[4], ds[4];
                // element e: (key, query) = (key_a, qa), (key_a, qb), (key_b, qa), (key_b, qb)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = (e < 2) ? key_a : key_b;
                    const int qi = (e & 1) ? qb : qa;
                    const float lq = (e & 1) ? lqb : lqa, Dq = (e & 1) ? Dqb : Dqa;
                    float pe = 0.f, de = 0.f;
                    if (key <= qi && qi < L && key < L) {
                        const float p = ex2((st[e] - lq) * LOG2E);
                        bool kp = true;
                        if (dc.train) {
                            const uint32_t r = rng4(dc.seed, site, ((bhL + qi) * Lp + key) >> 2);
                            kp = rng_keep(r, key & 3, dc.thr16);
                        }
                        pe = kp ? p * sc : 0.f;
                        de = p * ((kp ? dpt[e] * sc : 0.f) - Dq);
                    }
                    pd[e] = pe; ds[e] = de;
                }
                mma_px(dva, pd, gs, i0, L, g, t);          // dv[key] += Pd^T[key][query] dO[query]
                mma_px(dka, ds, qs, i0, L, g, t);          // dk[key] += dS^T[key][query] q[query]
            }
            if (key_a < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt) {
                    *reinterpret_cast<float2*>(dk + base + (size_t)key_a * D + 8 * dt + 2 * t) = make_float2(dka[dt][0], dka[dt][1]);
                    *reinterpret_cast<float2*>(dv + base + (size_t)key_a * D + 8 * dt + 2 * t) = make_float2(dva[dt][0], dva[dt][1]);
                }
            }
            if (key_b < L) {
#pragma unroll
                for (int dt = 0; dt < 2; ++dt) {
                    *reinterpret_cast<float2*>(dk + base + (size_t)key_b * D + 8 * dt + 2 * t) = make_float2(dka[dt][2], dka[dt][3]);
                    *reinterpret_cast<float2*>(dv + base + (size_t)key_b * D + 8 * dt + 2 * t) = make_float2(dva[dt][2], dva[dt][3]);
                }
            }
        }
    }
}

}  // namespace attn
}  // namespace amid
