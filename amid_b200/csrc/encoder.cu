// SASRec encoder (Log2feats, model_seq.py:331-387) forward and backward, exact-fp32 path.
//
// Per block, three fused kernels forward:
//   k_ln_qkv   : Qn = LN1(x); q = 0.25 (Qn Wq^T + bq); k = x Wk^T + bk; v = x Wv^T + bv
//   k_attn_fwd : causal softmax(q k^T) (+dropout) v, one CTA per (sample, head), K/V in smem
//   k_proj_ffn : x1 = Qn + o Wo^T + bo; y = LN2(x1); h = relu(drop(y W1^T + b1));
//                xout = (drop(h W2^T + b2) + y) * ~tmask   [+ last LayerNorm on the last block]
// and backward: k_ffn_bwd, k_attn_bwd, k_qkv_bwd, k_wgrad (+ tiny reductions).
// A 128-token tile stays in shared memory across the chained GEMMs of a kernel, so each
// activation crosses HBM once per kernel instead of once per op.
#include "tile.cuh"
#include "encoder_tc.cuh"
#include "attn_mma.cuh"
#include "encoder_tc16.cuh"
#include "x3.cuh"
#include "attn_tc.cuh"
#include "attn_p.cuh"
#include <algorithm>

namespace amid {

// ----------------------------------------------------------------------------------
// weight transposes: Wt[k][n] = W[n][k] for the forward GEMMs (12 matrices per encoder)
// ----------------------------------------------------------------------------------
struct TransJobs {
    const float* src[12];
};
__global__ void k_transpose128(TransJobs jobs, float* __restrict__ dst) {
    __shared__ float t[32][33];
    const float* s = jobs.src[blockIdx.z];
    float* d = dst + (size_t)blockIdx.z * D * D;
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) t[j][threadIdx.x] = s[(size_t)(by + j) * D + bx + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) d[(size_t)(bx + j) * D + by + threadIdx.x] = t[threadIdx.x][j];
}

// ----------------------------------------------------------------------------------
// forward kernel 1: LN1 + QKV projections
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_ln_qkv(const float* __restrict__ x, int M, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
         const float* __restrict__ WqT, const float* __restrict__ WkT, const float* __restrict__ WvT,
         const float* __restrict__ in_b, float* __restrict__ qn, float* __restrict__ st1,
         float* __restrict__ q, float* __restrict__ k, float* __restrict__ v) {
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;
    float* Qs = smem + TILE_FLOATS;
    float* Ws = smem + 2 * TILE_FLOATS;
    const int row0 = blockIdx.x * TM;
    Frag f;
    load_tile(Xs, x, row0, M);
    __syncthreads();
    ln_tile(Xs, Qs, ln_w, ln_b, row0, M, qn, st1);
    __syncthreads();
    float acc[8][8];
    const float* As[3] = {Qs, Xs, Xs};
    const float* Bs[3] = {WqT, WkT, WvT};
    float* outs[3] = {q, k, v};
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
        tile_gemm<false>(As[g], Bs[g], Ws, acc);
        const float sc = g == 0 ? 0.25f : 1.0f;  // q * sqrt(1/head_dim), functional.py:6632
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(in_b + g * D + f.c0()));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(in_b + g * D + f.c1()));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int gr = row0 + f.row(i);
            if (gr < M) {
                st4(outs[g] + (size_t)gr * D + f.c0(), scale4(add4(f4(acc[i], 0), b0), sc));
                st4(outs[g] + (size_t)gr * D + f.c1(), scale4(add4(f4(acc[i], 1), b1), sc));
            }
        }
    }
}

// ----------------------------------------------------------------------------------
// forward kernel 3: out-proj + residual + LN2 + FFN + mask (+ last LN)
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_proj_ffn(const float* __restrict__ o, const float* __restrict__ qn, int M,
           const float* __restrict__ WoT, const float* __restrict__ bo,
           const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
           const float* __restrict__ W1T, const float* __restrict__ b1,
           const float* __restrict__ W2T, const float* __restrict__ b2,
           const uint32_t* __restrict__ tmask, DropCfg dc, uint32_t site1, uint32_t site2,
           float* __restrict__ x1, float* __restrict__ st2, float* __restrict__ y, float* __restrict__ h,
           float* __restrict__ xout,
           const float* __restrict__ ln3_w, const float* __restrict__ ln3_b, float* __restrict__ enc,
           float* __restrict__ st3) {
    extern __shared__ __align__(16) float smem[];
    float* T0 = smem;
    float* T1 = smem + TILE_FLOATS;
    float* Ws = smem + 2 * TILE_FLOATS;
    const int row0 = blockIdx.x * TM;
    Frag f;
    load_tile(T0, o, row0, M);
    load_tile(T1, qn, row0, M);
    __syncthreads();
    float acc[8][8];
    // x1 = Qn + o Wo^T + bo   (model_seq.py:378)
    tile_gemm<false>(T0, WoT, Ws, acc);
    {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(bo + f.c0()));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(bo + f.c1()));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = f.row(i), gr = row0 + r;
            float4 v0 = add4(add4(f4(acc[i], 0), c0), ld4(T1 + r * LDA + f.c0()));
            float4 v1 = add4(add4(f4(acc[i], 1), c1), ld4(T1 + r * LDA + f.c1()));
            st4(T0 + r * LDA + f.c0(), v0);
            st4(T0 + r * LDA + f.c1(), v1);
            if (gr < M) {
                st4(x1 + (size_t)gr * D + f.c0(), v0);
                st4(x1 + (size_t)gr * D + f.c1(), v1);
            }
        }
    }
    __syncthreads();
    // y = LN2(x1)  -> T1
    ln_tile(T0, T1, ln2_w, ln2_b, row0, M, y, st2);
    __syncthreads();
    // h = relu(dropout1(y W1^T + b1))   (model_seq.py:323: dropout sits between conv1 and ReLU)
    tile_gemm<false>(T1, W1T, Ws, acc);
    {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(b1 + f.c0()));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(b1 + f.c1()));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = f.row(i), gr = row0 + r;
            float4 v0 = add4(f4(acc[i], 0), c0);
            float4 v1 = add4(f4(acc[i], 1), c1);
            if (dc.train) {
                v0 = drop4(v0, dc, site1, (uint64_t)(gr + dc.tok_off) * D + f.c0());
                v1 = drop4(v1, dc, site1, (uint64_t)(gr + dc.tok_off) * D + f.c1());
            }
            v0 = make_float4(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f), fmaxf(v0.z, 0.f), fmaxf(v0.w, 0.f));
            v1 = make_float4(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f), fmaxf(v1.z, 0.f), fmaxf(v1.w, 0.f));
            st4(T0 + r * LDA + f.c0(), v0);
            st4(T0 + r * LDA + f.c1(), v1);
            if (gr < M) {
                st4(h + (size_t)gr * D + f.c0(), v0);
                st4(h + (size_t)gr * D + f.c1(), v1);
            }
        }
    }
    // xout = (dropout2(h W2^T + b2) + y) * ~tmask   (model_seq.py:323-325, :383)
    tile_gemm<false>(T0, W2T, Ws, acc);
    {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(b2 + f.c0()));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(b2 + f.c1()));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = f.row(i), gr = row0 + r;
            float4 v0 = add4(f4(acc[i], 0), c0);
            float4 v1 = add4(f4(acc[i], 1), c1);
            if (dc.train) {
                v0 = drop4(v0, dc, site2, (uint64_t)(gr + dc.tok_off) * D + f.c0());
                v1 = drop4(v1, dc, site2, (uint64_t)(gr + dc.tok_off) * D + f.c1());
            }
            v0 = add4(v0, ld4(T1 + r * LDA + f.c0()));
            v1 = add4(v1, ld4(T1 + r * LDA + f.c1()));
            if (gr < M) {
                const uint4 tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr);
                v0 = apply_tmask(v0, tw, f.tn);
                v1 = apply_tmask(v1, tw, 16 + f.tn);
                st4(xout + (size_t)gr * D + f.c0(), v0);
                st4(xout + (size_t)gr * D + f.c1(), v1);
            }
            if (enc) {  // uniform across the CTA
                st4(T0 + r * LDA + f.c0(), v0);
                st4(T0 + r * LDA + f.c1(), v1);
            }
        }
    }
    if (enc) {  // last_layernorm (model_seq.py:385)
        __syncthreads();
        ln_tile(T0, T0, ln3_w, ln3_b, row0, M, enc, st3);
    }
}

// ----------------------------------------------------------------------------------
// attention forward: one CTA per (sample, head); thread t owns query rows t and L-1-t
// (balanced causal work); K/V of the head staged in shared memory; online softmax.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float dot16(const float (&a)[16], const float* __restrict__ b) {
    const float4 b0 = ld4(b), b1 = ld4(b + 4), b2 = ld4(b + 8), b3 = ld4(b + 12);
    float s = a[0] * b0.x;
    s = fmaf(a[1], b0.y, s); s = fmaf(a[2], b0.z, s); s = fmaf(a[3], b0.w, s);
    s = fmaf(a[4], b1.x, s); s = fmaf(a[5], b1.y, s); s = fmaf(a[6], b1.z, s); s = fmaf(a[7], b1.w, s);
    s = fmaf(a[8], b2.x, s); s = fmaf(a[9], b2.y, s); s = fmaf(a[10], b2.z, s); s = fmaf(a[11], b2.w, s);
    s = fmaf(a[12], b3.x, s); s = fmaf(a[13], b3.y, s); s = fmaf(a[14], b3.z, s); s = fmaf(a[15], b3.w, s);
    return s;
}
__device__ __forceinline__ void axpy16(float (&acc)[16], float p, const float* __restrict__ b) {
    const float4 b0 = ld4(b), b1 = ld4(b + 4), b2 = ld4(b + 8), b3 = ld4(b + 12);
    acc[0] = fmaf(p, b0.x, acc[0]); acc[1] = fmaf(p, b0.y, acc[1]); acc[2] = fmaf(p, b0.z, acc[2]); acc[3] = fmaf(p, b0.w, acc[3]);
    acc[4] = fmaf(p, b1.x, acc[4]); acc[5] = fmaf(p, b1.y, acc[5]); acc[6] = fmaf(p, b1.z, acc[6]); acc[7] = fmaf(p, b1.w, acc[7]);
    acc[8] = fmaf(p, b2.x, acc[8]); acc[9] = fmaf(p, b2.y, acc[9]); acc[10] = fmaf(p, b2.z, acc[10]); acc[11] = fmaf(p, b2.w, acc[11]);
    acc[12] = fmaf(p, b3.x, acc[12]); acc[13] = fmaf(p, b3.y, acc[13]); acc[14] = fmaf(p, b3.z, acc[14]); acc[15] = fmaf(p, b3.w, acc[15]);
}
__device__ __forceinline__ void load16(float (&a)[16], const float* __restrict__ g) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float4 t = ld4(g + c * 4);
        a[c * 4] = t.x; a[c * 4 + 1] = t.y; a[c * 4 + 2] = t.z; a[c * 4 + 3] = t.w;
    }
}
__device__ __forceinline__ void store16(float* __restrict__ g, const float (&a)[16], float s) {
#pragma unroll
    for (int c = 0; c < 4; ++c) st4(g + c * 4, make_float4(a[c * 4] * s, a[c * 4 + 1] * s, a[c * 4 + 2] * s, a[c * 4 + 3] * s));
}
// stage the head slice [L,16] of a [M,128] tensor into smem
__device__ __forceinline__ void stage_head(float* s, const float* __restrict__ g, int L) {
    for (int idx = threadIdx.x; idx < L * 4; idx += blockDim.x) {
        int r = idx >> 2, c4 = idx & 3;
        st4(s + r * DH + c4 * 4, __ldg(reinterpret_cast<const float4*>(g + (size_t)r * D) + c4));
    }
}

__global__ void k_attn_fwd(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                           float* __restrict__ o, float* __restrict__ lse, int L, DropCfg dc, uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    float* ks = smem;
    float* vs = smem + L * DH;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    stage_head(ks, k + base, L);
    stage_head(vs, v + base, L);
    __syncthreads();
    const int half = (L + 1) / 2, t = threadIdx.x;
    const int Lp = (L + 3) & ~3;
    // the j loop bound is made warp-uniform so the K/V reads are broadcasts
    for (int pass = 0; pass < 2; ++pass) {
        int i = pass == 0 ? t : L - 1 - t;
        bool active = t < half && !(pass == 1 && i == t);
        int imax = active ? i : -1;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, of));
        if (imax < 0) continue;
        if (!active) i = 0;
        float qr[16], acc[16];
        load16(qr, q + base + (size_t)i * D);
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        float m = -INFINITY, l = 0.f;
        for (int j0 = 0; j0 <= imax; j0 += 4) {
            if (j0 > i) continue;  // this lane is done; keep looping for the warp-uniform bound
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int j = j0 + u;
                s[u] = (j <= i) ? dot16(qr, ks + j * DH) : -INFINITY;
            }
            const float mn = fmaxf(fmaxf(m, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
            const float corr = expf(m - mn);
            l *= corr;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] *= corr;
            uint32_t r = 0;
            if (dc.train) r = rng4(dc.seed, site, (((uint64_t)(bh + dc.bh_off) * L + i) * Lp + j0) >> 2);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int j = j0 + u;
                if (j <= i) {
                    float p = expf(s[u] - mn);
                    l += p;
                    if (dc.train) p = rng_keep(r, u, dc.thr16) ? p * dc.scale : 0.f;
                    axpy16(acc, p, vs + j * DH);
                }
            }
            m = mn;
        }
        if (active) {
            store16(o + base + (size_t)i * D, acc, 1.0f / l);
            lse[(size_t)bh * L + i] = m + logf(l);
        }
    }
}

// ----------------------------------------------------------------------------------
// attention backward: recompute P from q,k and the saved log-sum-exp.
//   pass A (thread per query row): dq
//   pass B (thread per key column): dk, dv
// ----------------------------------------------------------------------------------
__global__ void k_attn_bwd(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                           const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
                           float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, DropCfg dc,
                           uint32_t site) {
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;
    float* ks = qs + L * DH;
    float* vs = ks + L * DH;
    float* ds = vs + L * DH;   // dO
    float* Dv = ds + L * DH;   // D_i = <dO_i, O_i>
    float* ls = Dv + L;        // lse
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    stage_head(qs, q + base, L);
    stage_head(ks, k + base, L);
    stage_head(vs, v + base, L);
    stage_head(ds, dO + base, L);
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        float a[16];
        load16(a, o + base + (size_t)i * D);
        Dv[i] = dot16(a, dO + base + (size_t)i * D);
        ls[i] = lse[(size_t)bh * L + i];
    }
    __syncthreads();
    const int half = (L + 1) / 2, t = threadIdx.x;
    const int Lp = (L + 3) & ~3;
    // ---- pass A: dq_i = 0.25 * sum_{j<=i} dS_ij k_j
    for (int pass = 0; pass < 2; ++pass) {
        int i = pass == 0 ? t : L - 1 - t;
        bool active = t < half && !(pass == 1 && i == t);
        int imax = active ? i : -1;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, of));
        if (imax < 0) continue;
        if (!active) i = 0;
        float qr[16], dor[16], acc[16];
        load16(qr, qs + i * DH);
        load16(dor, ds + i * DH);
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        const float li = ls[i], Di = Dv[i];
        for (int j0 = 0; j0 <= imax; j0 += 4) {
            if (j0 > i) continue;
            uint32_t r = 0;
            if (dc.train) r = rng4(dc.seed, site, (((uint64_t)(bh + dc.bh_off) * L + i) * Lp + j0) >> 2);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int j = j0 + u;
                if (j <= i) {
                    const float p = expf(dot16(qr, ks + j * DH) - li);
                    float dp = dot16(dor, vs + j * DH);
                    if (dc.train) dp = rng_keep(r, u, dc.thr16) ? dp * dc.scale : 0.f;
                    axpy16(acc, p * (dp - Di), ks + j * DH);
                }
            }
        }
        if (active) store16(dq + base + (size_t)i * D, acc, 0.25f);
    }
    // ---- pass B: dk_j = sum_{i>=j} dS_ij q_i ; dv_j = sum_{i>=j} Pd_ij dO_i
    for (int pass = 0; pass < 2; ++pass) {
        int j = pass == 0 ? t : L - 1 - t;
        bool active = t < half && !(pass == 1 && j == t);
        int jmin = active ? j : L;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, of));
        if (jmin >= L) continue;
        if (!active) j = L;  // never satisfies i >= j
        float kr[16], vr[16], ak[16], av[16];
        const int jj = active ? j : 0;
        load16(kr, ks + jj * DH);
        load16(vr, vs + jj * DH);
#pragma unroll
        for (int c = 0; c < 16; ++c) { ak[c] = 0.f; av[c] = 0.f; }
        for (int i = jmin; i < L; ++i) {
            if (i < j) continue;
            const float p = expf(dot16(kr, qs + i * DH) - ls[i]);
            float dp = dot16(vr, ds + i * DH);
            float pd = p;
            if (dc.train) {
                const uint32_t r = rng4(dc.seed, site, (((uint64_t)(bh + dc.bh_off) * L + i) * Lp + j) >> 2);
                const bool kp = rng_keep(r, j & 3, dc.thr16);
                dp = kp ? dp * dc.scale : 0.f;
                pd = kp ? p * dc.scale : 0.f;
            }
            axpy16(ak, p * (dp - Dv[i]), qs + i * DH);
            axpy16(av, pd, ds + i * DH);
        }
        if (active) {
            store16(dk + base + (size_t)j * D, ak, 1.0f);
            store16(dv + base + (size_t)j * D, av, 1.0f);
        }
    }
}

// ----------------------------------------------------------------------------------
// backward kernel: last LayerNorm (row-wise), also used nowhere else
// ----------------------------------------------------------------------------------
// dx = rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat)); per-CTA partial dw, db.
__global__ void __launch_bounds__(NT)
k_ln_bwd(const float* __restrict__ dy, const float* __restrict__ xin, const float* __restrict__ st,
         const float* __restrict__ w, int M, float* __restrict__ dx, float* __restrict__ part /*[grid][256]*/) {
    __shared__ float red[NT / 32][2 * D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + lane);
    float4 dw = make_float4(0, 0, 0, 0), db = make_float4(0, 0, 0, 0);
    const int row0 = blockIdx.x * TM;
    for (int r = warp; r < TM; r += NT / 32) {
        const int gr = row0 + r;
        if (gr >= M) break;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (size_t)gr * D) + lane);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xin + (size_t)gr * D) + lane);
        const float mean = st[(size_t)gr * 2], rstd = st[(size_t)gr * 2 + 1];
        const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        const float4 gw = mul4(g, wv);
        const float c1 = warp_sum(sum4(gw)) * (1.0f / D);
        const float c2 = warp_sum(sum4(mul4(gw, xh))) * (1.0f / D);
        st4(dx + (size_t)gr * D + lane * 4,
            make_float4(rstd * (gw.x - c1 - xh.x * c2), rstd * (gw.y - c1 - xh.y * c2),
                        rstd * (gw.z - c1 - xh.z * c2), rstd * (gw.w - c1 - xh.w * c2)));
        dw = add4(dw, mul4(g, xh));
        db = add4(db, g);
    }
    st4(&red[warp][lane * 4], dw);
    st4(&red[warp][D + lane * 4], db);
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int wv_ = 0; wv_ < NT / 32; ++wv_) s += red[wv_][threadIdx.x];
    part[(size_t)blockIdx.x * 2 * D + threadIdx.x] = s;
}

// ----------------------------------------------------------------------------------
// backward kernel: FFN + LN2 + out-proj input gradient
//   in : dxo (grad of block output), h, x1, st2, tmask
//   out: do2, dhpre (for the weight-grad GEMMs), dx1, dO, LN2 partial dw/db
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_ffn_bwd(const float* __restrict__ dxo, const float* __restrict__ h, const float* __restrict__ x1,
          const float* __restrict__ st2, const uint32_t* __restrict__ tmask, int M,
          const float* __restrict__ W2, const float* __restrict__ W1, const float* __restrict__ Wo,
          const float* __restrict__ ln2_w, DropCfg dc, uint32_t site1, uint32_t site2,
          float* __restrict__ do2, float* __restrict__ dhpre, float* __restrict__ dx1, float* __restrict__ dO,
          float* __restrict__ ln_part /*[tiles][256]*/) {
    extern __shared__ __align__(16) float smem[];
    float* T0 = smem;
    float* T1 = smem + TILE_FLOATS;
    float* Ws = smem + 2 * TILE_FLOATS;
    const int row0 = blockIdx.x * TM;
    Frag f;
    // T0 = do2 = dropout2-mask * g,  g = dxo * ~tmask
    for (int idx = threadIdx.x; idx < TM * (D / 4); idx += NT) {
        int r = idx >> 5, c4 = idx & 31, gr = row0 + r;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < M) {
            g = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)gr * D) + c4);
            g = apply_tmask(g, __ldg(reinterpret_cast<const uint4*>(tmask) + gr), c4);
            if (dc.train) g = drop4(g, dc, site2, (uint64_t)(gr + dc.tok_off) * D + c4 * 4);
            st4(do2 + (size_t)gr * D + c4 * 4, g);
        }
        st4(T0 + r * LDA + c4 * 4, g);
    }
    __syncthreads();
    float acc[8][8];
    // dh = do2 W2 ; dhpre = dh * scale * [h > 0]
    tile_gemm<false>(T0, W2, Ws, acc);
    {
        const float sc = dc.train ? dc.scale : 1.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = f.row(i), gr = row0 + r;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (gr < M) {
                const float4 h0 = __ldg(reinterpret_cast<const float4*>(h + (size_t)gr * D + f.c0()));
                const float4 h1 = __ldg(reinterpret_cast<const float4*>(h + (size_t)gr * D + f.c1()));
                v0 = make_float4(h0.x > 0.f ? acc[i][0] * sc : 0.f, h0.y > 0.f ? acc[i][1] * sc : 0.f,
                                 h0.z > 0.f ? acc[i][2] * sc : 0.f, h0.w > 0.f ? acc[i][3] * sc : 0.f);
                v1 = make_float4(h1.x > 0.f ? acc[i][4] * sc : 0.f, h1.y > 0.f ? acc[i][5] * sc : 0.f,
                                 h1.z > 0.f ? acc[i][6] * sc : 0.f, h1.w > 0.f ? acc[i][7] * sc : 0.f);
                st4(dhpre + (size_t)gr * D + f.c0(), v0);
                st4(dhpre + (size_t)gr * D + f.c1(), v1);
            }
            st4(T1 + r * LDA + f.c0(), v0);
            st4(T1 + r * LDA + f.c1(), v1);
        }
    }
    // dy = dhpre W1 + g   -> T0
    tile_gemm<false>(T1, W1, Ws, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = f.row(i), gr = row0 + r;
        float4 v0 = f4(acc[i], 0), v1 = f4(acc[i], 1);
        if (gr < M) {
            const uint4 tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr);
            v0 = add4(v0, apply_tmask(__ldg(reinterpret_cast<const float4*>(dxo + (size_t)gr * D + f.c0())), tw, f.tn));
            v1 = add4(v1, apply_tmask(__ldg(reinterpret_cast<const float4*>(dxo + (size_t)gr * D + f.c1())), tw, 16 + f.tn));
        }
        st4(T0 + r * LDA + f.c0(), v0);
        st4(T0 + r * LDA + f.c1(), v1);
    }
    __syncthreads();
    // LN2 backward (warp per row): dx1 -> T1 and global
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(ln2_w) + lane);
        float4 dw = make_float4(0, 0, 0, 0), db = make_float4(0, 0, 0, 0);
        for (int r = warp; r < TM; r += NT / 32) {
            const int gr = row0 + r;
            float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < M) {  // warp-uniform
                const float4 g = ld4(T0 + r * LDA + lane * 4);
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)gr * D) + lane);
                const float mean = st2[(size_t)gr * 2], rstd = st2[(size_t)gr * 2 + 1];
                const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                const float4 gw = mul4(g, wv);
                const float c1 = warp_sum(sum4(gw)) * (1.0f / D);
                const float c2 = warp_sum(sum4(mul4(gw, xh))) * (1.0f / D);
                dx = make_float4(rstd * (gw.x - c1 - xh.x * c2), rstd * (gw.y - c1 - xh.y * c2),
                                 rstd * (gw.z - c1 - xh.z * c2), rstd * (gw.w - c1 - xh.w * c2));
                st4(dx1 + (size_t)gr * D + lane * 4, dx);
                dw = add4(dw, mul4(g, xh));
                db = add4(db, g);
            }
            st4(T1 + r * LDA + lane * 4, dx);
        }
        // cross-warp reduction of the LN parameter partials through the (idle) W buffer
        st4(Ws + warp * 2 * D + lane * 4, dw);
        st4(Ws + warp * 2 * D + D + lane * 4, db);
        __syncthreads();
        float s = 0.f;
#pragma unroll
        for (int wv_ = 0; wv_ < NT / 32; ++wv_) s += Ws[wv_ * 2 * D + threadIdx.x];
        ln_part[(size_t)blockIdx.x * 2 * D + threadIdx.x] = s;
        __syncthreads();
    }
    // dO = dx1 Wo
    tile_gemm<false>(T1, Wo, Ws, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gr = row0 + f.row(i);
        if (gr < M) {
            st4(dO + (size_t)gr * D + f.c0(), f4(acc[i], 0));
            st4(dO + (size_t)gr * D + f.c1(), f4(acc[i], 1));
        }
    }
}

// ----------------------------------------------------------------------------------
// backward kernel: QKV projections + LN1
//   dQn = dx1 + dq Wq ; dx_in = dk Wk + dv Wv + LN1bwd(dQn)
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_qkv_bwd(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
          const float* __restrict__ dx1, const float* __restrict__ xin, const float* __restrict__ st1, int M,
          const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
          const float* __restrict__ ln1_w, float* __restrict__ dxin, float* __restrict__ ln_part) {
    extern __shared__ __align__(16) float smem[];
    float* T0 = smem;
    float* T1 = smem + TILE_FLOATS;
    float* Ws = smem + 2 * TILE_FLOATS;
    const int row0 = blockIdx.x * TM;
    Frag f;
    float acc[8][8];
    load_tile(T0, dq, row0, M);
    __syncthreads();
    tile_gemm<false>(T0, Wq, Ws, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // T1 = dQn
        int r = f.row(i), gr = row0 + r;
        float4 v0 = f4(acc[i], 0), v1 = f4(acc[i], 1);
        if (gr < M) {
            v0 = add4(v0, __ldg(reinterpret_cast<const float4*>(dx1 + (size_t)gr * D + f.c0())));
            v1 = add4(v1, __ldg(reinterpret_cast<const float4*>(dx1 + (size_t)gr * D + f.c1())));
        }
        st4(T1 + r * LDA + f.c0(), v0);
        st4(T1 + r * LDA + f.c1(), v1);
    }
    load_tile(T0, dk, row0, M);   // T0 is free: tile_gemm ended with a barrier
    __syncthreads();
    tile_gemm<false>(T0, Wk, Ws, acc);
    load_tile(T0, dv, row0, M);
    __syncthreads();
    tile_gemm<true>(T0, Wv, Ws, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // T0 = dk Wk + dv Wv
        int r = f.row(i);
        st4(T0 + r * LDA + f.c0(), f4(acc[i], 0));
        st4(T0 + r * LDA + f.c1(), f4(acc[i], 1));
    }
    __syncthreads();
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(ln1_w) + lane);
        float4 dw = make_float4(0, 0, 0, 0), db = make_float4(0, 0, 0, 0);
        for (int r = warp; r < TM; r += NT / 32) {
            const int gr = row0 + r;
            if (gr >= M) break;
            const float4 g = ld4(T1 + r * LDA + lane * 4);
            const float4 xv = __ldg(reinterpret_cast<const float4*>(xin + (size_t)gr * D) + lane);
            const float mean = st1[(size_t)gr * 2], rstd = st1[(size_t)gr * 2 + 1];
            const float4 xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            const float4 gw = mul4(g, wv);
            const float c1 = warp_sum(sum4(gw)) * (1.0f / D);
            const float c2 = warp_sum(sum4(mul4(gw, xh))) * (1.0f / D);
            const float4 base = ld4(T0 + r * LDA + lane * 4);
            st4(dxin + (size_t)gr * D + lane * 4,
                make_float4(base.x + rstd * (gw.x - c1 - xh.x * c2), base.y + rstd * (gw.y - c1 - xh.y * c2),
                            base.z + rstd * (gw.z - c1 - xh.z * c2), base.w + rstd * (gw.w - c1 - xh.w * c2)));
            dw = add4(dw, mul4(g, xh));
            db = add4(db, g);
        }
        st4(Ws + warp * 2 * D + lane * 4, dw);
        st4(Ws + warp * 2 * D + D + lane * 4, db);
        __syncthreads();
        float s = 0.f;
#pragma unroll
        for (int wv_ = 0; wv_ < NT / 32; ++wv_) s += Ws[wv_ * 2 * D + threadIdx.x];
        ln_part[(size_t)blockIdx.x * 2 * D + threadIdx.x] = s;
    }
}

// ----------------------------------------------------------------------------------
// weight gradients: dW[n][k] = sum_m dY[m][n] X[m][k], db[n] = sum_m dY[m][n]
// split over row chunks (blockIdx.x) -> partials, reduced in fixed order afterwards.
// ----------------------------------------------------------------------------------
struct WgradJobs {
    const float* dY[6];
    const float* X[6];
};
constexpr int WG_ROWS = 32;
constexpr size_t WG_SMEM_BYTES = (size_t)4 * WG_ROWS * D * sizeof(float);  // 2 stages x (Y,X)

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(NT, 2)
k_wgrad(WgradJobs jobs, int M, int rows_per_cta, float* __restrict__ wpart /*[jobs][S][128*128]*/,
        float* __restrict__ bpart /*[jobs][S][128]*/) {
    extern __shared__ __align__(16) float smem[];
    const float* __restrict__ dY = jobs.dY[blockIdx.y];
    const float* __restrict__ X = jobs.X[blockIdx.y];
    const int S = gridDim.x;
    const int m0 = blockIdx.x * rows_per_cta;
    const int m1 = min(M, m0 + rows_per_cta);
    Frag f;
    float acc[8][8], bs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bs[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const int nch = (max(m1 - m0, 0) + WG_ROWS - 1) / WG_ROWS;
    auto issue = [&](int ch, int st) {
        float* Ys = smem + st * 2 * WG_ROWS * D;
        float* Xs = Ys + WG_ROWS * D;
        const int r0 = m0 + ch * WG_ROWS;
        for (int idx = threadIdx.x; idx < WG_ROWS * (D / 4); idx += NT) {
            int r = idx >> 5, c4 = idx & 31;
            bool ok = r0 + r < m1;
            size_t go = ok ? ((size_t)(r0 + r) * D + c4 * 4) : 0;
            cp_async16_zfill(Ys + r * D + c4 * 4, dY + go, ok);
            cp_async16_zfill(Xs + r * D + c4 * 4, X + go, ok);
        }
        cp_async_commit();
    };
    if (nch > 0) issue(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            issue(ch + 1, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* Ys = smem + (ch & 1) * 2 * WG_ROWS * D;
        const float* Xs = Ys + WG_ROWS * D;
#pragma unroll 4
        for (int mm = 0; mm < WG_ROWS; ++mm) {
            const float4 a0 = ld4(Ys + mm * D + f.tm * 8), a1 = ld4(Ys + mm * D + f.tm * 8 + 4);
            const float4 b0 = ld4(Xs + mm * D + f.c0()), b1 = ld4(Xs + mm * D + f.c1());
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bs[i] += a[i];
                acc[i][0] = fmaf(a[i], b0.x, acc[i][0]); acc[i][1] = fmaf(a[i], b0.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b0.z, acc[i][2]); acc[i][3] = fmaf(a[i], b0.w, acc[i][3]);
                acc[i][4] = fmaf(a[i], b1.x, acc[i][4]); acc[i][5] = fmaf(a[i], b1.y, acc[i][5]);
                acc[i][6] = fmaf(a[i], b1.z, acc[i][6]); acc[i][7] = fmaf(a[i], b1.w, acc[i][7]);
            }
        }
        __syncthreads();
    }
    float* wp = wpart + ((size_t)blockIdx.y * S + blockIdx.x) * D * D;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        st4(wp + (size_t)f.row(i) * D + f.c0(), f4(acc[i], 0));
        st4(wp + (size_t)f.row(i) * D + f.c1(), f4(acc[i], 1));
    }
    if (f.tn == 0) {
        float* bp = bpart + ((size_t)blockIdx.y * S + blockIdx.x) * D;
#pragma unroll
        for (int i = 0; i < 8; ++i) bp[f.row(i)] = bs[i];
    }
}

// out[j][e] = sum_s part[j][s][e]   (fixed order => deterministic)
struct ReduceJobs {
    float* out[8];
};
__global__ void k_reduce_partials(const float* __restrict__ part, int S, int n, ReduceJobs outs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float* p = part + (size_t)blockIdx.y * S * n + e;
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < S; ++i) s += p[(size_t)i * n];
    outs.out[blockIdx.y][e] = s;
}
// LN partials are [S][256] = (dw[128], db[128]) per tile.  16 CTAs x 1024 threads: a CTA owns 16
// columns, its 64 row groups sum their rows in order, then the group sums are added in fixed order.
__global__ void __launch_bounds__(1024)
k_reduce_ln(const float* __restrict__ part, int S, float* __restrict__ dw, float* __restrict__ db) {
    // 16 columns x 64 row groups per CTA (1024 threads): short dependent chains, fixed summation order
    __shared__ float red[64][16];
    const int c = threadIdx.x & 15, g = threadIdx.x >> 4;
    const int e = blockIdx.x * 16 + c;
    float s = 0.f;
#pragma unroll 8
    for (int i = g; i < S; i += 64) s += part[(size_t)i * 2 * D + e];
    red[g][c] = s;
    __syncthreads();
    if (g == 0) {
        float t = red[0][c];
#pragma unroll
        for (int k = 1; k < 64; ++k) t += red[k][c];
        if (e < D) dw[e] = t; else db[e - D] = t;
    }
}

// ----------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------
static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
static int ensure_smem(const void* fn, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return 0;
}

static int check_encoder_args(int B, int L, int b_off = 0) {
    AMID_REQUIRE(B > 0 && L > 0, "encoder: B=%d L=%d must be positive", B, L);
    AMID_REQUIRE(L <= 512, "encoder: L=%d > 512 unsupported (per-CTA shared-memory attention)", L);
    AMID_REQUIRE((int64_t)B * L < (1ll << 31) / D, "encoder: B*L too large");
    // the tensor-core attention kernels index the dropout stream of a site with 32 bits
    AMID_REQUIRE((int64_t)(B + b_off) * 8 * L * ((L + 3) / 4) < (1ll << 32),
                 "encoder: (batch_offset+B)*8*L*ceil(L/4) must be below 2^32");
    return 0;
}

static int wgrad_chunks(int M, int* rows_per_cta) {
    int S = (M + TM - 1) / TM;
    if (S > 74) S = 74;
    int rp = (int)round_up((M + S - 1) / S, WG_ROWS);
    S = (M + rp - 1) / rp;
    *rows_per_cta = rp;
    return S;
}

}  // namespace amid

using namespace amid;

extern "C" int64_t amid_encoder_fwd_workspace_bytes(int32_t, int32_t) { return (int64_t)(12 * D * D + 64) * sizeof(float); }

static int encoder_fwd_impl(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                            int32_t L, const amid_dropout* drop, amid_encoder_saved* S, float* enc_out,
                            void* workspace, int64_t workspace_bytes, amid_stream_t stream_, int mode) {
    const bool use_tc = mode != 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_encoder_args(B, L, drop ? drop->batch_offset : 0)) return rc;
    AMID_REQUIRE(P && S && x0 && tmask && enc_out && workspace, "encoder_fwd: null argument");
    AMID_REQUIRE(workspace_bytes >= amid_encoder_fwd_workspace_bytes(B, L), "encoder_fwd: workspace too small");
    AMID_REQUIRE(aligned16(x0) && aligned16(enc_out) && aligned16(workspace) && aligned16(tmask), "encoder_fwd: misaligned buffer");
    const int M = B * L;
    const int tiles = (M + TM - 1) / TM;
    const DropCfg dc = with_offsets(make_drop(drop), L);
    const size_t attn_smem = (size_t)2 * L * DH * sizeof(float);
    if (int rc = ensure_smem((const void*)k_attn_fwd, attn_smem)) return rc;
    const int attn_threads = (int)round_up((L + 1) / 2, 32);
    if (mode == 3) {   // split-operand fp32-accurate path (x3.cuh): FP16-pair weight images, token tile in tensor memory
        if (int rc = ensure_smem((const void*)x3::k_ln_qkv_x3, x3::CHAINX_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)x3::k_proj_ffn_x3, x3::CHAINX_SMEM)) return rc;
        const size_t mma_smem = (size_t)2 * ((L + 15) / 16 * 16) * attn::LDS * sizeof(float);
        if (int rc = ensure_smem((const void*)attn::k_attn_fwd_mma<true>, mma_smem)) return rc;
        uint8_t* img = (uint8_t*)workspace;
        float* winv = (float*)(img + (size_t)12 * x3::WIMG_BYTES);
        x3::PrepJobsX pj;
        for (int i = 0; i < 2; ++i) {
            pj.src[i * 6 + 0] = P->in_w[i];
            pj.src[i * 6 + 1] = P->in_w[i] + D * D;
            pj.src[i * 6 + 2] = P->in_w[i] + 2 * D * D;
            pj.src[i * 6 + 3] = P->out_w[i];
            pj.src[i * 6 + 4] = P->c1_w[i];
            pj.src[i * 6 + 5] = P->c2_w[i];
        }
        AMID_K("k_prep_wx3", stream);
        x3::k_prep_wx3<<<12, 256, 0, stream>>>(pj, img, winv, 0);
        AMID_LAUNCH_CHECK("k_prep_wx3");
        const float* xin = x0;
        for (int i = 0; i < 2; ++i) {
            const uint8_t* W = img + (size_t)i * 6 * x3::WIMG_BYTES;
            const float* wi = winv + i * 6;
            AMID_K("k_ln_qkv_x3", stream);
            x3::k_ln_qkv_x3<<<tiles, 256, x3::CHAINX_SMEM, stream>>>(xin, M, P->ln1_w[i], P->ln1_b[i], W, W + x3::WIMG_BYTES,
                                                                    W + 2 * x3::WIMG_BYTES, wi, P->in_b[i], S->qn[i], S->st1[i],
                                                                    S->q[i], S->k[i], S->v[i]);
            AMID_LAUNCH_CHECK("k_ln_qkv_x3");
            if (L >= attn_p::PMINL && L <= attn_tc::MAXL) {      // scores in tensor memory
                auto kfn = dc.train ? attn_p::k_attn_fwd_p<true> : attn_p::k_attn_fwd_p<false>;
                if (int rc = ensure_smem((const void*)kfn, attn_p::PFWD_SMEM)) return rc;
                AMID_K("k_attn_fwd_p", stream);
                kfn<<<B * H, 256, attn_p::PFWD_SMEM, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], L, dc, dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_fwd_p");
            } else {
                AMID_K("k_attn_fwd_mma3", stream);
                attn::k_attn_fwd_mma<true><<<B * H, attn::NW * 32, mma_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i],
                                                                                       L, dc, dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_fwd_mma3");
            }
            const bool last = i == 1;
            AMID_K("k_proj_ffn_x3", stream);
            x3::k_proj_ffn_x3<<<tiles, 256, x3::CHAINX_SMEM, stream>>>(
                S->o[i], S->qn[i], M, W + 3 * x3::WIMG_BYTES, W + 4 * x3::WIMG_BYTES, W + 5 * x3::WIMG_BYTES, wi + 3, P->out_b[i],
                P->ln2_w[i], P->ln2_b[i], P->c1_b[i], P->c2_b[i], tmask, dc, dc.site_base + site_ffn1(i),
                dc.site_base + site_ffn2(i), S->x1[i], S->st2[i], S->y[i], S->h[i], S->xout[i], last ? P->ln3_w : nullptr,
                last ? P->ln3_b : nullptr, last ? enc_out : nullptr, last ? S->st3 : nullptr);
            AMID_LAUNCH_CHECK("k_proj_ffn_x3");
            xin = S->xout[i];
        }
        return 0;
    }
    if (mode == 2) {   // BF16 operands: convert the 12 weight matrices once, then 2 CTAs/SM chain kernels
        if (int rc = ensure_smem((const void*)tc16::k_ln_qkv_16, tc16::CHAIN16_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tc16::k_proj_ffn_16, tc16::CHAIN16_SMEM)) return rc;
        const size_t mma_smem = (size_t)2 * ((L + 15) / 16 * 16) * attn::LDS * sizeof(float);
        if (int rc = ensure_smem((const void*)attn::k_attn_fwd_mma<false>, mma_smem)) return rc;
        uint16_t* w16 = (uint16_t*)workspace;
        tc16::PrepJobs pj;
        for (int i = 0; i < 2; ++i) {
            pj.src[i * 6 + 0] = P->in_w[i];
            pj.src[i * 6 + 1] = P->in_w[i] + D * D;
            pj.src[i * 6 + 2] = P->in_w[i] + 2 * D * D;
            pj.src[i * 6 + 3] = P->out_w[i];
            pj.src[i * 6 + 4] = P->c1_w[i];
            pj.src[i * 6 + 5] = P->c2_w[i];
        }
        AMID_K("k_prep_w16", stream);
        tc16::k_prep_w16<<<dim3(4, 4, 12), dim3(32, 8), 0, stream>>>(pj, w16, 0);
        AMID_LAUNCH_CHECK("k_prep_w16");
        const float* xin = x0;
        for (int i = 0; i < 2; ++i) {
            const uint16_t* W = w16 + (size_t)i * 6 * D * D;
            AMID_K("k_ln_qkv_16", stream);
            tc16::k_ln_qkv_16<<<tiles, 256, tc16::CHAIN16_SMEM, stream>>>(xin, M, P->ln1_w[i], P->ln1_b[i], W, W + D * D,
                                                                         W + 2 * D * D, P->in_b[i], S->qn[i], S->st1[i],
                                                                         S->q[i], S->k[i], S->v[i]);
            AMID_LAUNCH_CHECK("k_ln_qkv_16");
            AMID_K("k_attn_fwd_mma", stream);
            attn::k_attn_fwd_mma<false><<<B * H, attn::NW * 32, mma_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], L,
                                                                             dc, dc.site_base + site_attn(i));
            AMID_LAUNCH_CHECK("k_attn_fwd_mma");
            const bool last = i == 1;
            AMID_K("k_proj_ffn_16", stream);
            tc16::k_proj_ffn_16<<<tiles, 256, tc16::CHAIN16_SMEM, stream>>>(
                S->o[i], S->qn[i], M, W + 3 * D * D, P->out_b[i], P->ln2_w[i], P->ln2_b[i], W + 4 * D * D, P->c1_b[i],
                W + 5 * D * D, P->c2_b[i], tmask, dc, dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), S->x1[i],
                S->st2[i], S->y[i], S->h[i], S->xout[i], last ? P->ln3_w : nullptr, last ? P->ln3_b : nullptr,
                last ? enc_out : nullptr, last ? S->st3 : nullptr);
            AMID_LAUNCH_CHECK("k_proj_ffn_16");
            xin = S->xout[i];
        }
        return 0;
    }
    if (use_tc) {   // tcgen05 path: the weights are consumed K-major in their natural [out][in] layout
        if (int rc = ensure_smem((const void*)tcenc::k_ln_qkv_tc, tcenc::CHAIN_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tcenc::k_proj_ffn_tc, tcenc::CHAIN_SMEM)) return rc;
        const size_t mma_smem = (size_t)2 * ((L + 15) / 16 * 16) * attn::LDS * sizeof(float);
        if (int rc = ensure_smem((const void*)attn::k_attn_fwd_mma<false>, mma_smem)) return rc;
        const float* xin = x0;
        for (int i = 0; i < 2; ++i) {
            AMID_K("k_ln_qkv_tc", stream);
            tcenc::k_ln_qkv_tc<<<tiles, 256, tcenc::CHAIN_SMEM, stream>>>(
                xin, M, P->ln1_w[i], P->ln1_b[i], P->in_w[i], P->in_w[i] + D * D, P->in_w[i] + 2 * D * D, P->in_b[i],
                S->qn[i], S->st1[i], S->q[i], S->k[i], S->v[i]);
            AMID_LAUNCH_CHECK("k_ln_qkv_tc");
            AMID_K("k_attn_fwd_mma", stream);
            attn::k_attn_fwd_mma<false><<<B * H, attn::NW * 32, mma_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], L,
                                                                             dc, dc.site_base + site_attn(i));
            AMID_LAUNCH_CHECK("k_attn_fwd_mma");
            const bool last = i == 1;
            AMID_K("k_proj_ffn_tc", stream);
            tcenc::k_proj_ffn_tc<<<tiles, 256, tcenc::CHAIN_SMEM, stream>>>(
                S->o[i], S->qn[i], M, P->out_w[i], P->out_b[i], P->ln2_w[i], P->ln2_b[i], P->c1_w[i], P->c1_b[i],
                P->c2_w[i], P->c2_b[i], tmask, dc, dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), S->x1[i],
                S->st2[i], S->y[i], S->h[i], S->xout[i], last ? P->ln3_w : nullptr, last ? P->ln3_b : nullptr,
                last ? enc_out : nullptr, last ? S->st3 : nullptr);
            AMID_LAUNCH_CHECK("k_proj_ffn_tc");
            xin = S->xout[i];
        }
        return 0;
    }
    float* wt = (float*)workspace;
    TransJobs tj;
    for (int i = 0; i < 2; ++i) {
        tj.src[i * 6 + 0] = P->in_w[i];
        tj.src[i * 6 + 1] = P->in_w[i] + D * D;
        tj.src[i * 6 + 2] = P->in_w[i] + 2 * D * D;
        tj.src[i * 6 + 3] = P->out_w[i];
        tj.src[i * 6 + 4] = P->c1_w[i];
        tj.src[i * 6 + 5] = P->c2_w[i];
    }
    AMID_K("k_transpose128", stream);
    k_transpose128<<<dim3(4, 4, 12), dim3(32, 8), 0, stream>>>(tj, wt);
    AMID_LAUNCH_CHECK("k_transpose128");
    if (int rc = ensure_smem((const void*)k_ln_qkv, ENC_SMEM_BYTES)) return rc;
    if (int rc = ensure_smem((const void*)k_proj_ffn, ENC_SMEM_BYTES)) return rc;
    const float* xin = x0;
    for (int i = 0; i < 2; ++i) {
        const float* W = wt + (size_t)i * 6 * D * D;
        AMID_K("k_ln_qkv", stream);
        k_ln_qkv<<<tiles, NT, ENC_SMEM_BYTES, stream>>>(xin, M, P->ln1_w[i], P->ln1_b[i], W, W + D * D, W + 2 * D * D,
                                                        P->in_b[i], S->qn[i], S->st1[i], S->q[i], S->k[i], S->v[i]);
        AMID_LAUNCH_CHECK("k_ln_qkv");
        AMID_K("k_attn_fwd", stream);
        k_attn_fwd<<<B * H, attn_threads, attn_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], L, dc,
                                                               dc.site_base + site_attn(i));
        AMID_LAUNCH_CHECK("k_attn_fwd");
        const bool last = i == 1;
        AMID_K("k_proj_ffn", stream);
        k_proj_ffn<<<tiles, NT, ENC_SMEM_BYTES, stream>>>(
            S->o[i], S->qn[i], M, W + 3 * D * D, P->out_b[i], P->ln2_w[i], P->ln2_b[i], W + 4 * D * D, P->c1_b[i],
            W + 5 * D * D, P->c2_b[i], tmask, dc, dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), S->x1[i],
            S->st2[i], S->y[i], S->h[i], S->xout[i], last ? P->ln3_w : nullptr, last ? P->ln3_b : nullptr,
            last ? enc_out : nullptr, last ? S->st3 : nullptr);
        AMID_LAUNCH_CHECK("k_proj_ffn");
        xin = S->xout[i];
    }
    return 0;
}

extern "C" int amid_encoder_fwd(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                int32_t L, const amid_dropout* drop, amid_encoder_saved* S, float* enc_out,
                                void* workspace, int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_fwd_impl(P, x0, tmask, B, L, drop, S, enc_out, workspace, workspace_bytes, stream, 0);
}
extern "C" int amid_encoder_fwd_tc(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                   int32_t L, const amid_dropout* drop, amid_encoder_saved* S, float* enc_out,
                                   void* workspace, int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_fwd_impl(P, x0, tmask, B, L, drop, S, enc_out, workspace, workspace_bytes, stream, 1);
}
extern "C" int amid_encoder_fwd_bf16(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                     int32_t L, const amid_dropout* drop, amid_encoder_saved* S, float* enc_out,
                                     void* workspace, int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_fwd_impl(P, x0, tmask, B, L, drop, S, enc_out, workspace, workspace_bytes, stream, 2);
}

extern "C" int amid_encoder_fwd_x3(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                   int32_t L, const amid_dropout* drop, amid_encoder_saved* S, float* enc_out,
                                   void* workspace, int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_fwd_impl(P, x0, tmask, B, L, drop, S, enc_out, workspace, workspace_bytes, stream, 3);
}

constexpr int WG_TC_S = 24;   // CTAs per weight-gradient job on the tensor-core path (6 jobs -> 144 CTAs)

extern "C" int64_t amid_encoder_bwd_workspace_bytes(int32_t B, int32_t L) {
    const int64_t M = (int64_t)B * L;
    const int64_t tiles = (M + TM - 1) / TM;
    int rp;
    int S = wgrad_chunks((int)M, &rp);
    if (S < 2 * WG_TC_S) S = 2 * WG_TC_S;
    int64_t fl = 9 * M * D;                  // dxa, dxb, do2, dhpre, dx1, dO, dq, dk, dv
    fl += (int64_t)6 * S * (D * D + D);      // weight / bias partials
    fl += 2 * tiles * 2 * D;                 // LN partials (two in flight)
    fl += (int64_t)12 * D * D + 64;          // transposed weights / weight images + inverse scales (tensor-core paths)
    return fl * (int64_t)sizeof(float) + 256;
}

static int encoder_bwd_impl(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                            int32_t L, const amid_dropout* drop, const amid_encoder_saved* S, const float* enc_out,
                            const float* d_enc, amid_encoder_tensors* G, float* dx0, void* workspace,
                            int64_t workspace_bytes, amid_stream_t stream_, int mode) {
    const bool use_tc = mode != 0;
    (void)enc_out;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_encoder_args(B, L, drop ? drop->batch_offset : 0)) return rc;
    AMID_REQUIRE(P && S && G && x0 && tmask && d_enc && dx0 && workspace, "encoder_bwd: null argument");
    AMID_REQUIRE(workspace_bytes >= amid_encoder_bwd_workspace_bytes(B, L), "encoder_bwd: workspace too small");
    AMID_REQUIRE(aligned16(d_enc) && aligned16(dx0) && aligned16(workspace), "encoder_bwd: misaligned buffer");
    const int M = B * L;
    const int tiles = (M + TM - 1) / TM;
    const DropCfg dc = with_offsets(make_drop(drop), L);
    int rp;
    int SW = wgrad_chunks(M, &rp);
    const int SWmax = SW < 2 * WG_TC_S ? 2 * WG_TC_S : SW;
    if (use_tc) SW = tiles < WG_TC_S ? tiles : WG_TC_S;
    if (mode == 2) SW = tiles < 2 * WG_TC_S ? tiles : 2 * WG_TC_S;   // 2 CTAs/SM -> 48 x 6 CTAs
    if (mode == 3) SW = tiles < WG_TC_S ? tiles : WG_TC_S;           // 1 CTA/SM (192 KB of piece buffers) -> 24 x 6 CTAs
    float* w = (float*)workspace;
    const size_t MD = (size_t)M * D;
    float* dxa = w;            // gradient of the current block output
    float* dxb = w + MD;       // gradient of the current block input
    float* do2 = w + 2 * MD;
    float* dhp = w + 3 * MD;
    float* dx1 = w + 4 * MD;
    float* dO = w + 5 * MD;
    float* dq = w + 6 * MD;
    float* dk = w + 7 * MD;
    float* dv = w + 8 * MD;
    float* wpart = w + 9 * MD;
    float* bpart = wpart + (size_t)6 * SWmax * D * D;
    float* lnp0 = bpart + (size_t)6 * SWmax * D;
    float* lnp1 = lnp0 + (size_t)tiles * 2 * D;
    float* wtr = lnp1 + (size_t)tiles * 2 * D;   // 12 transposed weights: per block W2t, W1t, Wot, Wqt, Wkt, Wvt
    uint8_t* ximg = (uint8_t*)wtr;
    float* xwinv = (float*)(ximg + (size_t)12 * x3::WIMG_BYTES);
    if (mode == 3) {
        if (int rc = ensure_smem((const void*)x3::k_ffn_bwd_x3, x3::CHAINX_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)x3::k_qkv_bwd_x3, x3::CHAINX_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)x3::k_wgrad_x3, x3::WGRADX_SMEM)) return rc;
        x3::PrepJobsX pj;
        for (int i = 0; i < 2; ++i) {
            pj.src[i * 6 + 0] = P->c2_w[i];
            pj.src[i * 6 + 1] = P->c1_w[i];
            pj.src[i * 6 + 2] = P->out_w[i];
            pj.src[i * 6 + 3] = P->in_w[i];
            pj.src[i * 6 + 4] = P->in_w[i] + D * D;
            pj.src[i * 6 + 5] = P->in_w[i] + 2 * D * D;
        }
        AMID_K("k_prep_wx3", stream);
        x3::k_prep_wx3<<<12, 256, 0, stream>>>(pj, ximg, xwinv, 1);
        AMID_LAUNCH_CHECK("k_prep_wx3");
    } else if (mode == 2) {
        if (int rc = ensure_smem((const void*)tc16::k_ffn_bwd_16, tc16::CHAIN16_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tc16::k_qkv_bwd_16, tc16::CHAIN16_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tc16::k_wgrad_16, tc16::WGRAD16_SMEM)) return rc;
        tc16::PrepJobs pj;
        for (int i = 0; i < 2; ++i) {
            pj.src[i * 6 + 0] = P->c2_w[i];
            pj.src[i * 6 + 1] = P->c1_w[i];
            pj.src[i * 6 + 2] = P->out_w[i];
            pj.src[i * 6 + 3] = P->in_w[i];
            pj.src[i * 6 + 4] = P->in_w[i] + D * D;
            pj.src[i * 6 + 5] = P->in_w[i] + 2 * D * D;
        }
        AMID_K("k_prep_w16", stream);
        tc16::k_prep_w16<<<dim3(4, 4, 12), dim3(32, 8), 0, stream>>>(pj, (uint16_t*)wtr, 1);
        AMID_LAUNCH_CHECK("k_prep_w16");
    } else if (use_tc) {
        if (int rc = ensure_smem((const void*)tcenc::k_ffn_bwd_tc, tcenc::CHAIN_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tcenc::k_qkv_bwd_tc, tcenc::CHAIN_SMEM)) return rc;
        if (int rc = ensure_smem((const void*)tcenc::k_wgrad_tc, tcenc::WGRAD_SMEM)) return rc;
        TransJobs tj;
        for (int i = 0; i < 2; ++i) {
            tj.src[i * 6 + 0] = P->c2_w[i];
            tj.src[i * 6 + 1] = P->c1_w[i];
            tj.src[i * 6 + 2] = P->out_w[i];
            tj.src[i * 6 + 3] = P->in_w[i];
            tj.src[i * 6 + 4] = P->in_w[i] + D * D;
            tj.src[i * 6 + 5] = P->in_w[i] + 2 * D * D;
        }
        AMID_K("k_transpose128", stream);
        k_transpose128<<<dim3(4, 4, 12), dim3(32, 8), 0, stream>>>(tj, wtr);
        AMID_LAUNCH_CHECK("k_transpose128");
    }

    if (int rc = ensure_smem((const void*)k_ffn_bwd, ENC_SMEM_BYTES)) return rc;
    if (int rc = ensure_smem((const void*)k_qkv_bwd, ENC_SMEM_BYTES)) return rc;
    if (int rc = ensure_smem((const void*)k_wgrad, WG_SMEM_BYTES)) return rc;
    const size_t attn_smem = (size_t)(4 * L * DH + 2 * L) * sizeof(float);
    if (int rc = ensure_smem((const void*)k_attn_bwd, attn_smem)) return rc;
    const int attn_threads = (int)round_up((L + 1) / 2, 32);

    // last LayerNorm
    AMID_K("k_ln_bwd", stream);
    k_ln_bwd<<<tiles, NT, 0, stream>>>(d_enc, S->xout[1], S->st3, P->ln3_w, M, dxa, lnp0);
    AMID_LAUNCH_CHECK("k_ln_bwd");
    AMID_K("k_reduce_ln", stream);
    k_reduce_ln<<<16, 1024, 0, stream>>>(lnp0, tiles, G->ln3_w, G->ln3_b);
    AMID_LAUNCH_CHECK("k_reduce_ln");

    for (int i = 1; i >= 0; --i) {
        const float* xin = i == 0 ? x0 : S->xout[0];
        float* dxin = i == 0 ? dx0 : dxb;
        const float* Wt = wtr + (size_t)i * 6 * D * D;
        if (mode == 3) {
            const uint8_t* W = ximg + (size_t)i * 6 * x3::WIMG_BYTES;
            AMID_K("k_ffn_bwd_x3", stream);
            x3::k_ffn_bwd_x3<<<tiles, 256, x3::CHAINX_SMEM, stream>>>(
                dxa, S->h[i], S->x1[i], S->st2[i], tmask, M, W, W + x3::WIMG_BYTES, W + 2 * x3::WIMG_BYTES, xwinv + i * 6,
                P->ln2_w[i], dc, dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), do2, dhp, dx1, dO, lnp0);
            AMID_LAUNCH_CHECK("k_ffn_bwd_x3");
        } else if (mode == 2) {
            const uint16_t* W16 = (const uint16_t*)wtr + (size_t)i * 6 * D * D;
            AMID_K("k_ffn_bwd_16", stream);
            tc16::k_ffn_bwd_16<<<tiles, 256, tc16::CHAIN16_SMEM, stream>>>(
                dxa, S->h[i], S->x1[i], S->st2[i], tmask, M, W16, W16 + D * D, W16 + 2 * D * D, P->ln2_w[i], dc,
                dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), do2, dhp, dx1, dO, lnp0);
            AMID_LAUNCH_CHECK("k_ffn_bwd_16");
        } else if (use_tc) {
            AMID_K("k_ffn_bwd_tc", stream);
            tcenc::k_ffn_bwd_tc<<<tiles, 256, tcenc::CHAIN_SMEM, stream>>>(
                dxa, S->h[i], S->x1[i], S->st2[i], tmask, M, Wt, Wt + D * D, Wt + 2 * D * D, P->ln2_w[i], dc,
                dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), do2, dhp, dx1, dO, lnp0);
            AMID_LAUNCH_CHECK("k_ffn_bwd_tc");
        } else {
        AMID_K("k_ffn_bwd", stream);
        k_ffn_bwd<<<tiles, NT, ENC_SMEM_BYTES, stream>>>(dxa, S->h[i], S->x1[i], S->st2[i], tmask, M, P->c2_w[i],
                                                         P->c1_w[i], P->out_w[i], P->ln2_w[i], dc,
                                                         dc.site_base + site_ffn1(i), dc.site_base + site_ffn2(i), do2,
                                                         dhp, dx1, dO, lnp0);
        AMID_LAUNCH_CHECK("k_ffn_bwd");
        }
        AMID_K("k_reduce_ln", stream);
        k_reduce_ln<<<16, 1024, 0, stream>>>(lnp0, tiles, G->ln2_w[i], G->ln2_b[i]);
        AMID_LAUNCH_CHECK("k_reduce_ln");
        if (mode == 3) {
            const size_t mma_smem = (size_t)((L + 15) / 16 * 16) * (4 * attn::LDS + 2) * sizeof(float);
            if (int rc = ensure_smem((const void*)attn::k_attn_bwd_mma<true>, mma_smem)) return rc;
            if (L >= attn_tc::MINL && L <= attn_tc::MAXL) {       // two-pass kernel, two CTAs per SM (741 us at the C3 shape)
                auto kfn = dc.train ? attn_p::k_attn_bwd_t2<true> : attn_p::k_attn_bwd_t2<false>;
                if (int rc = ensure_smem((const void*)kfn, attn_p::T2_SMEM)) return rc;
                AMID_K("k_attn_bwd_t2", stream);
                kfn<<<B * H, 256, attn_p::T2_SMEM, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], dO, dq, dk, dv, L, dc,
                                                             dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_bwd_t2");
            } else if (L >= attn_p::PMINL && L <= attn_p::PMAXL) {       // single-pass pipelined kernel, persistent CTAs (805 us)
                auto kfn = dc.train ? attn_p::k_attn_bwd_p<true> : attn_p::k_attn_bwd_p<false>;
                if (int rc = ensure_smem((const void*)kfn, attn_p::PBWD_SMEM)) return rc;
                AMID_K("k_attn_bwd_p", stream);
                kfn<<<std::min(B * H, sm_count()), attn_p::NTH, attn_p::PBWD_SMEM, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], dO,
                                                                                             dq, dk, dv, L, B * H, dc,
                                                                                             dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_bwd_p");
            } else if (L >= attn_tc::MINL && L <= attn_tc::MAXL) {
                if (int rc = ensure_smem((const void*)attn_tc::k_attn_bwd_tc, attn_tc::BWD_SMEM)) return rc;
                AMID_K("k_attn_bwd_tc", stream);
                attn_tc::k_attn_bwd_tc<<<B * H, 256, attn_tc::BWD_SMEM, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], dO, dq,
                                                                                  dk, dv, L, dc, dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_bwd_tc");
            } else {
                AMID_K("k_attn_bwd_mma3", stream);
                attn::k_attn_bwd_mma<true><<<B * H, attn::NWB * 32, mma_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i],
                                                                                        dO, dq, dk, dv, L, dc,
                                                                                        dc.site_base + site_attn(i));
                AMID_LAUNCH_CHECK("k_attn_bwd_mma3");
            }
        } else if (use_tc) {
            const size_t mma_smem = (size_t)((L + 15) / 16 * 16) * (4 * attn::LDS + 2) * sizeof(float);
            if (int rc = ensure_smem((const void*)attn::k_attn_bwd_mma<false>, mma_smem)) return rc;
            AMID_K("k_attn_bwd_mma", stream);
            attn::k_attn_bwd_mma<false><<<B * H, attn::NWB * 32, mma_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], dO,
                                                                             dq, dk, dv, L, dc, dc.site_base + site_attn(i));
            AMID_LAUNCH_CHECK("k_attn_bwd_mma");
        } else {
        AMID_K("k_attn_bwd", stream);
        k_attn_bwd<<<B * H, attn_threads, attn_smem, stream>>>(S->q[i], S->k[i], S->v[i], S->o[i], S->lse[i], dO, dq, dk,
                                                               dv, L, dc, dc.site_base + site_attn(i));
        AMID_LAUNCH_CHECK("k_attn_bwd");
        }
        if (mode == 3) {
            const uint8_t* W = ximg + (size_t)(i * 6 + 3) * x3::WIMG_BYTES;
            AMID_K("k_qkv_bwd_x3", stream);
            x3::k_qkv_bwd_x3<<<tiles, 256, x3::CHAINX_SMEM, stream>>>(dq, dk, dv, dx1, xin, S->st1[i], M, W, W + x3::WIMG_BYTES,
                                                                     W + 2 * x3::WIMG_BYTES, xwinv + i * 6 + 3, P->ln1_w[i], dxin,
                                                                     lnp1);
            AMID_LAUNCH_CHECK("k_qkv_bwd_x3");
        } else if (mode == 2) {
            const uint16_t* W16 = (const uint16_t*)wtr + (size_t)i * 6 * D * D;
            AMID_K("k_qkv_bwd_16", stream);
            tc16::k_qkv_bwd_16<<<tiles, 256, tc16::CHAIN16_SMEM, stream>>>(dq, dk, dv, dx1, xin, S->st1[i], M, W16 + 3 * D * D,
                                                                          W16 + 4 * D * D, W16 + 5 * D * D, P->ln1_w[i], dxin,
                                                                          lnp1);
            AMID_LAUNCH_CHECK("k_qkv_bwd_16");
        } else if (use_tc) {
            AMID_K("k_qkv_bwd_tc", stream);
            tcenc::k_qkv_bwd_tc<<<tiles, 256, tcenc::CHAIN_SMEM, stream>>>(dq, dk, dv, dx1, xin, S->st1[i], M, Wt + 3 * D * D,
                                                                          Wt + 4 * D * D, Wt + 5 * D * D, P->ln1_w[i], dxin,
                                                                          lnp1);
            AMID_LAUNCH_CHECK("k_qkv_bwd_tc");
        } else {
        AMID_K("k_qkv_bwd", stream);
        k_qkv_bwd<<<tiles, NT, ENC_SMEM_BYTES, stream>>>(dq, dk, dv, dx1, xin, S->st1[i], M, P->in_w[i],
                                                         P->in_w[i] + D * D, P->in_w[i] + 2 * D * D, P->ln1_w[i], dxin,
                                                         lnp1);
        AMID_LAUNCH_CHECK("k_qkv_bwd");
        }
        AMID_K("k_reduce_ln", stream);
        k_reduce_ln<<<16, 1024, 0, stream>>>(lnp1, tiles, G->ln1_w[i], G->ln1_b[i]);
        AMID_LAUNCH_CHECK("k_reduce_ln");
        // weight gradients of the block, one launch: W2, W1, Wo, Wq, Wk, Wv
        WgradJobs wj;
        wj.dY[0] = do2; wj.X[0] = S->h[i];
        wj.dY[1] = dhp; wj.X[1] = S->y[i];
        wj.dY[2] = dx1; wj.X[2] = S->o[i];
        wj.dY[3] = dq;  wj.X[3] = S->qn[i];
        wj.dY[4] = dk;  wj.X[4] = xin;
        wj.dY[5] = dv;  wj.X[5] = xin;
        if (mode == 3) {
            x3::WgradJobsX wx;
            for (int j = 0; j < 6; ++j) { wx.dY[j] = wj.dY[j]; wx.X[j] = wj.X[j]; }
            AMID_K("k_wgrad_x3", stream);
            x3::k_wgrad_x3<<<dim3(SW, 6), x3::WGX_THREADS, x3::WGRADX_SMEM, stream>>>(wx, M, wpart, bpart);
            AMID_LAUNCH_CHECK("k_wgrad_x3");
        } else if (mode == 2) {
            tc16::WgradJobs16 w16;
            for (int j = 0; j < 6; ++j) { w16.dY[j] = wj.dY[j]; w16.X[j] = wj.X[j]; }
            AMID_K("k_wgrad_16", stream);
            tc16::k_wgrad_16<<<dim3(SW, 6), 256, tc16::WGRAD16_SMEM, stream>>>(w16, M, wpart, bpart);
            AMID_LAUNCH_CHECK("k_wgrad_16");
        } else if (use_tc) {
            tcenc::WgradJobsTc wt6;
            for (int j = 0; j < 6; ++j) { wt6.dY[j] = wj.dY[j]; wt6.X[j] = wj.X[j]; }
            AMID_K("k_wgrad_tc", stream);
            tcenc::k_wgrad_tc<<<dim3(SW, 6), 256, tcenc::WGRAD_SMEM, stream>>>(wt6, M, wpart, bpart);
            AMID_LAUNCH_CHECK("k_wgrad_tc");
        } else {
        AMID_K("k_wgrad", stream);
        k_wgrad<<<dim3(SW, 6), NT, WG_SMEM_BYTES, stream>>>(wj, M, rp, wpart, bpart);
        AMID_LAUNCH_CHECK("k_wgrad");
        }
        ReduceJobs rw, rb;
        rw.out[0] = G->c2_w[i]; rb.out[0] = G->c2_b[i];
        rw.out[1] = G->c1_w[i]; rb.out[1] = G->c1_b[i];
        rw.out[2] = G->out_w[i]; rb.out[2] = G->out_b[i];
        rw.out[3] = G->in_w[i]; rb.out[3] = G->in_b[i];
        rw.out[4] = G->in_w[i] + D * D; rb.out[4] = G->in_b[i] + D;
        rw.out[5] = G->in_w[i] + 2 * D * D; rb.out[5] = G->in_b[i] + 2 * D;
        AMID_K("k_reduce_partials", stream);
        k_reduce_partials<<<dim3(D * D / 256, 6), 256, 0, stream>>>(wpart, SW, D * D, rw);
        AMID_LAUNCH_CHECK("k_reduce_partials(w)");
        AMID_K("k_reduce_partials", stream);
        k_reduce_partials<<<dim3(1, 6), 128, 0, stream>>>(bpart, SW, D, rb);
        AMID_LAUNCH_CHECK("k_reduce_partials(b)");
        // the input gradient of block 1 is the output gradient of block 0
        if (i == 1) {
            float* t = dxa; dxa = dxb; dxb = t;
        }
    }
    return 0;
}

extern "C" int amid_encoder_bwd(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                int32_t L, const amid_dropout* drop, const amid_encoder_saved* S, const float* enc_out,
                                const float* d_enc, amid_encoder_tensors* G, float* dx0, void* workspace,
                                int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_bwd_impl(P, x0, tmask, B, L, drop, S, enc_out, d_enc, G, dx0, workspace, workspace_bytes, stream, 0);
}
extern "C" int amid_encoder_bwd_tc(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                   int32_t L, const amid_dropout* drop, const amid_encoder_saved* S, const float* enc_out,
                                   const float* d_enc, amid_encoder_tensors* G, float* dx0, void* workspace,
                                   int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_bwd_impl(P, x0, tmask, B, L, drop, S, enc_out, d_enc, G, dx0, workspace, workspace_bytes, stream, 1);
}
extern "C" int amid_encoder_bwd_bf16(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                     int32_t L, const amid_dropout* drop, const amid_encoder_saved* S, const float* enc_out,
                                     const float* d_enc, amid_encoder_tensors* G, float* dx0, void* workspace,
                                     int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_bwd_impl(P, x0, tmask, B, L, drop, S, enc_out, d_enc, G, dx0, workspace, workspace_bytes, stream, 2);
}
extern "C" int amid_encoder_bwd_x3(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask, int32_t B,
                                   int32_t L, const amid_dropout* drop, const amid_encoder_saved* S, const float* enc_out,
                                   const float* d_enc, amid_encoder_tensors* G, float* dx0, void* workspace,
                                   int64_t workspace_bytes, amid_stream_t stream) {
    return encoder_bwd_impl(P, x0, tmask, B, L, drop, S, enc_out, d_enc, G, dx0, workspace, workspace_bytes, stream, 3);
}

#include "x3_test.cuh"
