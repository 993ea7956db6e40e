// Causal multi-head attention (head_dim 16, 64 <= L <= 224) on tcgen05 tensor cores with the scores in TENSOR MEMORY,
// at fp32-level accuracy through the FP16-pair split of x3.cuh (x = s^-1 (h0 + h1), products h0h0 + h1h0 + h0h1).
//
// One CTA per (sample, head), 256 threads, two CTAs per SM.  The head slices of q / k / v (and dO in the backward) are
// read ONCE from HBM (64-byte row segments), scaled by a power of two per head slice and stored as FP16 pairs in
// K-major SWIZZLE_128B operand tiles; v / k / q / dO are additionally stored TRANSPOSED ([16 features][positions]) as
// the B operands of the second GEMMs.
//
// forward, per 128-query tile:  S[128 x Nk] = Q K^T   (3 MMAs, K = 16, fp32 accumulator in TMEM, Nk <= 224 columns)
//   thread = (query row, parity of the 32-column chunks): row max -> P = 2^(S f - m) -> row sum -> dropout ->
//   FP16 pair split -> written back IN PLACE over the consumed score columns (chunk c: P0 in [32c, 32c+16), P1 in
//   [32c+16, 32c+32)) -> O[128 x 16] += P V as MMAs with the A operand read from tensor memory -> O / l.
// backward: two passes over <= 112-column blocks, S and dP side by side in TMEM (2 x 112 columns):
//   pass A (rows = queries): dS = P o (dPd - delta) -> in place -> dQ += dS K
//   pass B (rows = keys):    S^T, dP^T recomputed; Pd^T, dS^T in place -> dV += Pd^T dO, dK += dS^T Q.
// Arithmetic follows torch/nn/functional.py:6630-6647 (q pre-scaled by 0.25, -inf above the diagonal, softmax,
// dropout without renormalisation, P v); the dropout bits are the same counter hash as every other attention kernel.
#pragma once
#include "x3.cuh"
#include "attn_mma.cuh"

namespace amid {
namespace attn_tc {
using namespace tc;
using x3::idesc_f16;
using x3::mma_f16_ss;
using x3::mma_f16_ts;
using x3::pow2_scale;
using x3::split_f16x2;
using x3::tmem_st16;
using x3::tmem_st_wait;
using tcenc::align1k;
using attn::ex2;
using attn::LN2;
using attn::LOG2E;

constexpr int MAXL = 224;                       // score columns of a query tile must fit the TMEM allocation
constexpr int MINL = 64;                        // below this the 128-row tiles are mostly padding: mma.sync path
constexpr int QP_BYTES = 2 * 128 * 128;         // two query tiles, rows of 128 B: [q0 (16 fp16) | q1 | unused]
constexpr int KP_BYTES = MAXL * 128;            // [k0 | k1 | unused]
constexpr int VT_BYTES = 4 * 2048;              // one transposed piece: 4 chunks of [16 rows x 64 keys]
constexpr size_t FWD_SMEM = (size_t)QP_BYTES + KP_BYTES + 2 * VT_BYTES + 1024;
constexpr uint32_t O_COL = 240;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// byte offset of 16-byte unit u of row r in a K-major SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t row_unit(int r, int u) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((u ^ r) & 7) << 4));
}
// byte offset of element (feature d, position j) in a transposed piece [16][<= 256], K-major, K = positions
__device__ __forceinline__ uint32_t t_off(int d, int j) {
    return (uint32_t)((j >> 6) * 2048 + (d >> 3) * 1024 + (d & 7) * 128 + (((((j & 63) >> 3) ^ d) & 7) << 4) + (j & 7) * 2);
}
// 4 consecutive features (scaled) -> 8 bytes of piece 0 and 8 bytes of piece 1
__device__ __forceinline__ void split4(const float4 v, float s, uint2& p0, uint2& p1) {
    split_f16x2(v.x * s, v.y * s, p0.x, p1.x);
    split_f16x2(v.z * s, v.w * s, p0.y, p1.y);
}
__device__ __forceinline__ float amax4(const float4 v, float m) {
    return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// the four features of one position, both pieces, into a transposed tile pair
__device__ __forceinline__ void store_t(uint8_t* t0, uint8_t* t1, int d0, int j, const uint2 p0, const uint2 p1) {
    const uint32_t a[4] = {p0.x & 0xFFFFu, p0.x >> 16, p0.y & 0xFFFFu, p0.y >> 16};
    const uint32_t b[4] = {p1.x & 0xFFFFu, p1.x >> 16, p1.y & 0xFFFFu, p1.y >> 16};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t off = t_off(d0 + e, j);
        *reinterpret_cast<uint16_t*>(t0 + off) = (uint16_t)a[e];
        *reinterpret_cast<uint16_t*>(t1 + off) = (uint16_t)b[e];
    }
}
struct BlockMax {
    float red[4][8];
    // maxima of up to four per-thread values over the CTA (256 threads); contains one __syncthreads
    template <int N>
    __device__ __forceinline__ void run(float (&m)[N]) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float w = warp_max(m[k]);
            if (lane == 0) red[k][warp] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float r = red[k][0];
#pragma unroll
            for (int w = 1; w < 8; ++w) r = fmaxf(r, red[k][w]);
            m[k] = r;
        }
    }
};
// dropout keep decisions of 32 consecutive keys (8 hash groups starting at group g0) applied to p
__device__ __forceinline__ void drop32(float (&p)[32], const DropCfg& dc, uint32_t site, uint32_t g0) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const uint32_t r = rng4(dc.seed, site, (uint64_t)(g0 + g));
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (((r >> (8 * e)) & 0xFFu) < dc.thr16) p[4 * g + e] = 0.f;
    }
}

struct ShF {
    uint64_t bar;
    uint32_t tmem;
    BlockMax bm;
    float xm[2][128];
    float xl[2][128];
};

__global__ void __launch_bounds__(256, 2)
k_attn_fwd_tc(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, float* __restrict__ o,
              float* __restrict__ lse, int L, DropCfg dc, uint32_t site) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ ShF sh;
    uint8_t* Qp = align1k(smem_raw);
    uint8_t* Kp = Qp + QP_BYTES;
    uint8_t* V0 = Kp + KP_BYTES;
    uint8_t* V1 = V0 + VT_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    const int ntile = (L + 127) >> 7, NKP = (L + 15) & ~15, Lp4 = ((L + 3) & ~3) >> 2;
    if (warp == 0) tmem_alloc(&sh.tmem, 256);
    if (tid == 0) { mbar_init(&sh.bar, 1); fence_barrier_init(); }
    // ---- head slices -> registers, per-slice maxima, FP16 pair tiles
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
    sh.bm.run(mx);
    float sq, iq, sk, ik, sv, iv;
    pow2_scale(mx[0], sq, iq); pow2_scale(mx[1], sk, ik); pow2_scale(mx[2], sv, iv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        uint2 p0, p1;
        split4(vq[i], sq, p0, p1);
        {
            uint8_t* t = Qp + (R >> 7) * (128 * 128);
            const int rr = R & 127;
            *reinterpret_cast<uint2*>(t + row_unit(rr, c4 >> 1) + (c4 & 1) * 8) = p0;
            *reinterpret_cast<uint2*>(t + row_unit(rr, 2 + (c4 >> 1)) + (c4 & 1) * 8) = p1;
        }
        if (R < MAXL) {
            split4(vk[i], sk, p0, p1);
            *reinterpret_cast<uint2*>(Kp + row_unit(R, c4 >> 1) + (c4 & 1) * 8) = p0;
            *reinterpret_cast<uint2*>(Kp + row_unit(R, 2 + (c4 >> 1)) + (c4 & 1) * 8) = p1;
        }
        split4(vv[i], sv, p0, p1);
        store_t(V0, V1, 4 * c4, R, p0, p1);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = sh.tmem;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int half = warp >> 2, row = 32 * (warp & 3) + lane;
    const float f = iq * ik * LOG2E;               // raw accumulator -> scores in the log2 domain
    const float osc = (dc.train ? dc.scale : 1.0f) * iv;
    uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        const int Nk = min(NKP, 128 * (t + 1));
        if (tid == 0) {
            const uint32_t qa = smem_u32(Qp) + t * (128 * 128), ka = smem_u32(Kp);
            const uint32_t id = idesc_f16(Nk, false, false);
            mma_f16_ss(tmem, make_desc(qa, 16, 1024), make_desc(ka, 16, 1024), id, 0u);            // q0 k0
            mma_f16_ss(tmem, make_desc(qa + 32, 16, 1024), make_desc(ka, 16, 1024), id, 1u);       // q1 k0
            mma_f16_ss(tmem, make_desc(qa, 16, 1024), make_desc(ka + 32, 16, 1024), id, 1u);       // q0 k1
            mma_commit(&sh.bar);
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
                uint32_t p0[16], p1[16];
                if (c <= cdiag) {
                    float a[32];
                    tmem_ld32(tl + 32 * c, a);
#pragma unroll
                    for (int e = 0; e < 32; ++e) a[e] = ex2(fmaf(a[e], f, -m));
                    if (c == cdiag) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (e > lane) a[e] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) l += a[e];
                    if (dc.train) drop32(a, dc, site, rb4 + 8 * c);
#pragma unroll
                    for (int e = 0; e < 16; ++e) split_f16x2(a[2 * e], a[2 * e + 1], p0[e], p1[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) p0[e] = p1[e] = 0u;
                }
                tmem_st16(tl + 32 * c, p0);
                tmem_st16(tl + 32 * c + 16, p1);
            }
        }
        sh.xl[half][row] = l;
        tmem_st_wait();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            const uint32_t v0 = smem_u32(V0), v1 = smem_u32(V1);
            constexpr uint32_t id = idesc_f16(16, false, false);
            for (int kk = 0; kk < Nk / 16; ++kk) {
                const uint32_t a0 = tmem + 32 * (kk >> 1) + 8 * (kk & 1), a1 = a0 + 16;
                const uint32_t bo = (kk >> 2) * 2048 + (kk & 3) * 32;
                mma_f16_ts(tmem + O_COL, a0, make_desc(v0 + bo, 16, 1024), id, kk ? 1u : 0u);
                mma_f16_ts(tmem + O_COL, a1, make_desc(v0 + bo, 16, 1024), id, 1u);
                mma_f16_ts(tmem + O_COL, a0, make_desc(v1 + bo, 16, 1024), id, 1u);
            }
            mma_commit(&sh.bar);
        }
        mbar_wait(&sh.bar, phase);
        phase ^= 1;
        fence_after();
        if (half == 0 && wvalid) {               // warp-uniform: tcgen05.ld is a warp-collective instruction
            float a[16];
            tmem_ld16(tl + O_COL, a);
            if (i < L) {
                const float lt = sh.xl[0][row] + sh.xl[1][row];
                const float sc = osc / lt;
                float4* dst = reinterpret_cast<float4*>(o + base + (size_t)i * D);
#pragma unroll
                for (int u = 0; u < 4; ++u) dst[u] = make_float4(a[4 * u] * sc, a[4 * u + 1] * sc, a[4 * u + 2] * sc, a[4 * u + 3] * sc);
                lse[(size_t)bh * L + i] = m * LN2 + logf(lt);
            }
        }
        fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ================================================================================================
// backward
// ================================================================================================
constexpr int BW = 112;                         // score-block width: S and dP blocks side by side in 256 TMEM columns
constexpr int RT_BYTES = MAXL * 128;            // row tiles: [q0 | q1 | g0 | g1] per query, [k0 | k1 | v0 | v1] per key
constexpr size_t BWD_SMEM = 2 * (size_t)RT_BYTES + 6 * VT_BYTES + 1024;
constexpr uint32_t X_COL = 0, Y_COL = BW, A1_COL = 224, A2_COL = 240;

struct ShB {
    uint64_t bar;
    uint32_t tmem;
    BlockMax bm;
    float ls[256];          // lse * log2(e) per query (+inf for rows >= L)
    float dl[256];          // delta_i = <dO_i, O_i>
};
// 16 consecutive columns -> FP16 pair pieces, 8 packed columns each
__device__ __forceinline__ void put16(uint32_t taddr, const float (&v)[16]) {
    uint32_t p0[8], p1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_f16x2(v[2 * e], v[2 * e + 1], p0[e], p1[e]);
    tmem_st8(taddr, p0);
    tmem_st8(taddr + 8, p1);
}
__device__ __forceinline__ void zero16(uint32_t taddr) {
    const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(taddr, z);
    tmem_st8(taddr + 8, z);
}
// byte offset of the 16-position k-step starting at position pos (multiple of 16) inside a transposed piece
__device__ __forceinline__ uint32_t t_kstep(int pos) { return (uint32_t)((pos >> 6) * 2048 + ((pos & 63) >> 4) * 32); }

__global__ void __launch_bounds__(256, 2)
k_attn_bwd_tc(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
              const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dO,
              float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int L, DropCfg dc, uint32_t site) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ ShB sh;
    uint8_t* QG = align1k(smem_raw);
    uint8_t* KV = QG + RT_BYTES;
    uint8_t* KT0 = KV + RT_BYTES;
    uint8_t* KT1 = KT0 + VT_BYTES;
    uint8_t* QT0 = KT1 + VT_BYTES;
    uint8_t* QT1 = QT0 + VT_BYTES;
    uint8_t* GT0 = QT1 + VT_BYTES;
    uint8_t* GT1 = GT0 + VT_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x, b = bh / H, hd = bh % H;
    const size_t base = (size_t)b * L * D + hd * DH;
    const int ntile = (L + 127) >> 7, NKP = (L + 15) & ~15, nblk = (NKP + BW - 1) / BW, Lp4 = ((L + 3) & ~3) >> 2;
    if (warp == 0) tmem_alloc(&sh.tmem, 256);
    if (tid == 0) { mbar_init(&sh.bar, 1); fence_barrier_init(); }
    const int c4 = tid & 3, r0 = tid >> 2;
    float4 vq[4], vk[4], vv[4], vg[4];
    float mx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        vq[i] = vk[i] = vv[i] = vg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float dsum = 0.f;
        if (R < L) {
            const size_t off = base + (size_t)R * D + 4 * c4;
            vq[i] = __ldg(reinterpret_cast<const float4*>(q + off));
            vk[i] = __ldg(reinterpret_cast<const float4*>(k + off));
            vv[i] = __ldg(reinterpret_cast<const float4*>(v + off));
            vg[i] = __ldg(reinterpret_cast<const float4*>(dO + off));
            const float4 oo = __ldg(reinterpret_cast<const float4*>(o + off));
            dsum = vg[i].x * oo.x + vg[i].y * oo.y + vg[i].z * oo.z + vg[i].w * oo.w;
        }
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        if (c4 == 0) {
            sh.dl[R] = dsum;
            sh.ls[R] = R < L ? lse[(size_t)bh * L + R] * LOG2E : INFINITY;
        }
        mx[0] = amax4(vq[i], mx[0]); mx[1] = amax4(vk[i], mx[1]); mx[2] = amax4(vv[i], mx[2]); mx[3] = amax4(vg[i], mx[3]);
    }
    sh.bm.run(mx);
    float sq, iq, sk, ik, sv, iv, sg, ig, sds, ids;
    pow2_scale(mx[0], sq, iq); pow2_scale(mx[1], sk, ik); pow2_scale(mx[2], sv, iv); pow2_scale(mx[3], sg, ig);
    const float dsc = dc.train ? dc.scale : 1.0f;
    pow2_scale(64.0f * mx[3] * mx[2] * dsc, sds, ids);        // |dS| <= |dPd| + |delta| <= 2 * 16 gmax vmax scale
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = r0 + 64 * i;
        uint2 a0, a1, b0, b1;
        split4(vq[i], sq, a0, a1);
        split4(vg[i], sg, b0, b1);
        if (R < MAXL) {
            const uint32_t h8 = (c4 & 1) * 8;
            *reinterpret_cast<uint2*>(QG + row_unit(R, c4 >> 1) + h8) = a0;
            *reinterpret_cast<uint2*>(QG + row_unit(R, 2 + (c4 >> 1)) + h8) = a1;
            *reinterpret_cast<uint2*>(QG + row_unit(R, 4 + (c4 >> 1)) + h8) = b0;
            *reinterpret_cast<uint2*>(QG + row_unit(R, 6 + (c4 >> 1)) + h8) = b1;
        }
        store_t(QT0, QT1, 4 * c4, R, a0, a1);
        store_t(GT0, GT1, 4 * c4, R, b0, b1);
        split4(vk[i], sk, a0, a1);
        split4(vv[i], sv, b0, b1);
        if (R < MAXL) {
            const uint32_t h8 = (c4 & 1) * 8;
            *reinterpret_cast<uint2*>(KV + row_unit(R, c4 >> 1) + h8) = a0;
            *reinterpret_cast<uint2*>(KV + row_unit(R, 2 + (c4 >> 1)) + h8) = a1;
            *reinterpret_cast<uint2*>(KV + row_unit(R, 4 + (c4 >> 1)) + h8) = b0;
            *reinterpret_cast<uint2*>(KV + row_unit(R, 6 + (c4 >> 1)) + h8) = b1;
        }
        store_t(KT0, KT1, 4 * c4, R, a0, a1);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = sh.tmem;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int half = warp >> 2;
    const float f = iq * ik * LOG2E;               // raw S -> log2 domain
    const float fdp = ig * iv * dsc;               // raw dP -> dPd where kept
    const uint32_t qg = smem_u32(QG), kv = smem_u32(KV);
    constexpr uint32_t id16 = idesc_f16(16, false, false);
    uint32_t phase = 0;

    // ---------------- pass A: rows = queries; dQ += dS K
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
            const int key0 = BW * kb;
            if (key0 > last_q) break;
            const int Nb = min(BW, NKP - key0);
            if (tid == 0) {
                const uint32_t id = idesc_f16(Nb, false, false);
                const uint32_t a = qg + (uint32_t)(128 * t) * 128, bb = kv + (uint32_t)key0 * 128;
                mma_f16_ss(tmem + X_COL, make_desc(a, 16, 1024), make_desc(bb, 16, 1024), id, 0u);             // q0 k0
                mma_f16_ss(tmem + X_COL, make_desc(a + 32, 16, 1024), make_desc(bb, 16, 1024), id, 1u);        // q1 k0
                mma_f16_ss(tmem + X_COL, make_desc(a, 16, 1024), make_desc(bb + 32, 16, 1024), id, 1u);        // q0 k1
                mma_f16_ss(tmem + Y_COL, make_desc(a + 64, 16, 1024), make_desc(bb + 64, 16, 1024), id, 0u);   // g0 v0
                mma_f16_ss(tmem + Y_COL, make_desc(a + 96, 16, 1024), make_desc(bb + 64, 16, 1024), id, 1u);   // g1 v0
                mma_f16_ss(tmem + Y_COL, make_desc(a + 64, 16, 1024), make_desc(bb + 96, 16, 1024), id, 1u);   // g0 v1
                mma_commit(&sh.bar);
            }
            mbar_wait(&sh.bar, phase);
            phase ^= 1;
            fence_after();
            if (wvalid) {
#pragma unroll 1
                for (int ch = half; ch < Nb / 16; ch += 2) {
                    const int j0 = key0 + 16 * ch;
                    if (j0 > Rw + 31) { zero16(tl + X_COL + 16 * ch); continue; }
                    float sx[16], dp[16];
                    tmem_ld16(tl + X_COL + 16 * ch, sx);
                    tmem_ld16(tl + Y_COL + 16 * ch, dp);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t r = 0xFFFFFFFFu;
                        if (dc.train) r = rng4(dc.seed, site, (uint64_t)(rb4 + (uint32_t)(j0 >> 2) + g));
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 4 * g + e;
                            const float p = ex2(fmaf(sx[c], f, -li));
                            const bool keep = !dc.train || ((r >> (8 * e)) & 0xFFu) >= dc.thr16;
                            const float dpd = keep ? dp[c] * fdp : 0.f;
                            float ds = p * (dpd - Di) * sds;
                            if (j0 + c > i) ds = 0.f;
                            sx[c] = ds;
                        }
                    }
                    put16(tl + X_COL + 16 * ch, sx);
                }
            }
            tmem_st_wait();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                const uint32_t k0 = smem_u32(KT0), k1 = smem_u32(KT1);
                for (int kk = 0; kk < Nb / 16; ++kk) {
                    const uint32_t a0 = tmem + X_COL + 16 * kk, a1 = a0 + 8;
                    const uint32_t bo = t_kstep(key0 + 16 * kk);
                    mma_f16_ts(tmem + A1_COL, a0, make_desc(k0 + bo, 16, 1024), id16, (first && kk == 0) ? 0u : 1u);
                    mma_f16_ts(tmem + A1_COL, a1, make_desc(k0 + bo, 16, 1024), id16, 1u);
                    mma_f16_ts(tmem + A1_COL, a0, make_desc(k1 + bo, 16, 1024), id16, 1u);
                }
                mma_commit(&sh.bar);
            }
            first = false;
            mbar_wait(&sh.bar, phase);
            phase ^= 1;
            fence_after();
            __syncthreads();      // no thread may still be in this wait when the next commit on the same barrier is issued
        }
        if (half == 0 && wvalid) {
            float a[16];
            tmem_ld16(tl + A1_COL, a);
            if (i < L) {
                const float sc = 0.25f * ids * ik;
                float4* dst = reinterpret_cast<float4*>(dq + base + (size_t)i * D);
#pragma unroll
                for (int u = 0; u < 4; ++u) dst[u] = make_float4(a[4 * u] * sc, a[4 * u + 1] * sc, a[4 * u + 2] * sc, a[4 * u + 3] * sc);
            }
        }
        fence_before();
        __syncthreads();
    }

    // ---------------- pass B: rows = keys; dV += Pd^T dO, dK += dS^T Q
#pragma unroll 1
    for (int t = 0; t < ntile; ++t) {
        const int Jw = 128 * t + 32 * (warp & 3), j = Jw + lane;
        const bool wvalid = Jw < L;
        const uint32_t jg = (uint32_t)(j >> 2);
        const int jb = 8 * (j & 3);
        bool first = true;
#pragma unroll 1
        for (int qb = 0; qb < nblk; ++qb) {
            const int q0 = BW * qb;
            const int Nb = min(BW, NKP - q0);
            if (q0 + Nb <= 128 * t) continue;              // every query of the block precedes every key of the tile
            if (tid == 0) {
                const uint32_t id = idesc_f16(Nb, false, false);
                const uint32_t a = kv + (uint32_t)(128 * t) * 128, bb = qg + (uint32_t)q0 * 128;
                mma_f16_ss(tmem + X_COL, make_desc(a, 16, 1024), make_desc(bb, 16, 1024), id, 0u);             // k0 q0
                mma_f16_ss(tmem + X_COL, make_desc(a + 32, 16, 1024), make_desc(bb, 16, 1024), id, 1u);        // k1 q0
                mma_f16_ss(tmem + X_COL, make_desc(a, 16, 1024), make_desc(bb + 32, 16, 1024), id, 1u);        // k0 q1
                mma_f16_ss(tmem + Y_COL, make_desc(a + 64, 16, 1024), make_desc(bb + 64, 16, 1024), id, 0u);   // v0 g0
                mma_f16_ss(tmem + Y_COL, make_desc(a + 96, 16, 1024), make_desc(bb + 64, 16, 1024), id, 1u);   // v1 g0
                mma_f16_ss(tmem + Y_COL, make_desc(a + 64, 16, 1024), make_desc(bb + 96, 16, 1024), id, 1u);   // v0 g1
                mma_commit(&sh.bar);
            }
            mbar_wait(&sh.bar, phase);
            phase ^= 1;
            fence_after();
            if (wvalid) {
#pragma unroll 1
                for (int ch = half; ch < Nb / 16; ch += 2) {
                    const int i0 = q0 + 16 * ch;
                    if (i0 + 15 < Jw) { zero16(tl + X_COL + 16 * ch); zero16(tl + Y_COL + 16 * ch); continue; }
                    float sx[16], dp[16];
                    tmem_ld16(tl + X_COL + 16 * ch, sx);
                    tmem_ld16(tl + Y_COL + 16 * ch, dp);
                    // the four keys of a hash group sit in four neighbouring lanes: lane x of the quad hashes the
                    // queries e = x, x+4, x+8, x+12 of the chunk and the quad exchanges them
                    uint32_t hv[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
                    if (dc.train) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int qi = min(i0 + (lane & 3) + 4 * u, L - 1);
                            hv[u] = rng4(dc.seed, site, (uint64_t)((((uint32_t)bh + dc.bh_off) * (uint32_t)L + (uint32_t)qi) * (uint32_t)Lp4 + jg));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int i = i0 + e;
                        const uint32_t r = __shfl_sync(0xffffffffu, hv[e >> 2], (lane & ~3) | (e & 3));
                        float p = ex2(fmaf(sx[e], f, -sh.ls[i]));
                        if (j > i) p = 0.f;
                        const bool keep = !dc.train || ((r >> jb) & 0xFFu) >= dc.thr16;
                        const float dpd = keep ? dp[e] * fdp : 0.f;
                        sx[e] = keep ? p * dsc : 0.f;
                        dp[e] = p * (dpd - sh.dl[i]) * sds;
                    }
                    put16(tl + X_COL + 16 * ch, sx);
                    put16(tl + Y_COL + 16 * ch, dp);
                }
            }
            tmem_st_wait();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                const uint32_t g0 = smem_u32(GT0), g1 = smem_u32(GT1), t0 = smem_u32(QT0), t1 = smem_u32(QT1);
                for (int kk = 0; kk < Nb / 16; ++kk) {
                    const uint32_t ax = tmem + X_COL + 16 * kk, ay = tmem + Y_COL + 16 * kk;
                    const uint32_t bo = t_kstep(q0 + 16 * kk);
                    const uint32_t acc = (first && kk == 0) ? 0u : 1u;
                    mma_f16_ts(tmem + A1_COL, ax, make_desc(g0 + bo, 16, 1024), id16, acc);
                    mma_f16_ts(tmem + A1_COL, ax + 8, make_desc(g0 + bo, 16, 1024), id16, 1u);
                    mma_f16_ts(tmem + A1_COL, ax, make_desc(g1 + bo, 16, 1024), id16, 1u);
                    mma_f16_ts(tmem + A2_COL, ay, make_desc(t0 + bo, 16, 1024), id16, acc);
                    mma_f16_ts(tmem + A2_COL, ay + 8, make_desc(t0 + bo, 16, 1024), id16, 1u);
                    mma_f16_ts(tmem + A2_COL, ay, make_desc(t1 + bo, 16, 1024), id16, 1u);
                }
                mma_commit(&sh.bar);
            }
            first = false;
            mbar_wait(&sh.bar, phase);
            phase ^= 1;
            fence_after();
            __syncthreads();      // no thread may still be in this wait when the next commit on the same barrier is issued
        }
        if (wvalid) {                                    // half 0 stores dv, half 1 stores dk
            float a[16];
            tmem_ld16(tl + (half == 0 ? A1_COL : A2_COL), a);
            if (j < L) {
                const float sc = half == 0 ? ig : ids * iq;
                float4* dst = reinterpret_cast<float4*>((half == 0 ? dv : dk) + base + (size_t)j * D);
#pragma unroll
                for (int u = 0; u < 4; ++u) dst[u] = make_float4(a[4 * u] * sc, a[4 * u + 1] * sc, a[4 * u + 2] * sc, a[4 * u + 3] * sc);
            }
        }
        fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace attn_tc
}  // namespace amid
