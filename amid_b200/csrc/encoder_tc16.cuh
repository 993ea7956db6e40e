// BF16-operand version of the tcgen05 encoder kernels (precision = "bf16").
// Operand tiles are BF16 (32 KB each, tc.cuh), accumulation is fp32 in TMEM, all epilogue math and
// every tensor in HBM stay fp32.  Half-size tiles bring a CTA to ~100 KB of shared memory and <= 128
// registers per thread, so TWO CTAs are resident per SM: one CTA's loads / epilogue overlap the other's
// MMAs, which is what the latency-bound TF32 version (one 192 KB CTA per SM) lacked.  The weight
// gradients read the row-major token tiles through MN-major descriptors (dW = dY^T X with no
// transposes) and keep their accumulators in TMEM across all the tiles of a CTA.
#pragma once
#include "encoder_tc.cuh"

namespace amid {
namespace tc16 {
using namespace tc;
using tcenc::Epi;
using tcenc::Shared;
using tcenc::align1k;
using tcenc::flush_ln_partials;
using tcenc::row_sum;
using tcenc::rows_valid;
using tcenc::setup;
using tcenc::teardown;
using tcenc::warp_colsum32;
using tcenc::warp_load32;
using tcenc::warp_load32_cg;
using tcenc::warp_store32;
using tcenc::WSTAGE_FLOATS;

constexpr size_t CHAIN16_SMEM = 2 * (size_t)TILE16_BYTES + 8 * WSTAGE_FLOATS * 4 + 1024;   // 97 KB
constexpr int ONES16_BYTES = 16 * 128 * 2;
constexpr size_t WGRAD16_SMEM = 2 * (size_t)TILE16_BYTES + ONES16_BYTES + 1024;

struct Chain16 {
    uint8_t* A;
    uint8_t* W;
    float* stage;
    __device__ Chain16(uint8_t* raw) {
        A = align1k(raw);
        W = A + TILE16_BYTES;
        stage = reinterpret_cast<float*>(W + TILE16_BYTES) + (threadIdx.x >> 5) * WSTAGE_FLOATS;
    }
};
// bf16 weight [128 n][128 k] (row-major) -> swizzled K-major tile, asynchronously
__device__ __forceinline__ void load_w16_async(uint8_t* buf, const uint16_t* __restrict__ W) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < 128 * 16; idx += 256) {
        const int r = idx >> 4, u = idx & 15;
        cp_async16(buf + tile16_off8(r, u), W + (size_t)r * D + u * 8);
    }
    cp_async_commit();
}
__device__ __forceinline__ void run_gemm16(Shared& sh, uint32_t acc_col, const uint8_t* A, const uint8_t* W, bool accumulate,
                                           uint32_t& phase) {
    cp_async_wait<0>();
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after();
        issue_gemm16_kk(sh.tmem + acc_col, smem_u32(A), smem_u32(W), accumulate);
        mma_commit(&sh.bar);
    }
    mbar_wait(&sh.bar, phase);
    phase ^= 1;
    fence_after();
}
// two-pass LayerNorm statistics of a 128-wide row whose halves sit in two threads (64 registers each)
__device__ __forceinline__ void ln_stats(Shared& sh, const Epi& e, const float (&xr)[64], float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += xr[i];
    mean = row_sum(sh, e, 0, s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) { const float a = xr[i] - mean; ss = fmaf(a, a, ss); }
    rstd = 1.0f / sqrtf(row_sum(sh, e, 1, ss) * (1.0f / D) + LN_EPS);
}

// weights fp32 [128][128] -> bf16 [128][128], optionally transposed; grid (4,4,n)
struct PrepJobs {
    const float* src[12];
};
__global__ void k_prep_w16(PrepJobs jobs, uint16_t* __restrict__ dst, int transpose) {
    __shared__ float t[32][33];
    const float* s = jobs.src[blockIdx.z];
    uint16_t* d = dst + (size_t)blockIdx.z * D * D;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) t[j][threadIdx.x] = s[(size_t)(by + j) * D + bx + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        if (transpose)
            d[(size_t)(bx + j) * D + by + threadIdx.x] = (uint16_t)(pack_bf16(t[threadIdx.x][j], 0.f) & 0xFFFFu);
        else
            d[(size_t)(by + j) * D + bx + threadIdx.x] = (uint16_t)(pack_bf16(t[j][threadIdx.x], 0.f) & 0xFFFFu);
    }
}

// ----------------------------------------------------------------------------------------------
// forward 1: k, v from x; q from LN1(x)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_ln_qkv_16(const float* __restrict__ x, int M, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
            const uint16_t* __restrict__ Wq, const uint16_t* __restrict__ Wk, const uint16_t* __restrict__ Wv,
            const float* __restrict__ in_b, float* __restrict__ qn, float* __restrict__ st1, float* __restrict__ q,
            float* __restrict__ k, float* __restrict__ v) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    Chain16 sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w16_async(sm.W, Wk);
    fill_tile16(sm.A, x, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    float* outs[2] = {k, v};
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        run_gemm16(sh, 0, sm.A, sm.W, false, phase);
        load_w16_async(sm.W, g == 0 ? Wv : Wq);
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            const float* bb = in_b + (g + 1) * D + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += __ldg(bb + i);
            warp_store32(sm.stage, e.lane, a, outs[g] + wbase + c0, rv);
        }
    }
    {   // LN1 from the fp32 input (not from the rounded tile); result becomes the new operand tile
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            warp_load32(sm.stage, e.lane, x + wbase + e.cb + half * 32, rv, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[half * 32 + i] = a[i];
        }
        float mean, rstd;
        ln_stats(sh, e, xr, mean, rstd);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf((xr[half * 32 + i] - mean) * rstd, __ldg(ln_w + c0 + i), __ldg(ln_b + c0 + i));
            tile16_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, qn + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st1[(size_t)gr * 2] = mean; st1[(size_t)gr * 2 + 1] = rstd; }
    }
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = (a[i] + __ldg(in_b + c0 + i)) * 0.25f;
            warp_store32(sm.stage, e.lane, a, q + wbase + c0, rv);
        }
    }
    teardown(sh, sh.tmem, 128);
}

// ----------------------------------------------------------------------------------------------
// forward 2: out-proj + residual + LN2 + FFN + mask (+ last LN)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_proj_ffn_16(const float* __restrict__ o, const float* __restrict__ qn, int M, const uint16_t* __restrict__ Wo,
              const float* __restrict__ bo, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
              const uint16_t* __restrict__ W1, const float* __restrict__ b1, const uint16_t* __restrict__ W2,
              const float* __restrict__ b2, const uint32_t* __restrict__ tmask, DropCfg dc, uint32_t site1, uint32_t site2,
              float* __restrict__ x1, float* __restrict__ st2, float* __restrict__ y, float* __restrict__ h,
              float* __restrict__ xout, const float* __restrict__ ln3_w, const float* __restrict__ ln3_b,
              float* __restrict__ enc, float* __restrict__ st3) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    Chain16 sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w16_async(sm.W, Wo);
    fill_tile16(sm.A, o, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    // ---- x1 = Qn + o Wo^T + bo ; y = LN2(x1)
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);
    load_w16_async(sm.W, W1);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], r[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_load32(sm.stage, e.lane, qn + wbase + c0, rv, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += __ldg(bo + c0 + i) + r[i];
            warp_store32(sm.stage, e.lane, a, x1 + wbase + c0, rv);
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[half * 32 + i] = a[i];
        }
        float mean, rstd;
        ln_stats(sh, e, xr, mean, rstd);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf((xr[half * 32 + i] - mean) * rstd, __ldg(ln2_w + c0 + i), __ldg(ln2_b + c0 + i));
            tile16_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, y + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st2[(size_t)gr * 2] = mean; st2[(size_t)gr * 2 + 1] = rstd; }
    }
    // ---- h = relu(dropout1(y W1^T + b1))
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);
    load_w16_async(sm.W, W2);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b1 + c0 + i));
                float4 t = make_float4(a[i] + b4.x, a[i + 1] + b4.y, a[i + 2] + b4.z, a[i + 3] + b4.w);
                if (dc.train) t = drop4(t, dc, site1, (uint64_t)(gr + dc.tok_off) * D + c0 + i);
                a[i] = fmaxf(t.x, 0.f); a[i + 1] = fmaxf(t.y, 0.f); a[i + 2] = fmaxf(t.z, 0.f); a[i + 3] = fmaxf(t.w, 0.f);
            }
            tile16_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, h + wbase + c0, rv);
        }
    }
    // ---- xout = (dropout2(h W2^T + b2) + y) * ~tmask  (+ last LayerNorm)
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr);
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], yy[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_load32_cg(sm.stage, e.lane, y + wbase + c0, rv, yy);   // written by this CTA above: coherent loads
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + c0 + i));
                float4 t = make_float4(a[i] + b4.x, a[i + 1] + b4.y, a[i + 2] + b4.z, a[i + 3] + b4.w);
                if (dc.train) t = drop4(t, dc, site2, (uint64_t)(gr + dc.tok_off) * D + c0 + i);
                t = make_float4(t.x + yy[i], t.y + yy[i + 1], t.z + yy[i + 2], t.w + yy[i + 3]);
                t = apply_tmask(t, tw, (c0 + i) >> 2);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
            warp_store32(sm.stage, e.lane, a, xout + wbase + c0, rv);
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[half * 32 + i] = a[i];
        }
        if (enc) {
            float mean, rstd;
            ln_stats(sh, e, xr, mean, rstd);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float a[32];
                const int c0 = e.cb + half * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = fmaf((xr[half * 32 + i] - mean) * rstd, __ldg(ln3_w + c0 + i), __ldg(ln3_b + c0 + i));
                warp_store32(sm.stage, e.lane, a, enc + wbase + c0, rv);
            }
            if (valid && e.cb == 0) { st3[(size_t)gr * 2] = mean; st3[(size_t)gr * 2 + 1] = rstd; }
        }
    }
    teardown(sh, sh.tmem, 128);
}

// LayerNorm backward of one row (two threads, 64 columns each), in two sweeps over the half rows so that
// at most ~3 x 32 values are live.  get(half, dy[32], xh[32]) produces the upstream gradient and x-hat.
template <class Get, class Put>
__device__ __forceinline__ void ln_bwd_row(Shared& sh, const Epi& e, const float* __restrict__ w, float rstd, bool valid,
                                           Get get, Put put) {
    float p1 = 0.f, p2 = 0.f;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float dy[32], xh[32];
        const int c0 = e.cb + half * 32;
        get(half, c0, dy, xh);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float dw = dy[i] * __ldg(w + c0 + i);
            p1 += dw;
            p2 = fmaf(dw, xh[i], p2);
        }
    }
    const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
    const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float dy[32], xh[32], dx[32];
        const int c0 = e.cb + half * 32;
        get(half, c0, dy, xh);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            dx[i] = valid ? rstd * (dy[i] * __ldg(w + c0 + i) - c1 - xh[i] * c2) : 0.f;
            xh[i] *= dy[i];                       // dw terms
        }
        put(half, c0, dx);
        const float sw = warp_colsum32(xh, e.lane);
        const float sb = warp_colsum32(dy, e.lane);
        sh.lnacc[e.warp][0][half * 32 + e.lane] = sw;
        sh.lnacc[e.warp][1][half * 32 + e.lane] = sb;
    }
}

// ----------------------------------------------------------------------------------------------
// backward 1: FFN + LN2 + out-proj input gradient (bf16 weights passed TRANSPOSED)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_ffn_bwd_16(const float* __restrict__ dxo, const float* __restrict__ h, const float* __restrict__ x1,
             const float* __restrict__ st2, const uint32_t* __restrict__ tmask, int M, const uint16_t* __restrict__ W2t,
             const uint16_t* __restrict__ W1t, const uint16_t* __restrict__ Wot, const float* __restrict__ ln2_w, DropCfg dc,
             uint32_t site1, uint32_t site2, float* __restrict__ do2, float* __restrict__ dhpre, float* __restrict__ dx1,
             float* __restrict__ dO, float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    Chain16 sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w16_async(sm.W, W2t);
    {   // A = do2 = dropout2-mask * (dxo * ~tmask): warp per row, all 16 row loads of a thread in flight at once
        const int c4 = threadIdx.x & 31, rb = threadIdx.x >> 5;
        float4 g[16];
        uint4 twl = make_uint4(0u, 0u, 0u, 0u);
        if (c4 < 16 && row0 + c4 * 8 + rb < M) twl = __ldg(reinterpret_cast<const uint4*>(tmask) + row0 + c4 * 8 + rb);   // lane i: row of iteration i
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int grr = row0 + it * 8 + rb;
            g[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grr < M) g[it] = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)grr * D) + c4);
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int r = it * 8 + rb, grr = row0 + r;
            uint4 tw;
            tw.x = __shfl_sync(0xffffffffu, twl.x, it); tw.y = __shfl_sync(0xffffffffu, twl.y, it);
            tw.z = __shfl_sync(0xffffffffu, twl.z, it); tw.w = __shfl_sync(0xffffffffu, twl.w, it);
            float4 gg = apply_tmask(g[it], tw, c4);
            if (grr < M) {
                if (dc.train) gg = drop4(gg, dc, site2, (uint64_t)(grr + dc.tok_off) * D + c4 * 4);
                *(reinterpret_cast<float4*>(do2 + (size_t)grr * D) + c4) = gg;
            }
            *reinterpret_cast<uint2*>(sm.A + tile16_off8(r, c4 >> 1) + (c4 & 1) * 8) = make_uint2(pack_bf16(gg.x, gg.y), pack_bf16(gg.z, gg.w));
        }
    }
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    const float sc = dc.train ? dc.scale : 1.0f;
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);          // dh = do2 W2
    load_w16_async(sm.W, W1t);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32], hh[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_load32(sm.stage, e.lane, h + wbase + c0, rv, hh);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = hh[i] > 0.f ? a[i] * sc : 0.f;
            tile16_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, dhpre + wbase + c0, rv);
        }
    }
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);          // dy = dhpre W1 + g
    load_w16_async(sm.W, Wot);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) { mean = st2[(size_t)gr * 2]; rstd = st2[(size_t)gr * 2 + 1]; tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr); }
        auto get = [&](int half, int c0, float (&dy)[32], float (&xh)[32]) {
            float g[32];
            tmem_ld32(tm + c0, dy);
            warp_load32(sm.stage, e.lane, dxo + wbase + c0, rv, g);
            warp_load32(sm.stage, e.lane, x1 + wbase + c0, rv, xh);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 gm = apply_tmask(make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]), tw, (c0 + i) >> 2);
                dy[i] += gm.x; dy[i + 1] += gm.y; dy[i + 2] += gm.z; dy[i + 3] += gm.w;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) xh[i] = valid ? (xh[i] - mean) * rstd : 0.f;
        };
        auto put = [&](int half, int c0, float (&dx)[32]) {
            tile16_store32(sm.A, e.row, c0, dx);
            warp_store32(sm.stage, e.lane, dx, dx1 + wbase + c0, rv);
        };
        ln_bwd_row(sh, e, ln2_w, rstd, valid, get, put);
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);          // dO = dx1 Wo
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_store32(sm.stage, e.lane, a, dO + wbase + c0, rv);
        }
    }
    teardown(sh, sh.tmem, 128);
}

// ----------------------------------------------------------------------------------------------
// backward 2: dQn = dx1 + dq Wq ; dx_in = dk Wk + dv Wv + LN1bwd(dQn)   (bf16 weights transposed)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_qkv_bwd_16(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
             const float* __restrict__ dx1, const float* __restrict__ xin, const float* __restrict__ st1, int M,
             const uint16_t* __restrict__ Wqt, const uint16_t* __restrict__ Wkt, const uint16_t* __restrict__ Wvt,
             const float* __restrict__ ln1_w, float* __restrict__ dxin, float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    Chain16 sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 256);
    load_w16_async(sm.W, Wqt);
    fill_tile16(sm.A, dq, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    run_gemm16(sh, 0, sm.A, sm.W, false, phase);
    load_w16_async(sm.W, Wkt);
    fill_tile16(sm.A, dk, row0, M);
    run_gemm16(sh, 128, sm.A, sm.W, false, phase);
    load_w16_async(sm.W, Wvt);
    fill_tile16(sm.A, dv, row0, M);
    run_gemm16(sh, 128, sm.A, sm.W, true, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        if (valid) { mean = st1[(size_t)gr * 2]; rstd = st1[(size_t)gr * 2 + 1]; }
        auto get = [&](int half, int c0, float (&dy)[32], float (&xh)[32]) {
            float r[32];
            tmem_ld32(tm + c0, dy);
            warp_load32(sm.stage, e.lane, dx1 + wbase + c0, rv, r);
            warp_load32(sm.stage, e.lane, xin + wbase + c0, rv, xh);
#pragma unroll
            for (int i = 0; i < 32; ++i) { dy[i] += r[i]; xh[i] = valid ? (xh[i] - mean) * rstd : 0.f; }
        };
        auto put = [&](int half, int c0, float (&dx)[32]) {
            float base[32];
            tmem_ld32(tm + 128 + c0, base);
#pragma unroll
            for (int i = 0; i < 32; ++i) dx[i] += base[i];
            warp_store32(sm.stage, e.lane, dx, dxin + wbase + c0, rv);
        };
        ln_bwd_row(sh, e, ln1_w, rstd, valid, get, put);
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    teardown(sh, sh.tmem, 256);
}

// ----------------------------------------------------------------------------------------------
// weight gradients: dW[n][k] = sum_m dY[m][n] X[m][k] via MN-major views of the row-major token
// tiles (no transposes); db[n] through one more MMA against a K-major tile of ones.  grid = (S, 6).
// ----------------------------------------------------------------------------------------------
struct WgradJobs16 {
    const float* dY[6];
    const float* X[6];
};
__global__ void __launch_bounds__(256, 2)
k_wgrad_16(WgradJobs16 jobs, int M, float* __restrict__ wpart /*[6][S][128*128]*/, float* __restrict__ bpart /*[6][S][128]*/) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* At = align1k(smem_raw);
    uint8_t* Xt = At + TILE16_BYTES;
    uint8_t* Ones = At + 2 * TILE16_BYTES;      // [16 rows][128 tokens] bf16, K-major: 2 chunks x (16 rows x 128 B)
    const float* __restrict__ dY = jobs.dY[blockIdx.y];
    const float* __restrict__ X = jobs.X[blockIdx.y];
    const int S = gridDim.x;
    setup(sh, 256);
    for (int i = threadIdx.x; i < ONES16_BYTES / 4; i += 256) reinterpret_cast<uint32_t*>(Ones)[i] = 0x3F803F80u;   // bf16 1.0 x2
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = sh.tmem;
    const int tiles = (M + 127) / 128;
    uint32_t phase = 0;
    bool first = true;
    constexpr uint32_t id_w = idesc_bf16(128, true, true);
    constexpr uint32_t id_b = idesc_bf16(16, true, false);
    for (int t = blockIdx.x; t < tiles; t += S) {
        fill_tile16(At, dY, t * 128, M);
        fill_tile16(Xt, X, t * 128, M);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
            const uint32_t a = smem_u32(At), x = smem_u32(Xt), o = smem_u32(Ones);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {       // 16 token rows per MMA
                const uint32_t acc = (!first || ks) ? 1u : 0u;
                mma_bf16(tmem, desc16_mn(a, ks), desc16_mn(x, ks), id_w, acc);
                mma_bf16(tmem + 128, desc16_mn(a, ks), make_desc(o + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), id_b, acc);
            }
            mma_commit(&sh.bar);
        }
        mbar_wait(&sh.bar, phase);
        phase ^= 1;
        first = false;
    }
    fence_after();
    Epi e;
    float* wp = wpart + ((size_t)blockIdx.y * S + blockIdx.x) * D * D;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float a[32];
        const int c0 = e.cb + half * 32;
        if (!first) {
            tmem_ld32(tmem + e.lane_addr + c0, a);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(wp + (size_t)e.row * D + c0 + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
    }
    if (e.cb == 0) {
        float a[32];
        if (!first) {
            tmem_ld32(tmem + e.lane_addr + 128, a);
        } else {
            a[0] = 0.f;
        }
        bpart[((size_t)blockIdx.y * S + blockIdx.x) * D + e.row] = a[0];
    }
    teardown(sh, tmem, 256);
}

}  // namespace tc16
}  // namespace amid
