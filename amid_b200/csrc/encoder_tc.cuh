// Tensor-core (tcgen05 + TMEM, TF32 operands / fp32 accumulate) versions of the fused encoder
// tile kernels.  Same fusion boundaries and the same HBM tensors as the exact-fp32 kernels in
// encoder.cu; the 128x128x128 stages run as tcgen05.mma with the token tile and the weight tile
// in SWIZZLE_128B shared memory (tc.cuh), accumulators in TMEM, and the epilogues (bias, LayerNorm,
// dropout, ReLU, residual, mask, LayerNorm backward) executed by 256 threads straight out of TMEM:
// thread = (row, 64-column half), so row reductions are an exchange between two threads.
#pragma once
#include "tc.cuh"
#include "tile.cuh"

namespace amid {
namespace tcenc {
using namespace tc;

constexpr size_t CHAIN_SMEM = 3 * (size_t)TILE_BYTES + 1024;               // A tile + 2 weight buffers
constexpr int ONES_BYTES = 16 * 128 * 4;                                   // [16 rows][128] K-major
constexpr size_t WGRAD_SMEM = 2 * (size_t)TILE_BYTES + ONES_BYTES + 1024;

struct Shared {
    uint64_t bar;
    uint32_t tmem;
    float xch[2][2][128];       // [value][half][row]
    float lnacc[8][2][64];      // [warp][dw|db][col in half]
};

struct Epi {
    int warp, lane, row, cb;
    uint32_t lane_addr;
    __device__ Epi() {
        warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
        row = 32 * (warp & 3) + lane; cb = 64 * (warp >> 2);
        lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    }
};

__device__ __forceinline__ uint8_t* align1k(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}
__device__ __forceinline__ void setup(Shared& sh, int tmem_cols) {
    if ((threadIdx.x >> 5) == 0) tmem_alloc(&sh.tmem, tmem_cols);
    if (threadIdx.x == 0) { mbar_init(&sh.bar, 1); fence_barrier_init(); }
}
__device__ __forceinline__ void teardown(Shared& sh, uint32_t tmem, int tmem_cols) {
    fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem, tmem_cols);
}
// weight [128 n][128 k] (row-major, k contiguous) -> swizzled K-major tile, asynchronously
__device__ __forceinline__ void load_w_async(uint8_t* buf, const float* __restrict__ W) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < 128 * 32; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31;
        cp_async16(buf + tile_off4(r, c4), W + (size_t)r * D + c4 * 4);
    }
    cp_async_commit();
}
// make smem operands visible to the tensor core, issue one 128x128x128 GEMM, wait for it
template <int PENDING>
__device__ __forceinline__ void run_gemm(Shared& sh, uint32_t acc_col, const uint8_t* A, const uint8_t* W, bool accumulate,
                                         uint32_t& phase) {
    cp_async_wait<PENDING>();
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after();
        issue_gemm_kk(sh.tmem + acc_col, smem_u32(A), smem_u32(W), accumulate);   // sh.tmem is valid after the barrier
        mma_commit(&sh.bar);
    }
    mbar_wait(&sh.bar, phase);
    phase ^= 1;
    fence_after();
}
// sum of a per-thread partial over the two threads that share a row (fixed order)
__device__ __forceinline__ float row_sum(Shared& sh, const Epi& e, int slot, float partial) {
    sh.xch[slot][e.cb >> 6][e.row] = partial;
    __syncthreads();
    const float t = sh.xch[slot][0][e.row] + sh.xch[slot][1][e.row];
    return t;
}
__device__ __forceinline__ void st_tile4(uint8_t* tile, int r, int c, float4 v) {
    *reinterpret_cast<float4*>(tile + tile_off4(r, c >> 2)) = v;
}
__device__ __forceinline__ float4 ld_tile4(const uint8_t* tile, int r, int c) {
    return *reinterpret_cast<const float4*>(tile + tile_off4(r, c >> 2));
}
// 32 lanes x 32 columns -> lane l ends with the column-l sum over the 32 lanes (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float recv = __shfl_xor_sync(0xffffffffu, send, s);
            v[i] = (up ? v[i + s] : v[i]) + recv;
        }
    }
    return v[0];
}
// per-tile LayerNorm parameter-gradient partials: lnacc[warp] -> part[tile][dw 128 | db 128]
__device__ __forceinline__ void flush_ln_partials(Shared& sh, float* __restrict__ part_tile) {
    __syncthreads();
    const int t = threadIdx.x, arr = t >> 7, c = t & 127, hb = c >> 6, cc = c & 63;
    float s = 0.f;
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) s += sh.lnacc[hb * 4 + rg][arr][cc];
    part_tile[arr * D + c] = s;
}

// ----------------------------------------------------------------------------------------------
// forward 1: k, v from x; LN1 in place; q from LN1(x)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_ln_qkv_tc(const float* __restrict__ x, int M, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
            const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
            const float* __restrict__ in_b, float* __restrict__ qn, float* __restrict__ st1, float* __restrict__ q,
            float* __restrict__ k, float* __restrict__ v) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* A = align1k(smem_raw);
    uint8_t* Wb[2] = {A + TILE_BYTES, A + 2 * TILE_BYTES};
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(Wb[0], Wk);
    fill_tile(A, x, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    uint32_t phase = 0;
    // ---- k and v from the raw tile
    float* outs[2] = {k, v};
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        load_w_async(Wb[(g + 1) & 1], g == 0 ? Wv : Wq);
        run_gemm<1>(sh, 0, A, Wb[g & 1], false, phase);
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            if (valid) {
                float* dst = outs[g] + (size_t)gr * D + c0;
                const float* bb = in_b + (g + 1) * D + c0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bb + i));
                    *reinterpret_cast<float4*>(dst + i) = make_float4(a[i] + b4.x, a[i + 1] + b4.y, a[i + 2] + b4.z, a[i + 3] + b4.w);
                }
            }
        }
    }
    // ---- LN1 in place on the tile (the v GEMM has completed, the tile is free)
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) { const float4 t = ld_tile4(A, e.row, e.cb + i); s += (t.x + t.y) + (t.z + t.w); }
        const float mean = row_sum(sh, e, 0, s) * (1.0f / D);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 t = ld_tile4(A, e.row, e.cb + i);
            const float a0 = t.x - mean, a1 = t.y - mean, a2 = t.z - mean, a3 = t.w - mean;
            ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
        const float rstd = 1.0f / sqrtf(row_sum(sh, e, 1, ss) * (1.0f / D) + LN_EPS);
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 t = ld_tile4(A, e.row, e.cb + i);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln_w + e.cb + i));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ln_b + e.cb + i));
            const float4 y = make_float4(fmaf((t.x - mean) * rstd, w4.x, b4.x), fmaf((t.y - mean) * rstd, w4.y, b4.y),
                                         fmaf((t.z - mean) * rstd, w4.z, b4.z), fmaf((t.w - mean) * rstd, w4.w, b4.w));
            st_tile4(A, e.row, e.cb + i, y);
            if (valid) *reinterpret_cast<float4*>(qn + (size_t)gr * D + e.cb + i) = y;
        }
        if (valid && e.cb == 0) { st1[(size_t)gr * 2] = mean; st1[(size_t)gr * 2 + 1] = rstd; }
    }
    // ---- q = 0.25 (LN1(x) Wq^T + bq)
    run_gemm<0>(sh, 0, A, Wb[0], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            if (valid) {
                float* dst = q + (size_t)gr * D + c0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(in_b + c0 + i));
                    *reinterpret_cast<float4*>(dst + i) = make_float4((a[i] + b4.x) * 0.25f, (a[i + 1] + b4.y) * 0.25f,
                                                                      (a[i + 2] + b4.z) * 0.25f, (a[i + 3] + b4.w) * 0.25f);
                }
            }
        }
    }
    teardown(sh, sh.tmem, 128);
}

// LayerNorm of a row held as 2 x 32 registers by this thread and 64 more by its partner
__device__ __forceinline__ void ln_rows64(Shared& sh, const Epi& e, float (&xr)[64], const float* __restrict__ w,
                                          const float* __restrict__ b, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += xr[i];
    mean = row_sum(sh, e, 0, s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) { const float a = xr[i] - mean; ss = fmaf(a, a, ss); }
    rstd = 1.0f / sqrtf(row_sum(sh, e, 1, ss) * (1.0f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < 64; ++i) xr[i] = fmaf((xr[i] - mean) * rstd, __ldg(w + e.cb + i), __ldg(b + e.cb + i));
}

// ----------------------------------------------------------------------------------------------
// forward 2: out-proj + residual + LN2 + FFN + mask (+ last LN)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_proj_ffn_tc(const float* __restrict__ o, const float* __restrict__ qn, int M, const float* __restrict__ Wo,
              const float* __restrict__ bo, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
              const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
              const float* __restrict__ b2, const uint32_t* __restrict__ tmask, DropCfg dc, uint32_t site1, uint32_t site2,
              float* __restrict__ x1, float* __restrict__ st2, float* __restrict__ y, float* __restrict__ h,
              float* __restrict__ xout, const float* __restrict__ ln3_w, const float* __restrict__ ln3_b,
              float* __restrict__ enc, float* __restrict__ st3) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* A = align1k(smem_raw);
    uint8_t* Wb[2] = {A + TILE_BYTES, A + 2 * TILE_BYTES};
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(Wb[0], Wo);
    fill_tile(A, o, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    uint32_t phase = 0;
    // ---- x1 = Qn + o Wo^T + bo ; y = LN2(x1)
    load_w_async(Wb[1], W1);
    run_gemm<1>(sh, 0, A, Wb[0], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            tmem_ld32(tm + e.cb + half * 32, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[half * 32 + i] = a[i];
        }
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bo + e.cb + i));
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) r4 = __ldg(reinterpret_cast<const float4*>(qn + (size_t)gr * D + e.cb + i));
            xr[i] += b4.x + r4.x; xr[i + 1] += b4.y + r4.y; xr[i + 2] += b4.z + r4.z; xr[i + 3] += b4.w + r4.w;
            if (valid) *reinterpret_cast<float4*>(x1 + (size_t)gr * D + e.cb + i) = make_float4(xr[i], xr[i + 1], xr[i + 2], xr[i + 3]);
        }
        float mean, rstd;
        ln_rows64(sh, e, xr, ln2_w, ln2_b, mean, rstd);
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 t = make_float4(xr[i], xr[i + 1], xr[i + 2], xr[i + 3]);
            st_tile4(A, e.row, e.cb + i, t);
            if (valid) *reinterpret_cast<float4*>(y + (size_t)gr * D + e.cb + i) = t;
        }
        if (valid && e.cb == 0) { st2[(size_t)gr * 2] = mean; st2[(size_t)gr * 2 + 1] = rstd; }
    }
    // ---- h = relu(dropout1(y W1^T + b1))
    load_w_async(Wb[0], W2);
    run_gemm<1>(sh, 0, A, Wb[1], false, phase);
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
                if (dc.train) t = drop4(t, dc, site1, (uint64_t)gr * D + c0 + i);
                t = make_float4(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f), fmaxf(t.z, 0.f), fmaxf(t.w, 0.f));
                st_tile4(A, e.row, c0 + i, t);
                if (valid) *reinterpret_cast<float4*>(h + (size_t)gr * D + c0 + i) = t;
            }
        }
    }
    // ---- xout = (dropout2(h W2^T + b2) + y) * ~tmask  (+ last LayerNorm)
    run_gemm<0>(sh, 0, A, Wb[0], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float xr[64];
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + c0 + i));
                float4 t = make_float4(a[i] + b4.x, a[i + 1] + b4.y, a[i + 2] + b4.z, a[i + 3] + b4.w);
                if (dc.train) t = drop4(t, dc, site2, (uint64_t)gr * D + c0 + i);
                float4 yy = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) yy = *reinterpret_cast<const float4*>(y + (size_t)gr * D + c0 + i);
                t = make_float4(t.x + yy.x, t.y + yy.y, t.z + yy.z, t.w + yy.w);
                t = apply_tmask(t, tw, (c0 + i) >> 2);
                if (valid) *reinterpret_cast<float4*>(xout + (size_t)gr * D + c0 + i) = t;
                xr[half * 32 + i] = t.x; xr[half * 32 + i + 1] = t.y; xr[half * 32 + i + 2] = t.z; xr[half * 32 + i + 3] = t.w;
            }
        }
        if (enc) {   // uniform
            float mean, rstd;
            ln_rows64(sh, e, xr, ln3_w, ln3_b, mean, rstd);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 64; i += 4)
                    *reinterpret_cast<float4*>(enc + (size_t)gr * D + e.cb + i) = make_float4(xr[i], xr[i + 1], xr[i + 2], xr[i + 3]);
                if (e.cb == 0) { st3[(size_t)gr * 2] = mean; st3[(size_t)gr * 2 + 1] = rstd; }
            }
        }
    }
    teardown(sh, sh.tmem, 128);
}

// ----------------------------------------------------------------------------------------------
// backward 1: FFN + LN2 + out-proj input gradient (weights passed TRANSPOSED: Wt[k_in][n_out])
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_ffn_bwd_tc(const float* __restrict__ dxo, const float* __restrict__ h, const float* __restrict__ x1,
             const float* __restrict__ st2, const uint32_t* __restrict__ tmask, int M, const float* __restrict__ W2t,
             const float* __restrict__ W1t, const float* __restrict__ Wot, const float* __restrict__ ln2_w, DropCfg dc,
             uint32_t site1, uint32_t site2, float* __restrict__ do2, float* __restrict__ dhpre, float* __restrict__ dx1,
             float* __restrict__ dO, float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* A = align1k(smem_raw);
    uint8_t* Wb[2] = {A + TILE_BYTES, A + 2 * TILE_BYTES};
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(Wb[0], W2t);
    // A = do2 = dropout2-mask * (dxo * ~tmask)
#pragma unroll 2
    for (int idx = threadIdx.x; idx < 128 * 32; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31, grr = row0 + r;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grr < M) {
            g = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)grr * D) + c4);
            g = apply_tmask(g, __ldg(reinterpret_cast<const uint4*>(tmask) + grr), c4);
            if (dc.train) g = drop4(g, dc, site2, (uint64_t)grr * D + c4 * 4);
            *(reinterpret_cast<float4*>(do2 + (size_t)grr * D) + c4) = g;
        }
        *reinterpret_cast<float4*>(A + tile_off4(r, c4)) = g;
    }
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    uint32_t phase = 0;
    const float sc = dc.train ? dc.scale : 1.0f;
    // ---- dhpre = (do2 W2) * scale * [h > 0]
    load_w_async(Wb[1], W1t);
    run_gemm<1>(sh, 0, A, Wb[0], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float4 hh = __ldg(reinterpret_cast<const float4*>(h + (size_t)gr * D + c0 + i));
                    t = make_float4(hh.x > 0.f ? a[i] * sc : 0.f, hh.y > 0.f ? a[i + 1] * sc : 0.f,
                                    hh.z > 0.f ? a[i + 2] * sc : 0.f, hh.w > 0.f ? a[i + 3] * sc : 0.f);
                    *reinterpret_cast<float4*>(dhpre + (size_t)gr * D + c0 + i) = t;
                }
                st_tile4(A, e.row, c0 + i, t);
            }
        }
    }
    // ---- dy = dhpre W1 + g ; LN2 backward -> dx1
    load_w_async(Wb[0], Wot);
    run_gemm<1>(sh, 0, A, Wb[1], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) { mean = st2[(size_t)gr * 2]; rstd = st2[(size_t)gr * 2 + 1]; tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr); }
        float p1 = 0.f, p2 = 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {     // pass 1: row reductions
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 g = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)gr * D + c0 + i));
                    g = apply_tmask(g, tw, (c0 + i) >> 2);
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)gr * D + c0 + i));
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln2_w + c0 + i));
                    const float d0 = (a[i] + g.x) * w4.x, d1 = (a[i + 1] + g.y) * w4.y, d2 = (a[i + 2] + g.z) * w4.z, d3 = (a[i + 3] + g.w) * w4.w;
                    p1 += (d0 + d1) + (d2 + d3);
                    p2 += d0 * ((xv.x - mean) * rstd) + d1 * ((xv.y - mean) * rstd) + d2 * ((xv.z - mean) * rstd) + d3 * ((xv.w - mean) * rstd);
                }
            }
        }
        const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
        const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {     // pass 2: dx1 and the LN parameter partials
            float a[32], pw[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 dyv = dx, xh = dx;
                if (valid) {
                    float4 g = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)gr * D + c0 + i));
                    g = apply_tmask(g, tw, (c0 + i) >> 2);
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)gr * D + c0 + i));
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln2_w + c0 + i));
                    dyv = make_float4(a[i] + g.x, a[i + 1] + g.y, a[i + 2] + g.z, a[i + 3] + g.w);
                    xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                    dx = make_float4(rstd * (dyv.x * w4.x - c1 - xh.x * c2), rstd * (dyv.y * w4.y - c1 - xh.y * c2),
                                     rstd * (dyv.z * w4.z - c1 - xh.z * c2), rstd * (dyv.w * w4.w - c1 - xh.w * c2));
                    *reinterpret_cast<float4*>(dx1 + (size_t)gr * D + c0 + i) = dx;
                }
                st_tile4(A, e.row, c0 + i, dx);
                a[i] = dyv.x; a[i + 1] = dyv.y; a[i + 2] = dyv.z; a[i + 3] = dyv.w;                 // db terms
                pw[i] = dyv.x * xh.x; pw[i + 1] = dyv.y * xh.y; pw[i + 2] = dyv.z * xh.z; pw[i + 3] = dyv.w * xh.w;   // dw terms
            }
            const float sw = warp_colsum32(pw, e.lane);
            const float sb = warp_colsum32(a, e.lane);
            sh.lnacc[e.warp][0][half * 32 + e.lane] = sw;
            sh.lnacc[e.warp][1][half * 32 + e.lane] = sb;
        }
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    // ---- dO = dx1 Wo
    run_gemm<0>(sh, 0, A, Wb[0], false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dO + (size_t)gr * D + c0 + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
            }
        }
    }
    teardown(sh, sh.tmem, 128);
}

// ----------------------------------------------------------------------------------------------
// backward 2: dQn = dx1 + dq Wq ; dx_in = dk Wk + dv Wv + LN1bwd(dQn)   (weights transposed)
// two accumulators in TMEM: [0,128) = dq Wq, [128,256) = dk Wk + dv Wv
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_qkv_bwd_tc(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
             const float* __restrict__ dx1, const float* __restrict__ xin, const float* __restrict__ st1, int M,
             const float* __restrict__ Wqt, const float* __restrict__ Wkt, const float* __restrict__ Wvt,
             const float* __restrict__ ln1_w, float* __restrict__ dxin, float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* A = align1k(smem_raw);
    uint8_t* Wb[2] = {A + TILE_BYTES, A + 2 * TILE_BYTES};
    const int row0 = blockIdx.x * 128;
    setup(sh, 256);
    load_w_async(Wb[0], Wqt);
    fill_tile(A, dq, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    uint32_t phase = 0;
    load_w_async(Wb[1], Wkt);
    run_gemm<1>(sh, 0, A, Wb[0], false, phase);
    fill_tile(A, dk, row0, M);
    load_w_async(Wb[0], Wvt);
    run_gemm<1>(sh, 128, A, Wb[1], false, phase);
    fill_tile(A, dv, row0, M);
    run_gemm<0>(sh, 128, A, Wb[0], true, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        if (valid) { mean = st1[(size_t)gr * 2]; rstd = st1[(size_t)gr * 2 + 1]; }
        float p1 = 0.f, p2 = 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(dx1 + (size_t)gr * D + c0 + i));
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(xin + (size_t)gr * D + c0 + i));
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln1_w + c0 + i));
                    const float d0 = (a[i] + r4.x) * w4.x, d1 = (a[i + 1] + r4.y) * w4.y, d2 = (a[i + 2] + r4.z) * w4.z, d3 = (a[i + 3] + r4.w) * w4.w;
                    p1 += (d0 + d1) + (d2 + d3);
                    p2 += d0 * ((xv.x - mean) * rstd) + d1 * ((xv.y - mean) * rstd) + d2 * ((xv.z - mean) * rstd) + d3 * ((xv.w - mean) * rstd);
                }
            }
        }
        const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
        const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32], pw[32], base[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            tmem_ld32(tm + 128 + c0, base);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 dyv = make_float4(0.f, 0.f, 0.f, 0.f), xh = dyv;
                if (valid) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(dx1 + (size_t)gr * D + c0 + i));
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(xin + (size_t)gr * D + c0 + i));
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln1_w + c0 + i));
                    dyv = make_float4(a[i] + r4.x, a[i + 1] + r4.y, a[i + 2] + r4.z, a[i + 3] + r4.w);
                    xh = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                    *reinterpret_cast<float4*>(dxin + (size_t)gr * D + c0 + i) =
                        make_float4(base[i] + rstd * (dyv.x * w4.x - c1 - xh.x * c2), base[i + 1] + rstd * (dyv.y * w4.y - c1 - xh.y * c2),
                                    base[i + 2] + rstd * (dyv.z * w4.z - c1 - xh.z * c2), base[i + 3] + rstd * (dyv.w * w4.w - c1 - xh.w * c2));
                }
                a[i] = dyv.x; a[i + 1] = dyv.y; a[i + 2] = dyv.z; a[i + 3] = dyv.w;
                pw[i] = dyv.x * xh.x; pw[i + 1] = dyv.y * xh.y; pw[i + 2] = dyv.z * xh.z; pw[i + 3] = dyv.w * xh.w;
            }
            const float sw = warp_colsum32(pw, e.lane);
            const float sb = warp_colsum32(a, e.lane);
            sh.lnacc[e.warp][0][half * 32 + e.lane] = sw;
            sh.lnacc[e.warp][1][half * 32 + e.lane] = sb;
        }
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    teardown(sh, sh.tmem, 256);
}

// ----------------------------------------------------------------------------------------------
// weight gradients on the tensor core: dW[n][k] = sum_m dY[m][n] X[m][k], db[n] = sum_m dY[m][n].
// Both operands are staged TRANSPOSED ([n][m] and [k][m], K-major with K = tokens); the bias
// gradient is one more MMA against a tile of ones.  The accumulators stay in TMEM across all the
// token tiles of a CTA; grid = (S, 6 jobs).
// ----------------------------------------------------------------------------------------------
struct WgradJobsTc {
    const float* dY[6];
    const float* X[6];
};
__device__ __forceinline__ void fill_tile_transposed(uint8_t* tile, const float* __restrict__ g, int row0, int M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = 32 * (warp & 3) + lane;
    const int cbeg = 16 * (warp >> 2);
    const bool ok = row0 + m < M;
    const float4* src = reinterpret_cast<const float4*>(g + (size_t)(row0 + m) * D);
#pragma unroll 4
    for (int c4 = cbeg; c4 < cbeg + 16; ++c4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(src + c4);
        *reinterpret_cast<float*>(tile + tile_off(4 * c4 + 0, m)) = v.x;
        *reinterpret_cast<float*>(tile + tile_off(4 * c4 + 1, m)) = v.y;
        *reinterpret_cast<float*>(tile + tile_off(4 * c4 + 2, m)) = v.z;
        *reinterpret_cast<float*>(tile + tile_off(4 * c4 + 3, m)) = v.w;
    }
}
__global__ void __launch_bounds__(256, 1)
k_wgrad_tc(WgradJobsTc jobs, int M, float* __restrict__ wpart /*[6][S][128*128]*/, float* __restrict__ bpart /*[6][S][128]*/) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* At = align1k(smem_raw);
    uint8_t* Xt = At + TILE_BYTES;
    uint8_t* Ones = At + 2 * TILE_BYTES;      // [16][128] K-major: 4 chunks x (16 rows x 128 B)
    const float* __restrict__ dY = jobs.dY[blockIdx.y];
    const float* __restrict__ X = jobs.X[blockIdx.y];
    const int S = gridDim.x;
    setup(sh, 256);
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += 256) reinterpret_cast<float*>(Ones)[i] = 1.0f;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = sh.tmem;
    const int tiles = (M + 127) / 128;
    uint32_t phase = 0;
    bool first = true;
    constexpr uint32_t id_w = idesc_tf32(128, false, false);
    constexpr uint32_t id_b = idesc_tf32(16, false, false);
    for (int t = blockIdx.x; t < tiles; t += S) {
        fill_tile_transposed(At, dY, t * 128, M);
        fill_tile_transposed(Xt, X, t * 128, M);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
            const uint32_t a = smem_u32(At), x = smem_u32(Xt), o = smem_u32(Ones);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t acc = (!first || c || ks) ? 1u : 0u;
                    mma_tf32(tmem, desc_kmajor(a, c, ks), desc_kmajor(x, c, ks), id_w, acc);
                    mma_tf32(tmem + 128, desc_kmajor(a, c, ks), make_desc(o + c * 2048 + ks * 32, 16, 1024), id_b, acc);
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
            tmem_ld32(tmem + e.lane_addr + 128, a);    // 16 valid columns, all equal to db[row]
        } else {
            a[0] = 0.f;
        }
        bpart[((size_t)blockIdx.y * S + blockIdx.x) * D + e.row] = a[0];
    }
    teardown(sh, tmem, 256);
}

}  // namespace tcenc
}  // namespace amid
