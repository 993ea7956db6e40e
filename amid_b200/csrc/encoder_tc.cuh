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

constexpr int WSTAGE_FLOATS = 32 * 32;                                     // per-warp transposing stage (4 KB)
constexpr size_t CHAIN_SMEM = 2 * (size_t)TILE_BYTES + 8 * WSTAGE_FLOATS * 4 + 1024;   // A tile + weight tile + 8 warp stages
constexpr int ONES_BYTES = 16 * 128 * 4;                                   // [16 rows][128] K-major
constexpr size_t WGRAD_SMEM = 2 * (size_t)TILE_BYTES + ONES_BYTES + 1024;

struct Shared {
    uint64_t bar;
    uint32_t tmem;
    float xch[2][2][128];       // [value][half][row]
    float lnacc[8][2][64];      // [warp][dw|db][col in half]
};

struct Epi {
    int warp, lane, row, cb, wrow0;
    uint32_t lane_addr;
    __device__ Epi() {
        warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
        wrow0 = 32 * (warp & 3);
        row = wrow0 + lane; cb = 64 * (warp >> 2);
        lane_addr = (uint32_t)wrow0 << 16;
    }
};

// (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the shared address
// space and emits LDS / STS instead of generic LD / ST for everything derived from the result)
__device__ __forceinline__ uint8_t* align1k(uint8_t* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    return p + ((1024u - (a & 1023u)) & 1023u);
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
__device__ __forceinline__ void run_gemm(Shared& sh, uint32_t acc_col, const uint8_t* A, const uint8_t* W, bool accumulate,
                                         uint32_t& phase) {
    cp_async_wait<0>();
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
// ---- coalesced global I/O for the thread-per-row epilogues: a 32x32 chunk (this warp's 32 rows x
// 32 columns) goes through a warp-private swizzled stage so that the global side is 128-byte
// row segments (4 rows per instruction) while the register side stays one row per thread.
__device__ __forceinline__ int ws_off(int r, int u) { return r * 32 + ((u ^ (r & 7)) << 2); }
// registers (lane = row) -> global: g points at (first row of the warp, first column of the chunk)
__device__ __forceinline__ void warp_store32(float* buf, int lane, const float (&v)[32], float* __restrict__ g, int rows_valid) {
#pragma unroll
    for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(buf + ws_off(lane, u)) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), u = lane & 7;
        const float4 x = *reinterpret_cast<const float4*>(buf + ws_off(r, u));
        if (r < rows_valid) *reinterpret_cast<float4*>(g + (size_t)r * D + u * 4) = x;
    }
    __syncwarp();
}
// global -> registers (lane = row); rows >= rows_valid read as zero
__device__ __forceinline__ void warp_load32(float* buf, int lane, const float* __restrict__ g, int rows_valid, float (&v)[32]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), u = lane & 7;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid) x = *reinterpret_cast<const float4*>(g + (size_t)r * D + u * 4);
        *reinterpret_cast<float4*>(buf + ws_off(r, u)) = x;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(buf + ws_off(lane, u));
        v[4 * u] = x.x; v[4 * u + 1] = x.y; v[4 * u + 2] = x.z; v[4 * u + 3] = x.w;
    }
    __syncwarp();
}
// the same for data written earlier in this kernel by other threads (coherent L2 loads)
__device__ __forceinline__ void warp_load32_cg(float* buf, int lane, const float* g, int rows_valid, float (&v)[32]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), u = lane & 7;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid) x = __ldcg(reinterpret_cast<const float4*>(g + (size_t)r * D + u * 4));
        *reinterpret_cast<float4*>(buf + ws_off(r, u)) = x;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(buf + ws_off(lane, u));
        v[4 * u] = x.x; v[4 * u + 1] = x.y; v[4 * u + 2] = x.z; v[4 * u + 3] = x.w;
    }
    __syncwarp();
}
// registers (lane = row) -> the operand tile A, columns [c0, c0+32)
__device__ __forceinline__ void tile_store32(uint8_t* A, int row, int c0, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(A + tile_off4(row, (c0 + i) >> 2)) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
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
// Shared structure of the four chain kernels
//   smem: A (operand tile) | W (weight tile) | 8 warp stages.  The next stage's weight is fetched
//   with cp.async right after the current MMA has completed, so it overlaps the epilogue.
// ----------------------------------------------------------------------------------------------
struct ChainSmem {
    uint8_t* A;
    uint8_t* W;
    float* stage;     // this warp's 32x32 transposing stage
    __device__ ChainSmem(uint8_t* raw) {
        A = align1k(raw);
        W = A + TILE_BYTES;
        stage = reinterpret_cast<float*>(W + TILE_BYTES) + (threadIdx.x >> 5) * WSTAGE_FLOATS;
    }
};
__device__ __forceinline__ int rows_valid(int row0, const Epi& e, int M) { return max(0, min(32, M - (row0 + e.wrow0))); }

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
    ChainSmem sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(sm.W, Wk);
    fill_tile(sm.A, x, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    float* outs[2] = {k, v};
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {        // k and v from the raw tile
        run_gemm(sh, 0, sm.A, sm.W, false, phase);
        load_w_async(sm.W, g == 0 ? Wv : Wq);
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
    // LN1 in place on the tile (the v GEMM has completed, the tile is free)
    float mean, rstd;
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) { const float4 t = ld_tile4(sm.A, e.row, e.cb + i); s += (t.x + t.y) + (t.z + t.w); }
        mean = row_sum(sh, e, 0, s) * (1.0f / D);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 t = ld_tile4(sm.A, e.row, e.cb + i);
            const float a0 = t.x - mean, a1 = t.y - mean, a2 = t.z - mean, a3 = t.w - mean;
            ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
        rstd = 1.0f / sqrtf(row_sum(sh, e, 1, ss) * (1.0f / D) + LN_EPS);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 t = ld_tile4(sm.A, e.row, c0 + i);
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(ln_w + c0 + i));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + i));
                a[i] = fmaf((t.x - mean) * rstd, w4.x, b4.x); a[i + 1] = fmaf((t.y - mean) * rstd, w4.y, b4.y);
                a[i + 2] = fmaf((t.z - mean) * rstd, w4.z, b4.z); a[i + 3] = fmaf((t.w - mean) * rstd, w4.w, b4.w);
            }
            tile_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, qn + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st1[(size_t)gr * 2] = mean; st1[(size_t)gr * 2 + 1] = rstd; }
    }
    run_gemm(sh, 0, sm.A, sm.W, false, phase);      // q = 0.25 (LN1(x) Wq^T + bq)
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

// LayerNorm statistics of a row held as 64 registers by this thread and 64 more by its partner
__device__ __forceinline__ void ln_stats64(Shared& sh, const Epi& e, const float (&xr)[64], float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += xr[i];
    mean = row_sum(sh, e, 0, s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) { const float a = xr[i] - mean; ss = fmaf(a, a, ss); }
    rstd = 1.0f / sqrtf(row_sum(sh, e, 1, ss) * (1.0f / D) + LN_EPS);
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
    ChainSmem sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(sm.W, Wo);
    fill_tile(sm.A, o, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    float yr[64];                                   // LN2 output of this thread's half row, kept for the residual
    // ---- x1 = Qn + o Wo^T + bo ; y = LN2(x1)
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    load_w_async(sm.W, W1);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
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
            for (int i = 0; i < 32; ++i) yr[half * 32 + i] = a[i];
        }
        float mean, rstd;
        ln_stats64(sh, e, yr, mean, rstd);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                a[i] = fmaf((yr[half * 32 + i] - mean) * rstd, __ldg(ln2_w + c0 + i), __ldg(ln2_b + c0 + i));
                yr[half * 32 + i] = a[i];
            }
            tile_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, y + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st2[(size_t)gr * 2] = mean; st2[(size_t)gr * 2 + 1] = rstd; }
    }
    // ---- h = relu(dropout1(y W1^T + b1))
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    load_w_async(sm.W, W2);
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
            tile_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, h + wbase + c0, rv);
        }
    }
    // ---- xout = (dropout2(h W2^T + b2) + y) * ~tmask  (+ last LayerNorm)
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
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
                if (dc.train) t = drop4(t, dc, site2, (uint64_t)(gr + dc.tok_off) * D + c0 + i);
                t = make_float4(t.x + yr[half * 32 + i], t.y + yr[half * 32 + i + 1], t.z + yr[half * 32 + i + 2], t.w + yr[half * 32 + i + 3]);
                t = apply_tmask(t, tw, (c0 + i) >> 2);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
            warp_store32(sm.stage, e.lane, a, xout + wbase + c0, rv);
#pragma unroll
            for (int i = 0; i < 32; ++i) yr[half * 32 + i] = a[i];
        }
        if (enc) {   // uniform: last_layernorm (model_seq.py:385)
            float mean, rstd;
            ln_stats64(sh, e, yr, mean, rstd);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float a[32];
                const int c0 = e.cb + half * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = fmaf((yr[half * 32 + i] - mean) * rstd, __ldg(ln3_w + c0 + i), __ldg(ln3_b + c0 + i));
                warp_store32(sm.stage, e.lane, a, enc + wbase + c0, rv);
            }
            if (valid && e.cb == 0) { st3[(size_t)gr * 2] = mean; st3[(size_t)gr * 2 + 1] = rstd; }
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
    ChainSmem sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 128);
    load_w_async(sm.W, W2t);
    // A = do2 = dropout2-mask * (dxo * ~tmask)   (warp per row: coalesced)
#pragma unroll 2
    for (int idx = threadIdx.x; idx < 128 * 32; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31, grr = row0 + r;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grr < M) {
            g = __ldg(reinterpret_cast<const float4*>(dxo + (size_t)grr * D) + c4);
            g = apply_tmask(g, __ldg(reinterpret_cast<const uint4*>(tmask) + grr), c4);
            if (dc.train) g = drop4(g, dc, site2, (uint64_t)(grr + dc.tok_off) * D + c4 * 4);
            *(reinterpret_cast<float4*>(do2 + (size_t)grr * D) + c4) = g;
        }
        *reinterpret_cast<float4*>(sm.A + tile_off4(r, c4)) = g;
    }
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    const float sc = dc.train ? dc.scale : 1.0f;
    // ---- dhpre = (do2 W2) * scale * [h > 0]
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    load_w_async(sm.W, W1t);
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
            tile_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, dhpre + wbase + c0, rv);
        }
    }
    // ---- dy = dhpre W1 + g ; LN2 backward -> dx1
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    load_w_async(sm.W, Wot);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) { mean = st2[(size_t)gr * 2]; rstd = st2[(size_t)gr * 2 + 1]; tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr); }
        float dy[64], xh[64];
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], g[32], xv[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_load32(sm.stage, e.lane, dxo + wbase + c0, rv, g);
            warp_load32(sm.stage, e.lane, x1 + wbase + c0, rv, xv);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 gm = apply_tmask(make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]), tw, (c0 + i) >> 2);
                const float gg[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float d = a[i + u] + gg[u];
                    const float xx = valid ? (xv[i + u] - mean) * rstd : 0.f;
                    const float dw = d * __ldg(ln2_w + c0 + i + u);
                    dy[half * 32 + i + u] = d;
                    xh[half * 32 + i + u] = xx;
                    p1 += dw;
                    p2 = fmaf(dw, xx, p2);
                }
            }
        }
        const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
        const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], pw[32], pb[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = dy[half * 32 + i], xx = xh[half * 32 + i];
                a[i] = valid ? rstd * (d * __ldg(ln2_w + c0 + i) - c1 - xx * c2) : 0.f;
                pw[i] = d * xx;
                pb[i] = d;
            }
            tile_store32(sm.A, e.row, c0, a);
            warp_store32(sm.stage, e.lane, a, dx1 + wbase + c0, rv);
            const float sw = warp_colsum32(pw, e.lane);
            const float sb = warp_colsum32(pb, e.lane);
            sh.lnacc[e.warp][0][half * 32 + e.lane] = sw;
            sh.lnacc[e.warp][1][half * 32 + e.lane] = sb;
        }
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    // ---- dO = dx1 Wo
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
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
    ChainSmem sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    setup(sh, 256);
    load_w_async(sm.W, Wqt);
    fill_tile(sm.A, dq, row0, M);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    uint32_t phase = 0;
    run_gemm(sh, 0, sm.A, sm.W, false, phase);
    load_w_async(sm.W, Wkt);
    fill_tile(sm.A, dk, row0, M);
    run_gemm(sh, 128, sm.A, sm.W, false, phase);
    load_w_async(sm.W, Wvt);
    fill_tile(sm.A, dv, row0, M);
    run_gemm(sh, 128, sm.A, sm.W, true, phase);
    {
        const uint32_t tm = sh.tmem + e.lane_addr;
        float mean = 0.f, rstd = 0.f;
        if (valid) { mean = st1[(size_t)gr * 2]; rstd = st1[(size_t)gr * 2 + 1]; }
        float dy[64], xh[64];
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], r[32], xv[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + c0, a);
            warp_load32(sm.stage, e.lane, dx1 + wbase + c0, rv, r);
            warp_load32(sm.stage, e.lane, xin + wbase + c0, rv, xv);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = a[i] + r[i];
                const float xx = valid ? (xv[i] - mean) * rstd : 0.f;
                const float dw = d * __ldg(ln1_w + c0 + i);
                dy[half * 32 + i] = d;
                xh[half * 32 + i] = xx;
                p1 += dw;
                p2 = fmaf(dw, xx, p2);
            }
        }
        const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
        const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], pw[32], pb[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tm + 128 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = dy[half * 32 + i], xx = xh[half * 32 + i];
                a[i] += valid ? rstd * (d * __ldg(ln1_w + c0 + i) - c1 - xx * c2) : 0.f;
                pw[i] = d * xx;
                pb[i] = d;
            }
            warp_store32(sm.stage, e.lane, a, dxin + wbase + c0, rv);
            const float sw = warp_colsum32(pw, e.lane);
            const float sb = warp_colsum32(pb, e.lane);
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
