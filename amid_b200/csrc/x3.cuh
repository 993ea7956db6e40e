// Split-operand tensor-core arithmetic at fp32-level accuracy (precision = "x3", the parity-grade default).
//
// tcgen05 has no fp32 MMA.  An fp32 operand x is therefore carried as a short sum of 16-bit pieces and the
// product of two operands as the sum of the piece products that matter, accumulated in fp32 in TMEM:
//
//   chain GEMMs (token tile x weight):  x = s^-1 (h0 + h1), h = FP16, s = a power of two chosen per ROW of the
//     token tile (row max -> [2^14, 2^15)) and per weight matrix.  22 mantissa bits per operand; the products
//     h0 w0 + h1 w0 + h0 w1 leave a relative error of 2^-22 (the dropped h1 w1 term), i.e. fp32 level.  Three
//     kind::f16 MMAs at the full 16-bit rate cost half of a 3xTF32 split and the pieces are half the bytes, which is
//     what lets TWO CTAs share an SM: the token-tile pieces live in TENSOR MEMORY (64 columns each, written by the
//     epilogue threads with tcgen05.st and read by tcgen05.mma as the A operand), the weight pieces arrive in
//     shared memory as ONE bulk copy (cp.async.bulk, completion on an mbarrier) of an image that the weight prep
//     kernel stored pre-swizzled in the canonical K-major SWIZZLE_128B layout.
//   weight gradients (tokens are the contraction index, so no per-row scale can be factored out):
//     x = b0 + b1 + b2, b = BF16 (full fp32 range, 24 bits), six products b0b0 + b0b1 + b1b0 + b0b2 + b1b1 + b2b0.
//
// TMEM map of a chain CTA (256 columns): [0,128) fp32 accumulator, [128,192) A piece 0, [192,256) A piece 1
// (row = lane, two consecutive k per 32-bit column).
#pragma once
#include <cuda_fp16.h>

#include "encoder_tc16.cuh"

namespace amid {
namespace x3 {
using namespace tc;
using tcenc::align1k;
using tcenc::Epi;
using tcenc::rows_valid;
using tcenc::warp_colsum32;
using tcenc::warp_load32;
using tcenc::warp_load32_cg;
using tcenc::warp_store32;
using tcenc::WSTAGE_FLOATS;

constexpr int PIECE_BYTES = TILE16_BYTES;            // one 16-bit piece of a [128][128] tile: 32 KB
constexpr int WIMG_BYTES = 2 * PIECE_BYTES;          // both pieces of a weight: 64 KB, contiguous in global memory
constexpr uint32_t ACC_COL = 0, A0_COL = 128, A1_COL = 192;
constexpr int CHAIN_TMEM_COLS = 256;
constexpr size_t CHAINX_SMEM = (size_t)WIMG_BYTES + 8 * WSTAGE_FLOATS * 4 + 1024;   // 97 KB -> two CTAs per SM

// ---- instruction descriptor: kind::f16 with FP16 operands, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one lane of a converged warp (the pattern ptxas recognises: MMAs behind it issue without a per-instruction election loop)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
    return pred != 0;
}
// rows [row0, row0+128) of a [M,128] fp32 tensor into L2 (512 lines of 128 B, two per thread of a 256-thread CTA), issued
// right before a GEMM wait for the tile the NEXT stage of the chain reads: that load then pays an L2 hit instead of an HBM
// round trip in the middle of the chain.  (Prefetching everything at kernel start was measured slower: the requests compete
// with the first demand loads.)
__device__ __forceinline__ void prefetch_tile_l2(const float* __restrict__ g, int row0, int M) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int idx = threadIdx.x * 2 + j, r = row0 + (idx >> 2);
        if (r < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + (size_t)r * D + (idx & 3) * 32));
    }
}

// CTAs are dispatched in index order and two are resident per SM: the tile this SM slot will most probably process next
constexpr int NEXT_SLOT_TILES = 2 * 148;

// ---- bulk asynchronous copy global -> shared (TMA engine, 1-D), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- the FP16 pair split
// scale = 2^k with amax * scale in [2^14, 2^15); k is clamped to [-60, 60] so that products of a few scales and their
// inverses stay finite (rows below 2^-46 simply keep fewer bits relative to their own tiny maximum)
__device__ __forceinline__ void pow2_scale(float amax, float& scale, float& inv) {
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xFFu);
    int sb = e == 0 ? 127 : 268 - e;
    sb = min(max(sb, 67), 187);
    scale = __uint_as_float((uint32_t)sb << 23);
    inv = __uint_as_float((uint32_t)(254 - sb) << 23);
}
// (a, b) already scaled -> packed FP16 pairs: p0 = {h0(a), h0(b)}, p1 = {h1(a), h1(b)}; low half = a
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& p0, uint32_t& p1) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 r = __floats2half2_rn(a - f.x, b - f.y);
    p0 = *reinterpret_cast<const uint32_t*>(&h);
    p1 = *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ float absmax32(const float (&v)[32], float m) {
#pragma unroll
    for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(v[i]));
    return m;
}
// 32 consecutive columns [c0, c0+32) of this thread's row (already multiplied by the row scale) -> both A pieces in TMEM
__device__ __forceinline__ void put_a32(uint32_t tmem_lane, int c0, const float (&v)[32], float scale) {
    uint32_t p0[16], p1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) split_f16x2(v[2 * i] * scale, v[2 * i + 1] * scale, p0[i], p1[i]);
    tmem_st16(tmem_lane + A0_COL + (c0 >> 1), p0);
    tmem_st16(tmem_lane + A1_COL + (c0 >> 1), p1);
}

struct SharedX {
    uint64_t bar_mma;           // tcgen05.commit of the current GEMM
    uint64_t bar_w;             // bulk copy of the current weight image
    uint32_t tmem;
    float xch[3][2][128];       // [slot][half][row]
    float lnacc[8][2][64];      // [warp][dw|db][col in half]
};
__device__ __forceinline__ void setup(SharedX& sh, int tmem_cols) {
    if ((threadIdx.x >> 5) == 0) tmem_alloc(&sh.tmem, tmem_cols);
    if (threadIdx.x == 0) { mbar_init(&sh.bar_mma, 1); mbar_init(&sh.bar_w, 1); fence_barrier_init(); }
}
__device__ __forceinline__ void teardown(SharedX& sh, int tmem_cols) {
    fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(sh.tmem, tmem_cols);
}
// combine a per-thread partial over the two threads that share a row
__device__ __forceinline__ float row_sum(SharedX& sh, const Epi& e, int slot, float partial) {
    sh.xch[slot][e.cb >> 6][e.row] = partial;
    __syncthreads();
    return sh.xch[slot][0][e.row] + sh.xch[slot][1][e.row];
}
__device__ __forceinline__ float row_max(SharedX& sh, const Epi& e, int slot, float partial) {
    sh.xch[slot][e.cb >> 6][e.row] = partial;
    __syncthreads();
    return fmaxf(sh.xch[slot][0][e.row], sh.xch[slot][1][e.row]);
}
__device__ __forceinline__ void flush_ln_partials(SharedX& sh, float* __restrict__ part_tile) {
    __syncthreads();
    const int t = threadIdx.x, arr = t >> 7, c = t & 127, hb = c >> 6, cc = c & 63;
    float s = 0.f;
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) s += sh.lnacc[hb * 4 + rg][arr][cc];
    part_tile[arr * D + c] = s;
}
// one thread: fetch the 64 KB image of a weight (both pieces) into the W buffer
__device__ __forceinline__ void load_w_bulk(SharedX& sh, uint8_t* Wbuf, const uint8_t* __restrict__ img) {
    if (threadIdx.x == 0) {
        mbar_expect_tx(&sh.bar_w, WIMG_BYTES);
        bulk_g2s(Wbuf, img, PIECE_BYTES, &sh.bar_w);
        bulk_g2s(Wbuf + PIECE_BYTES, img + PIECE_BYTES, PIECE_BYTES, &sh.bar_w);
    }
}
// acc[ACC_COL + acc_off] (+)= A(tmem pieces) * W^T : 8 k-steps x {a0 w0, a1 w0, a0 w1}
__device__ __forceinline__ void issue_gemm_x3(uint32_t tmem, uint32_t acc_off, uint32_t w_addr, bool accumulate) {
    constexpr uint32_t id = idesc_f16(128, false, false);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const uint64_t w0 = desc16_k(w_addr, ks >> 2, ks & 3), w1 = desc16_k(w_addr + PIECE_BYTES, ks >> 2, ks & 3);
        const uint32_t a0 = tmem + A0_COL + 8 * ks, a1 = tmem + A1_COL + 8 * ks;
        mma_f16_ts(tmem + ACC_COL + acc_off, a0, w0, id, (accumulate || ks) ? 1u : 0u);
        mma_f16_ts(tmem + ACC_COL + acc_off, a1, w0, id, 1u);
        mma_f16_ts(tmem + ACC_COL + acc_off, a0, w1, id, 1u);
    }
}
// every thread has stored its share of the A pieces: publish them, run the GEMM against the weight image whose bulk
// copy is in flight, wait for the accumulator.  The weight buffer is free again on return.
__device__ __forceinline__ void run_gemm_x3(SharedX& sh, uint32_t acc_off, const uint8_t* Wbuf, bool accumulate,
                                            uint32_t& ph_mma, uint32_t& ph_w) {
    tmem_st_wait();
    fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {          // warp-uniform issue path: elect.sync lets ptxas emit back-to-back UTCHMMA
        mbar_wait(&sh.bar_w, ph_w);
        fence_after();
        if (elect_one_sync()) {
            issue_gemm_x3(sh.tmem, acc_off, smem_u32(Wbuf), accumulate);
            mma_commit(&sh.bar_mma);
        }
        __syncwarp();
    }
    ph_w ^= 1;
    mbar_wait(&sh.bar_mma, ph_mma);
    ph_mma ^= 1;
    fence_after();
}

// ---- weight preparation: fp32 [128][128] (optionally transposed) -> per-matrix power-of-two scale + FP16 pair image,
// pre-swizzled (canonical K-major SWIZZLE_128B, piece 0 then piece 1).  One CTA per matrix.
struct PrepJobsX {
    const float* src[12];
};
__global__ void __launch_bounds__(256)
k_prep_wx3(PrepJobsX jobs, uint8_t* __restrict__ img /*[n][64 KB]*/, float* __restrict__ inv_scale /*[n]*/, int transpose) {
    __shared__ float red[8];
    __shared__ float s_scale;
    const float* __restrict__ s = jobs.src[blockIdx.x];
    uint8_t* out = img + (size_t)blockIdx.x * WIMG_BYTES;
    float m = 0.f;
    for (int i = threadIdx.x; i < D * D / 4; i += 256) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = red[0];
        for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
        float sc, inv;
        pow2_scale(r, sc, inv);
        s_scale = sc;
        inv_scale[blockIdx.x] = inv;
    }
    __syncthreads();
    const float sc = s_scale;
    // output element (n, k) = W[n][k] (or W[k][n] when transposing); 16-byte unit = 8 consecutive k
    for (int idx = threadIdx.x; idx < 128 * 16; idx += 256) {
        const int n = transpose ? (idx & 127) : (idx >> 4), u = transpose ? (idx >> 7) : (idx & 15);
        uint32_t p0[4], p1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = u * 8 + 2 * j;
            const float a = transpose ? s[(size_t)k * D + n] : s[(size_t)n * D + k];
            const float b = transpose ? s[(size_t)(k + 1) * D + n] : s[(size_t)n * D + k + 1];
            split_f16x2(a * sc, b * sc, p0[j], p1[j]);
        }
        const uint32_t off = tile16_off8(n, u);
        *reinterpret_cast<uint4*>(out + off) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
        *reinterpret_cast<uint4*>(out + PIECE_BYTES + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    }
}

// fp32 parking in the accumulator columns (a thread only ever touches its own lane and its own 64 columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
          "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
          "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
          "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31]) : "memory");
}
// the 64 parked values of this thread -> A pieces (after the row scale is known)
__device__ __forceinline__ void parked_to_a(uint32_t tl, const Epi& e, float sc) {
    tmem_st_wait();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float a[32];
        const int c0 = e.cb + half * 32;
        tmem_ld32(tl + ACC_COL + c0, a);
        put_a32(tl, c0, a, sc);
    }
}
// multiply the 64 parked values of this thread by f (a power of two)
__device__ __forceinline__ void parked_scale(uint32_t tl, const Epi& e, float f) {
    tmem_st_wait();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float a[32];
        const int c0 = e.cb + half * 32;
        tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] *= f;
        tmem_st32(tl + ACC_COL + c0, a);
    }
}
struct ChainX {
    uint8_t* W;
    float* stage;
    __device__ ChainX(uint8_t* raw) {
        W = align1k(raw);
        stage = reinterpret_cast<float*>(W + WIMG_BYTES) + (threadIdx.x >> 5) * WSTAGE_FLOATS;
    }
};
__device__ __forceinline__ void begin(SharedX& sh, uint8_t* Wbuf, const uint8_t* __restrict__ first_img) {
    setup(sh, CHAIN_TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    load_w_bulk(sh, Wbuf, first_img);
}
// a token tile from global memory -> A pieces (row scale from the row maximum); returns 1 / scale
__device__ __forceinline__ float global_to_a(SharedX& sh, const Epi& e, float* stage, uint32_t tl, const float* __restrict__ g,
                                             size_t wbase, int rv) {
    float v0[32], v1[32];
    warp_load32(stage, e.lane, g + wbase + e.cb, rv, v0);
    warp_load32(stage, e.lane, g + wbase + e.cb + 32, rv, v1);
    float sc, inv;
    pow2_scale(row_max(sh, e, 2, absmax32(v1, absmax32(v0, 0.f))), sc, inv);
    put_a32(tl, e.cb, v0, sc);
    put_a32(tl, e.cb + 32, v1, sc);
    return inv;
}
__device__ __forceinline__ void ln_stats(SharedX& sh, const Epi& e, const float (&xr)[64], float& mean, float& rstd) {
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
// forward 1: k, v from x; q from LN1(x).   wimg = images of Wk, Wv, Wq (in that order), winv their inverse scales
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_ln_qkv_x3(const float* __restrict__ x, int M, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
            const uint8_t* __restrict__ img_q, const uint8_t* __restrict__ img_k, const uint8_t* __restrict__ img_v,
            const float* __restrict__ winv /*q,k,v*/, const float* __restrict__ in_b, float* __restrict__ qn,
            float* __restrict__ st1, float* __restrict__ q, float* __restrict__ k, float* __restrict__ v) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedX sh;
    ChainX sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    begin(sh, sm.W, img_k);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    const uint32_t tl = sh.tmem + e.lane_addr;
    uint32_t ph_mma = 0, ph_w = 0;
    const float inv_x = global_to_a(sh, e, sm.stage, tl, x, wbase, rv);
    float* outs[2] = {k, v};
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {        // the raw tile stays in tensor memory for both
        run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
        load_w_bulk(sh, sm.W, g == 0 ? img_v : img_q);
        const float f = inv_x * __ldg(winv + 1 + g);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            const float* bb = in_b + (g + 1) * D + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], f, __ldg(bb + i));
            warp_store32(sm.stage, e.lane, a, outs[g] + wbase + c0, rv);
        }
    }
    float inv_qn;
    {   // LN1 of the fp32 input; the result replaces the operand pieces (the v GEMM has completed)
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
        float am = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            xr[i] = fmaf((xr[i] - mean) * rstd, __ldg(ln_w + e.cb + i), __ldg(ln_b + e.cb + i));
            am = fmaxf(am, fabsf(xr[i]));
        }
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_qn);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = xr[half * 32 + i];
            put_a32(tl, c0, a, sc);
            warp_store32(sm.stage, e.lane, a, qn + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st1[(size_t)gr * 2] = mean; st1[(size_t)gr * 2 + 1] = rstd; }
    }
    prefetch_tile_l2(x, row0 + NEXT_SLOT_TILES * 128, M);       // the first load of the CTA that follows on this slot
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    {
        const float f = inv_qn * __ldg(winv);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], f, __ldg(in_b + c0 + i)) * 0.25f;
            warp_store32(sm.stage, e.lane, a, q + wbase + c0, rv);
        }
    }
    teardown(sh, CHAIN_TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// forward 2: out-proj + residual + LN2 + FFN + mask (+ last LN).   images / winv: Wo, W1, W2
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_proj_ffn_x3(const float* __restrict__ o, const float* __restrict__ qn, int M, const uint8_t* __restrict__ img_o,
              const uint8_t* __restrict__ img_1, const uint8_t* __restrict__ img_2, const float* __restrict__ winv /*o,1,2*/,
              const float* __restrict__ bo, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
              const float* __restrict__ b1, const float* __restrict__ b2, const uint32_t* __restrict__ tmask, DropCfg dc,
              uint32_t site1, uint32_t site2, float* __restrict__ x1, float* __restrict__ st2, float* __restrict__ y,
              float* __restrict__ h, float* __restrict__ xout, const float* __restrict__ ln3_w,
              const float* __restrict__ ln3_b, float* __restrict__ enc, float* __restrict__ st3) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedX sh;
    ChainX sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    begin(sh, sm.W, img_o);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    const uint32_t tl = sh.tmem + e.lane_addr;
    uint32_t ph_mma = 0, ph_w = 0;
    float inv_a = global_to_a(sh, e, sm.stage, tl, o, wbase, rv);
    // ---- x1 = Qn + o Wo^T + bo ; y = LN2(x1)
    prefetch_tile_l2(qn, row0, M);
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    load_w_bulk(sh, sm.W, img_1);
    {
        const float f = inv_a * __ldg(winv);
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], r[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            warp_load32(sm.stage, e.lane, qn + wbase + c0, rv, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], f, __ldg(bo + c0 + i)) + r[i];
            warp_store32(sm.stage, e.lane, a, x1 + wbase + c0, rv);
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[half * 32 + i] = a[i];
        }
        float mean, rstd;
        ln_stats(sh, e, xr, mean, rstd);
        float am = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            xr[i] = fmaf((xr[i] - mean) * rstd, __ldg(ln2_w + e.cb + i), __ldg(ln2_b + e.cb + i));
            am = fmaxf(am, fabsf(xr[i]));
        }
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_a);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = xr[half * 32 + i];
            put_a32(tl, c0, a, sc);
            warp_store32(sm.stage, e.lane, a, y + wbase + c0, rv);
        }
        if (valid && e.cb == 0) { st2[(size_t)gr * 2] = mean; st2[(size_t)gr * 2 + 1] = rstd; }
    }
    // ---- h = relu(dropout1(y W1^T + b1))
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    load_w_bulk(sh, sm.W, img_2);
    {
        const float f = inv_a * __ldg(winv + 1);
        float am = 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b1 + c0 + i));
                float4 t = make_float4(fmaf(a[i], f, b4.x), fmaf(a[i + 1], f, b4.y), fmaf(a[i + 2], f, b4.z), fmaf(a[i + 3], f, b4.w));
                if (dc.train) t = drop4(t, dc, site1, (uint64_t)(gr + dc.tok_off) * D + c0 + i);
                a[i] = fmaxf(t.x, 0.f); a[i + 1] = fmaxf(t.y, 0.f); a[i + 2] = fmaxf(t.z, 0.f); a[i + 3] = fmaxf(t.w, 0.f);
            }
            am = absmax32(a, am);
            tmem_st32(tl + ACC_COL + c0, a);                 // parked until the row scale is known
            warp_store32(sm.stage, e.lane, a, h + wbase + c0, rv);
        }
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_a);
        parked_to_a(tl, e, sc);
    }
    // ---- xout = (dropout2(h W2^T + b2) + y) * ~tmask  (+ last LayerNorm)
    prefetch_tile_l2(o, row0 + NEXT_SLOT_TILES * 128, M);
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    {
        const float f = inv_a * __ldg(winv + 2);
        uint4 tw = make_uint4(0u, 0u, 0u, 0u);
        if (valid) tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr);
        float xr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a[32], yy[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            warp_load32_cg(sm.stage, e.lane, y + wbase + c0, rv, yy);   // written by this CTA above: coherent loads
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + c0 + i));
                float4 t = make_float4(fmaf(a[i], f, b4.x), fmaf(a[i + 1], f, b4.y), fmaf(a[i + 2], f, b4.z), fmaf(a[i + 3], f, b4.w));
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
    teardown(sh, CHAIN_TMEM_COLS);
}

// LayerNorm backward of the row whose upstream gradient dy is PARKED in the accumulator columns.  xhat(c0, out[32]) yields
// x-hat of 32 columns.  On return the parked values are dx = rstd (dy w - c1 - xhat c2); the per-tile LN parameter
// partials are in sh.lnacc; returns max |dx| over this thread's 64 columns.
template <class XHat>
__device__ __forceinline__ float ln_bwd_parked(SharedX& sh, const Epi& e, uint32_t tl, const float* __restrict__ w, float rstd,
                                               bool valid, XHat xhat) {
    float p1 = 0.f, p2 = 0.f;
    tmem_st_wait();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float dy[32], xh[32];
        const int c0 = e.cb + half * 32;
        tmem_ld32(tl + ACC_COL + c0, dy);
        xhat(c0, xh);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float dw = dy[i] * __ldg(w + c0 + i);
            p1 += dw;
            p2 = fmaf(dw, xh[i], p2);
        }
    }
    const float c1 = row_sum(sh, e, 0, p1) * (1.0f / D);
    const float c2 = row_sum(sh, e, 1, p2) * (1.0f / D);
    float am = 0.f;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float dy[32], xh[32], dx[32];
        const int c0 = e.cb + half * 32;
        tmem_ld32(tl + ACC_COL + c0, dy);
        xhat(c0, xh);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            dx[i] = valid ? rstd * (dy[i] * __ldg(w + c0 + i) - c1 - xh[i] * c2) : 0.f;
            xh[i] *= dy[i];                       // dw terms
        }
        am = absmax32(dx, am);
        tmem_st32(tl + ACC_COL + c0, dx);
        const float sw = warp_colsum32(xh, e.lane);
        const float sb = warp_colsum32(dy, e.lane);
        sh.lnacc[e.warp][0][half * 32 + e.lane] = sw;
        sh.lnacc[e.warp][1][half * 32 + e.lane] = sb;
    }
    return am;
}

// ----------------------------------------------------------------------------------------------
// backward 1: FFN + LN2 + out-proj input gradient.   images / winv: W2^T, W1^T, Wo^T (transposed images)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_ffn_bwd_x3(const float* __restrict__ dxo, const float* __restrict__ h, const float* __restrict__ x1,
             const float* __restrict__ st2, const uint32_t* __restrict__ tmask, int M, const uint8_t* __restrict__ img_2t,
             const uint8_t* __restrict__ img_1t, const uint8_t* __restrict__ img_ot, const float* __restrict__ winv /*2,1,o*/,
             const float* __restrict__ ln2_w, DropCfg dc, uint32_t site1, uint32_t site2, float* __restrict__ do2,
             float* __restrict__ dhpre, float* __restrict__ dx1, float* __restrict__ dO, float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedX sh;
    ChainX sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    begin(sh, sm.W, img_2t);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    const uint32_t tl = sh.tmem + e.lane_addr;
    uint32_t ph_mma = 0, ph_w = 0;
    const float dsc = dc.train ? dc.scale : 1.0f;
    uint4 tw = make_uint4(0u, 0u, 0u, 0u);
    float mean = 0.f, rstd = 0.f;
    if (valid) { tw = __ldg(reinterpret_cast<const uint4*>(tmask) + gr); mean = st2[(size_t)gr * 2]; rstd = st2[(size_t)gr * 2 + 1]; }
    float inv_a;
    {   // A = do2 = dropout2-mask * (dxo * ~tmask)
        float g[2][32];
        float am = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c0 = e.cb + half * 32;
            warp_load32(sm.stage, e.lane, dxo + wbase + c0, rv, g[half]);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 t = apply_tmask(make_float4(g[half][i], g[half][i + 1], g[half][i + 2], g[half][i + 3]), tw, (c0 + i) >> 2);
                if (dc.train) t = drop4(t, dc, site2, (uint64_t)(gr + dc.tok_off) * D + c0 + i);
                g[half][i] = t.x; g[half][i + 1] = t.y; g[half][i + 2] = t.z; g[half][i + 3] = t.w;
            }
            am = absmax32(g[half], am);
            warp_store32(sm.stage, e.lane, g[half], do2 + wbase + c0, rv);
        }
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_a);
        put_a32(tl, e.cb, g[0], sc);
        put_a32(tl, e.cb + 32, g[1], sc);
    }
    // ---- dhpre = (do2 W2) * scale * [h > 0]
    prefetch_tile_l2(h, row0, M);
    prefetch_tile_l2(x1, row0, M);
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    load_w_bulk(sh, sm.W, img_1t);
    {
        const float f = inv_a * __ldg(winv) * dsc;
        float am = 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32], hh[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            warp_load32(sm.stage, e.lane, h + wbase + c0, rv, hh);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = hh[i] > 0.f ? a[i] * f : 0.f;
            am = absmax32(a, am);
            tmem_st32(tl + ACC_COL + c0, a);
            warp_store32(sm.stage, e.lane, a, dhpre + wbase + c0, rv);
        }
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_a);
        parked_to_a(tl, e, sc);
    }
    // ---- dy = dhpre W1 + g ; LN2 backward -> dx1
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    load_w_bulk(sh, sm.W, img_ot);
    {
        const float f = inv_a * __ldg(winv + 1);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {          // park dy = acc f + masked upstream gradient
            float a[32], g[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            warp_load32(sm.stage, e.lane, dxo + wbase + c0, rv, g);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 gm = apply_tmask(make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]), tw, (c0 + i) >> 2);
                a[i] = fmaf(a[i], f, gm.x); a[i + 1] = fmaf(a[i + 1], f, gm.y);
                a[i + 2] = fmaf(a[i + 2], f, gm.z); a[i + 3] = fmaf(a[i + 3], f, gm.w);
            }
            tmem_st32(tl + ACC_COL + c0, a);
        }
        auto xhat = [&](int c0, float (&xh)[32]) {
            warp_load32(sm.stage, e.lane, x1 + wbase + c0, rv, xh);
#pragma unroll
            for (int i = 0; i < 32; ++i) xh[i] = valid ? (xh[i] - mean) * rstd : 0.f;
        };
        const float am = ln_bwd_parked(sh, e, tl, ln2_w, rstd, valid, xhat);
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
        float sc;
        pow2_scale(row_max(sh, e, 2, am), sc, inv_a);
        tmem_st_wait();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            put_a32(tl, c0, a, sc);
            warp_store32(sm.stage, e.lane, a, dx1 + wbase + c0, rv);
        }
    }
    // ---- dO = dx1 Wo
    prefetch_tile_l2(dxo, row0 + NEXT_SLOT_TILES * 128, M);
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    {
        const float f = inv_a * __ldg(winv + 2);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] *= f;
            warp_store32(sm.stage, e.lane, a, dO + wbase + c0, rv);
        }
    }
    teardown(sh, CHAIN_TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// backward 2: dQn = dx1 + dq Wq ; dx_in = LN1bwd(dQn) + dk Wk + dv Wv.   images / winv: Wq^T, Wk^T, Wv^T
// One accumulator: the LayerNorm-backward result is parked in it, rescaled (exactly, by powers of two) to the scale of
// the next GEMM's operands, and the dk / dv products accumulate on top.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_qkv_bwd_x3(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
             const float* __restrict__ dx1, const float* __restrict__ xin, const float* __restrict__ st1, int M,
             const uint8_t* __restrict__ img_qt, const uint8_t* __restrict__ img_kt, const uint8_t* __restrict__ img_vt,
             const float* __restrict__ winv /*q,k,v*/, const float* __restrict__ ln1_w, float* __restrict__ dxin,
             float* __restrict__ ln_part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedX sh;
    ChainX sm(smem_raw);
    const int row0 = blockIdx.x * 128;
    begin(sh, sm.W, img_qt);
    Epi e;
    const int gr = row0 + e.row;
    const bool valid = gr < M;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    const uint32_t tl = sh.tmem + e.lane_addr;
    uint32_t ph_mma = 0, ph_w = 0;
    float mean = 0.f, rstd = 0.f;
    if (valid) { mean = st1[(size_t)gr * 2]; rstd = st1[(size_t)gr * 2 + 1]; }
    const float inv_q = global_to_a(sh, e, sm.stage, tl, dq, wbase, rv);
    prefetch_tile_l2(dx1, row0, M);
    prefetch_tile_l2(xin, row0, M);
    run_gemm_x3(sh, 0, sm.W, false, ph_mma, ph_w);
    load_w_bulk(sh, sm.W, img_kt);
    prefetch_tile_l2(dk, row0, M);
    {
        const float f = inv_q * __ldg(winv);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {          // park dQn = dx1 + dq Wq
            float a[32], r[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
            warp_load32(sm.stage, e.lane, dx1 + wbase + c0, rv, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], f, r[i]);
            tmem_st32(tl + ACC_COL + c0, a);
        }
        auto xhat = [&](int c0, float (&xh)[32]) {
            warp_load32(sm.stage, e.lane, xin + wbase + c0, rv, xh);
#pragma unroll
            for (int i = 0; i < 32; ++i) xh[i] = valid ? (xh[i] - mean) * rstd : 0.f;
        };
        ln_bwd_parked(sh, e, tl, ln1_w, rstd, valid, xhat);
        flush_ln_partials(sh, ln_part + (size_t)blockIdx.x * 2 * D);
    }
    float cur = 1.0f;          // the parked values are (true value) * cur
    const float* srcs[2] = {dk, dv};
    float inv_a = 1.0f;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        float v0[32], v1[32];
        warp_load32(sm.stage, e.lane, srcs[g] + wbase + e.cb, rv, v0);
        warp_load32(sm.stage, e.lane, srcs[g] + wbase + e.cb + 32, rv, v1);
        float sc;
        pow2_scale(row_max(sh, e, 2, absmax32(v1, absmax32(v0, 0.f))), sc, inv_a);
        const float wi = __ldg(winv + 1 + g);
        const float target = sc * (1.0f / wi);            // scale of this GEMM's products (powers of two: exact)
        parked_scale(tl, e, target / cur);
        cur = target;
        put_a32(tl, e.cb, v0, sc);
        put_a32(tl, e.cb + 32, v1, sc);
        if (g == 0) prefetch_tile_l2(dv, row0, M);
        else prefetch_tile_l2(dq, row0 + NEXT_SLOT_TILES * 128, M);
        run_gemm_x3(sh, 0, sm.W, true, ph_mma, ph_w);
        if (g == 0) load_w_bulk(sh, sm.W, img_vt);
        inv_a *= wi;
    }
    {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float a[32];
            const int c0 = e.cb + half * 32;
            tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] *= inv_a;
            warp_store32(sm.stage, e.lane, a, dxin + wbase + c0, rv);
        }
    }
    teardown(sh, CHAIN_TMEM_COLS);
}

// ================================================================================================
// weight gradients with BF16 triples: dW[n][k] = sum_m dY[m][n] X[m][k] through MN-major views of the row-major token
// tiles; 64-token sub-tiles, double buffered (the MMAs of one sub-tile run while the next one is loaded and split).
// ================================================================================================
constexpr int SUB_ROWS = 64;
constexpr int SUBP_BYTES = SUB_ROWS * 128 * 2;       // one BF16 piece of a [64][128] sub-tile: 16 KB (2 chunks of 8 KB)
constexpr int SUB_CHUNK = SUB_ROWS * 128;            // bytes between the two 64-feature chunks
constexpr int WG_BUF_BYTES = 6 * SUBP_BYTES;         // dY pieces 0..2, X pieces 0..2
constexpr size_t WGRADX_SMEM = 2 * (size_t)WG_BUF_BYTES + 1024;    // 193 KB
__device__ __forceinline__ uint32_t sub_off8(int r, int u) {
    return (uint32_t)((u >> 3) * SUB_CHUNK + (r >> 3) * 1024 + (r & 7) * 128 + (((u & 7) ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint64_t desc_sub_mn(uint32_t tile_addr, int kstep) {
    return make_desc(tile_addr + kstep * 2048, SUB_CHUNK, 1024);
}
// x = b0 + b1 + b2 (each rounded to nearest BF16; the residuals are exact in fp32)
__device__ __forceinline__ void split_bf16x3(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    p0 = pack_bf16(a, b);
    const float ra = a - __uint_as_float(p0 << 16), rb = b - __uint_as_float(p0 & 0xFFFF0000u);
    p1 = pack_bf16(ra, rb);
    const float sa = ra - __uint_as_float(p1 << 16), sb = rb - __uint_as_float(p1 & 0xFFFF0000u);
    p2 = pack_bf16(sa, sb);
}
constexpr int WGX_THREADS = 512;                     // 16 warps: the split / store work of a sub-tile is issue-bound
constexpr int WGX_IT = SUB_ROWS / (WGX_THREADS / 32);
// rows [row0, row0+64) of g[M,128] -> three BF16 piece sub-tiles at dst, dst + 16 KB, dst + 32 KB
__device__ __forceinline__ void fill_sub3(uint8_t* dst, const float4 (&v)[WGX_IT]) {
    const int c4 = threadIdx.x & 31, rb = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < WGX_IT; ++it) {
        const int r = it * (WGX_THREADS / 32) + rb;
        uint32_t a0, a1, a2, b0, b1, b2;
        split_bf16x3(v[it].x, v[it].y, a0, a1, a2);
        split_bf16x3(v[it].z, v[it].w, b0, b1, b2);
        const uint32_t off = sub_off8(r, c4 >> 1) + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(dst + off) = make_uint2(a0, b0);
        *reinterpret_cast<uint2*>(dst + SUBP_BYTES + off) = make_uint2(a1, b1);
        *reinterpret_cast<uint2*>(dst + 2 * SUBP_BYTES + off) = make_uint2(a2, b2);
    }
}
__device__ __forceinline__ void load_sub(float4 (&v)[WGX_IT], const float* __restrict__ g, int row0, int M) {
    const int c4 = threadIdx.x & 31, rb = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < WGX_IT; ++it) {
        const int r = row0 + it * (WGX_THREADS / 32) + rb;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < M) v[it] = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * D) + c4);
    }
}
struct WgradJobsX {
    const float* dY[6];
    const float* X[6];
};
__global__ void __launch_bounds__(WGX_THREADS, 1)
k_wgrad_x3(WgradJobsX jobs, int M, float* __restrict__ wpart /*[6][S][128*128]*/, float* __restrict__ bpart /*[6][S][128]*/) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];           // tcgen05.commit: the MMAs that read buffer b are complete
    __shared__ uint32_t tmem_s;
    uint8_t* buf0 = align1k(smem_raw);
    const float* __restrict__ dY = jobs.dY[blockIdx.y];
    const float* __restrict__ X = jobs.X[blockIdx.y];
    const int S = gridDim.x;
    if ((threadIdx.x >> 5) == 0) tmem_alloc(&tmem_s, 256);
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_s;
    const int subs = (M + SUB_ROWS - 1) / SUB_ROWS;
    uint32_t phase[2] = {0u, 0u};
    constexpr uint32_t id_w = idesc_bf16(128, true, true);
    // sub-tiles are dealt in pairs so that a CTA streams 128 consecutive tokens at a time; the global loads of sub-tile
    // i+1 are issued before sub-tile i is split and stored, so their latency hides behind the conversion and the MMAs
    const int pairs = (subs + 1) / 2;
    const int my_pairs = blockIdx.x < pairs ? (pairs - 1 - blockIdx.x) / S + 1 : 0;
    auto sub_of = [&](int i) { return 2 * (blockIdx.x + (i >> 1) * S) + (i & 1); };
    int n_it = 2 * my_pairs;
    if (n_it && sub_of(n_it - 1) >= subs) --n_it;
    float4 vy[WGX_IT], vx[WGX_IT];
    // bias gradient = column sums of dY: accumulated on the CUDA cores from the rows this thread converts anyway (thread =
    // 4 columns, fixed row order) instead of three more N = 16 MMAs per k-step against a tile of ones -- the MMAs of this kernel
    // are bound by their 8 KB of shared-memory operand reads, and those three read the 4 KB A operand once more each
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_it) { load_sub(vy, dY, sub_of(0) * SUB_ROWS, M); load_sub(vx, X, sub_of(0) * SUB_ROWS, M); }
    int it = 0;
    for (; it < n_it; ++it) {
        const int b = it & 1;
        if (it >= 2) { mbar_wait(&bars[b], phase[b]); phase[b] ^= 1; }      // MMAs that read this buffer are done
        uint8_t* buf = buf0 + b * WG_BUF_BYTES;
        fill_sub3(buf, vy);
        fill_sub3(buf + 3 * SUBP_BYTES, vx);
#pragma unroll
        for (int r = 0; r < WGX_IT; ++r) { bsum.x += vy[r].x; bsum.y += vy[r].y; bsum.z += vy[r].z; bsum.w += vy[r].w; }
        if (it + 1 < n_it) { load_sub(vy, dY, sub_of(it + 1) * SUB_ROWS, M); load_sub(vx, X, sub_of(it + 1) * SUB_ROWS, M); }
        fence_async_smem();
        __syncthreads();          // (an mbarrier hand-off that lets the other 15 warps run ahead was measured 40 % slower)
        if ((threadIdx.x >> 5) == 0 && elect_one_sync()) {
            fence_after();
            const uint32_t a = smem_u32(buf), x = a + 3 * SUBP_BYTES;
#pragma unroll
            for (int ks = 0; ks < SUB_ROWS / 16; ++ks) {
                const uint32_t acc = (it || ks) ? 1u : 0u;
                const uint64_t a0 = desc_sub_mn(a, ks), a1 = desc_sub_mn(a + SUBP_BYTES, ks), a2 = desc_sub_mn(a + 2 * SUBP_BYTES, ks);
                const uint64_t x0 = desc_sub_mn(x, ks), x1 = desc_sub_mn(x + SUBP_BYTES, ks), x2 = desc_sub_mn(x + 2 * SUBP_BYTES, ks);
                mma_bf16(tmem, a0, x0, id_w, acc);
                mma_bf16(tmem, a0, x1, id_w, 1u);
                mma_bf16(tmem, a1, x0, id_w, 1u);
                mma_bf16(tmem, a0, x2, id_w, 1u);
                mma_bf16(tmem, a1, x1, id_w, 1u);
                mma_bf16(tmem, a2, x0, id_w, 1u);
            }
            mma_commit(&bars[b]);
        }
    }
    // drain: the last one or two commits
    if (it >= 2) { const int b = it & 1; mbar_wait(&bars[b], phase[b]); phase[b] ^= 1; }
    if (it >= 1) { const int b = (it - 1) & 1; mbar_wait(&bars[b], phase[b]); phase[b] ^= 1; }
    fence_after();
    if (threadIdx.x < 256) {             // the first 8 warps read the accumulators out (lane quarter = warp & 3)
    Epi e;
    float* wp = wpart + ((size_t)blockIdx.y * S + blockIdx.x) * D * D;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float a[32];
        const int c0 = e.cb + half * 32;
        if (it) {
            tmem_ld32(tmem + e.lane_addr + c0, a);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(wp + (size_t)e.row * D + c0 + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
    }
    }
    {   // bias partials: 16 row groups x 128 columns through shared memory (the operand buffers are free now), fixed order
        float* bred = reinterpret_cast<float*>(buf0);
        *reinterpret_cast<float4*>(bred + (threadIdx.x >> 5) * D + 4 * (threadIdx.x & 31)) = bsum;
        __syncthreads();
        if (threadIdx.x < D) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < WGX_THREADS / 32; ++w) sacc += bred[w * D + threadIdx.x];
            bpart[((size_t)blockIdx.y * S + blockIdx.x) * D + threadIdx.x] = sacc;
        }
    }
    fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem, 256);
}

}  // namespace x3
}  // namespace amid
