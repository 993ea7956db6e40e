// 128x128x128 fp32 tile GEMM on CUDA cores, the building block of the exact-fp32 encoder
// path.  A tile lives in shared memory row-major with a padded stride; the B operand
// ([k][n], n contiguous) is streamed from global/L2 through a cp.async double buffer.
#pragma once
#include "common.cuh"

namespace amid {

constexpr int TM = 128;   // rows per tile
constexpr int LDA = 132;  // smem row stride in floats (16B aligned, conflict-free row broadcast)
constexpr int KC = 32;    // k-chunk streamed per stage
constexpr int NT = 256;   // threads per CTA
constexpr int TILE_FLOATS = TM * LDA;
constexpr int WBUF_FLOATS = 2 * KC * D;
constexpr size_t ENC_SMEM_BYTES = (size_t)(2 * TILE_FLOATS + WBUF_FLOATS) * sizeof(float);  // 167,936

// thread -> micro-tile: rows tm*8..+7 ; cols tn*4..+3 and 64+tn*4..+3
struct Frag {
    int tm, tn;
    __device__ Frag() {
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        tn = lane & 15;
        tm = warp * 2 + (lane >> 4);
    }
    __device__ __forceinline__ int row(int i) const { return tm * 8 + i; }
    __device__ __forceinline__ int c0() const { return tn * 4; }
    __device__ __forceinline__ int c1() const { return 64 + tn * 4; }
};

__device__ __forceinline__ void load_w_chunk(float* Wbuf, const float* __restrict__ Bg, int chunk) {
    const float4* src = reinterpret_cast<const float4*>(Bg + (size_t)chunk * KC * D);
#pragma unroll
    for (int i = threadIdx.x; i < KC * D / 4; i += NT) cp_async16(Wbuf + i * 4, src + i);
}

// acc (+)= As[128 x 128] * Bg[128(k) x 128(n)]
template <bool ACCUM>
__device__ __forceinline__ void tile_gemm(const float* As, const float* __restrict__ Bg, float* Ws,
                                          float (&acc)[8][8]) {
    Frag f;
    if (!ACCUM) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    load_w_chunk(Ws, Bg, 0);
    cp_async_commit();
#pragma unroll 1
    for (int c = 0; c < D / KC; ++c) {
        if (c + 1 < D / KC) {
            load_w_chunk(Ws + ((c + 1) & 1) * KC * D, Bg, c + 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* W = Ws + (c & 1) * KC * D;
        const float* A = As + f.tm * 8 * LDA + c * KC;
#pragma unroll 2
        for (int kk = 0; kk < KC; kk += 4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + i * LDA + kk);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 b0 = *reinterpret_cast<const float4*>(W + (kk + u) * D + f.c0());
                const float4 b1 = *reinterpret_cast<const float4*>(W + (kk + u) * D + f.c1());
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = u == 0 ? a[i].x : (u == 1 ? a[i].y : (u == 2 ? a[i].z : a[i].w));
                    acc[i][0] = fmaf(av, b0.x, acc[i][0]);
                    acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, b0.z, acc[i][2]);
                    acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, b1.x, acc[i][4]);
                    acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, b1.z, acc[i][6]);
                    acc[i][7] = fmaf(av, b1.w, acc[i][7]);
                }
            }
        }
        __syncthreads();
    }
}

// global [M,128] rows row0.. -> smem tile (rows >= M zero-filled).  Caller syncs.
__device__ __forceinline__ void load_tile(float* T, const float* __restrict__ g, int row0, int M) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < TM * (D / 4); idx += NT) {
        int r = idx >> 5, c4 = idx & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)(row0 + r) * D) + c4);
        *reinterpret_cast<float4*>(T + r * LDA + c4 * 4) = v;
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(const float (&a)[8], int half) {
    return half == 0 ? make_float4(a[0], a[1], a[2], a[3]) : make_float4(a[4], a[5], a[6], a[7]);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float sum4(float4 a) { return (a.x + a.y) + (a.z + a.w); }

// timeline-mask bits of a row: word e bit j <-> column 4*j+e.  Zero the masked lanes of
// the float4 that starts at column 4*j.
__device__ __forceinline__ float4 apply_tmask(float4 v, uint4 tw, int j) {
    if ((tw.x >> j) & 1u) v.x = 0.f;
    if ((tw.y >> j) & 1u) v.y = 0.f;
    if ((tw.z >> j) & 1u) v.z = 0.f;
    if ((tw.w >> j) & 1u) v.w = 0.f;
    return v;
}

// LayerNorm of every row of a smem tile (warp per row).  dst may alias src.
__device__ __forceinline__ void ln_tile(const float* src, float* dst, const float* __restrict__ w,
                                        const float* __restrict__ b, int row0, int M,
                                        float* __restrict__ out_g, float* __restrict__ stats_g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + lane);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(b) + lane);
    for (int r = warp; r < TM; r += NT / 32) {
        float4 x = ld4(src + r * LDA + lane * 4);
        const float mean = warp_sum(sum4(x)) * (1.0f / D);
        x = make_float4(x.x - mean, x.y - mean, x.z - mean, x.w - mean);
        const float var = warp_sum(sum4(mul4(x, x))) * (1.0f / D);
        const float rstd = 1.0f / sqrtf(var + LN_EPS);
        float4 y = make_float4(fmaf(x.x * rstd, wv.x, bv.x), fmaf(x.y * rstd, wv.y, bv.y),
                               fmaf(x.z * rstd, wv.z, bv.z), fmaf(x.w * rstd, wv.w, bv.w));
        st4(dst + r * LDA + lane * 4, y);
        if (row0 + r < M) {
            if (out_g) st4(out_g + (size_t)(row0 + r) * D + lane * 4, y);
            if (stats_g && lane == 0) {
                stats_g[(size_t)(row0 + r) * 2] = mean;
                stats_g[(size_t)(row0 + r) * 2 + 1] = rstd;
            }
        }
    }
}

}  // namespace amid
