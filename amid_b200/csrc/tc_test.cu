// Bring-up / unit-test entry points of the tcgen05 tile pipeline (tc.cuh): a plain linear layer
// (K-major operands).  The fused encoder kernels are built from the same pieces.
#include "tc.cuh"

namespace amid {
using namespace tc;

constexpr size_t TC_TEST_SMEM = 2 * TILE_BYTES + 1024;

// y[M,128] = x[M,128] * w[128,128]^T + b
__global__ void __launch_bounds__(256, 1)
k_tc_linear(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int M,
            float* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint8_t* At = base;
    uint8_t* Bt = base + TILE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fill_tile(At, x, row0, M);
    fill_tile(Bt, w, 0, 128);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        issue_gemm_kk(tmem, smem_u32(At), smem_u32(Bt), false);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    const int row = 32 * (warp & 3) + lane;
    const int cb = 64 * (warp >> 2);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb + half * 32, v);
        if (row0 + row < M) {
            float* dst = y + (size_t)(row0 + row) * D + cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b + cb + half * 32 + i));
                *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- BF16 bring-up: y = x w^T + b (K-major) and part[cta] = dy^T x (MN-major, accumulator resident in TMEM)
constexpr size_t TC16_TEST_SMEM = 2 * TILE16_BYTES + 1024;

__global__ void __launch_bounds__(256, 2)
k_tc_linear16(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int M,
              float* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint8_t* At = base;
    uint8_t* Bt = base + TILE16_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fill_tile16(At, x, row0, M);
    fill_tile16(Bt, w, 0, 128);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        issue_gemm16_kk(tmem, smem_u32(At), smem_u32(Bt), false);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    const int row = 32 * (warp & 3) + lane;
    const int cb = 64 * (warp >> 2);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb + half * 32, v);
        if (row0 + row < M) {
            float* dst = y + (size_t)(row0 + row) * D + cb + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b + cb + half * 32 + i));
                *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(256, 2)
k_tc_wgrad16(const float* __restrict__ dy, const float* __restrict__ x, int M, float* __restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint8_t* At = base;
    uint8_t* Bt = base + TILE16_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    const int tiles = (M + 127) / 128;
    uint32_t phase = 0;
    bool first = true;
    constexpr uint32_t id = idesc_bf16(128, true, true);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        fill_tile16(At, dy, t * 128, M);
        fill_tile16(Bt, x, t * 128, M);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
            const uint32_t a = smem_u32(At), bb = smem_u32(Bt);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) mma_bf16(tmem, desc16_mn(a, ks), desc16_mn(bb, ks), id, (!first || ks) ? 1u : 0u);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        first = false;
    }
    fence_after();
    const int row = 32 * (warp & 3) + lane;
    const int cb = 64 * (warp >> 2);
    float* out = part + (size_t)blockIdx.x * D * D;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float v[32];
        if (!first) {
            tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb + half * 32, v);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        float* dst = out + (size_t)row * D + cb + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace amid

using namespace amid;

extern "C" int amid_tc_linear16_test(const float* x, const float* w, const float* b, int32_t M, float* y, amid_stream_t s_) {
    AMID_REQUIRE(x && w && b && y && M > 0, "tc_linear16_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_tc_linear16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC16_TEST_SMEM);
    if (e != cudaSuccess) return set_error(-3, "tc_linear16_test: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_tc_linear16", s_);
    k_tc_linear16<<<(M + 127) / 128, 256, TC16_TEST_SMEM, (cudaStream_t)s_>>>(x, w, b, M, y);
    AMID_LAUNCH_CHECK("k_tc_linear16");
    return 0;
}
extern "C" int amid_tc_wgrad16_test(const float* dy, const float* x, int32_t M, float* part, int32_t n_ctas, amid_stream_t s_) {
    AMID_REQUIRE(dy && x && part && M > 0 && n_ctas > 0, "tc_wgrad16_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_tc_wgrad16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC16_TEST_SMEM);
    if (e != cudaSuccess) return set_error(-3, "tc_wgrad16_test: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_tc_wgrad16", s_);
    k_tc_wgrad16<<<n_ctas, 256, TC16_TEST_SMEM, (cudaStream_t)s_>>>(dy, x, M, part);
    AMID_LAUNCH_CHECK("k_tc_wgrad16");
    return 0;
}


extern "C" int amid_tc_linear_test(const float* x, const float* w, const float* b, int32_t M, float* y, amid_stream_t s_) {
    AMID_REQUIRE(x && w && b && y && M > 0, "tc_linear_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_tc_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_TEST_SMEM);
    if (e != cudaSuccess) return set_error(-3, "tc_linear_test: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_tc_linear", s_);
    k_tc_linear<<<(M + 127) / 128, 256, TC_TEST_SMEM, (cudaStream_t)s_>>>(x, w, b, M, y);
    AMID_LAUNCH_CHECK("k_tc_linear");
    return 0;
}
