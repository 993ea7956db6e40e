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
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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

}  // namespace amid

using namespace amid;

extern "C" int amid_tc_linear_test(const float* x, const float* w, const float* b, int32_t M, float* y, amid_stream_t s_) {
    AMID_REQUIRE(x && w && b && y && M > 0, "tc_linear_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_tc_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_TEST_SMEM);
    if (e != cudaSuccess) return set_error(-3, "tc_linear_test: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_tc_linear", s_);
    k_tc_linear<<<(M + 127) / 128, 256, TC_TEST_SMEM, (cudaStream_t)s_>>>(x, w, b, M, y);
    AMID_LAUNCH_CHECK("k_tc_linear");
    return 0;
}
