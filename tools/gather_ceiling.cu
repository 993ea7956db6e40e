// Micro-benchmark: what can a 512 B-row gather reach on this GPU?  Compares a sequential copy with a pure
// random-row copy (no arithmetic) over the reference-sized table (894,820 x 128 fp32), for several
// rows-in-flight-per-warp settings.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_ceiling gather_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int RPW>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ table, const int64_t* __restrict__ ids, int64_t n,
                                                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = warp * RPW;
    if (r0 >= n) return;
    int64_t id = 0;
    if (lane < RPW && r0 + lane < n) id = ids[r0 + lane];
    float4 v[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(table + __shfl_sync(0xffffffffu, id, u) * 128) + lane);
#pragma unroll
    for (int u = 0; u < RPW; ++u)
        if (r0 + u < n) stg_stream(reinterpret_cast<float4*>(out + (r0 + u) * 128) + lane, v[u]);
}
template <int RPW>
float run(const float* table, const int64_t* ids, int64_t n, float* out, int iters) {
    const unsigned blocks = (unsigned)(((n + RPW - 1) / RPW * 32 + 255) / 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k_gather<RPW><<<blocks, 256>>>(table, ids, n, out);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) k_gather<RPW><<<blocks, 256>>>(table, ids + (i % 4) * n, n, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / iters;
}
int main() {
    const int64_t V = 894820, n = 1024 * 402;
    float *table, *out;
    int64_t* ids;
    cudaMalloc(&table, V * 512);
    cudaMalloc(&out, n * 512);
    cudaMalloc(&ids, 4 * n * 8);
    cudaMemset(table, 0, V * 512);
    std::vector<int64_t> h(4 * n);
    srand(1);
    const double bytes = (double)n * (1024 + 8);
    for (int mode = 0; mode < 2; ++mode) {
        for (auto& x : h) x = mode == 0 ? 0 : (int64_t)((((uint64_t)rand() << 16) ^ rand()) % V);
        if (mode == 0) for (int64_t i = 0; i < 4 * n; ++i) h[i] = (i % n) % V;     // sequential rows
        cudaMemcpy(ids, h.data(), 4 * n * 8, cudaMemcpyHostToDevice);
        const char* name = mode == 0 ? "sequential rows" : "uniform random rows";
        printf("%s: RPW4 %.1f us %.0f GB/s | RPW8 %.1f us %.0f GB/s | RPW16 %.1f us %.0f GB/s\n", name,
               run<4>(table, ids, n, out, 50), bytes / run<4>(table, ids, n, out, 50) / 1e3,
               run<8>(table, ids, n, out, 50), bytes / run<8>(table, ids, n, out, 50) / 1e3,
               run<16>(table, ids, n, out, 50), bytes / run<16>(table, ids, n, out, 50) / 1e3);
    }
    return 0;
}
