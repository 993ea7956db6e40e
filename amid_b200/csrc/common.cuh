// Shared device/host helpers for the amid_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/amid_b200.h"

namespace amid {

constexpr int D = AMID_D;       // embedding width
constexpr int H = AMID_HEADS;   // attention heads
constexpr int DH = D / H;       // 16
constexpr float LN_EPS = 1e-8f; // model_seq.py:342,345,352

// ------------------------------------------------------------------ error plumbing
int set_error(int code, const char* fmt, ...);
#define AMID_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::amid::set_error(-1, __VA_ARGS__);   \
    } while (0)
// every launch site: AMID_K(name, stream); kernel<<<...>>>(...); AMID_LAUNCH_CHECK(name);
// counts the launch and, in profile mode, brackets it with CUDA events on its stream.
void prof_begin(const char* name, cudaStream_t s);
void prof_end();
#define AMID_K(name, stream) ::amid::prof_begin(name, (cudaStream_t)(stream))
#define AMID_LAUNCH_CHECK(name)                                                         \
    do {                                                                                \
        ::amid::prof_end();                                                             \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return ::amid::set_error(-2, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// device-side "id out of range" flag shared by every kernel that indexes the table (gather.cu);
// read and cleared by amid_gather_error_host_sync()
int* err_flag();

// ------------------------------------------------------------------ counter-based dropout RNG
// One 32-bit integer hash yields four 8-bit lanes = the keep decisions of 4 consecutive elements
// (idx4 = element_index >> 2).  keep <=> lane >= thr8, thr8 = round(p * 256)  (p = 0.5 is exact).
__device__ __forceinline__ uint32_t rng4(uint32_t seed32, uint32_t site, uint64_t idx4) {
    uint32_t x = (uint32_t)idx4 * 0x9E3779B1u + (uint32_t)(idx4 >> 32) * 0x85EBCA77u + (seed32 ^ ((site + 1u) * 0xC2B2AE3Du));
    x ^= x >> 16; x *= 0x7FEB352Du;      // "lowbias32" finaliser
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool rng_keep(uint32_t r, int lane4, uint32_t thr8) {
    return ((r >> (8 * lane4)) & 0xFFu) >= thr8;
}

struct DropCfg {
    int train;
    uint32_t thr16;      // 8-bit threshold (name kept for brevity at the call sites)
    float scale;
    uint32_t seed;       // 32-bit digest of the 64-bit step seed
    uint32_t site_base;
    uint32_t b_off;      // first sample of this call in the global batch (data parallelism)
    uint32_t tok_off;    // = b_off * L   (set by the launcher once L is known: see with_offsets)
    uint32_t bh_off;     // = b_off * H
};
inline DropCfg make_drop(const amid_dropout* d) {
    DropCfg c{0, 0u, 1.0f, 0u, 0u, 0u, 0u, 0u};
    if (d && d->train && d->p > 0.f) {
        c.train = 1;
        double t = (double)d->p * 256.0 + 0.5;
        c.thr16 = (uint32_t)(t > 255.0 ? 255.0 : t);
        c.scale = 1.0f / (1.0f - d->p);
        uint64_t z = d->seed + 0x9E3779B97F4A7C15ull;           // splitmix64 digest on the host
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        c.seed = (uint32_t)(z ^ (z >> 32));
    }
    if (d) { c.site_base = d->site_base; c.b_off = d->batch_offset > 0 ? (uint32_t)d->batch_offset : 0u; }
    return c;
}
inline DropCfg with_offsets(DropCfg c, int L) {
    c.tok_off = c.b_off * (uint32_t)L;
    c.bh_off = c.b_off * (uint32_t)H;
    return c;
}

// apply dropout to 4 consecutive elements starting at element index e0 (multiple of 4)
__device__ __forceinline__ float4 drop4(float4 v, const DropCfg& c, uint32_t site, uint64_t e0) {
    const uint32_t r = rng4(c.seed, site, e0 >> 2);
    v.x = rng_keep(r, 0, c.thr16) ? v.x * c.scale : 0.f;
    v.y = rng_keep(r, 1, c.thr16) ? v.y * c.scale : 0.f;
    v.z = rng_keep(r, 2, c.thr16) ? v.z * c.scale : 0.f;
    v.w = rng_keep(r, 3, c.thr16) ? v.w * c.scale : 0.f;
    return v;
}

// sites inside one encoder (relative to site_base)
enum : uint32_t { SITE_EMB = 0, SITE_ATTN0 = 1, SITE_FFN1_0 = 2, SITE_FFN2_0 = 3 };
__host__ __device__ inline uint32_t site_attn(int blk) { return 1u + 3u * blk; }
__host__ __device__ inline uint32_t site_ffn1(int blk) { return 2u + 3u * blk; }
__host__ __device__ inline uint32_t site_ffn2(int blk) { return 3u + 3u * blk; }

// ------------------------------------------------------------------ cp.async
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace amid
