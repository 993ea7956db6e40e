// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) and the shared-memory operand
// layout used by the tensor-core path.
//
// Operand tile = [128 rows][128 fp32] stored as 4 column chunks of 32 floats (128 B); each
// chunk is [128 rows x 128 B] in the canonical SWIZZLE_128B layout (8-row groups of 1024 B,
// 16-byte units XOR-ed with row%8).  The SAME bytes are
//   * a K-major operand  (rows = M/N index, columns = K)    -> Y = X W^T, dX = dY W
//   * an MN-major operand (columns = M/N index, rows = K)   -> dW = dY^T X
// (the MN-major view needs the SWIZZLE_128B_BASE32B layout for 32-bit operands, so the weight-gradient
// kernel stages transposed K-major tiles instead).  kind::tf32 reads the fp32 bit patterns directly
// (10-bit mantissa), accumulators are fp32 in TMEM.
#pragma once
#include "common.cuh"

namespace amid {
namespace tc {

constexpr int TILE_BYTES = 128 * 128 * 4;   // 64 KB
constexpr int CHUNK_BYTES = 128 * 128;      // one 32-column chunk: 128 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, column k) inside a tile
__device__ __forceinline__ uint32_t tile_off(int r, int k) {
    return (uint32_t)((k >> 5) * CHUNK_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) +
                      (k & 3) * 4);
}
// byte offset of the 16-byte unit holding columns [4*c4, 4*c4+4) of row r
__device__ __forceinline__ uint32_t tile_off4(int r, int c4) {
    return (uint32_t)((c4 >> 3) * CHUNK_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + (((c4 & 7) ^ (r & 7)) << 4));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Every wait is bounded: a protocol error (a phase that never completes) must surface as a launch failure, not as a GPU that
// hangs until somebody resets it.  try_wait suspends the thread for a hardware-defined interval per attempt, so 2^26 failed
// attempts are many seconds -- orders of magnitude beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins == (1u << 26)) __trap();      // (no printf here: its stack frame costs registers in every kernel)
    } while (!ok);
}

// ---- proxies / fences
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors
// shared-memory matrix descriptor, SWIZZLE_128B; lbo/sbo in bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
// K-major view of a tile: k-step ks (8 tf32 = 32 B) of chunk c
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int c, int ks) {
    return make_desc(tile_addr + c * CHUNK_BYTES + ks * 32, 16, 1024);
}
// MN-major view of a tile: k-step = 8 rows (one 1024 B group); the 4 chunks are the MN atoms
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kgroup) {
    return make_desc(tile_addr + kgroup * 1024, CHUNK_BYTES, 1024);
}
// instruction descriptor: kind::tf32, fp32 accumulate, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_tf32(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- MMA issue (ONE thread) and completion
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[128 x 128] (+)= A_tile * B_tile^T, both K-major (K = 128 = 4 chunks x 4 k-steps)
__device__ __forceinline__ void issue_gemm_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
    constexpr uint32_t id = idesc_tf32(128, false, false);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            mma_tf32(tmem_d, desc_kmajor(a_addr, c, ks), desc_kmajor(b_addr, c, ks), id, (accumulate || c || ks) ? 1u : 0u);
}
// D[128(n) x 128(k)] (+)= A_tile^T * B_tile over `rows` token rows (multiple of 8), both MN-major
__device__ __forceinline__ void issue_gemm_mn(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int rows, bool accumulate) {
    constexpr uint32_t id = idesc_tf32(128, true, true);
    for (int g = 0; g < rows / 8; ++g)
        mma_tf32(tmem_d, desc_mnmajor(a_addr, g), desc_mnmajor(b_addr, g), id, (accumulate || g) ? 1u : 0u);
}

// ---- TMEM -> registers: 32 lanes (this warp's quarter) x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- cooperative tile fills (256 threads); rows >= M are zero-filled
// global [M,128] fp32 rows row0.. -> swizzled tile.  One warp reads one 512 B row per step.
__device__ __forceinline__ void fill_tile(uint8_t* tile, const float* __restrict__ g, int row0, int M) {
#pragma unroll 4
    for (int idx = threadIdx.x; idx < 128 * 32; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)(row0 + r) * D) + c4);
        *reinterpret_cast<float4*>(tile + tile_off4(r, c4)) = v;
    }
}


// ================================================================================================
// BF16 operand tiles: [128 rows][128 bf16] = 2 column chunks of 64 bf16 (128 B); chunk = [128 rows x
// 128 B] SWIZZLE_128B (32 KB per tile).  kind::f16 with BF16 inputs, fp32 accumulation, K = 16 per MMA.
// The same bytes serve as a K-major operand (rows = M/N) and as an MN-major operand (columns = M/N,
// rows = K): for 16-bit types the MN-major SWIZZLE_128B atom is 64 elements x 8 rows = one 1024 B group.
// ================================================================================================
constexpr int TILE16_BYTES = 128 * 128 * 2;
constexpr int CHUNK16_BYTES = 128 * 128;

// byte offset of the 16-byte unit holding columns [8*u, 8*u+8) of row r
__device__ __forceinline__ uint32_t tile16_off8(int r, int u) {
    return (uint32_t)((u >> 3) * CHUNK16_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + (((u & 7) ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half = a, high half = b
    return r;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// K-major view: k-step ks (16 bf16 = 32 B) of chunk c
__device__ __forceinline__ uint64_t desc16_k(uint32_t tile_addr, int c, int ks) {
    return make_desc(tile_addr + c * CHUNK16_BYTES + ks * 32, 16, 1024);
}
// MN-major view: k-step = 16 rows = two 1024 B groups (SBO apart); the 2 chunks are the MN atoms (LBO apart)
__device__ __forceinline__ uint64_t desc16_mn(uint32_t tile_addr, int kstep) {
    return make_desc(tile_addr + kstep * 2048, CHUNK16_BYTES, 1024);
}
// D[128 x 128] (+)= A * B^T, both K-major bf16 tiles (K = 128 = 2 chunks x 4 k-steps)
__device__ __forceinline__ void issue_gemm16_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
    constexpr uint32_t id = idesc_bf16(128, false, false);
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            mma_bf16(tmem_d, desc16_k(a_addr, c, ks), desc16_k(b_addr, c, ks), id, (accumulate || c || ks) ? 1u : 0u);
}
// global fp32 [M,128] rows row0.. -> bf16 tile (warp per row; rows >= M zero).  All 16 loads of a thread are
// issued before the first conversion so that a CTA keeps 64 KB in flight.
__device__ __forceinline__ void fill_tile16(uint8_t* tile, const float* __restrict__ g, int row0, int M) {
    const int c4 = threadIdx.x & 31, rb = threadIdx.x >> 5;
    float4 v[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int r = it * 8 + rb;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v[it] = __ldg(reinterpret_cast<const float4*>(g + (size_t)(row0 + r) * D) + c4);
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int r = it * 8 + rb;
        *reinterpret_cast<uint2*>(tile + tile16_off8(r, c4 >> 1) + (c4 & 1) * 8) =
            make_uint2(pack_bf16(v[it].x, v[it].y), pack_bf16(v[it].z, v[it].w));
    }
}
// registers (one row, 32 consecutive fp32 columns starting at c0, c0 % 32 == 0) -> bf16 tile
__device__ __forceinline__ void tile16_store32(uint8_t* tile, int row, int c0, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(tile + tile16_off8(row, (c0 >> 3) + j)) =
            make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
}

}  // namespace tc
}  // namespace amid
