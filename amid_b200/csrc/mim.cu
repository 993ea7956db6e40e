// Multi-interest information module: InterComp / InnerComp (model_seq.py:474-497 /
// 450-472) in the exact closed form of SURVEY.md section 8a-6.  The reference builds
// [bs,B,n,n] tensors whose outer axis is redundant; here the only O(n^2 d) work is one
// similarity-max per sample (k_mim_scores), everything else is O(B n d) or smaller.
#include "common.cuh"
#include "tc.cuh"

namespace amid {

constexpr int MT = 64;        // tile edge of the similarity GEMM
constexpr int MLD = D + 4;    // smem row stride
constexpr size_t MIM_SMEM = (size_t)2 * MT * MLD * sizeof(float);

// m[j] = max_{s,t} <a[j,s], b[j,t]>.  One CTA per sample, 256 threads, 4x4 micro-tiles.
__global__ void __launch_bounds__(256)
k_mim_scores(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ m) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float wmax[8];
    float* As = smem;
    float* Bs = smem + MT * MLD;
    const int j = blockIdx.x;
    const float* aj = a + (size_t)j * n * D;
    const float* bj = b + (size_t)j * n * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = lane & 15, ty = warp * 2 + (lane >> 4);   // rows ty+16*i of A, rows tx+16*jj of B
    float best = -INFINITY;
    for (int s0 = 0; s0 < n; s0 += MT) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < MT * (D / 4); idx += 256) {
            int r = idx >> 5, c4 = idx & 31;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + r < n) v = __ldg(reinterpret_cast<const float4*>(aj + (size_t)(s0 + r) * D) + c4);
            *reinterpret_cast<float4*>(As + r * MLD + c4 * 4) = v;
        }
        for (int t0 = 0; t0 < n; t0 += MT) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < MT * (D / 4); idx += 256) {
                int r = idx >> 5, c4 = idx & 31;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t0 + r < n) v = __ldg(reinterpret_cast<const float4*>(bj + (size_t)(t0 + r) * D) + c4);
                *reinterpret_cast<float4*>(Bs + r * MLD + c4 * 4) = v;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
#pragma unroll 4
            for (int k = 0; k < D; k += 4) {
                float4 av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * MLD + k);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) bv[jj] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * jj) * MLD + k);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        acc[i][jj] = fmaf(av[i].x, bv[jj].x, acc[i][jj]);
                        acc[i][jj] = fmaf(av[i].y, bv[jj].y, acc[i][jj]);
                        acc[i][jj] = fmaf(av[i].z, bv[jj].z, acc[i][jj]);
                        acc[i][jj] = fmaf(av[i].w, bv[jj].w, acc[i][jj]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (s0 + ty + 16 * i < n && t0 + tx + 16 * jj < n) best = fmaxf(best, acc[i][jj]);
        }
    }
    best = warp_max(best);
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = wmax[0];
        for (int w = 1; w < 8; ++w) r = fmaxf(r, wmax[w]);
        m[j] = r;
    }
}

// Tensor-core version: the 64x64 similarity block as mma.sync m16n8k8 with the 3xTF32 split
// (a = a_hi + a_lo, products a_hi b_hi + a_hi b_lo + a_lo b_hi), which keeps the scores at fp32
// accuracy -- the hard gate p > ts downstream must not depend on the precision mode.
constexpr int MLD2 = D + 4;    // 132: fragment loads are bank-conflict free (4g + t distinct)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256)
k_mim_scores_mma(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ m) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float wmax[8];
    float* As = smem;
    float* Bs = smem + MT * MLD2;
    const int j = blockIdx.x;
    const float* aj = a + (size_t)j * n * D;
    const float* bj = b + (size_t)j * n * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int rw = (warp & 3) * 16;      // this warp's 16 rows of the 64-row block
    const int cw = (warp >> 2) * 32;     // and its 32 columns (4 n-tiles)
    float best = -INFINITY;
    for (int s0 = 0; s0 < n; s0 += MT) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < MT * (D / 4); idx += 256) {
            int r = idx >> 5, c4 = idx & 31;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + r < n) v = __ldg(reinterpret_cast<const float4*>(aj + (size_t)(s0 + r) * D) + c4);
            *reinterpret_cast<float4*>(As + r * MLD2 + c4 * 4) = v;
        }
        for (int t0 = 0; t0 < n; t0 += MT) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < MT * (D / 4); idx += 256) {
                int r = idx >> 5, c4 = idx & 31;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t0 + r < n) v = __ldg(reinterpret_cast<const float4*>(bj + (size_t)(t0 + r) * D) + c4);
                *reinterpret_cast<float4*>(Bs + r * MLD2 + c4 * 4) = v;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            const float* pa = As + (rw + g) * MLD2 + t;
#pragma unroll 4
            for (int k0 = 0; k0 < D; k0 += 8) {
                uint32_t ah[4], al[4];
                split_tf32(pa[k0], ah[0], al[0]);
                split_tf32(pa[8 * MLD2 + k0], ah[1], al[1]);
                split_tf32(pa[k0 + 4], ah[2], al[2]);
                split_tf32(pa[8 * MLD2 + k0 + 4], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float* pb = Bs + (cw + 8 * nt + g) * MLD2 + k0 + t;
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(pb[0], bh0, bl0);
                    split_tf32(pb[4], bh1, bl1);
                    mma8(acc[nt], al, bh0, bh1);
                    mma8(acc[nt], ah, bl0, bl1);
                    mma8(acc[nt], ah, bh0, bh1);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int ra = s0 + rw + g, rb = ra + 8, c = t0 + cw + 8 * nt + 2 * t;
                if (ra < n && c < n) best = fmaxf(best, acc[nt][0]);
                if (ra < n && c + 1 < n) best = fmaxf(best, acc[nt][1]);
                if (rb < n && c < n) best = fmaxf(best, acc[nt][2]);
                if (rb < n && c + 1 < n) best = fmaxf(best, acc[nt][3]);
            }
        }
    }
    best = warp_max(best);
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = wmax[0];
        for (int w = 1; w < 8; ++w) r = fmaxf(r, wmax[w]);
        m[j] = r;
    }
}

// tcgen05 version (n <= 256): one CTA per sample, 2 CTAs per SM.  The 128-feature contraction is streamed in four
// 32-feature chunks; per chunk the CTA stages hi / lo TF32 halves of 128 rows of a and of all (padded) rows of b as
// K-major SWIZZLE_128B operands and one thread issues 4 k-steps x {lo*hi, hi*lo, hi*hi} into a [128 x NB] fp32
// accumulator in TMEM.  The epilogue takes the masked maximum straight out of TMEM.
namespace mimtc {
using namespace tc;
constexpr int A_BYTES = 128 * 128;            // one operand half: [128 rows][32 fp32]
__host__ __device__ inline int nb_rows(int n) { return (n + 31) & ~31; }
__host__ __device__ inline size_t smem_bytes(int n) { return (size_t)2 * A_BYTES + (size_t)2 * nb_rows(n) * 128 + 1024; }
__device__ __forceinline__ uint32_t chunk_off4(int r, int c4) {      // 16-byte unit c4 (0..7) of row r
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c4 ^ (r & 7)) << 4));
}
__device__ __forceinline__ void store_split(uint8_t* hi, uint8_t* lo, uint32_t off, const float4 v) {
    uint4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
}
// rows [row0, row0 + rows) x features [32 kc, 32 kc + 32) of g[n,128] -> hi / lo chunk (rows >= n zero)
__device__ __forceinline__ void fill_split(uint8_t* hi, uint8_t* lo, const float* __restrict__ g, int row0, int rows,
                                           int n, int kc) {
    const int c4 = threadIdx.x & 7, rb = threadIdx.x >> 3;             // 32 rows x 8 units per pass
    for (int r = rb; r < rows; r += 128) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r + 32 * u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr < rows && row0 + rr < n) v[u] = __ldg(reinterpret_cast<const float4*>(g + (size_t)(row0 + rr) * D + kc * 32) + c4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r + 32 * u;
            if (rr < rows) store_split(hi, lo, chunk_off4(rr, c4), v[u]);
        }
    }
}
}  // namespace mimtc

__global__ void __launch_bounds__(256, 2)
k_mim_scores_tc5(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ m) {
    using namespace mimtc;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float wmax[8];
    uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    const int NB = nb_rows(n);
    uint8_t* Ahi = base;
    uint8_t* Alo = Ahi + A_BYTES;
    uint8_t* Bhi = Alo + A_BYTES;
    uint8_t* Blo = Bhi + NB * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x;
    const float* aj = a + (size_t)j * n * D;
    const float* bj = b + (size_t)j * n * D;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t id = idesc_tf32(NB, false, false);
    uint32_t phase = 0;
    float best = -INFINITY;
    for (int m0 = 0; m0 < n; m0 += 128) {
        for (int kc = 0; kc < 4; ++kc) {
            if (kc) { mbar_wait(&bar, phase); phase ^= 1; }          // the previous chunk's MMAs have read the stage
            fill_split(Ahi, Alo, aj, m0, 128, n, kc);
            fill_split(Bhi, Blo, bj, 0, NB, n, kc);
            fence_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                fence_after();
                const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), bh = smem_u32(Bhi), bl = smem_u32(Blo);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dah = make_desc(ah + ks * 32, 16, 1024), dal = make_desc(al + ks * 32, 16, 1024);
                    const uint64_t dbh = make_desc(bh + ks * 32, 16, 1024), dbl = make_desc(bl + ks * 32, 16, 1024);
                    mma_tf32(tmem, dal, dbh, id, (kc || ks) ? 1u : 0u);
                    mma_tf32(tmem, dah, dbl, id, 1u);
                    mma_tf32(tmem, dah, dbh, id, 1u);
                }
                mma_commit(&bar);
            }
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        fence_after();
        const int row = m0 + 32 * (warp & 3) + lane;
        for (int cg = (warp >> 2); cg * 32 < NB; cg += 2) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cg * 32, v);
            if (row < n) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (cg * 32 + i < n) best = fmaxf(best, v[i]);
            }
        }
        fence_before();
        __syncthreads();                                              // TMEM drained before the next row tile overwrites it
        fence_after();
    }
    best = warp_max(best);
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = wmax[0];
        for (int w = 1; w < 8; ++w) r = fmaxf(r, wmax[w]);
        m[j] = r;
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// softmax over the batch + hard gate + ordered compaction.  Single CTA, 1024 threads.
__global__ void __launch_bounds__(1024)
k_mim_gate(const float* __restrict__ m, const float* __restrict__ w_bs, int B, float ts, float* __restrict__ p,
           float* __restrict__ gate, float* __restrict__ coef, int* __restrict__ active, int* __restrict__ n_active,
           float* __restrict__ scal) {
    __shared__ float red[32];
    __shared__ int cnt[1024];
    __shared__ float bcast;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // max
    float mx = -INFINITY;
    for (int j = t; j < B; j += 1024) mx = fmaxf(mx, m[j]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (t == 0) { float r = red[0]; for (int w = 1; w < 32; ++w) r = fmaxf(r, red[w]); bcast = r; }
    __syncthreads();
    mx = bcast;
    __syncthreads();
    // sum of exp, fixed order: per-thread strided partial, warp tree, then warp 0 sequential
    float se = 0.f, sw = 0.f;
    for (int j = t; j < B; j += 1024) { se += expf(m[j] - mx); sw += w_bs[j]; }
    se = warp_sum(se);
    if (lane == 0) red[warp] = se;
    __syncthreads();
    if (t == 0) { float r = 0.f; for (int w = 0; w < 32; ++w) r += red[w]; bcast = r; }
    __syncthreads();
    const float denom = bcast;
    __syncthreads();
    sw = warp_sum(sw);
    if (lane == 0) red[warp] = sw;
    __syncthreads();
    if (t == 0) { float r = 0.f; for (int w = 0; w < 32; ++w) r += red[w]; scal[0] = r; }
    // gates; contiguous ranges per thread so the compaction keeps ascending order
    const int per = (B + 1023) / 1024;
    const int j0 = t * per, j1 = min(B, j0 + per);
    int c = 0;
    for (int j = j0; j < j1; ++j) {
        const float pj = expf(m[j] - mx) / denom;
        const bool g = pj > ts;               // getBinaryTensor, model_seq.py:445-448
        p[j] = pj;
        gate[j] = g ? 1.f : 0.f;
        coef[j] = g ? w_bs[j] : 0.f;
        c += g;
    }
    cnt[t] = c;
    __syncthreads();
    // exclusive scan of cnt (Hillis-Steele, 1024 entries)
    for (int off = 1; off < 1024; off <<= 1) {
        int v = t >= off ? cnt[t - off] : 0;
        __syncthreads();
        cnt[t] += v;
        __syncthreads();
    }
    int pos = cnt[t] - c;
    for (int j = j0; j < j1; ++j)
        if (gate[j] != 0.f) active[pos++] = j;
    if (t == 1023) n_active[0] = cnt[1023];
}

// Ssum[e] = sum over active local samples of coef * other  (float4 per thread)
__global__ void k_mim_aggregate(const float* __restrict__ other, const float* __restrict__ coef,
                                const int* __restrict__ active, const int* __restrict__ n_active, int j0, int Bl, int n,
                                float* __restrict__ Ssum) {
    const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 >= n * D / 4) return;
    const int na = n_active[0];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < na; ++a) {
        const int j = active[a];
        if (j < j0 || j >= j0 + Bl) continue;
        const float c = coef[j];
        const float4 v = __ldg(reinterpret_cast<const float4*>(other + (size_t)(j - j0) * n * D) + e4);
        s.x = fmaf(c, v.x, s.x); s.y = fmaf(c, v.y, s.y); s.z = fmaf(c, v.z, s.z); s.w = fmaf(c, v.w, s.w);
    }
    reinterpret_cast<float4*>(Ssum)[e4] = s;
}

// E[t][c] = sum_k Ssum[t][k] W[c][k] + scal*b_nn[c] + b_bs     (one CTA per row t, 128 threads)
__global__ void __launch_bounds__(128)
k_mim_project(const float* __restrict__ Ssum, const float* __restrict__ W, const float* __restrict__ b_nn,
              const float* __restrict__ b_bs, const float* __restrict__ scal, float* __restrict__ E) {
    __shared__ __align__(16) float row[D];
    const int t = blockIdx.x, c = threadIdx.x;
    row[c] = Ssum[(size_t)t * D + c];
    __syncthreads();
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)c * D);
    float s = 0.f;
#pragma unroll 8
    for (int k4 = 0; k4 < D / 4; ++k4) {
        const float4 w = __ldg(w4 + k4);
        const float4 r = *reinterpret_cast<const float4*>(row + k4 * 4);
        s = fmaf(r.x, w.x, s); s = fmaf(r.y, w.y, s); s = fmaf(r.z, w.z, s); s = fmaf(r.w, w.w, s);
    }
    E[(size_t)t * D + c] = s + scal[0] * b_nn[c] + b_bs[0];
}
// out[c] = scale * sum_t X[t][c]      (1 CTA, 128 threads)
__global__ void k_colsum(const float* __restrict__ X, int n, float* __restrict__ out) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int t = 0; t < n; ++t) s += X[(size_t)t * D + c];
    out[c] = s;
}

__global__ void k_mim_concat(const float* __restrict__ self_, const float* __restrict__ E, int64_t B, int n,
                             float* __restrict__ out) {
    const int64_t e4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)2 * n * D / 4;
    if (e4 >= B * per) return;
    const int64_t i = e4 / per, r = e4 % per;
    const int64_t half = (int64_t)n * D / 4;
    float4 v = r < half ? __ldg(reinterpret_cast<const float4*>(self_) + i * half + r)
                        : __ldg(reinterpret_cast<const float4*>(E) + (r - half));
    reinterpret_cast<float4*>(out)[e4] = v;
}

// ---- backward pieces
// dS[t][k] = sum_c dE[t][c] W[c][k]    (CTA per row t, thread k: coalesced rows of W)
__global__ void __launch_bounds__(128)
k_mim_dS(const float* __restrict__ dE, const float* __restrict__ W, float* __restrict__ dS) {
    __shared__ float row[D];
    const int t = blockIdx.x, k = threadIdx.x;
    row[k] = dE[(size_t)t * D + k];
    __syncthreads();
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < D; ++c) s = fmaf(row[c], __ldg(W + (size_t)c * D + k), s);
    dS[(size_t)t * D + k] = s;
}
// dW[c][k] = sum_t dE[t][c] Ssum[t][k] ; CTA per c, thread k.  Also db_nn, db_bs by CTA 0..: see below
__global__ void __launch_bounds__(128)
k_mim_dW(const float* __restrict__ dE, const float* __restrict__ Ssum, const float* __restrict__ scal, int n,
         float* __restrict__ dW, float* __restrict__ db_nn, float* __restrict__ db_bs) {
    __shared__ float red[4];
    const int c = blockIdx.x, k = threadIdx.x;
    float s = 0.f, cs = 0.f;
    for (int t = 0; t < n; ++t) {
        const float g = dE[(size_t)t * D + c];
        s = fmaf(g, Ssum[(size_t)t * D + k], s);
        cs += g;
    }
    dW[(size_t)c * D + k] = s;
    if (k == 0) db_nn[c] = scal[0] * cs;
    if (c == 0) {  // db_bs = sum of all dE: thread k sums column k, then fixed-order reduce
        float col = 0.f;
        for (int t = 0; t < n; ++t) col += dE[(size_t)t * D + k];
        col = warp_sum(col);
        if ((k & 31) == 0) red[k >> 5] = col;
        __syncthreads();
        if (k == 0) db_bs[0] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}
// per local sample j: dw_bs[j] = gate_j <dS, other[j]> + <colsum(dE), b_nn> ; d_other[j] (+)= coef_j dS if gate_j
__global__ void __launch_bounds__(256)
k_mim_dsample(const float* __restrict__ dS, const float* __restrict__ dE, const float* __restrict__ other,
              const float* __restrict__ b_nn, const float* __restrict__ coef, const float* __restrict__ gate, int j0,
              int n, int accumulate, float* __restrict__ dw_bs, float* __restrict__ d_other) {
    __shared__ float red[8];
    const int j = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool on = gate[j0 + j] != 0.f;
    // constant term <colsum(dE), b_nn> = sum_{t,c} dE[t][c] b_nn[c]
    float s = 0.f;
    const int total4 = n * D / 4;
    for (int e4 = t; e4 < total4; e4 += 256) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dE) + e4);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b_nn) + (e4 & 31));
        s += g.x * bb.x + g.y * bb.y + g.z * bb.z + g.w * bb.w;
        if (on) {
            const float4 ds = __ldg(reinterpret_cast<const float4*>(dS) + e4);
            const float4 o = __ldg(reinterpret_cast<const float4*>(other + (size_t)j * n * D) + e4);
            s += ds.x * o.x + ds.y * o.y + ds.z * o.z + ds.w * o.w;
            if (d_other) {
                const float c = coef[j0 + j];
                float4* dst = reinterpret_cast<float4*>(d_other + (size_t)j * n * D) + e4;
                float4 cur = accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
                cur.x = fmaf(c, ds.x, cur.x); cur.y = fmaf(c, ds.y, cur.y); cur.z = fmaf(c, ds.z, cur.z); cur.w = fmaf(c, ds.w, cur.w);
                *dst = cur;
            }
        }
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (t == 0) {
        float r = 0.f;
        for (int w = 0; w < 8; ++w) r += red[w];
        dw_bs[j] = r;
    }
}

// ---- mean pool
__global__ void __launch_bounds__(128)
k_meanpool_fwd(const float* __restrict__ enc, const float* __restrict__ esum, int n, float inv, float* __restrict__ u) {
    // CTA per sample, 4 warps stride over t, lanes over float4 columns
    __shared__ float4 red[4][32];
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = warp; t < n; t += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(enc + ((size_t)i * n + t) * D) + lane);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float4 r = red[0][lane];
        for (int w = 1; w < 4; ++w) { r.x += red[w][lane].x; r.y += red[w][lane].y; r.z += red[w][lane].z; r.w += red[w][lane].w; }
        if (esum) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(esum) + lane);
            r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
        }
        reinterpret_cast<float4*>(u + (size_t)i * D)[lane] = make_float4(r.x * inv, r.y * inv, r.z * inv, r.w * inv);
    }
}
__global__ void k_meanpool_bwd(const float* __restrict__ du, int64_t B, int n, float inv, int accumulate,
                               float* __restrict__ d_enc) {
    const int64_t e4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 >= B * n * (D / 4)) return;
    const int64_t i = e4 / ((int64_t)n * (D / 4));
    const int c4 = (int)(e4 & 31);
    const float4 g = __ldg(reinterpret_cast<const float4*>(du + i * D) + c4);
    float4* dst = reinterpret_cast<float4*>(d_enc) + e4;
    float4 cur = accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
    cur.x = fmaf(g.x, inv, cur.x); cur.y = fmaf(g.y, inv, cur.y); cur.z = fmaf(g.z, inv, cur.z); cur.w = fmaf(g.w, inv, cur.w);
    *dst = cur;
}
// dcol[c] = inv * sum_i du[i][c]   (1 CTA, 1024 threads = 8 row groups x 128 columns, fixed-order combine)
__global__ void __launch_bounds__(1024)
k_du_colsum(const float* __restrict__ du, int B, float inv, float* __restrict__ dcol) {
    __shared__ float red[8][D];
    const int c = threadIdx.x & (D - 1), g = threadIdx.x >> 7;
    float s = 0.f;
#pragma unroll 8
    for (int i = g; i < B; i += 8) s += du[(size_t)i * D + c];
    red[g][c] = s;
    __syncthreads();
    if (g == 0) {
        float t = red[0][c];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += red[k][c];
        dcol[c] = t * inv;
    }
}

}  // namespace amid

using namespace amid;

extern "C" int amid_mim_scores(const float* a, const float* b, int32_t B, int32_t n, float* m, amid_stream_t s_) {
    AMID_REQUIRE(a && b && m && B > 0 && n > 0, "mim_scores: bad argument");
    AMID_REQUIRE(aligned16(a) && aligned16(b), "mim_scores: misaligned buffer");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_mim_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIM_SMEM);
    if (e != cudaSuccess) return set_error(-3, "mim_scores: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_mim_scores", (cudaStream_t)s_);
    k_mim_scores<<<B, 256, MIM_SMEM, (cudaStream_t)s_>>>(a, b, n, m);
    AMID_LAUNCH_CHECK("k_mim_scores");
    return 0;
}

extern "C" int amid_mim_scores_tc(const float* a, const float* b, int32_t B, int32_t n, float* m, amid_stream_t s_) {
    AMID_REQUIRE(a && b && m && B > 0 && n > 0, "mim_scores_tc: bad argument");
    AMID_REQUIRE(aligned16(a) && aligned16(b), "mim_scores_tc: misaligned buffer");
    if (n <= 256) {                           // tcgen05: the whole key axis is one MMA N extent
        const size_t smem = mimtc::smem_bytes(n);
        cudaError_t e = cudaFuncSetAttribute((const void*)k_mim_scores_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-3, "mim_scores_tc: smem attribute: %s", cudaGetErrorString(e));
        AMID_K("k_mim_scores_tc5", s_);
        k_mim_scores_tc5<<<B, 256, smem, (cudaStream_t)s_>>>(a, b, n, m);
        AMID_LAUNCH_CHECK("k_mim_scores_tc5");
        return 0;
    }
    const size_t smem = (size_t)2 * MT * MLD2 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute((const void*)k_mim_scores_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-3, "mim_scores_tc: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_mim_scores_mma", s_);
    k_mim_scores_mma<<<B, 256, smem, (cudaStream_t)s_>>>(a, b, n, m);
    AMID_LAUNCH_CHECK("k_mim_scores_mma");
    return 0;
}

extern "C" int amid_mim_gate(const float* m, const float* w_bs, int32_t Bg, float ts, float* p, float* gate, float* coef,
                             int32_t* active, int32_t* n_active, float* scal, amid_stream_t s_) {
    AMID_REQUIRE(m && w_bs && p && gate && coef && active && n_active && scal && Bg > 0, "mim_gate: bad argument");
    AMID_K("k_mim_gate", (cudaStream_t)s_);
    k_mim_gate<<<1, 1024, 0, (cudaStream_t)s_>>>(m, w_bs, Bg, ts, p, gate, coef, active, n_active, scal);
    AMID_LAUNCH_CHECK("k_mim_gate");
    return 0;
}

extern "C" int amid_mim_aggregate(const float* other, const float* coef, const int32_t* active, const int32_t* n_active,
                                  int32_t j0, int32_t Bl, int32_t n, float* Ssum, amid_stream_t s_) {
    AMID_REQUIRE(other && coef && active && n_active && Ssum && n > 0 && Bl > 0, "mim_aggregate: bad argument");
    const int total4 = n * D / 4;
    AMID_K("k_mim_aggregate", (cudaStream_t)s_);
    k_mim_aggregate<<<(total4 + 127) / 128, 128, 0, (cudaStream_t)s_>>>(other, coef, active, n_active, j0, Bl, n, Ssum);
    AMID_LAUNCH_CHECK("k_mim_aggregate");
    return 0;
}

extern "C" int amid_mim_project(const float* Ssum, const float* w_nn, const float* b_nn, const float* b_bs,
                                const float* scal, int32_t n, float* E, float* esum, amid_stream_t s_) {
    AMID_REQUIRE(Ssum && w_nn && b_nn && b_bs && scal && E && n > 0, "mim_project: bad argument");
    AMID_K("k_mim_project", (cudaStream_t)s_);
    k_mim_project<<<n, 128, 0, (cudaStream_t)s_>>>(Ssum, w_nn, b_nn, b_bs, scal, E);
    AMID_LAUNCH_CHECK("k_mim_project");
    if (esum) {
        AMID_K("k_colsum", (cudaStream_t)s_);
        k_colsum<<<1, 128, 0, (cudaStream_t)s_>>>(E, n, esum);
        AMID_LAUNCH_CHECK("k_colsum");
    }
    return 0;
}

extern "C" int amid_mim_concat(const float* self_, const float* E, int32_t B, int32_t n, float* out, amid_stream_t s_) {
    AMID_REQUIRE(self_ && E && out && B > 0 && n > 0, "mim_concat: bad argument");
    const int64_t total4 = (int64_t)B * 2 * n * D / 4;
    AMID_K("k_mim_concat", (cudaStream_t)s_);
    k_mim_concat<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(self_, E, B, n, out);
    AMID_LAUNCH_CHECK("k_mim_concat");
    return 0;
}

extern "C" int amid_mim_bwd(const float* dE, const float* Ssum, const float* other, const float* w_nn, const float* b_nn,
                            const float* coef, const float* gate, const int32_t* active, const int32_t* n_active,
                            const float* scal, int32_t j0, int32_t Bl, int32_t n, float* dW_nn, float* db_nn,
                            float* db_bs, float* dw_bs, float* d_other, float* ws_dS, amid_stream_t s_) {
    (void)active; (void)n_active;
    AMID_REQUIRE(dE && Ssum && other && w_nn && b_nn && coef && gate && scal && dW_nn && db_nn && db_bs && dw_bs && ws_dS,
                 "mim_bwd: null argument");
    AMID_REQUIRE(Bl > 0 && n > 0, "mim_bwd: Bl=%d n=%d", Bl, n);
    cudaStream_t s = (cudaStream_t)s_;
    AMID_K("k_mim_dS", s);
    k_mim_dS<<<n, 128, 0, s>>>(dE, w_nn, ws_dS);
    AMID_LAUNCH_CHECK("k_mim_dS");
    AMID_K("k_mim_dW", s);
    k_mim_dW<<<D, 128, 0, s>>>(dE, Ssum, scal, n, dW_nn, db_nn, db_bs);
    AMID_LAUNCH_CHECK("k_mim_dW");
    AMID_K("k_mim_dsample", s);
    k_mim_dsample<<<Bl, 256, 0, s>>>(ws_dS, dE, other, b_nn, coef, gate, j0, n, d_other ? 1 : 0, dw_bs, d_other);
    AMID_LAUNCH_CHECK("k_mim_dsample");
    return 0;
}

extern "C" int amid_meanpool_fwd(const float* enc, const float* esum, int32_t B, int32_t n, float denom, float* u,
                                 amid_stream_t s_) {
    AMID_REQUIRE(enc && u && B > 0 && n > 0 && denom > 0.f, "meanpool_fwd: bad argument");
    AMID_K("k_meanpool_fwd", (cudaStream_t)s_);
    k_meanpool_fwd<<<B, 128, 0, (cudaStream_t)s_>>>(enc, esum, n, 1.0f / denom, u);
    AMID_LAUNCH_CHECK("k_meanpool_fwd");
    return 0;
}

extern "C" int amid_meanpool_bwd(const float* du, int32_t B, int32_t n, float denom, int32_t accumulate, float* d_enc,
                                 float* dcol, amid_stream_t s_) {
    AMID_REQUIRE(du && B > 0 && n > 0 && denom > 0.f, "meanpool_bwd: bad argument");
    cudaStream_t s = (cudaStream_t)s_;
    if (d_enc) {
        const int64_t total4 = (int64_t)B * n * (D / 4);
        AMID_K("k_meanpool_bwd", s);
        k_meanpool_bwd<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(du, B, n, 1.0f / denom, accumulate, d_enc);
        AMID_LAUNCH_CHECK("k_meanpool_bwd");
    }
    if (dcol) {
        AMID_K("k_du_colsum", s);
        k_du_colsum<<<1, 1024, 0, s>>>(du, B, 1.0f / denom, dcol);
        AMID_LAUNCH_CHECK("k_du_colsum");
    }
    return 0;
}
