// Item-embedding gather (model_seq.py:27-29, :418-421), fused with the Log2feats
// prologue (model_seq.py:360-366): positional add, feature-level timeline mask, dropout.
// HBM-bound: one warp moves one 512-byte row with a single 128-bit load per lane; four
// rows are in flight per warp so each SM keeps > 40 KB of loads outstanding.
#include "common.cuh"

namespace amid {

constexpr int ROWS_PER_WARP = 4;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// plain gather: out[r] = table[ids[r]]
__global__ void __launch_bounds__(256)
k_gather(const float* __restrict__ table, const int64_t* __restrict__ ids, int64_t n_rows, int64_t V,
         float* __restrict__ out, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = warp * ROWS_PER_WARP;
    if (r0 >= n_rows) return;
    int64_t id = 0;
    if (lane < ROWS_PER_WARP && r0 + lane < n_rows) id = __ldg(ids + r0 + lane);
    float4 v[ROWS_PER_WARP];
    bool ok[ROWS_PER_WARP];
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
        const int64_t idu = __shfl_sync(0xffffffffu, id, u);
        ok[u] = (r0 + u < n_rows);
        if (ok[u] && (idu < 0 || idu >= V)) { ok[u] = false; if (lane == 0) atomicExch(err, 1); }
        if (ok[u]) v[u] = ldg_stream(reinterpret_cast<const float4*>(table + idu * D) + lane);
    }
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u)
        if (ok[u]) stg_stream(reinterpret_cast<float4*>(out + (r0 + u) * D) + lane, v[u]);
}

// fused: x0 = dropout(src_row + pos[l]) * ~tmask ; tmask bits written per row
template <bool FROM_TABLE>
__global__ void __launch_bounds__(256)
k_seq_embed(const float* __restrict__ table, const int64_t* __restrict__ ids, const float* __restrict__ rows,
            const float* __restrict__ pos, int64_t n_rows, int L, int64_t V, float* __restrict__ x0,
            uint32_t* __restrict__ tmask, DropCfg dc, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = warp * ROWS_PER_WARP;
    if (r0 >= n_rows) return;
    int64_t id = 0;
    if (FROM_TABLE && lane < ROWS_PER_WARP && r0 + lane < n_rows) id = __ldg(ids + r0 + lane);
    float4 v[ROWS_PER_WARP];
    bool ok[ROWS_PER_WARP];
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
        ok[u] = (r0 + u < n_rows);
        if (FROM_TABLE) {
            const int64_t idu = __shfl_sync(0xffffffffu, id, u);
            if (ok[u] && (idu < 0 || idu >= V)) { ok[u] = false; if (lane == 0) atomicExch(err, 1); }
            if (ok[u]) v[u] = ldg_stream(reinterpret_cast<const float4*>(table + idu * D) + lane);
        } else {
            if (ok[u]) v[u] = ldg_stream(reinterpret_cast<const float4*>(rows + (r0 + u) * D) + lane);
        }
    }
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
        if (!ok[u]) continue;   // warp-uniform
        const int64_t r = r0 + u;
        const int l = (int)(r % L);
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)l * D) + lane);
        float4 x = make_float4(v[u].x + p.x, v[u].y + p.y, v[u].z + p.z, v[u].w + p.w);   // model_seq.py:362
        // timeline mask on the pre-dropout features (model_seq.py:365)
        const uint32_t w0 = __ballot_sync(0xffffffffu, x.x == 0.f);
        const uint32_t w1 = __ballot_sync(0xffffffffu, x.y == 0.f);
        const uint32_t w2 = __ballot_sync(0xffffffffu, x.z == 0.f);
        const uint32_t w3 = __ballot_sync(0xffffffffu, x.w == 0.f);
        if (lane == 0) *reinterpret_cast<uint4*>(tmask + r * 4) = make_uint4(w0, w1, w2, w3);
        if (dc.train) x = drop4(x, dc, dc.site_base + SITE_EMB, (uint64_t)(r + dc.tok_off) * D + lane * 4);   // :363
        // `seqs *= ~timeline_mask` (:366) is a value no-op: masked elements are already 0
        stg_stream(reinterpret_cast<float4*>(x0 + r * D) + lane, x);
    }
}

// One launch for every table row a training / eval step reads (model_seq.py:418-421 + the two Log2feats
// prologues :360-366): rows [0, n_items) = candidates (plain copy), then seq_d1, then seq_d2 (fused
// pos add, timeline-mask bits, dropout).  8 rows in flight per warp.
struct EmbedAll {
    const int64_t *ids0, *ids1, *ids2;
    const float *pos1, *pos2;
    float *out0, *out1, *out2;
    uint32_t *tm1, *tm2;
    uint32_t n0, n1, n2;          // rows per segment (each < 2^31)
};
constexpr int RPW2 = 8;
// CTAs never straddle a segment (each segment is rounded up to whole CTAs of 8 warps x 8 rows), so the segment,
// its pointers and the dropout site are block-uniform and the row loop is straight-line code.
__host__ __device__ inline uint32_t embed_all_ctas(uint32_t n) { return (n + 8 * RPW2 - 1) / (8 * RPW2); }
__global__ void __launch_bounds__(256)
k_embed_all(const float* __restrict__ table, int64_t V, EmbedAll ea, uint32_t L, DropCfg dc, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    uint32_t cta = blockIdx.x;
    const uint32_t c0 = embed_all_ctas(ea.n0), c1 = embed_all_ctas(ea.n1);
    uint32_t n;
    const int64_t* ids;
    float* out;
    const float* pos = nullptr;
    uint32_t* tmk = nullptr;
    uint32_t site = 0;
    if (cta < c0) { n = ea.n0; ids = ea.ids0; out = ea.out0; }
    else if (cta < c0 + c1) { cta -= c0; n = ea.n1; ids = ea.ids1; out = ea.out1; pos = ea.pos1; tmk = ea.tm1; site = SITE_EMB; }
    else { cta -= c0 + c1; n = ea.n2; ids = ea.ids2; out = ea.out2; pos = ea.pos2; tmk = ea.tm2; site = 8u + SITE_EMB; }
    const uint32_t r0 = (cta * 8 + (threadIdx.x >> 5)) * RPW2;
    if (r0 >= n) return;
    int64_t id = 0;
    if (lane < RPW2) id = __ldg(ids + min(r0 + lane, n - 1));          // tail rows re-read the last id, never stored
    const bool bad = id < 0 || id >= V;
    if (__any_sync(0xffffffffu, bad)) {                                  // host raises on the flag; keep the loads in range
        if (lane == 0) atomicExch(err, 1);
        if (bad) id = 0;
    }
    float4 v[RPW2];
#pragma unroll
    for (int u = 0; u < RPW2; ++u) {
        const int64_t idu = __shfl_sync(0xffffffffu, id, u);
        v[u] = ldg_stream(reinterpret_cast<const float4*>(table + idu * D) + lane);
    }
    uint32_t l = pos ? r0 % L : 0u;
#pragma unroll
    for (int u = 0; u < RPW2; ++u) {
        const uint32_t r = r0 + u;
        if (r >= n) break;                                               // warp-uniform, last warp of a segment only
        float4 x = v[u];
        if (pos) {                                                       // warp-uniform
            const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)l * D) + lane);
            x = make_float4(x.x + p.x, x.y + p.y, x.z + p.z, x.w + p.w);                       // model_seq.py:362
            // timeline-mask bits of the pre-dropout features (:365); exact zeros are rare, so vote once first
            uint4 bits = make_uint4(0u, 0u, 0u, 0u);
            if (__any_sync(0xffffffffu, x.x == 0.f || x.y == 0.f || x.z == 0.f || x.w == 0.f)) {
                bits.x = __ballot_sync(0xffffffffu, x.x == 0.f);
                bits.y = __ballot_sync(0xffffffffu, x.y == 0.f);
                bits.z = __ballot_sync(0xffffffffu, x.z == 0.f);
                bits.w = __ballot_sync(0xffffffffu, x.w == 0.f);
            }
            if (lane == 0) *reinterpret_cast<uint4*>(tmk + (size_t)r * 4) = bits;
            if (dc.train) x = drop4(x, dc, site, (uint64_t)(r + dc.tok_off) * D + lane * 4);                  // :363
            if (++l == L) l = 0;
        }
        stg_stream(reinterpret_cast<float4*>(out + (size_t)r * D) + lane, x);
    }
}

// backward: dx0 <- dx0 * ~tmask * keep*scale (in place); dpos[l] = sum_b dx0[b,l].
// One CTA of 32 warps per position l; the warps stride over the batch (4 rows in flight each), then a
// fixed-order cross-warp sum (deterministic, no atomics).
constexpr int SEB_WARPS = 32;
__global__ void __launch_bounds__(SEB_WARPS * 32)
k_seq_embed_bwd(float* __restrict__ dx0, const uint32_t* __restrict__ tmask, int B, int L, float* __restrict__ dpos,
                DropCfg dc) {
    __shared__ float4 red[SEB_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int b = warp; b < B; b += SEB_WARPS) {
        const int64_t r = (int64_t)b * L + l;
        float4 g = *(reinterpret_cast<const float4*>(dx0 + r * D) + lane);
        const uint4 tw = __ldg(reinterpret_cast<const uint4*>(tmask) + r);
        if ((tw.x >> lane) & 1u) g.x = 0.f;
        if ((tw.y >> lane) & 1u) g.y = 0.f;
        if ((tw.z >> lane) & 1u) g.z = 0.f;
        if ((tw.w >> lane) & 1u) g.w = 0.f;
        if (dc.train) g = drop4(g, dc, dc.site_base + SITE_EMB, (uint64_t)(r + dc.tok_off) * D + lane * 4);
        *(reinterpret_cast<float4*>(dx0 + r * D) + lane) = g;
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float4 s = red[0][lane];
#pragma unroll
        for (int w = 1; w < SEB_WARPS; ++w) { s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w; }
        *(reinterpret_cast<float4*>(dpos + (size_t)l * D) + lane) = s;
    }
}

__global__ void k_mask_feature(DropCfg dc, uint32_t site, int64_t n, uint8_t* __restrict__ out) {
    const int64_t e4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 * 4 >= n) return;
    const uint32_t r = rng4(dc.seed, site, (uint64_t)e4);
    for (int u = 0; u < 4; ++u)
        if (e4 * 4 + u < n) out[e4 * 4 + u] = dc.train ? (rng_keep(r, u, dc.thr16) ? 1 : 0) : 1;
}
__global__ void k_mask_attn(DropCfg dc, uint32_t site, int64_t BH, int L, uint8_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= BH * L * L) return;
    const int j = (int)(e % L);
    const int64_t bi = e / L;   // bh*L + i
    const int Lp = (L + 3) & ~3;
    const uint32_t r = rng4(dc.seed, site, ((uint64_t)(bi + (int64_t)dc.bh_off * L) * Lp + j) >> 2);
    out[e] = dc.train ? (rng_keep(r, j & 3, dc.thr16) ? 1 : 0) : 1;
}

// device-side error flag shared by the gather kernels (ids out of range)
static int* g_err_flag = nullptr;
int* err_flag() {
    if (!g_err_flag) {
        if (cudaMalloc(&g_err_flag, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(g_err_flag, 0, sizeof(int));
    }
    return g_err_flag;
}

}  // namespace amid

using namespace amid;

extern "C" int amid_emb_gather_fwd(const float* table, int64_t V, const int64_t* ids, int64_t n_rows, float* out,
                                   amid_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AMID_REQUIRE(table && ids && out, "emb_gather_fwd: null argument");
    AMID_REQUIRE(V > 0 && n_rows >= 0, "emb_gather_fwd: V=%lld n_rows=%lld", (long long)V, (long long)n_rows);
    AMID_REQUIRE(aligned16(table) && aligned16(out), "emb_gather_fwd: misaligned buffer");
    if (n_rows == 0) return 0;
    int* err = err_flag();
    AMID_REQUIRE(err, "emb_gather_fwd: cannot allocate error flag");
    const int64_t warps = (n_rows + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
    const int64_t blocks = (warps + 7) / 8;
    AMID_K("k_gather", stream);
    k_gather<<<(unsigned)blocks, 256, 0, stream>>>(table, ids, n_rows, V, out, err);
    AMID_LAUNCH_CHECK("k_gather");
    return 0;
}

// returns 1 if any gather since the last call saw an out-of-range id (synchronises the device)
extern "C" int amid_gather_error_host_sync(void) {
    int* err = err_flag();
    if (!err) return -1;
    int h = 0;
    cudaMemcpy(&h, err, sizeof(int), cudaMemcpyDeviceToHost);
    if (h) cudaMemset(err, 0, sizeof(int));
    return h;
}

extern "C" int amid_seq_embed_fwd(const float* table, int64_t V, const int64_t* ids, const float* rows,
                                  const float* pos, int32_t B, int32_t L, float* x0, uint32_t* tmask,
                                  const amid_dropout* drop, amid_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AMID_REQUIRE(pos && x0 && tmask, "seq_embed_fwd: null argument");
    AMID_REQUIRE((ids && table) || rows, "seq_embed_fwd: need (table, ids) or rows");
    AMID_REQUIRE(B > 0 && L > 0, "seq_embed_fwd: B=%d L=%d", B, L);
    AMID_REQUIRE(aligned16(pos) && aligned16(x0) && aligned16(tmask) && aligned16(table) && aligned16(rows),
                 "seq_embed_fwd: misaligned buffer");
    int* err = err_flag();
    AMID_REQUIRE(err, "seq_embed_fwd: cannot allocate error flag");
    const int64_t n_rows = (int64_t)B * L;
    const int64_t warps = (n_rows + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    const DropCfg dc = with_offsets(make_drop(drop), L);
    if (ids) {
        AMID_K("k_seq_embed", stream);
        k_seq_embed<true><<<blocks, 256, 0, stream>>>(table, ids, nullptr, pos, n_rows, L, V, x0, tmask, dc, err);
    } else {
        AMID_K("k_seq_embed", stream);
        k_seq_embed<false><<<blocks, 256, 0, stream>>>(nullptr, nullptr, rows, pos, n_rows, L, V, x0, tmask, dc, err);
    }
    AMID_LAUNCH_CHECK("k_seq_embed");
    return 0;
}

extern "C" int amid_embed_all_fwd(const float* table, int64_t V, const int64_t* ids_items, int64_t n_items,
                                  const int64_t* ids_d1, const int64_t* ids_d2, const float* pos_d1, const float* pos_d2,
                                  int32_t B, int32_t L, float* items, float* x0_d1, float* x0_d2, uint32_t* tmask_d1,
                                  uint32_t* tmask_d2, const amid_dropout* drop, amid_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AMID_REQUIRE(table && ids_items && ids_d1 && ids_d2 && pos_d1 && pos_d2 && items && x0_d1 && x0_d2 && tmask_d1 && tmask_d2,
                 "embed_all_fwd: null argument");
    AMID_REQUIRE(V > 0 && n_items >= 0 && B > 0 && L > 0, "embed_all_fwd: bad sizes");
    AMID_REQUIRE(aligned16(table) && aligned16(items) && aligned16(x0_d1) && aligned16(x0_d2) && aligned16(tmask_d1) &&
                 aligned16(tmask_d2) && aligned16(pos_d1) && aligned16(pos_d2), "embed_all_fwd: misaligned buffer");
    int* err = err_flag();
    AMID_REQUIRE(err, "embed_all_fwd: cannot allocate error flag");
    AMID_REQUIRE(n_items < (1ll << 30) && (int64_t)B * L < (1ll << 30), "embed_all_fwd: too many rows for one launch");
    EmbedAll ea;
    ea.ids0 = ids_items; ea.ids1 = ids_d1; ea.ids2 = ids_d2;
    ea.pos1 = pos_d1; ea.pos2 = pos_d2;
    ea.out0 = items; ea.out1 = x0_d1; ea.out2 = x0_d2;
    ea.tm1 = tmask_d1; ea.tm2 = tmask_d2;
    ea.n0 = (uint32_t)n_items; ea.n1 = (uint32_t)((int64_t)B * L); ea.n2 = ea.n1;
    const int64_t ctas = (int64_t)embed_all_ctas(ea.n0) + embed_all_ctas(ea.n1) + embed_all_ctas(ea.n2);
    const DropCfg dc = with_offsets(make_drop(drop), L);
    AMID_K("k_embed_all", stream);
    k_embed_all<<<(unsigned)ctas, 256, 0, stream>>>(table, V, ea, (uint32_t)L, dc, err);
    AMID_LAUNCH_CHECK("k_embed_all");
    return 0;
}

extern "C" int amid_seq_embed_bwd(float* dx0, const uint32_t* tmask, int32_t B, int32_t L, float* dpos,
                                  const amid_dropout* drop, amid_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AMID_REQUIRE(dx0 && tmask && dpos, "seq_embed_bwd: null argument");
    AMID_REQUIRE(B > 0 && L > 0, "seq_embed_bwd: B=%d L=%d", B, L);
    const DropCfg dc = with_offsets(make_drop(drop), L);
    AMID_K("k_seq_embed_bwd", stream);
    k_seq_embed_bwd<<<L, SEB_WARPS * 32, 0, stream>>>(dx0, tmask, B, L, dpos, dc);
    AMID_LAUNCH_CHECK("k_seq_embed_bwd");
    return 0;
}

extern "C" int amid_dropout_mask_feature(const amid_dropout* drop, uint32_t site, int64_t rows, uint8_t* out,
                                         amid_stream_t stream_) {
    const DropCfg dc = make_drop(drop);            // rows are indexed from 0: pass the rows of the global batch and slice
    const int64_t n = rows * D;
    AMID_K("k_mask_feature", (cudaStream_t)stream_);
    k_mask_feature<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, (cudaStream_t)stream_>>>(dc, site, n, out);
    AMID_LAUNCH_CHECK("k_mask_feature");
    return 0;
}
extern "C" int amid_dropout_mask_attn(const amid_dropout* drop, uint32_t site, int32_t B, int32_t L, uint8_t* out,
                                      amid_stream_t stream_) {
    const DropCfg dc = with_offsets(make_drop(drop), L);
    const int64_t n = (int64_t)B * H * L * L;
    AMID_K("k_mask_attn", (cudaStream_t)stream_);
    k_mask_attn<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(dc, site, (int64_t)B * H, L, out);
    AMID_LAUNCH_CHECK("k_mask_attn");
    return 0;
}
