// Embedding backward as a deterministic sort-by-index segmented reduction, plus Adam.
// Replaces aten::embedding_dense_backward + torch.optim.Adam on the [V,128] table
// (train_sr.py:213-215 with the dense-gradient table of model_seq.py:25).
//
//   keys = ids (uint32), vals = source row  --radix sort (stable)-->  runs of equal id
//   --run-length encode--> unique ids + counts --scan--> offsets
//   --k_segreduce--> one summed gradient row per unique id, rows added in ascending source
//   order; segments longer than LONG_SEG are summed as 128 fixed chunks so the hot pad row
//   (75% of all positions on the real data) is reduced by many warps yet in a fixed order.
#include <cub/cub.cuh>

#include "common.cuh"

namespace amid {

constexpr int LONG_SEG = 512;
constexpr int LONG_CHUNKS = 128;
constexpr int MAX_LONG = 64;     // long segments reduced by the parallel path (others: same order, one warp)

// ids outside [0, V) (the reference would raise an index error) get the key V: they sort behind every valid row, come out
// of the reduction with the id -1, which the scatter / Adam kernels skip, and raise the device error flag that the host
// polls (amid_gather_error_host_sync)
__global__ void k_make_keys(const int64_t* __restrict__ ids, int64_t n, int64_t V, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals, int* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t id = ids[i];
    const bool ok = id >= 0 && id < V;
    if (!ok) atomicOr(err, 2);
    keys[i] = ok ? (uint32_t)id : (uint32_t)V;
    vals[i] = (uint32_t)i;
}

__device__ __forceinline__ float4 ld_row4(const float* g, uint32_t row, int lane) {
    return __ldg(reinterpret_cast<const float4*>(g + (size_t)row * D) + lane);
}
__device__ __forceinline__ void acc4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// s += rows vals[lo..hi) in index order (4 independent loads in flight)
__device__ __forceinline__ void sum_range_into(float4& s, const float* __restrict__ grads, const uint32_t* __restrict__ vals,
                                               int lo, int hi, int lane) {
    int p = lo;
    for (; p + 4 <= hi; p += 4) {
        const float4 a = ld_row4(grads, vals[p], lane), b = ld_row4(grads, vals[p + 1], lane);
        const float4 c = ld_row4(grads, vals[p + 2], lane), d = ld_row4(grads, vals[p + 3], lane);
        acc4(s, a); acc4(s, b); acc4(s, c); acc4(s, d);
    }
    for (; p < hi; ++p) acc4(s, ld_row4(grads, vals[p], lane));
}
__device__ __forceinline__ float4 sum_range(const float* __restrict__ grads, const uint32_t* __restrict__ vals, int lo,
                                            int hi, int lane) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    sum_range_into(s, grads, vals, lo, hi, lane);
    return s;
}
__device__ __forceinline__ void chunk_bounds(int off, int cnt, int ch, int* lo, int* hi) {
    const int per = (cnt + LONG_CHUNKS - 1) / LONG_CHUNKS;
    *lo = off + min(cnt, ch * per);
    *hi = off + min(cnt, (ch + 1) * per);
}

// One warp per SEG_PER_WARP consecutive unique ids: the segment descriptors and the first source row of every
// segment are fetched together (most segments have one or two rows), the remaining rows are added in order.
constexpr int SEG_PER_WARP = 4;
__global__ void __launch_bounds__(256)
k_segreduce(const float* __restrict__ grads, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ ukeys,
            const int* __restrict__ counts, const int* __restrict__ offsets, const int* __restrict__ n_uniq,
            int64_t* __restrict__ uniq_ids, float* __restrict__ uniq_grads, int* __restrict__ long_list,
            int* __restrict__ n_long, uint32_t V) {
    const int lane = threadIdx.x & 31;
    const int u0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * SEG_PER_WARP;
    const int nu = n_uniq[0];
    if (u0 >= nu) return;
    int off = 0, cnt = 0;
    uint32_t first = 0;
    if (lane < SEG_PER_WARP && u0 + lane < nu) {
        off = offsets[u0 + lane];
        cnt = counts[u0 + lane];
        uniq_ids[u0 + lane] = ukeys[u0 + lane] < V ? (int64_t)ukeys[u0 + lane] : (int64_t)-1;
        first = vals[off];
    }
    float4 s[SEG_PER_WARP];
    int offs[SEG_PER_WARP], cnts[SEG_PER_WARP];
#pragma unroll
    for (int j = 0; j < SEG_PER_WARP; ++j) {
        offs[j] = __shfl_sync(0xffffffffu, off, j);
        cnts[j] = __shfl_sync(0xffffffffu, cnt, j);
        s[j] = ld_row4(grads, __shfl_sync(0xffffffffu, first, j), lane);      // past-the-end slots read row 0, unused
    }
#pragma unroll
    for (int j = 0; j < SEG_PER_WARP; ++j) {
        const int u = u0 + j;
        if (u >= nu) break;
        if (cnts[j] <= LONG_SEG) {
            sum_range_into(s[j], grads, vals, offs[j] + 1, offs[j] + cnts[j], lane);
        } else {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(n_long, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < MAX_LONG) {           // handed to the parallel path (identical summation order)
                if (lane == 0) long_list[slot] = u;
                continue;
            }
            s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ch = 0; ch < LONG_CHUNKS; ++ch) {
                int lo, hi;
                chunk_bounds(offs[j], cnts[j], ch, &lo, &hi);
                acc4(s[j], sum_range(grads, vals, lo, hi, lane));
            }
        }
        reinterpret_cast<float4*>(uniq_grads + (size_t)u * D)[lane] = s[j];
    }
}
// grid (LONG_CHUNKS, MAX_LONG), one warp per chunk of a long segment
__global__ void __launch_bounds__(32)
k_long_partial(const float* __restrict__ grads, const uint32_t* __restrict__ vals, const int* __restrict__ counts,
               const int* __restrict__ offsets, const int* __restrict__ long_list, const int* __restrict__ n_long,
               float* __restrict__ lpart /*[MAX_LONG][LONG_CHUNKS][128]*/) {
    const int li = blockIdx.y, ch = blockIdx.x, lane = threadIdx.x;
    if (li >= min(n_long[0], MAX_LONG)) return;
    const int u = long_list[li];
    int lo, hi;
    chunk_bounds(offsets[u], counts[u], ch, &lo, &hi);
    reinterpret_cast<float4*>(lpart + ((size_t)li * LONG_CHUNKS + ch) * D)[lane] = sum_range(grads, vals, lo, hi, lane);
}
__global__ void __launch_bounds__(32)
k_long_final(const float* __restrict__ lpart, const int* __restrict__ long_list, const int* __restrict__ n_long,
             float* __restrict__ uniq_grads) {
    const int li = blockIdx.x, lane = threadIdx.x;
    if (li >= min(n_long[0], MAX_LONG)) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ch = 0; ch < LONG_CHUNKS; ++ch)
        acc4(s, reinterpret_cast<const float4*>(lpart + ((size_t)li * LONG_CHUNKS + ch) * D)[lane]);
    reinterpret_cast<float4*>(uniq_grads + (size_t)long_list[li] * D)[lane] = s;
}

__global__ void __launch_bounds__(256)
k_scatter_dense(const int64_t* __restrict__ uniq_ids, const float* __restrict__ uniq_grads,
                const int* __restrict__ n_uniq, float* __restrict__ dense, int64_t V) {
    const int lane = threadIdx.x & 31;
    const int u = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (u >= n_uniq[0]) return;
    const int64_t id = uniq_ids[u];
    if (id < 0 || id >= V) return;
    reinterpret_cast<float4*>(dense + id * D)[lane] = reinterpret_cast<const float4*>(uniq_grads + (size_t)u * D)[lane];
}

// dense[ids[u] - id_offset, :] += rows[u, :] for u < n; ids are unique within a call (one warp per row, no atomics), rows whose id
// falls outside [id_offset, id_offset + rows_dense) are skipped.  The owner side of the sparse reduce-scatter of the table
// gradient: one call per sending rank, in rank order, gives a deterministic sum.
__global__ void __launch_bounds__(256)
k_scatter_add_dense(const int64_t* __restrict__ ids, const float* __restrict__ rows, int64_t n, int64_t id_offset,
                    float* __restrict__ dense, int64_t rows_dense) {
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n) return;
    const int64_t r = ids[u] - id_offset;
    if (r < 0 || r >= rows_dense) return;
    float4* dst = reinterpret_cast<float4*>(dense + r * D) + lane;
    const float4 a = *dst, b = reinterpret_cast<const float4*>(rows + (size_t)u * D)[lane];
    *dst = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ------------------------------------------------------------------ Adam
struct AdamStep {
    float w1;        // 1 - beta1   (lerp weight)
    float beta2;
    float omb2;      // 1 - beta2
    float step_size; // lr / (1 - beta1^t)
    float bc2_sqrt;  // sqrt(1 - beta2^t)
    float eps;
};
static AdamStep adam_consts(int step, float lr, float b1, float b2, float eps) {
    AdamStep a;
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    a.w1 = (float)(1.0 - (double)b1);
    a.beta2 = b2;
    a.omb2 = (float)(1.0 - (double)b2);
    a.step_size = (float)((double)lr / bc1);
    a.bc2_sqrt = (float)sqrt(bc2);
    a.eps = eps;
    return a;
}
// one element, same op order as torch.optim.Adam (_single_tensor_adam / _multi_tensor_adam)
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamStep& a) {
    m = m + a.w1 * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * a.beta2;                              // exp_avg_sq.mul_(beta2)
    v = v + (a.omb2 * g) * g;                     //   .addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p + (-a.step_size) * (m / denom);         // param.addcdiv_(exp_avg, denom, value = -step_size)
}
__global__ void __launch_bounds__(256)
k_adam_dense(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, AdamStep a) {
    const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = i4 * 4;
    if (i + 4 <= n) {
        float4 P = reinterpret_cast<float4*>(p)[i4], G = reinterpret_cast<const float4*>(g)[i4];
        float4 M = reinterpret_cast<float4*>(m)[i4], Vv = reinterpret_cast<float4*>(v)[i4];
        adam1(P.x, G.x, M.x, Vv.x, a); adam1(P.y, G.y, M.y, Vv.y, a);
        adam1(P.z, G.z, M.z, Vv.z, a); adam1(P.w, G.w, M.w, Vv.w, a);
        reinterpret_cast<float4*>(p)[i4] = P;
        reinterpret_cast<float4*>(m)[i4] = M;
        reinterpret_cast<float4*>(v)[i4] = Vv;
    } else {
        for (int64_t j = i; j < n; ++j) adam1(p[j], g[j], m[j], v[j], a);
    }
}

// Host-computed constants of the most recent steps, so that the replay of a short gap needs no FP64 on the device.
constexpr int ADAM_HIST = 48;
struct AdamHist {
    int first_step;                 // hist[i] belongs to step first_step + i
    float step_size[ADAM_HIST];
    float bc2_sqrt[ADAM_HIST];
};
static AdamHist adam_hist(int step, float lr, float b1, float b2) {
    AdamHist h;
    h.first_step = step - ADAM_HIST + 1;
    for (int i = 0; i < ADAM_HIST; ++i) {
        const int s = h.first_step + i;
        if (s < 1) { h.step_size[i] = 0.f; h.bc2_sqrt[i] = 1.f; continue; }
        h.step_size[i] = (float)((double)lr / (1.0 - pow((double)b1, (double)s)));
        h.bc2_sqrt[i] = (float)sqrt(1.0 - pow((double)b2, (double)s));
    }
    return h;
}
// replay of zero-gradient steps s0..s1 (inclusive) for one float4 of a row
__device__ __forceinline__ void adam_replay(float4& P, float4& M, float4& Vv, int s0, int s1, float lr, double b1,
                                            double b2, float eps, const AdamHist& hist) {
    if (s0 > s1) return;
    if (s0 >= hist.first_step && s0 >= 1) {            // common case: the gap is covered by the host table
        AdamStep a;
        a.w1 = (float)(1.0 - b1);
        a.beta2 = (float)b2;
        a.omb2 = (float)(1.0 - b2);
        a.eps = eps;
        for (int s = s0; s <= s1; ++s) {
            a.step_size = hist.step_size[s - hist.first_step];
            a.bc2_sqrt = hist.bc2_sqrt[s - hist.first_step];
            adam1(P.x, 0.f, M.x, Vv.x, a); adam1(P.y, 0.f, M.y, Vv.y, a);
            adam1(P.z, 0.f, M.z, Vv.z, a); adam1(P.w, 0.f, M.w, Vv.w, a);
        }
        return;
    }
    double p1 = pow(b1, (double)(s0 - 1)), p2 = pow(b2, (double)(s0 - 1));
    AdamStep a;
    a.w1 = (float)(1.0 - b1);
    a.beta2 = (float)b2;
    a.omb2 = (float)(1.0 - b2);
    a.eps = eps;
    for (int s = s0; s <= s1; ++s) {
        p1 *= b1;
        p2 *= b2;
        a.step_size = (float)((double)lr / (1.0 - p1));
        a.bc2_sqrt = (float)sqrt(1.0 - p2);
        adam1(P.x, 0.f, M.x, Vv.x, a); adam1(P.y, 0.f, M.y, Vv.y, a);
        adam1(P.z, 0.f, M.z, Vv.z, a); adam1(P.w, 0.f, M.w, Vv.w, a);
    }
}
__global__ void __launch_bounds__(256)
k_adam_rows_lazy(float* __restrict__ table, float* __restrict__ m, float* __restrict__ v, int* __restrict__ last_step,
                 const int64_t* __restrict__ uniq_ids, const float* __restrict__ uniq_grads,
                 const int* __restrict__ n_uniq, int step, float lr, float b1, float b2, float eps, AdamStep a,
                 const __grid_constant__ AdamHist hist) {
    // one warp per unique row (the kernel is bound by the IEEE divide / sqrt of the replayed steps, not by its
    // four 512 B row reads, so more rows per warp only costs occupancy)
    const int lane = threadIdx.x & 31;
    const int u = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (u >= n_uniq[0]) return;
    const int64_t id = uniq_ids[u];
    if (id < 0) return;                              // out-of-range ids of the step (flagged by k_make_keys)
    float4* Pp = reinterpret_cast<float4*>(table + id * D) + lane;
    float4* Mp = reinterpret_cast<float4*>(m + id * D) + lane;
    float4* Vp = reinterpret_cast<float4*>(v + id * D) + lane;
    float4 P = *Pp, M = *Mp, Vv = *Vp;
    const float4 G = reinterpret_cast<const float4*>(uniq_grads + (size_t)u * D)[lane];
    const int last = last_step[id];
    if (last > 0) adam_replay(P, M, Vv, last + 1, step - 1, lr, (double)b1, (double)b2, eps, hist);
    adam1(P.x, G.x, M.x, Vv.x, a); adam1(P.y, G.y, M.y, Vv.y, a);
    adam1(P.z, G.z, M.z, Vv.z, a); adam1(P.w, G.w, M.w, Vv.w, a);
    *Pp = P; *Mp = M; *Vp = Vv;
    __syncwarp();
    if (lane == 0) last_step[id] = step;
}
__global__ void __launch_bounds__(256)
k_adam_rows_flush(float* __restrict__ table, float* __restrict__ m, float* __restrict__ v, int* __restrict__ last_step,
                  int64_t V, int step, float lr, float b1, float b2, float eps, const __grid_constant__ AdamHist hist) {
    const int lane = threadIdx.x & 31;
    const int64_t id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (id >= V) return;
    const int last = last_step[id];
    if (last <= 0 || last >= step) return;
    float4* Pp = reinterpret_cast<float4*>(table + id * D) + lane;
    float4* Mp = reinterpret_cast<float4*>(m + id * D) + lane;
    float4* Vp = reinterpret_cast<float4*>(v + id * D) + lane;
    float4 P = *Pp, M = *Mp, Vv = *Vp;
    adam_replay(P, M, Vv, last + 1, step, lr, (double)b1, (double)b2, eps, hist);
    *Pp = P; *Mp = M; *Vp = Vv;
    __syncwarp();
    if (lane == 0) last_step[id] = step;
}

// workspace carve-up for the segmented reduction
struct SegWs {
    uint32_t *keys_in, *vals_in, *keys_out, *vals_out, *ukeys;
    int *counts, *offsets, *long_list, *n_long;
    float* lpart;
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
};
static size_t cub_temp_bytes(int64_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 32);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (int*)nullptr, (int*)nullptr, (int)n);
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}
static SegWs carve(void* base, int64_t n) {
    SegWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? (char*)base + off : nullptr;
        off += (size_t)round_up((int64_t)bytes, 256);
        return p;
    };
    w.keys_in = (uint32_t*)take(n * 4);
    w.vals_in = (uint32_t*)take(n * 4);
    w.keys_out = (uint32_t*)take(n * 4);
    w.vals_out = (uint32_t*)take(n * 4);
    w.ukeys = (uint32_t*)take(n * 4);
    w.counts = (int*)take(n * 4);
    w.offsets = (int*)take(n * 4);
    w.long_list = (int*)take(MAX_LONG * 4);
    w.n_long = (int*)take(4);
    w.lpart = (float*)take((size_t)MAX_LONG * LONG_CHUNKS * D * 4);
    w.cub_bytes = cub_temp_bytes(n);
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

}  // namespace amid

using namespace amid;

extern "C" int64_t amid_embgrad_workspace_bytes(int64_t n_rows) {
    if (n_rows <= 0) return 256;
    return (int64_t)carve(nullptr, n_rows).total;
}

extern "C" int amid_embgrad_segreduce(const int64_t* ids, const float* grad_rows, int64_t n, int64_t V,
                                      int64_t* uniq_ids, float* uniq_grads, int32_t* n_uniq, void* workspace,
                                      int64_t workspace_bytes, amid_stream_t s_) {
    cudaStream_t s = (cudaStream_t)s_;
    AMID_REQUIRE(ids && grad_rows && uniq_ids && uniq_grads && n_uniq && workspace, "embgrad_segreduce: null argument");
    AMID_REQUIRE(n > 0 && n < (1ll << 31), "embgrad_segreduce: n_rows=%lld", (long long)n);
    AMID_REQUIRE(V > 0 && V < 0xFFFFFFFFll, "embgrad_segreduce: V=%lld does not fit 32-bit keys", (long long)V);
    int* err = err_flag();
    AMID_REQUIRE(err, "embgrad_segreduce: cannot allocate error flag");
    AMID_REQUIRE(aligned16(grad_rows) && aligned16(uniq_grads) && ((uintptr_t)workspace & 255) == 0, "embgrad_segreduce: misaligned buffer");
    SegWs w = carve(workspace, n);
    AMID_REQUIRE((int64_t)w.total <= workspace_bytes, "embgrad_segreduce: workspace too small (%lld < %zu)",
                 (long long)workspace_bytes, w.total);
    AMID_K("k_make_keys", s);
    k_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, n, V, w.keys_in, w.vals_in, err);
    AMID_LAUNCH_CHECK("k_make_keys");
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (unsigned long long)V + 1) ++end_bit;     // keys 0..V (V = invalid id)
    size_t tb = w.cub_bytes;
    AMID_K("cub_radix_sort_pairs", s);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)n,
                                                    0, end_bit, s);
    ::amid::prof_end();
    if (e != cudaSuccess) return set_error(-2, "embgrad: radix sort: %s", cudaGetErrorString(e));
    // the scan below runs over all n slots while the encoder writes only the first n_uniq counts: define the rest
    e = cudaMemsetAsync(w.counts, 0, (size_t)n * sizeof(int), s);
    if (e != cudaSuccess) return set_error(-2, "embgrad: memset: %s", cudaGetErrorString(e));
    AMID_K("cub_rle_scan", s);
    tb = w.cub_bytes;
    e = cub::DeviceRunLengthEncode::Encode(w.cub_tmp, tb, w.keys_out, w.ukeys, w.counts, n_uniq, (int)n, s);
    if (e != cudaSuccess) return set_error(-2, "embgrad: run-length encode: %s", cudaGetErrorString(e));
    tb = w.cub_bytes;
    e = cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.counts, w.offsets, (int)n, s);
    ::amid::prof_end();
    if (e != cudaSuccess) return set_error(-2, "embgrad: scan: %s", cudaGetErrorString(e));
    e = cudaMemsetAsync(w.n_long, 0, 4, s);
    if (e != cudaSuccess) return set_error(-2, "embgrad: memset: %s", cudaGetErrorString(e));
    const unsigned blocks = (unsigned)(((n + SEG_PER_WARP - 1) / SEG_PER_WARP * 32 + 255) / 256);
    AMID_K("k_segreduce", s);
    k_segreduce<<<blocks, 256, 0, s>>>(grad_rows, w.vals_out, w.ukeys, w.counts, w.offsets, n_uniq, uniq_ids, uniq_grads,
                                       w.long_list, w.n_long, (uint32_t)V);
    AMID_LAUNCH_CHECK("k_segreduce");
    AMID_K("k_long_partial", s);
    k_long_partial<<<dim3(LONG_CHUNKS, MAX_LONG), 32, 0, s>>>(grad_rows, w.vals_out, w.counts, w.offsets, w.long_list,
                                                               w.n_long, w.lpart);
    AMID_LAUNCH_CHECK("k_long_partial");
    AMID_K("k_long_final", s);
    k_long_final<<<MAX_LONG, 32, 0, s>>>(w.lpart, w.long_list, w.n_long, uniq_grads);
    AMID_LAUNCH_CHECK("k_long_final");
    return 0;
}

extern "C" int amid_embgrad_scatter_dense(const int64_t* uniq_ids, const float* uniq_grads, const int32_t* n_uniq,
                                          int64_t max_rows, float* dense, int64_t V, amid_stream_t s_) {
    AMID_REQUIRE(uniq_ids && uniq_grads && n_uniq && dense && max_rows > 0 && V > 0, "embgrad_scatter_dense: bad argument");
    AMID_K("k_scatter_dense", (cudaStream_t)s_);
    k_scatter_dense<<<(unsigned)((max_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(uniq_ids, uniq_grads, n_uniq,
                                                                                            dense, V);
    AMID_LAUNCH_CHECK("k_scatter_dense");
    return 0;
}

extern "C" int amid_embgrad_scatter_add(const int64_t* ids, const float* rows, int64_t n, int64_t id_offset, float* dense,
                                       int64_t rows_dense, amid_stream_t s_) {
    AMID_REQUIRE(dense && rows_dense > 0 && n >= 0 && (n == 0 || (ids && rows)), "embgrad_scatter_add: bad argument");
    if (n == 0) return 0;
    AMID_K("k_scatter_add_dense", (cudaStream_t)s_);
    k_scatter_add_dense<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(ids, rows, n, id_offset, dense, rows_dense);
    AMID_LAUNCH_CHECK("k_scatter_add_dense");
    return 0;
}

extern "C" int amid_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                               float beta1, float beta2, float eps, amid_stream_t s_) {
    AMID_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam_dense: bad argument");
    AMID_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adam_dense: misaligned buffer");
    const AdamStep a = adam_consts(step, lr, beta1, beta2, eps);
    const int64_t n4 = (n + 3) / 4;
    AMID_K("k_adam_dense", (cudaStream_t)s_);
    k_adam_dense<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(p, g, m, v, n, a);
    AMID_LAUNCH_CHECK("k_adam_dense");
    return 0;
}

extern "C" int amid_adam_rows_lazy(float* table, float* m, float* v, int32_t* last_step, const int64_t* uniq_ids,
                                   const float* uniq_grads, const int32_t* n_uniq, int64_t max_rows, int32_t step,
                                   float lr, float beta1, float beta2, float eps, amid_stream_t s_) {
    AMID_REQUIRE(table && m && v && last_step && uniq_ids && uniq_grads && n_uniq && max_rows > 0 && step >= 1,
                 "adam_rows_lazy: bad argument");
    const AdamStep a = adam_consts(step, lr, beta1, beta2, eps);
    AMID_K("k_adam_rows_lazy", (cudaStream_t)s_);
    k_adam_rows_lazy<<<(unsigned)((max_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(
        table, m, v, last_step, uniq_ids, uniq_grads, n_uniq, step, lr, beta1, beta2, eps, a, adam_hist(step, lr, beta1, beta2));
    AMID_LAUNCH_CHECK("k_adam_rows_lazy");
    return 0;
}

extern "C" int amid_adam_rows_flush(float* table, float* m, float* v, int32_t* last_step, int64_t V, int32_t step,
                                    float lr, float beta1, float beta2, float eps, amid_stream_t s_) {
    AMID_REQUIRE(table && m && v && last_step && V > 0 && step >= 0, "adam_rows_flush: bad argument");
    if (step == 0) return 0;
    AMID_K("k_adam_rows_flush", (cudaStream_t)s_);
    k_adam_rows_flush<<<(unsigned)((V * 32 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(table, m, v, last_step, V, step,
                                                                                       lr, beta1, beta2, eps,
                                                                                       adam_hist(step, lr, beta1, beta2));
    AMID_LAUNCH_CHECK("k_adam_rows_flush");
    return 0;
}

// ------------------------------------------------------------------ row-sharded table: the lookup plan of one step
// (BASELINE config 4; amid_b200/sharded.py).  owner = id mod G, owner-local row = id div G.  One sort of the keys
// owner * Vs + local puts the step's ids in bucket order (by owner, ascending id inside); run-length encoding yields the
// unique rows, an exclusive scan their first positions, and two small kernels emit, entirely on the device,
//   uniq_local[u]   owner-local row index of unique row u (bucket order = the order of the step table)
//   virtual_ids[p]  row of the step table for every requested position p
//   send_counts[o]  unique rows requested from owner o
//   flags           {number of unique rows (a trailing group of out-of-range ids included), out-of-range seen}
namespace amid {
__global__ void k_plan_keys(const int64_t* __restrict__ ids, int64_t n, int64_t V, uint32_t G, uint32_t Vs,
                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t id = ids[i];
    const bool ok = id >= 0 && id < V;
    if (!ok) atomicOr(flags + 1, 1);
    keys[i] = ok ? (uint32_t)(id % G) * Vs + (uint32_t)(id / G) : G * Vs;
    vals[i] = (uint32_t)i;
}
__global__ void k_plan_emit(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ ukeys, const int* __restrict__ offsets,
                            const int* __restrict__ n_uniq, int64_t n, uint32_t Vs, int64_t* __restrict__ uniq_local,
                            int64_t* __restrict__ virtual_ids, int* __restrict__ flags) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int nu = n_uniq[0];
    int lo = 0, hi = nu;                       // last u with offsets[u] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid;
    }
    virtual_ids[vals[p]] = lo;
    if (offsets[lo] == p) uniq_local[lo] = (int64_t)(ukeys[lo] % Vs);
    if (p == 0) flags[0] = nu;
}
__global__ void k_plan_counts(const uint32_t* __restrict__ ukeys, const int* __restrict__ n_uniq, uint32_t G, uint32_t Vs,
                              int64_t* __restrict__ send_counts) {
    const uint32_t o = threadIdx.x;
    if (o >= G) return;
    const int nu = n_uniq[0];
    auto lower = [&](uint32_t key) {
        int lo = 0, hi = nu;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    send_counts[o] = (int64_t)(lower((o + 1) * Vs) - lower(o * Vs));
}
}  // namespace amid

extern "C" int64_t amid_shard_plan_workspace_bytes(int64_t n) {
    if (n <= 0) return -1;
    return (int64_t)carve(nullptr, n).total;
}
extern "C" int amid_shard_plan(const int64_t* ids, int64_t n, int64_t V, int32_t G, int64_t* uniq_local, int64_t* virtual_ids,
                               int32_t* flags, int64_t* send_counts, void* workspace, int64_t workspace_bytes, amid_stream_t s_) {
    cudaStream_t s = (cudaStream_t)s_;
    AMID_REQUIRE(ids && uniq_local && virtual_ids && flags && send_counts && workspace, "shard_plan: null argument");
    AMID_REQUIRE(n > 0 && n < (1ll << 31) && V > 0 && G >= 1 && G <= 1024, "shard_plan: bad sizes");
    const int64_t Vs = (V + G - 1) / G;
    AMID_REQUIRE(Vs * G < 0xFFFFFFFFll, "shard_plan: V=%lld does not fit 32-bit keys", (long long)V);
    AMID_REQUIRE(((uintptr_t)workspace & 255) == 0, "shard_plan: misaligned workspace");
    SegWs w = carve(workspace, n);
    AMID_REQUIRE((int64_t)w.total <= workspace_bytes, "shard_plan: workspace too small");
    cudaError_t e = cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), s);
    if (e != cudaSuccess) return set_error(-2, "shard_plan: memset: %s", cudaGetErrorString(e));
    AMID_K("k_plan_keys", s);
    k_plan_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, n, V, (uint32_t)G, (uint32_t)Vs, w.keys_in, w.vals_in, flags);
    AMID_LAUNCH_CHECK("k_plan_keys");
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (unsigned long long)(Vs * G) + 1) ++end_bit;
    size_t tb = w.cub_bytes;
    AMID_K("cub_radix_sort_pairs", s);
    e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)n, 0, end_bit, s);
    ::amid::prof_end();
    if (e != cudaSuccess) return set_error(-2, "shard_plan: radix sort: %s", cudaGetErrorString(e));
    int* n_uniq = w.n_long;                   // scratch int of the carve-up
    e = cudaMemsetAsync(w.counts, 0, (size_t)n * sizeof(int), s);
    if (e != cudaSuccess) return set_error(-2, "shard_plan: memset: %s", cudaGetErrorString(e));
    AMID_K("cub_rle_scan", s);
    tb = w.cub_bytes;
    e = cub::DeviceRunLengthEncode::Encode(w.cub_tmp, tb, w.keys_out, w.ukeys, w.counts, n_uniq, (int)n, s);
    if (e != cudaSuccess) return set_error(-2, "shard_plan: run-length encode: %s", cudaGetErrorString(e));
    tb = w.cub_bytes;
    e = cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.counts, w.offsets, (int)n, s);
    ::amid::prof_end();
    if (e != cudaSuccess) return set_error(-2, "shard_plan: scan: %s", cudaGetErrorString(e));
    AMID_K("k_plan_emit", s);
    k_plan_emit<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.vals_out, w.ukeys, w.offsets, n_uniq, n, (uint32_t)Vs, uniq_local,
                                                            virtual_ids, flags);
    AMID_LAUNCH_CHECK("k_plan_emit");
    AMID_K("k_plan_counts", s);
    k_plan_counts<<<1, 1024, 0, s>>>(w.ukeys, n_uniq, (uint32_t)G, (uint32_t)Vs, send_counts);
    AMID_LAUNCH_CHECK("k_plan_counts");
    return 0;
}
