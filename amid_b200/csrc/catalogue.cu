// Full-catalogue evaluation (BASELINE config 5; SURVEY.md 8d/8e "C5"): every user of an eval batch against every
// item of the target-domain pool, without materialising the user x item score matrix.
//
// predictModule (model_seq.py:40-54) couples user and item inside the ReLU, so only the two halves
//   A[u]  = W0[:, :128] u            (once per user and domain)
//   Bc[i] = W0[:, 128:] item_i + b0  (once per catalogue item, independent of the users: cached per evaluation)
// are contractions; the U x I stage is 32-wide add / max / multiply per pair on the fp32 ALUs, followed by the
// sigmoid -- ranking has to use the fp32 post-sigmoid value because ties from saturation are part of the
// reference result (utils.py:296-301).  What leaves the kernel per user is the number of pool items that score
// higher than / equal to the positive (with and without the 1e-7 fix of train_sr.py:114): HR@k, NDCG@k and MRR
// are functions of that rank alone.
//
// Every score is computed in exactly the op order of k_score_fwd (score.cu) -- sequential fma chains for the two
// halves, the xor-butterfly order of its warp reduction for the 32 hidden units -- so a pair scored here is
// bit-identical to the same pair scored by the sampled-candidate path (tests/test_gpu_catalogue.py).
#include "common.cuh"

namespace amid {

constexpr int CAT_HID = 32;       // hidden width the butterfly order is written for (run.sh: hid_dim 32)
constexpr int CAT_TU = 128;       // users per CTA, one per thread
constexpr int CAT_TI = 64;        // catalogue items per shared-memory tile
constexpr int PROJ_ITEMS = 64;    // items per CTA in the projection kernel

// Bc[i][hh] = b0[hh] + sum_k item_i[k] * W0[hh][128 + k]; item_i = table[ids[i]] (ids may be null: rows 0..I-1)
__global__ void __launch_bounds__(128)
k_item_proj(const float* __restrict__ table, const int64_t* __restrict__ ids, int64_t I, int64_t V,
            const float* __restrict__ w0, const float* __restrict__ b0, float* __restrict__ Bc, int* __restrict__ err) {
    __shared__ float WiT[D * CAT_HID];            // [k][hh]
    __shared__ __align__(16) float IT[4][D];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int idx = t; idx < CAT_HID * D; idx += 128) {
        const int hh = idx / D, k = idx % D;
        WiT[k * CAT_HID + hh] = __ldg(w0 + (size_t)hh * 2 * D + D + k);
    }
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * PROJ_ITEMS;
    float* it = IT[warp];
    for (int j = warp; j < PROJ_ITEMS; j += 4) {
        const int64_t i = i0 + j;
        if (i >= I) break;                         // warp-uniform
        int64_t row = ids ? __ldg(ids + i) : i;
        if (row < 0 || row >= V) { if (lane == 0) atomicExch(err, 1); row = 0; }
        const float4 v = __ldg(reinterpret_cast<const float4*>(table + row * D) + lane);
        __syncwarp();
        *reinterpret_cast<float4*>(it + lane * 4) = v;
        __syncwarp();
        float bc = __ldg(b0 + lane);
        const float* w = WiT + lane;
#pragma unroll 8
        for (int k = 0; k < D; ++k) bc = fmaf(it[k], w[k * CAT_HID], bc);
        Bc[i * CAT_HID + lane] = bc;
    }
}

// A[b][dom][hh] = <W0[hh][0:128], u_dom[b]>
__global__ void __launch_bounds__(64)
k_user_proj(const float* __restrict__ u1, const float* __restrict__ u2, int B, const float* __restrict__ w0,
            float* __restrict__ A) {
    __shared__ __align__(16) float U[2 * D];
    const int b = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < 2 * D; i += 64) U[i] = i < D ? u1[(size_t)b * D + i] : u2[(size_t)b * D + i - D];
    __syncthreads();
    const int dom = t >> 5, hh = t & 31;
    const float4* w = reinterpret_cast<const float4*>(w0 + (size_t)hh * 2 * D);
    const float4* uu = reinterpret_cast<const float4*>(U + dom * D);
    float s = 0.f;
    for (int k4 = 0; k4 < D / 4; ++k4) {
        const float4 a = __ldg(w + k4), x = uu[k4];
        s = fmaf(a.x, x.x, s); s = fmaf(a.y, x.y, s); s = fmaf(a.z, x.z, s); s = fmaf(a.w, x.w, s);
    }
    A[((size_t)b * 2 + dom) * CAT_HID + hh] = s;
}

// one (user, item) score from the user half in registers and the item half behind a float4 pointer
__device__ __forceinline__ float pair_score(const float (&a)[CAT_HID], const float (&w)[CAT_HID], const float4* __restrict__ bc4, float b2) {
    float t[CAT_HID];
#pragma unroll
    for (int q = 0; q < CAT_HID / 4; ++q) {
        const float4 b = bc4[q];
        t[4 * q + 0] = __fmul_rn(w[4 * q + 0], fmaxf(__fadd_rn(a[4 * q + 0], b.x), 0.f));
        t[4 * q + 1] = __fmul_rn(w[4 * q + 1], fmaxf(__fadd_rn(a[4 * q + 1], b.y), 0.f));
        t[4 * q + 2] = __fmul_rn(w[4 * q + 2], fmaxf(__fadd_rn(a[4 * q + 2], b.z), 0.f));
        t[4 * q + 3] = __fmul_rn(w[4 * q + 3], fmaxf(__fadd_rn(a[4 * q + 3], b.w), 0.f));
    }
    // the xor-butterfly of warp_sum (offsets 16, 8, 4, 2, 1) as seen from lane 0
#pragma unroll
    for (int o = CAT_HID / 2; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < o; ++h) t[h] = __fadd_rn(t[h], t[h + o]);
    return 1.0f / (1.0f + expf(-(t[0] + b2)));
}

// counts[u] = {#gt, #eq (threshold s_pos), #gt, #eq (threshold s_pos - fix)} over pool items [i_lo, i_hi) except
// the positive itself; grid (user tiles, item splits), the splits add into counts with integer atomics.
// WRITE: store the scores instead ([nu][i_hi - i_lo]; the positive's column keeps its score).
template <bool WRITE>
__global__ void __launch_bounds__(CAT_TU)
k_rank_full(const float* __restrict__ A, const int* __restrict__ user_rows, int nu, int dom, const float* __restrict__ Bc,
            int i_lo, int i_hi, const int* __restrict__ pos_idx, const float* __restrict__ w2, const float* __restrict__ b2p,
            float fix, int* __restrict__ counts, float* __restrict__ s_pos_out, float* __restrict__ scores) {
    __shared__ __align__(16) float tile[CAT_TI][CAT_HID];
    const int t = threadIdx.x;
    const int slot = blockIdx.x * CAT_TU + t;
    const bool valid = slot < nu;
    const int ur = user_rows[valid ? slot : nu - 1];          // row of this user in the eval batch
    float a[CAT_HID], w[CAT_HID];
    {
        const float4* ap = reinterpret_cast<const float4*>(A + ((size_t)ur * 2 + dom) * CAT_HID);
        const float4* wp = reinterpret_cast<const float4*>(w2);
#pragma unroll
        for (int q = 0; q < CAT_HID / 4; ++q) {
            const float4 x = __ldg(ap + q), y = __ldg(wp + q);
            a[4 * q] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
            w[4 * q] = y.x; w[4 * q + 1] = y.y; w[4 * q + 2] = y.z; w[4 * q + 3] = y.w;
        }
    }
    const float b2 = __ldg(b2p);
    const int pos = pos_idx[ur];
    const float s_pos = pair_score(a, w, reinterpret_cast<const float4*>(Bc + (size_t)pos * CAT_HID), b2);
    const float th0 = s_pos, th1 = s_pos - fix;               // fp32 subtraction, as numpy does (train_sr.py:114)
    const int n_items = i_hi - i_lo;
    const int per = ((n_items + (int)gridDim.y - 1) / (int)gridDim.y + CAT_TI - 1) / CAT_TI * CAT_TI;
    const int c0 = i_lo + (int)blockIdx.y * per, c1 = min(i_hi, c0 + per);
    int g0 = 0, e0 = 0, g1 = 0, e1 = 0;
    for (int base = c0; base < c1; base += CAT_TI) {
        const int nt = min(CAT_TI, c1 - base);
        __syncthreads();
        for (int idx = t; idx < nt * (CAT_HID / 4); idx += CAT_TU)
            reinterpret_cast<float4*>(&tile[0][0])[idx] = __ldg(reinterpret_cast<const float4*>(Bc + (size_t)base * CAT_HID) + idx);
        __syncthreads();
        for (int j = 0; j < nt; ++j) {
            const float p = pair_score(a, w, reinterpret_cast<const float4*>(tile[j]), b2);
            const int i = base + j;
            if (WRITE) {
                if (valid) scores[(size_t)slot * n_items + (i - i_lo)] = p;
            } else if (i != pos) {
                g0 += p > th0; e0 += p == th0;
                g1 += p > th1; e1 += p == th1;
            }
        }
    }
    if (!valid) return;
    if (blockIdx.y == 0) s_pos_out[slot] = s_pos;
    if (!WRITE) {
        atomicAdd(counts + 4 * slot + 0, g0);
        atomicAdd(counts + 4 * slot + 1, e0);
        atomicAdd(counts + 4 * slot + 2, g1);
        atomicAdd(counts + 4 * slot + 3, e1);
    }
}

}  // namespace amid

using namespace amid;

extern "C" int amid_catalogue_item_proj(const float* table, int64_t V, const int64_t* ids, int64_t n_items, const float* w0,
                                        const float* b0, int32_t hid, float* Bc, amid_stream_t s_) {
    AMID_REQUIRE(table && w0 && b0 && Bc && V > 0 && n_items >= 0, "catalogue_item_proj: bad argument");
    AMID_REQUIRE(hid == CAT_HID, "catalogue_item_proj: hid=%d, the full-catalogue path is written for hid=32", hid);
    AMID_REQUIRE(aligned16(table) && aligned16(w0), "catalogue_item_proj: misaligned buffer");
    AMID_REQUIRE(ids || n_items <= V, "catalogue_item_proj: n_items > V without an id list");
    if (n_items == 0) return 0;
    int* err = err_flag();
    AMID_REQUIRE(err, "catalogue_item_proj: cannot allocate error flag");
    AMID_K("k_item_proj", (cudaStream_t)s_);
    k_item_proj<<<(unsigned)((n_items + PROJ_ITEMS - 1) / PROJ_ITEMS), 128, 0, (cudaStream_t)s_>>>(table, ids, n_items, V, w0, b0, Bc, err);
    AMID_LAUNCH_CHECK("k_item_proj");
    return 0;
}

extern "C" int amid_catalogue_user_proj(const float* u1, const float* u2, int32_t B, const float* w0, int32_t hid, float* A,
                                        amid_stream_t s_) {
    AMID_REQUIRE(u1 && u2 && w0 && A && B > 0, "catalogue_user_proj: bad argument");
    AMID_REQUIRE(hid == CAT_HID, "catalogue_user_proj: hid=%d, the full-catalogue path is written for hid=32", hid);
    AMID_REQUIRE(aligned16(w0), "catalogue_user_proj: misaligned buffer");
    AMID_K("k_user_proj", (cudaStream_t)s_);
    k_user_proj<<<B, 64, 0, (cudaStream_t)s_>>>(u1, u2, B, w0, A);
    AMID_LAUNCH_CHECK("k_user_proj");
    return 0;
}

// enough item splits to fill the GPU (148 SMs x 4 resident CTAs) without making the per-CTA chunks tiny
static int item_splits(int tiles, int n_items) {
    int splits = (592 + tiles - 1) / tiles;
    const int max_splits = (n_items + 4 * CAT_TI - 1) / (4 * CAT_TI);
    if (splits > max_splits) splits = max_splits;
    return splits < 1 ? 1 : splits;
}

static int check_rank_full(const float* A, const int32_t* user_rows, int32_t n_users, int32_t dom, const float* Bc,
                           int32_t i_lo, int32_t i_hi, const int32_t* pos_idx, const float* w2, const float* b2) {
    AMID_REQUIRE(A && user_rows && Bc && pos_idx && w2 && b2, "catalogue_rank: null argument");
    AMID_REQUIRE(n_users >= 0 && (dom == 0 || dom == 1) && 0 <= i_lo && i_lo < i_hi, "catalogue_rank: bad sizes");
    AMID_REQUIRE(aligned16(A) && aligned16(Bc) && aligned16(w2), "catalogue_rank: misaligned buffer");
    return 0;
}

extern "C" int amid_catalogue_rank(const float* A, const int32_t* user_rows, int32_t n_users, int32_t dom, const float* Bc,
                                   int32_t i_lo, int32_t i_hi, const int32_t* pos_idx, const float* w2, const float* b2,
                                   float fix, int32_t* counts, float* s_pos, amid_stream_t s_) {
    if (int rc = check_rank_full(A, user_rows, n_users, dom, Bc, i_lo, i_hi, pos_idx, w2, b2)) return rc;
    AMID_REQUIRE(counts && s_pos, "catalogue_rank: null output");
    if (n_users == 0) return 0;
    cudaStream_t s = (cudaStream_t)s_;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)n_users * 4 * sizeof(int32_t), s);
    if (e != cudaSuccess) return set_error(-2, "catalogue_rank: memset: %s", cudaGetErrorString(e));
    const int tiles = (n_users + CAT_TU - 1) / CAT_TU;
    const int splits = item_splits(tiles, i_hi - i_lo);
    AMID_K("k_rank_full", s);
    k_rank_full<false><<<dim3(tiles, splits), CAT_TU, 0, s>>>(A, user_rows, n_users, dom, Bc, i_lo, i_hi, pos_idx, w2, b2, fix,
                                                               counts, s_pos, nullptr);
    AMID_LAUNCH_CHECK("k_rank_full");
    return 0;
}

extern "C" int amid_catalogue_scores(const float* A, const int32_t* user_rows, int32_t n_users, int32_t dom, const float* Bc,
                                     int32_t i_lo, int32_t i_hi, const int32_t* pos_idx, const float* w2, const float* b2,
                                     float* s_pos, float* scores, amid_stream_t s_) {
    if (int rc = check_rank_full(A, user_rows, n_users, dom, Bc, i_lo, i_hi, pos_idx, w2, b2)) return rc;
    AMID_REQUIRE(scores && s_pos, "catalogue_scores: null output");
    if (n_users == 0) return 0;
    const int tiles = (n_users + CAT_TU - 1) / CAT_TU;
    AMID_K("k_rank_full_scores", (cudaStream_t)s_);
    k_rank_full<true><<<dim3(tiles, item_splits(tiles, i_hi - i_lo)), CAT_TU, 0, (cudaStream_t)s_>>>(A, user_rows, n_users, dom, Bc, i_lo, i_hi, pos_idx, w2, b2,
                                                                        0.f, nullptr, s_pos, scores);
    AMID_LAUNCH_CHECK("k_rank_full_scores");
    return 0;
}
