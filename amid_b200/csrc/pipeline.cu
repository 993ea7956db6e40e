// Device-side batch construction: the per-row work of DualDomainSeqDataset.__getitem__
// (dataset_seq.py:177-236) and collate_fn_enhance (:252-274) for a whole batch in one launch, from histories that
// were tokenised once into CSR arrays in HBM.  Ids stay int64 end to end (the reference's collate routes them
// through float32, exact only below 2^24).
//
//   seq_dk[b, :]  = last L items of the processed history of row rows[b], left-padded with pad_id   (:12-22, 223-224)
//   i_node[b], domain_id[b], overlap_label[b], user_node[b], long_tail_mask_dk[b]                    (:181-190, 226-233)
//   neg_samples[b, 0:K] (optional) = K distinct items of the target domain's pool that are not in the user's full
//                  own-domain sequence (:188, 197-201): a random start + random stride coprime to the pool size walk
//                  the sorted pool (every pool position is visited exactly once), positions whose item is in the
//                  row's exclusion list are skipped.  One warp per row; lanes test 32 candidates per round.
// The "replay" mode of the Python layer skips the sampler and copies negatives drawn by the reference sampler.
#include "common.cuh"

namespace amid {

struct BatchSrc {
    const int64_t *h1_vals, *h2_vals, *ex_vals;      // processed histories (target removed) and exclusion lists
    const int64_t *h1_offs, *h2_offs, *ex_offs;      // [N+1]
    const int64_t *target, *user;                    // [N]
    const int32_t *domain, *overlap;                 // [N]
    const int64_t *pool1, *pool2;                    // sorted item ids of each domain
    int64_t n_pool1, n_pool2;
};
struct BatchDst {
    int64_t *seq_d1, *seq_d2;                        // [B, L]
    int64_t *i_node, *user_node, *domain_id, *overlap_label, *ltm1, *ltm2;   // [B]
    int64_t *neg;                                    // [B, K] or null
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
// Keyed bijection of [0, P): a 4-round Feistel network on 2 x h bits (2^(2h) >= P, < 4P) with cycle walking (a value that falls
// outside [0, P) is permuted again; the orbit of a point of [0, P) returns to [0, P), so the restriction is a bijection).
// perm(0), perm(1), ... is therefore a duplicate-free pseudo-random order of the pool, and its first K admissible elements are a
// pseudo-random K-subset -- the role random.sample plays in the reference sampler (dataset_seq.py:176-199).
struct Feistel {
    uint32_t key[4];
    int h;
    uint32_t mask;
};
__device__ __forceinline__ Feistel make_feistel(int64_t P, uint32_t k0, uint32_t k1) {
    Feistel f;
    int bits = 1;
    while (((int64_t)1 << bits) < P) ++bits;
    f.h = (bits + 1) >> 1;
    f.mask = (1u << f.h) - 1u;
    f.key[0] = k0; f.key[1] = k1; f.key[2] = mix32(k0 ^ 0x9E3779B1u); f.key[3] = mix32(k1 ^ 0x85EBCA77u);
    return f;
}
__device__ __forceinline__ int64_t feistel_perm(const Feistel& f, int64_t x, int64_t P) {
    do {
        uint32_t l = (uint32_t)(x >> f.h), r = (uint32_t)x & f.mask;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t t = l ^ (mix32(r ^ f.key[i]) & f.mask);
            l = r;
            r = t;
        }
        x = ((int64_t)l << f.h) | r;
    } while (x >= P);
    return x;
}

__global__ void __launch_bounds__(128)
k_build_batch(BatchSrc s, BatchDst d, const int64_t* __restrict__ rows, int64_t n_rows_total, int B, int L, int K,
              int long_length, int64_t pad_id, uint64_t seed, int* __restrict__ err) {
    const int b = blockIdx.x, t = threadIdx.x;
    const int64_t row = rows[b];
    if (row < 0 || row >= n_rows_total) { if (t == 0) atomicExch(err, 1); return; }
    const int64_t o1 = s.h1_offs[row], n1 = s.h1_offs[row + 1] - o1;
    const int64_t o2 = s.h2_offs[row], n2 = s.h2_offs[row + 1] - o2;
    for (int l = t; l < L; l += blockDim.x) {
        const int64_t p1 = n1 - L + l, p2 = n2 - L + l;
        d.seq_d1[(int64_t)b * L + l] = p1 >= 0 ? s.h1_vals[o1 + p1] : pad_id;
        d.seq_d2[(int64_t)b * L + l] = p2 >= 0 ? s.h2_vals[o2 + p2] : pad_id;
    }
    const int dom = s.domain[row];
    if (t == 0) {
        d.i_node[b] = s.target[row];
        d.user_node[b] = s.user[row];
        d.domain_id[b] = dom;
        d.overlap_label[b] = s.overlap[row];
        d.ltm1[b] = n1 >= long_length ? 1 : 0;
        d.ltm2[b] = n2 >= long_length ? 1 : 0;
    }
    if (!d.neg || K <= 0 || t >= 32) return;
    // ---- negative sampler (warp 0)
    const int lane = t;
    const int64_t* pool = dom == 0 ? s.pool1 : s.pool2;
    const int64_t P = dom == 0 ? s.n_pool1 : s.n_pool2;
    const int64_t eo = s.ex_offs[row], ne = s.ex_offs[row + 1] - eo;
    const uint32_t h1 = mix32((uint32_t)seed ^ mix32((uint32_t)row * 0x9E3779B1u + 0x85EBCA77u));
    const uint32_t h2 = mix32((uint32_t)(seed >> 32) ^ mix32((uint32_t)(row >> 32) + h1 + 0xC2B2AE3Du));
    const Feistel fe = make_feistel(P, h1, h2);
    int cnt = 0;
    for (int64_t j0 = 0; j0 < P && cnt < K; j0 += 32) {
        const int64_t j = j0 + lane;
        bool ok = j < P;
        int64_t cand = 0;
        if (ok) {
            cand = pool[feistel_perm(fe, j, P)];
            for (int64_t e = 0; e < ne; ++e)
                if (s.ex_vals[eo + e] == cand) { ok = false; break; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int slot = cnt + __popc(m & ((1u << lane) - 1u));
        if (ok && slot < K) d.neg[(int64_t)b * K + slot] = cand;
        cnt += __popc(m);
    }
    if (cnt < K && lane == 0) atomicExch(err, 2);                          // pool minus history smaller than K
}

}  // namespace amid

using namespace amid;

extern "C" int amid_batch_build(const amid_batch_source* src, const int64_t* rows, int32_t B, int32_t L, int32_t K,
                                int32_t long_length, int64_t pad_id, uint64_t seed, const amid_batch_out* out,
                                amid_stream_t s_) {
    AMID_REQUIRE(src && rows && out, "batch_build: null argument");
    AMID_REQUIRE(B > 0 && L > 0 && K >= 0, "batch_build: bad sizes B=%d L=%d K=%d", B, L, K);
    AMID_REQUIRE(src->hist_d1_offs && src->hist_d2_offs && src->target && src->user && src->domain && src->overlap && src->n_rows > 0,
                 "batch_build: incomplete source");
    AMID_REQUIRE(out->seq_d1 && out->seq_d2 && out->i_node && out->user_node && out->domain_id && out->overlap_label &&
                 out->long_tail_mask_d1 && out->long_tail_mask_d2, "batch_build: incomplete output");
    AMID_REQUIRE(K == 0 || !out->neg_samples || (src->excl_offs && src->pool_d1 && src->pool_d2 && src->n_pool_d1 > 0 && src->n_pool_d2 > 0),
                 "batch_build: the sampler needs both pools and the exclusion lists");
    int* err = err_flag();
    AMID_REQUIRE(err, "batch_build: cannot allocate error flag");
    BatchSrc s{src->hist_d1_vals, src->hist_d2_vals, src->excl_vals, src->hist_d1_offs, src->hist_d2_offs, src->excl_offs,
               src->target, src->user, src->domain, src->overlap, src->pool_d1, src->pool_d2, src->n_pool_d1, src->n_pool_d2};
    BatchDst d{out->seq_d1, out->seq_d2, out->i_node, out->user_node, out->domain_id, out->overlap_label,
               out->long_tail_mask_d1, out->long_tail_mask_d2, K > 0 ? out->neg_samples : nullptr};
    AMID_K("k_build_batch", (cudaStream_t)s_);
    k_build_batch<<<B, 128, 0, (cudaStream_t)s_>>>(s, d, rows, src->n_rows, B, L, K, long_length, pad_id, seed, err);
    AMID_LAUNCH_CHECK("k_build_batch");
    return 0;
}
