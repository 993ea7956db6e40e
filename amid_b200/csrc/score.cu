// predictModule scorer (model_seq.py:32-54) for 1 or 3 heads, the domain-masked BCE and
// doubly-robust losses (train_sr.py:205-212, train_sr_dr.py:212-221, 385-395) and the
// eval rank counts (utils.py:296-301).
//
// score = sigmoid(w2 . relu(W0 [u ; item] + b0) + b2) is split as A = W0[:, :128] u (once per
// user and domain) and Bc = W0[:, 128:] item + b0 (once per candidate, shared by both
// domains); the [B,C,256] concat of the reference is never materialised.
#include "common.cuh"

namespace amid {

constexpr int MAXH = 3;      // heads: predictModule, predict_ips, predict_gfunc
constexpr int MAXHID = 64;

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }   // keep float4 smem views 16B aligned

struct HeadPtrs {
    const float* w0[MAXH];
    const float* b0[MAXH];
    const float* w2[MAXH];
    const float* b2[MAXH];
};

// smem layout (floats): WiT[nh][128][hid] | A[nh][2][hid] | U[2][128] | item rows[4 warps][128]
__global__ void __launch_bounds__(128)
k_score_fwd(const float* __restrict__ u1, const float* __restrict__ u2, const float* __restrict__ items, HeadPtrs hp,
            int nh, int hid, int B, int C, float* __restrict__ probs) {
    extern __shared__ __align__(16) float smem[];
    float* WiT = smem;
    float* A = WiT + nh * D * hid;
    float* U = A + pad4(nh * 2 * hid);
    float* IT = U + 2 * D;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int idx = t; idx < nh * hid * D; idx += 128) {   // transpose the item half into smem
        const int hd = idx / (hid * D), rem = idx % (hid * D), hh = rem / D, k = rem % D;
        WiT[(hd * D + k) * hid + hh] = __ldg(hp.w0[hd] + (size_t)hh * 2 * D + D + k);
    }
    U[t] = u1[(size_t)b * D + t];
    U[D + t] = u2[(size_t)b * D + t];
    __syncthreads();
    for (int idx = t; idx < nh * 2 * hid; idx += 128) {   // A[hd][dom][hh] = <W0[hh][0:128], u_dom>
        const int hd = idx / (2 * hid), dom = (idx / hid) & 1, hh = idx % hid;
        const float4* w = reinterpret_cast<const float4*>(hp.w0[hd] + (size_t)hh * 2 * D);
        const float4* uu = reinterpret_cast<const float4*>(U + dom * D);
        float s = 0.f;
        for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 a = __ldg(w + k4), x = uu[k4];
            s = fmaf(a.x, x.x, s); s = fmaf(a.y, x.y, s); s = fmaf(a.z, x.z, s); s = fmaf(a.w, x.w, s);
        }
        A[idx] = s;
    }
    __syncthreads();
    float* it = IT + warp * D;
    for (int c = warp; c < C; c += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(items + ((size_t)b * C + c) * D) + lane);
        __syncwarp();
        *reinterpret_cast<float4*>(it + lane * 4) = v;
        __syncwarp();
        for (int hd = 0; hd < nh; ++hd) {
            float z1 = 0.f, z2 = 0.f;
            for (int hh = lane; hh < hid; hh += 32) {
                float bc = __ldg(hp.b0[hd] + hh);
                const float* w = WiT + hd * D * hid + hh;
#pragma unroll 8
                for (int k = 0; k < D; ++k) bc = fmaf(it[k], w[k * hid], bc);
                const float w2 = __ldg(hp.w2[hd] + hh);
                z1 = fmaf(w2, fmaxf(A[(hd * 2 + 0) * hid + hh] + bc, 0.f), z1);
                z2 = fmaf(w2, fmaxf(A[(hd * 2 + 1) * hid + hh] + bc, 0.f), z2);
            }
            z1 = warp_sum(z1);
            z2 = warp_sum(z2);
            if (lane == 0) {
                const float bb = __ldg(hp.b2[hd]);
                probs[((size_t)(hd * 2 + 0) * B + b) * C + c] = 1.0f / (1.0f + expf(-(z1 + bb)));
                probs[((size_t)(hd * 2 + 1) * B + b) * C + c] = 1.0f / (1.0f + expf(-(z2 + bb)));
            }
        }
    }
}

// ---- backward.  One CTA per group of SB samples; weight-gradient partial per CTA.
// partial layout per (group, head): dW0[hid][256] | db0[hid] | dw2[hid] | db2[1]
constexpr int SB = 4;
__host__ __device__ inline int head_grad_floats(int hid) { return hid * 2 * D + 2 * hid + 1; }

__global__ void __launch_bounds__(256)
k_score_bwd(const float* __restrict__ u1, const float* __restrict__ u2, const float* __restrict__ items, HeadPtrs hp,
            int nh, int hid, int B, int C, const float* __restrict__ probs, const float* __restrict__ dprobs,
            float* __restrict__ du1, float* __restrict__ du2, float* __restrict__ ditems, float* __restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    const int HG = head_grad_floats(hid);
    float* G = smem;                        // [nh][HG] gradient accumulators
    float* A = G + pad4(nh * HG);           // [nh][2][hid]
    float* dA = A + pad4(nh * 2 * hid);     // [nh][2][hid]
    float* dBc = dA + pad4(nh * 2 * hid);   // [nh][hid]
    float* PRE = dBc + pad4(nh * hid);      // [nh][hid] Bc of the current candidate
    float* U = PRE + pad4(nh * hid);        // [2][128]
    float* IT = U + 2 * D;                  // [128]
    float* DZ = IT + D;                     // [nh][2]
    const int t = threadIdx.x;
    for (int i = t; i < nh * HG; i += 256) G[i] = 0.f;
    const int b0 = blockIdx.x * SB, b1 = min(B, b0 + SB);
    for (int b = b0; b < b1; ++b) {
        __syncthreads();
        if (t < D) { U[t] = u1[(size_t)b * D + t]; U[D + t] = u2[(size_t)b * D + t]; }
        for (int i = t; i < nh * 2 * hid; i += 256) dA[i] = 0.f;
        __syncthreads();
        for (int idx = t; idx < nh * 2 * hid; idx += 256) {
            const int hd = idx / (2 * hid), dom = (idx / hid) & 1, hh = idx % hid;
            const float4* w = reinterpret_cast<const float4*>(hp.w0[hd] + (size_t)hh * 2 * D);
            const float4* uu = reinterpret_cast<const float4*>(U + dom * D);
            float s = 0.f;
            for (int k4 = 0; k4 < D / 4; ++k4) {
                const float4 a = __ldg(w + k4), x = uu[k4];
                s = fmaf(a.x, x.x, s); s = fmaf(a.y, x.y, s); s = fmaf(a.z, x.z, s); s = fmaf(a.w, x.w, s);
            }
            A[idx] = s;
        }
        for (int c = 0; c < C; ++c) {
            __syncthreads();
            if (t < D) IT[t] = items[((size_t)b * C + c) * D + t];
            if (t < nh * 2) {
                const int hd = t >> 1, dom = t & 1;
                const size_t o = ((size_t)(hd * 2 + dom) * B + b) * C + c;
                const float p = probs[o];
                DZ[t] = dprobs[o] * p * (1.0f - p);     // sigmoid backward from the saved output
            }
            __syncthreads();
            // Bc[hd][hh]
            for (int idx = t; idx < nh * hid; idx += 256) {
                const int hd = idx / hid, hh = idx % hid;
                const float4* w = reinterpret_cast<const float4*>(hp.w0[hd] + (size_t)hh * 2 * D + D);
                float s = __ldg(hp.b0[hd] + hh);
                for (int k4 = 0; k4 < D / 4; ++k4) {
                    const float4 a = __ldg(w + k4);
                    const float4 x = *reinterpret_cast<const float4*>(IT + k4 * 4);
                    s = fmaf(a.x, x.x, s); s = fmaf(a.y, x.y, s); s = fmaf(a.z, x.z, s); s = fmaf(a.w, x.w, s);
                }
                PRE[idx] = s;
            }
            __syncthreads();
            for (int idx = t; idx < nh * hid; idx += 256) {
                const int hd = idx / hid, hh = idx % hid;
                const float w2 = __ldg(hp.w2[hd] + hh);
                float* g = G + hd * HG;
                float dsum = 0.f;
#pragma unroll
                for (int dom = 0; dom < 2; ++dom) {
                    const float pre = A[(hd * 2 + dom) * hid + hh] + PRE[idx];
                    const float dz = DZ[hd * 2 + dom];
                    g[hid * 2 * D + hid + hh] += dz * fmaxf(pre, 0.f);          // dw2
                    const float dpre = pre > 0.f ? dz * w2 : 0.f;
                    dA[(hd * 2 + dom) * hid + hh] += dpre;
                    dsum += dpre;
                }
                dBc[idx] = dsum;
                g[hid * 2 * D + hh] += dsum;                                    // db0
            }
            if (t < nh) G[t * HG + hid * 2 * D + 2 * hid] += DZ[t * 2] + DZ[t * 2 + 1];   // db2
            __syncthreads();
            // ditem[k] = sum_{hd,hh} dBc W0[hh][128+k] ; dW0[hh][128+k] += dBc[hh] item[k]
            if (t < D) {
                float s = 0.f;
                for (int hd = 0; hd < nh; ++hd)
                    for (int hh = 0; hh < hid; ++hh)
                        s = fmaf(dBc[hd * hid + hh], __ldg(hp.w0[hd] + (size_t)hh * 2 * D + D + t), s);
                ditems[((size_t)b * C + c) * D + t] = s;
            }
            for (int idx = t; idx < nh * hid * D; idx += 256) {
                const int hd = idx / (hid * D), rem = idx % (hid * D), hh = rem / D, k = rem % D;
                G[hd * HG + hh * 2 * D + D + k] += dBc[hd * hid + hh] * IT[k];
            }
        }
        __syncthreads();
        // du_dom[k] = sum_{hd,hh} dA W0[hh][k] ; dW0[hh][k] += sum_dom dA u_dom[k]
        {
            const int dom = t >> 7, k = t & 127;
            float s = 0.f;
            for (int hd = 0; hd < nh; ++hd)
                for (int hh = 0; hh < hid; ++hh)
                    s = fmaf(dA[(hd * 2 + dom) * hid + hh], __ldg(hp.w0[hd] + (size_t)hh * 2 * D + k), s);
            (dom == 0 ? du1 : du2)[(size_t)b * D + k] = s;
        }
        for (int idx = t; idx < nh * hid * D; idx += 256) {
            const int hd = idx / (hid * D), rem = idx % (hid * D), hh = rem / D, k = rem % D;
            G[hd * HG + hh * 2 * D + k] += dA[(hd * 2 + 0) * hid + hh] * U[k] + dA[(hd * 2 + 1) * hid + hh] * U[D + k];
        }
    }
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * nh * HG;
    for (int i = t; i < nh * HG; i += 256) out[i] = G[i];
}

struct HeadGradPtrs {
    float* w0[MAXH];
    float* b0[MAXH];
    float* w2[MAXH];
    float* b2[MAXH];
};
__global__ void k_score_reduce(const float* __restrict__ part, int groups, int nh, int hid, HeadGradPtrs g) {
    const int HG = head_grad_floats(hid);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nh * HG) return;
    float s = 0.f;
    for (int i = 0; i < groups; ++i) s += part[(size_t)i * nh * HG + e];
    const int hd = e / HG, r = e % HG;
    if (r < hid * 2 * D) g.w0[hd][r] = s;
    else if (r < hid * 2 * D + hid) g.b0[hd][r - hid * 2 * D] = s;
    else if (r < hid * 2 * D + 2 * hid) g.w2[hd][r - hid * 2 * D - hid] = s;
    else g.b2[hd][0] = s;
}

// ---- losses: single CTA (B*C is small in training; eval sizes still finish in microseconds)
__device__ __forceinline__ float bce(float p, float y) {   // ATen binary_cross_entropy, log clamp -100
    return (y - 1.0f) * fmaxf(log1pf(-p), -100.0f) - y * fmaxf(logf(p), -100.0f);
}
__device__ __forceinline__ float dbce(float p, float y) {  // ATen binary_cross_entropy_backward, EPS 1e-12
    return (p - y) / fmaxf((1.0f - p) * p, 1e-12f);
}
__global__ void __launch_bounds__(1024)
k_loss(const float* __restrict__ probs, int nh, int B, int C, const float* __restrict__ labels,
       const int64_t* __restrict__ dom, const int64_t* __restrict__ ob, int mode, float w_e, float inv,
       float* __restrict__ losses, float* __restrict__ dprobs) {
    __shared__ float red[3][32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t BC = (size_t)B * C;
    float s_cls = 0.f, s_e = 0.f, s_r = 0.f;
    for (size_t e = t; e < BC; e += 1024) {
        const int b = (int)(e / C);
        const float y = labels[e];
        const float msk[2] = {(float)(1 - dom[b]), (float)dom[b]};
        const float obv = (mode == 2) ? (float)ob[b] : 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float p = probs[(size_t)(0 * 2 + k) * BC + e];
            const float l = bce(p, y), dl = dbce(p, y);
            float dp = 0.f, dips = 0.f, dg = 0.f;
            if (mode == 0 || mode == 1) {
                s_cls += l * msk[k];
                dp = msk[k] * dl * inv;
            }
            if (mode == 1) {
                const float ips = probs[(size_t)(1 * 2 + k) * BC + e], g = probs[(size_t)(2 * 2 + k) * BC + e];
                const float df = l - g;
                s_e += df * df / ips * msk[k];
                const float c = w_e * msk[k] * inv;
                dp += c * 2.0f * df / ips * dl;
                dg = -c * 2.0f * df / ips;
                dips = -c * df * df / (ips * ips);
            }
            if (mode == 2) {
                const float ips = probs[(size_t)(1 * 2 + k) * BC + e], g = probs[(size_t)(2 * 2 + k) * BC + e];
                const float q = l * l - g * g;
                s_r += (g * g + obv * (q * q) / ips) * msk[k];
                const float c = msk[k] * inv;
                dp = c * obv * 2.0f * q * 2.0f * l * dl / ips;
                dg = c * (2.0f * g - obv * 2.0f * q * 2.0f * g / ips);
                dips = -c * obv * q * q / (ips * ips);
            }
            dprobs[(size_t)(0 * 2 + k) * BC + e] = dp;
            if (nh == 3) {
                dprobs[(size_t)(1 * 2 + k) * BC + e] = dips;
                dprobs[(size_t)(2 * 2 + k) * BC + e] = dg;
            }
        }
    }
    s_cls = warp_sum(s_cls); s_e = warp_sum(s_e); s_r = warp_sum(s_r);
    if (lane == 0) { red[0][warp] = s_cls; red[1][warp] = s_e; red[2][warp] = s_r; }
    __syncthreads();
    if (t < 3) {
        float r = 0.f;
        for (int w = 0; w < 32; ++w) r += red[t][w];
        losses[t] = r * inv;
    }
}

// ---- eval rank counts: warp per row
__global__ void __launch_bounds__(256)
k_rank_counts(const float* __restrict__ scores, int64_t N, int C, float fix, int* __restrict__ ng, int* __restrict__ ne) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= N) return;
    const float* s = scores + r * C;
    const float s0 = s[0] - fix;     // fp32 subtraction as numpy does on a float32 array (train_sr.py:114)
    int g = 0, q = 0;
    for (int c = 1 + lane; c < C; c += 32) {
        const float v = s[c];
        g += v > s0;
        q += v == s0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g += __shfl_xor_sync(0xffffffffu, g, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) { ng[r] = g; ne[r] = q; }
}

static int fill_heads(const amid_head_tensors* heads, int nh, HeadPtrs* hp) {
    for (int i = 0; i < nh; ++i) {
        AMID_REQUIRE(heads[i].w0 && heads[i].b0 && heads[i].w2 && heads[i].b2, "score: head %d has null tensors", i);
        AMID_REQUIRE(aligned16(heads[i].w0), "score: head %d weight misaligned", i);
        hp->w0[i] = heads[i].w0; hp->b0[i] = heads[i].b0; hp->w2[i] = heads[i].w2; hp->b2[i] = heads[i].b2;
    }
    return 0;
}
static int check_score(int nh, int hid, int B, int C) {
    AMID_REQUIRE(nh >= 1 && nh <= MAXH, "score: n_heads=%d must be 1..3", nh);
    AMID_REQUIRE(hid >= 1 && hid <= MAXHID, "score: hid=%d must be 1..64", hid);
    AMID_REQUIRE(B > 0 && C > 0, "score: B=%d C=%d", B, C);
    return 0;
}

}  // namespace amid

using namespace amid;

extern "C" int amid_score_fwd(const float* u1, const float* u2, const float* items, const amid_head_tensors* heads,
                              int32_t nh, int32_t hid, int32_t B, int32_t C, float* probs, amid_stream_t s_) {
    if (int rc = check_score(nh, hid, B, C)) return rc;
    AMID_REQUIRE(u1 && u2 && items && heads && probs, "score_fwd: null argument");
    AMID_REQUIRE(aligned16(items) && aligned16(u1) && aligned16(u2), "score_fwd: misaligned buffer");
    HeadPtrs hp{};
    if (int rc = fill_heads(heads, nh, &hp)) return rc;
    const size_t smem = (size_t)(nh * D * hid + pad4(nh * 2 * hid) + 2 * D + 4 * D) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute((const void*)k_score_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-3, "score_fwd: smem attribute: %s", cudaGetErrorString(e));
    AMID_K("k_score_fwd", (cudaStream_t)s_);
    k_score_fwd<<<B, 128, smem, (cudaStream_t)s_>>>(u1, u2, items, hp, nh, hid, B, C, probs);
    AMID_LAUNCH_CHECK("k_score_fwd");
    return 0;
}

extern "C" int64_t amid_score_bwd_workspace_bytes(int32_t nh, int32_t hid, int32_t B, int32_t) {
    const int64_t groups = (B + SB - 1) / SB;
    return groups * nh * head_grad_floats(hid) * (int64_t)sizeof(float);
}

extern "C" int amid_score_bwd(const float* u1, const float* u2, const float* items, const amid_head_tensors* heads,
                              int32_t nh, int32_t hid, int32_t B, int32_t C, const float* probs, const float* dprobs,
                              float* du1, float* du2, float* ditems, amid_head_tensors* G, void* workspace,
                              int64_t workspace_bytes, amid_stream_t s_) {
    if (int rc = check_score(nh, hid, B, C)) return rc;
    AMID_REQUIRE(u1 && u2 && items && heads && probs && dprobs && du1 && du2 && ditems && G && workspace,
                 "score_bwd: null argument");
    AMID_REQUIRE(workspace_bytes >= amid_score_bwd_workspace_bytes(nh, hid, B, C), "score_bwd: workspace too small");
    HeadPtrs hp{};
    if (int rc = fill_heads(heads, nh, &hp)) return rc;
    const int HG = head_grad_floats(hid);
    const int groups = (B + SB - 1) / SB;
    const size_t smem = (size_t)(pad4(nh * HG) + 2 * pad4(nh * 2 * hid) + 2 * pad4(nh * hid) + 2 * D + D + nh * 2 + 8) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute((const void*)k_score_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-3, "score_bwd: smem attribute: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)s_;
    AMID_K("k_score_bwd", s);
    k_score_bwd<<<groups, 256, smem, s>>>(u1, u2, items, hp, nh, hid, B, C, probs, dprobs, du1, du2, ditems,
                                          (float*)workspace);
    AMID_LAUNCH_CHECK("k_score_bwd");
    HeadGradPtrs gp{};
    for (int i = 0; i < nh; ++i) {
        AMID_REQUIRE(G[i].w0 && G[i].b0 && G[i].w2 && G[i].b2, "score_bwd: grad head %d has null tensors", i);
        gp.w0[i] = G[i].w0; gp.b0[i] = G[i].b0; gp.w2[i] = G[i].w2; gp.b2[i] = G[i].b2;
    }
    AMID_K("k_score_reduce", s);
    k_score_reduce<<<(nh * HG + 255) / 256, 256, 0, s>>>((const float*)workspace, groups, nh, hid, gp);
    AMID_LAUNCH_CHECK("k_score_reduce");
    return 0;
}

extern "C" int amid_loss_fwd_bwd(const float* probs, int32_t nh, int32_t B, int32_t C, const float* labels,
                                 const int64_t* domain_id, const int64_t* ob_label, int32_t mode, float dr_e_w,
                                 float inv_count, float* losses, float* dprobs, amid_stream_t s_) {
    AMID_REQUIRE(probs && labels && domain_id && losses && dprobs, "loss: null argument");
    AMID_REQUIRE(mode >= 0 && mode <= 2, "loss: mode=%d", mode);
    AMID_REQUIRE(nh == 1 || nh == 3, "loss: n_heads=%d must be 1 or 3", nh);
    AMID_REQUIRE(mode == 0 || nh == 3, "loss: the doubly-robust losses need the 3 isDR heads");
    AMID_REQUIRE(mode != 2 || ob_label, "loss: mode 2 needs ob_label");
    AMID_K("k_loss", (cudaStream_t)s_);
    k_loss<<<1, 1024, 0, (cudaStream_t)s_>>>(probs, nh, B, C, labels, domain_id, ob_label, mode, dr_e_w, inv_count,
                                             losses, dprobs);
    AMID_LAUNCH_CHECK("k_loss");
    return 0;
}

extern "C" int amid_rank_counts(const float* scores, int64_t N, int32_t C, float fix, int32_t* n_greater,
                                int32_t* n_equal, amid_stream_t s_) {
    AMID_REQUIRE(scores && n_greater && n_equal && N >= 0 && C >= 1, "rank_counts: bad argument");
    if (N == 0) return 0;
    AMID_K("k_rank_counts", (cudaStream_t)s_);
    k_rank_counts<<<(unsigned)((N * 32 + 255) / 256), 256, 0, (cudaStream_t)s_>>>(scores, N, C, fix, n_greater, n_equal);
    AMID_LAUNCH_CHECK("k_rank_counts");
    return 0;
}
