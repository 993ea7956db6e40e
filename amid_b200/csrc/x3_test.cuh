// Bring-up / unit-test entry points of the split-operand ("x3") pipeline (x3.cuh); included by encoder.cu.
#pragma once
#include "x3.cuh"

// ---- split-operand ("x3") bring-up: y = x w^T + b with FP16 pair pieces, the token tile in TENSOR MEMORY (tcgen05.st +
// A-from-TMEM MMA), the weight image fetched with one bulk copy; and the BF16-triple weight-gradient kernel.

namespace amid {
__global__ void __launch_bounds__(256, 2)
k_x3_linear(const float* __restrict__ x, const uint8_t* __restrict__ wimg, const float* __restrict__ winv,
            const float* __restrict__ b, int M, float* __restrict__ y) {
    using namespace x3;
    extern __shared__ uint8_t smem_raw[];
    __shared__ SharedX sh;
    uint8_t* Wb = align1k(smem_raw);
    float* stage = reinterpret_cast<float*>(Wb + WIMG_BYTES) + (threadIdx.x >> 5) * WSTAGE_FLOATS;
    const int row0 = blockIdx.x * 128;
    setup(sh, CHAIN_TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    load_w_bulk(sh, Wb, wimg);
    Epi e;
    const int rv = rows_valid(row0, e, M);
    const size_t wbase = (size_t)(row0 + e.wrow0) * D;
    float v[2][32];
    float am = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        warp_load32(stage, e.lane, x + wbase + e.cb + half * 32, rv, v[half]);
        am = absmax32(v[half], am);
    }
    float sc, inv;
    pow2_scale(row_max(sh, e, 2, am), sc, inv);
    const uint32_t tl = sh.tmem + e.lane_addr;
    put_a32(tl, e.cb, v[0], sc);
    put_a32(tl, e.cb + 32, v[1], sc);
    uint32_t ph_mma = 0, ph_w = 0;
    run_gemm_x3(sh, 0, Wb, false, ph_mma, ph_w);
    const float f = inv * __ldg(winv);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float a[32];
        const int c0 = e.cb + half * 32;
        tmem_ld32(tl + ACC_COL + c0, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], f, __ldg(b + c0 + i));
        warp_store32(stage, e.lane, a, y + wbase + c0, rv);
    }
    teardown(sh, CHAIN_TMEM_COLS);
}
}  // namespace amid

// scratch: >= 64 KB + 4 B device buffer for the weight image and its inverse scale
extern "C" int amid_x3_linear_test(const float* x, const float* w, const float* b, int32_t M, float* y, void* scratch,
                                   amid_stream_t s_) {
    AMID_REQUIRE(x && w && b && y && scratch && M > 0, "x3_linear_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)k_x3_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)x3::CHAINX_SMEM);
    if (e != cudaSuccess) return set_error(-3, "x3_linear_test: smem attribute: %s", cudaGetErrorString(e));
    x3::PrepJobsX pj;
    pj.src[0] = w;
    uint8_t* img = (uint8_t*)scratch;
    float* inv = (float*)(img + x3::WIMG_BYTES);
    AMID_K("k_prep_wx3", s_);
    x3::k_prep_wx3<<<1, 256, 0, (cudaStream_t)s_>>>(pj, img, inv, 0);
    AMID_LAUNCH_CHECK("k_prep_wx3");
    AMID_K("k_x3_linear", s_);
    k_x3_linear<<<(M + 127) / 128, 256, x3::CHAINX_SMEM, (cudaStream_t)s_>>>(x, img, inv, b, M, y);
    AMID_LAUNCH_CHECK("k_x3_linear");
    return 0;
}
extern "C" int amid_x3_wgrad_test(const float* dy, const float* x, int32_t M, float* wpart, float* bpart, int32_t n_ctas,
                                  amid_stream_t s_) {
    AMID_REQUIRE(dy && x && wpart && bpart && M > 0 && n_ctas > 0, "x3_wgrad_test: bad argument");
    cudaError_t e = cudaFuncSetAttribute((const void*)x3::k_wgrad_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)x3::WGRADX_SMEM);
    if (e != cudaSuccess) return set_error(-3, "x3_wgrad_test: smem attribute: %s", cudaGetErrorString(e));
    x3::WgradJobsX j;
    for (int i = 0; i < 6; ++i) { j.dY[i] = dy; j.X[i] = x; }
    AMID_K("k_wgrad_x3", s_);
    x3::k_wgrad_x3<<<dim3(n_ctas, 1), x3::WGX_THREADS, x3::WGRADX_SMEM, (cudaStream_t)s_>>>(j, M, wpart, bpart);
    AMID_LAUNCH_CHECK("k_wgrad_x3");
    return 0;
}

// ---- attention kernels side by side (unit tests and micro-benchmarks): impl 0 = fp32 CUDA cores, 1 = mma.sync TF32,
// 2 = mma.sync 3xTF32, 3 = tcgen05 FP16-pair split with the scores in tensor memory
extern "C" int amid_attn_fwd_test(const float* q, const float* k, const float* v, float* o, float* lse, int32_t B, int32_t L,
                                  const amid_dropout* drop, uint32_t site, int32_t impl, amid_stream_t s_) {
    AMID_REQUIRE(q && k && v && o && lse && B > 0 && L > 0 && L <= 512, "attn_fwd_test: bad argument");
    cudaStream_t stream = (cudaStream_t)s_;
    const DropCfg dc = with_offsets(make_drop(drop), L);
    if (impl == 0) {
        const size_t smem = (size_t)2 * L * DH * sizeof(float);
        if (int rc = ensure_smem((const void*)k_attn_fwd, smem)) return rc;
        AMID_K("k_attn_fwd", stream);
        k_attn_fwd<<<B * H, (int)round_up((L + 1) / 2, 32), smem, stream>>>(q, k, v, o, lse, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_fwd");
    } else if (impl == 1 || impl == 2) {
        const size_t smem = (size_t)2 * ((L + 15) / 16 * 16) * attn::LDS * sizeof(float);
        if (impl == 1) {
            if (int rc = ensure_smem((const void*)attn::k_attn_fwd_mma<false>, smem)) return rc;
            AMID_K("k_attn_fwd_mma", stream);
            attn::k_attn_fwd_mma<false><<<B * H, attn::NW * 32, smem, stream>>>(q, k, v, o, lse, L, dc, site);
            AMID_LAUNCH_CHECK("k_attn_fwd_mma");
        } else {
            if (int rc = ensure_smem((const void*)attn::k_attn_fwd_mma<true>, smem)) return rc;
            AMID_K("k_attn_fwd_mma3", stream);
            attn::k_attn_fwd_mma<true><<<B * H, attn::NW * 32, smem, stream>>>(q, k, v, o, lse, L, dc, site);
            AMID_LAUNCH_CHECK("k_attn_fwd_mma3");
        }
    } else if (impl == 4) {
        AMID_REQUIRE(L >= attn_p::PMINL && L <= attn_tc::MAXL, "attn_fwd_test: pipelined tcgen05 path needs %d <= L <= %d", attn_p::PMINL, attn_tc::MAXL);
        auto kfn = dc.train ? attn_p::k_attn_fwd_p<true> : attn_p::k_attn_fwd_p<false>;
        if (int rc = ensure_smem((const void*)kfn, attn_p::PFWD_SMEM)) return rc;
        AMID_K("k_attn_fwd_p", stream);
        kfn<<<B * H, 256, attn_p::PFWD_SMEM, stream>>>(q, k, v, o, lse, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_fwd_p");
    } else {
        AMID_REQUIRE(L <= attn_tc::MAXL, "attn_fwd_test: tcgen05 path needs L <= %d", attn_tc::MAXL);
        if (int rc = ensure_smem((const void*)attn_tc::k_attn_fwd_tc, attn_tc::FWD_SMEM)) return rc;
        AMID_K("k_attn_fwd_tc", stream);
        attn_tc::k_attn_fwd_tc<<<B * H, 256, attn_tc::FWD_SMEM, stream>>>(q, k, v, o, lse, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_fwd_tc");
    }
    return 0;
}

#ifdef AMID_ATTN_DBG
extern "C" int amid_attn_dbg_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, attn_p::g_dbg, sizeof(long long) * 20 * 256);
}
#endif
extern "C" int amid_attn_bwd_test(const float* q, const float* k, const float* v, const float* o, const float* lse,
                                  const float* dO, float* dq, float* dk, float* dv, int32_t B, int32_t L,
                                  const amid_dropout* drop, uint32_t site, int32_t impl, amid_stream_t s_) {
    AMID_REQUIRE(q && k && v && o && lse && dO && dq && dk && dv && B > 0 && L > 0 && L <= 512, "attn_bwd_test: bad argument");
    cudaStream_t stream = (cudaStream_t)s_;
    const DropCfg dc = with_offsets(make_drop(drop), L);
    if (impl == 0) {
        const size_t smem = (size_t)(4 * L * DH + 2 * L) * sizeof(float);
        if (int rc = ensure_smem((const void*)k_attn_bwd, smem)) return rc;
        AMID_K("k_attn_bwd", stream);
        k_attn_bwd<<<B * H, (int)round_up((L + 1) / 2, 32), smem, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_bwd");
    } else if (impl == 1 || impl == 2) {
        const size_t smem = (size_t)((L + 15) / 16 * 16) * (4 * attn::LDS + 2) * sizeof(float);
        if (impl == 1) {
            if (int rc = ensure_smem((const void*)attn::k_attn_bwd_mma<false>, smem)) return rc;
            AMID_K("k_attn_bwd_mma", stream);
            attn::k_attn_bwd_mma<false><<<B * H, attn::NWB * 32, smem, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, dc, site);
            AMID_LAUNCH_CHECK("k_attn_bwd_mma");
        } else {
            if (int rc = ensure_smem((const void*)attn::k_attn_bwd_mma<true>, smem)) return rc;
            AMID_K("k_attn_bwd_mma3", stream);
            attn::k_attn_bwd_mma<true><<<B * H, attn::NWB * 32, smem, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, dc, site);
            AMID_LAUNCH_CHECK("k_attn_bwd_mma3");
        }
    } else if (impl == 4) {
        AMID_REQUIRE(L >= attn_p::PMINL && L <= attn_p::PMAXL, "attn_bwd_test: pipelined tcgen05 path needs %d <= L <= %d", attn_p::PMINL, attn_p::PMAXL);
        auto kfn = dc.train ? attn_p::k_attn_bwd_p<true> : attn_p::k_attn_bwd_p<false>;
        if (int rc = ensure_smem((const void*)kfn, attn_p::PBWD_SMEM)) return rc;
        AMID_K("k_attn_bwd_p", stream);
        kfn<<<std::min(B * H, sm_count()), attn_p::NTH, attn_p::PBWD_SMEM, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, B * H, dc, site);
        AMID_LAUNCH_CHECK("k_attn_bwd_p");
    } else if (impl == 5) {
        AMID_REQUIRE(L <= attn_tc::MAXL, "attn_bwd_test: two-pass tcgen05 path needs L <= %d", attn_tc::MAXL);
        auto kfn = dc.train ? attn_p::k_attn_bwd_t2<true> : attn_p::k_attn_bwd_t2<false>;
        if (int rc = ensure_smem((const void*)kfn, attn_p::T2_SMEM)) return rc;
        AMID_K("k_attn_bwd_t2", stream);
        kfn<<<B * H, 256, attn_p::T2_SMEM, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_bwd_t2");
    } else {
        AMID_REQUIRE(L <= attn_tc::MAXL, "attn_bwd_test: tcgen05 path needs L <= %d", attn_tc::MAXL);
        if (int rc = ensure_smem((const void*)attn_tc::k_attn_bwd_tc, attn_tc::BWD_SMEM)) return rc;
        AMID_K("k_attn_bwd_tc", stream);
        attn_tc::k_attn_bwd_tc<<<B * H, 256, attn_tc::BWD_SMEM, stream>>>(q, k, v, o, lse, dO, dq, dk, dv, L, dc, site);
        AMID_LAUNCH_CHECK("k_attn_bwd_tc");
    }
    return 0;
}
