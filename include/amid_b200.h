/* amid_b200 -- C ABI of the B200-native AMID SASRec hot path.
 *
 * The reference (WujiangXu/AMID) is pure Python/PyTorch and has no FFI of its own; its
 * "plugin API" for this path is the nn.Module contract of model_seq.SASRec
 * (model_seq.py:390-443).  amid_b200/model_seq.py keeps that contract and binds the
 * entry points below with ctypes.  Each entry point names the
 * reference code it replaces.
 *
 * Conventions: raw device pointers + sizes only; the caller owns every buffer,
 * including workspaces (sizes via the *_workspace_bytes functions); no allocation and no
 * device synchronisation inside (the two *_host_sync helpers excepted and named so);
 * kernels are launched on the passed stream; return value 0 = ok, negative = error and
 * amid_last_error() (thread local) says why.  Shapes: d = 128, heads = 8 (the reference
 * hard-codes 8 heads at model_seq.py:348; d = 128 is the run.sh/argparse default),
 * hid <= 64.  All floats are fp32, ids are int64, row-major contiguous, 16-byte aligned.
 * There is no CPU fallback.
 */
#ifndef AMID_B200_H
#define AMID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* amid_stream_t; /* cudaStream_t */

#define AMID_D 128
#define AMID_HEADS 8
#define AMID_BLOCKS 2

/* dropout description; train = 0 disables every dropout site (model.eval()). */
typedef struct {
    int32_t train;      /* 0 = eval */
    float p;            /* drop probability (reference: 0.5 everywhere, model_seq.py:335,350,355) */
    uint64_t seed;      /* per-step seed */
    uint32_t site_base; /* first site id used by this call (an encoder uses 7 sites) */
    int32_t batch_offset; /* position of this call's first sample in the GLOBAL batch: the keep bits are indexed by
                           * global sample, so a data-parallel rank draws exactly the bits a single GPU would draw for
                           * its slice (0 on a single GPU) */
} amid_dropout;

/* One Log2feats encoder (model_seq.py:331-357).  Used for parameters (read) and for
 * gradients (written).  Index [i] = block i. */
typedef struct {
    float* pos_emb;                  /* pos_emb.weight                [Lmax,128] */
    float* ln1_w[AMID_BLOCKS];       /* attention_layernorms.i.weight [128] */
    float* ln1_b[AMID_BLOCKS];
    float* in_w[AMID_BLOCKS];        /* attention_layers.i.in_proj_weight [384,128] */
    float* in_b[AMID_BLOCKS];        /* in_proj_bias [384] */
    float* out_w[AMID_BLOCKS];       /* out_proj.weight [128,128] */
    float* out_b[AMID_BLOCKS];
    float* ln2_w[AMID_BLOCKS];       /* forward_layernorms.i */
    float* ln2_b[AMID_BLOCKS];
    float* c1_w[AMID_BLOCKS];        /* forward_layers.i.conv1.weight [128,128,1] */
    float* c1_b[AMID_BLOCKS];
    float* c2_w[AMID_BLOCKS];
    float* c2_b[AMID_BLOCKS];
    float* ln3_w;                    /* last_layernorm */
    float* ln3_b;
} amid_encoder_tensors;

/* Activations the forward keeps for the backward, all [B*L,128] unless noted. */
typedef struct {
    float* qn[AMID_BLOCKS];   /* LN1 output (the "Q" of model_seq.py:373) */
    float* q[AMID_BLOCKS];    /* scaled query projection */
    float* k[AMID_BLOCKS];
    float* v[AMID_BLOCKS];
    float* o[AMID_BLOCKS];    /* attention output before out_proj */
    float* lse[AMID_BLOCKS];  /* [B,8,L] log-sum-exp of the attention rows */
    float* x1[AMID_BLOCKS];   /* Q + mha  (model_seq.py:378) */
    float* y[AMID_BLOCKS];    /* LN2 output */
    float* h[AMID_BLOCKS];    /* relu(dropout1(conv1)) */
    float* xout[AMID_BLOCKS]; /* block output after the timeline mask (model_seq.py:383) */
    float* st1[AMID_BLOCKS];  /* [B*L,2] mean, rstd of LN1 */
    float* st2[AMID_BLOCKS];
    float* st3;               /* last_layernorm stats */
} amid_encoder_saved;

const char* amid_last_error(void);
int amid_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t amid_launch_count(void);
/* optional per-kernel CUDA-event profiler: enable(1) starts recording every launch on its
 * own stream, report() synchronises and writes "kernel count total_ms" lines. */
int amid_profile_enable(int32_t on);
int64_t amid_profile_report_host_sync(char* buf, int64_t cap);

/* ---- a1: embItemLayerEnhance.forward (model_seq.py:27-29, calls :418-421) ---------- */
/* out[r,:] = table[ids[r],:]  for r < n_rows.  Bit-exact copy. */
int amid_emb_gather_fwd(const float* table, int64_t V, const int64_t* ids, int64_t n_rows,
                        float* out, amid_stream_t stream);
/* 1 if a gather since the last call saw an id outside [0,V) (such rows are skipped);
 * synchronises the device -- for tests / debugging, never on the hot path. */
int amid_gather_error_host_sync(void);

/* ---- a1+a2: gather fused with the Log2feats prologue (model_seq.py:360-366) -------- */
/* x0[b,l,:] = dropout(table[ids[b,l]] + pos[l]) * ~tmask,  tmask = (table[id]+pos == 0)
 * element-wise, bit-packed: tmask[(b*L+l)*4 + e] bit j  <->  column 4*j+e.
 * If ids == NULL the rows are read from `rows` ([B,L,128], the InnerComp path
 * model_seq.py:422-424) instead of the table. */
int amid_seq_embed_fwd(const float* table, int64_t V, const int64_t* ids, const float* rows,
                       const float* pos, int32_t B, int32_t L, float* x0, uint32_t* tmask,
                       const amid_dropout* drop, amid_stream_t stream);
/* All table reads of one step in ONE launch: the candidate rows (plain copy, a1) and both domains'
 * sequences with the fused prologue (a1+a2).  Dropout sites: domain 1 uses site 0, domain 2 site 8
 * (the site_base of `drop` is ignored).  This launch is the "emb-gather HBM GB/s" metric. */
int amid_embed_all_fwd(const float* table, int64_t V, const int64_t* ids_items, int64_t n_items,
                       const int64_t* ids_d1, const int64_t* ids_d2, const float* pos_d1, const float* pos_d2,
                       int32_t B, int32_t L, float* items, float* x0_d1, float* x0_d2, uint32_t* tmask_d1,
                       uint32_t* tmask_d2, const amid_dropout* drop, amid_stream_t stream);
/* backward of the above: dx0 <- dx0 * ~tmask * keep/(1-p) in place (this is then the
 * per-row table gradient), dpos[l,:] = sum_b dx0[b,l,:]. */
int amid_seq_embed_bwd(float* dx0, const uint32_t* tmask, int32_t B, int32_t L, float* dpos,
                       const amid_dropout* drop, amid_stream_t stream);

/* ---- a3-a5: Log2feats blocks + last LN (model_seq.py:371-385; MHA arithmetic from
 *      torch/nn/functional.py:5849-5856, 6630-6653; FFN model_seq.py:322-326) --------- */
int64_t amid_encoder_fwd_workspace_bytes(int32_t B, int32_t L);
int amid_encoder_fwd(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                     int32_t B, int32_t L, const amid_dropout* drop, amid_encoder_saved* S,
                     float* enc_out, void* workspace, int64_t workspace_bytes, amid_stream_t stream);
int64_t amid_encoder_bwd_workspace_bytes(int32_t B, int32_t L);
/* d_enc [B*L,128] is consumed; G receives the parameter gradients (overwritten, pos_emb
 * excluded -- see amid_seq_embed_bwd); dx0 [B*L,128] receives the input gradient. */
int amid_encoder_bwd(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                     int32_t B, int32_t L, const amid_dropout* drop, const amid_encoder_saved* S,
                     const float* enc_out, const float* d_enc, amid_encoder_tensors* G, float* dx0,
                     void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* Tensor-core variants (same arguments, same tensors): the 128x128x128 stages run as tcgen05.mma
 * with TF32 operands and fp32 accumulation in TMEM instead of fp32 CUDA-core tiles. */
int amid_encoder_fwd_tc(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                        int32_t B, int32_t L, const amid_dropout* drop, amid_encoder_saved* S,
                        float* enc_out, void* workspace, int64_t workspace_bytes, amid_stream_t stream);
int amid_encoder_bwd_tc(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                        int32_t B, int32_t L, const amid_dropout* drop, const amid_encoder_saved* S,
                        const float* enc_out, const float* d_enc, amid_encoder_tensors* G, float* dx0,
                        void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* BF16-operand variants (fp32 accumulate, fp32 tensors in HBM; 2 CTAs per SM). */
int amid_encoder_fwd_bf16(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                          int32_t B, int32_t L, const amid_dropout* drop, amid_encoder_saved* S,
                          float* enc_out, void* workspace, int64_t workspace_bytes, amid_stream_t stream);
int amid_encoder_bwd_bf16(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                          int32_t B, int32_t L, const amid_dropout* drop, const amid_encoder_saved* S,
                          const float* enc_out, const float* d_enc, amid_encoder_tensors* G, float* dx0,
                          void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* Split-operand variants (precision "x3", the parity-grade default): tcgen05 tensor cores at fp32-level accuracy.
 * Chain GEMMs multiply FP16 pair pieces (row-scaled token tile held in tensor memory x pre-swizzled weight images
 * fetched by bulk copy; products h0w0 + h1w0 + h0w1, relative error 2^-22), weight gradients multiply BF16 triples
 * (six products), attention runs 3xTF32.  Same arguments, tensors and tolerances as amid_encoder_fwd / _bwd
 * (model_seq.py:371-385, torch/nn/functional.py:5849-5856, 6630-6653). */
int amid_encoder_fwd_x3(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                        int32_t B, int32_t L, const amid_dropout* drop, amid_encoder_saved* S,
                        float* enc_out, void* workspace, int64_t workspace_bytes, amid_stream_t stream);
int amid_encoder_bwd_x3(const amid_encoder_tensors* P, const float* x0, const uint32_t* tmask,
                        int32_t B, int32_t L, const amid_dropout* drop, const amid_encoder_saved* S,
                        const float* enc_out, const float* d_enc, amid_encoder_tensors* G, float* dx0,
                        void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* ---- a6: InterComp / InnerComp in closed form (model_seq.py:474-497 / 450-472) ----- */
/* m[j] = max_{s,t} <a[j,s,:], b[j,t,:]>,  a,b: [B,n,128]  (the [bs,B,n,n] matmul+max of
 * model_seq.py:489-490 without its redundant outer axis). */
int amid_mim_scores(const float* a, const float* b, int32_t B, int32_t n, float* m, amid_stream_t stream);
/* the same on the tensor cores (mma.sync TF32 with the 3xTF32 split: fp32-accurate scores) */
int amid_mim_scores_tc(const float* a, const float* b, int32_t B, int32_t n, float* m, amid_stream_t stream);
/* p = softmax_j(m) over the (global) batch, g = 1[p > ts], coef[j] = w_bs[j]*g[j];
 * scal[0] = sum_j w_bs[j]; active list = indices with g=1 (ascending), n_active[0]. */
int amid_mim_gate(const float* m, const float* w_bs, int32_t Bglobal, float ts, float* p, float* gate, float* coef,
                  int32_t* active, int32_t* n_active, float* scal, amid_stream_t stream);
/* Ssum[n,128] = sum_{j active, j0<=j<j0+Blocal} coef[j]*other[j-j0]  (fixed ascending order). */
int amid_mim_aggregate(const float* other, const float* coef, const int32_t* active, const int32_t* n_active,
                       int32_t j0, int32_t Blocal, int32_t n, float* Ssum, amid_stream_t stream);
/* E[n,128] = Ssum W_nn^T + scal[0]*b_nn + b_bs ; esum[128] = sum_t E[t,:]. */
int amid_mim_project(const float* Ssum, const float* w_nn, const float* b_nn, const float* b_bs,
                     const float* scal, int32_t n, float* E, float* esum, amid_stream_t stream);
/* out[i] = cat(self[i], E) : [B,2n,128]  (only the InnerComp path materialises it). */
int amid_mim_concat(const float* self_, const float* E, int32_t B, int32_t n, float* out, amid_stream_t stream);
/* backward given dE [n,128]: dSsum = dE W_nn; dW_nn = dE^T Ssum; db_nn = scal[0]*colsum(dE);
 * db_bs = sum(dE); dw_bs[j] = g_j <dSsum, other[j]> + <colsum(dE), b_nn>;
 * d_other[j] (+)= coef[j]*dSsum for active j (accumulate != 0 adds, else rows of
 * inactive j are left untouched: the caller zero-fills or accumulates). */
int amid_mim_bwd(const float* dE, const float* Ssum, const float* other, const float* w_nn, const float* b_nn,
                 const float* coef, const float* gate, const int32_t* active, const int32_t* n_active,
                 const float* scal, int32_t j0, int32_t Blocal, int32_t n,
                 float* dW_nn, float* db_nn, float* db_bs, float* dw_bs /*[Blocal]*/, float* d_other,
                 float* ws_dS /*[n,128]*/, amid_stream_t stream);

/* ---- a7: mean pool (model_seq.py:432-434) ------------------------------------------ */
/* u[i,:] = (sum_t enc[i,t,:] + (esum ? esum[:] : 0)) / denom */
int amid_meanpool_fwd(const float* enc, const float* esum, int32_t B, int32_t n, float denom, float* u,
                      amid_stream_t stream);
/* d_enc[i,t,:] (+)= du[i,:]/denom ; dcol[:] = sum_i du[i,:]/denom (the dE row when ItC). */
int amid_meanpool_bwd(const float* du, int32_t B, int32_t n, float denom, int32_t accumulate, float* d_enc,
                      float* dcol, amid_stream_t stream);

/* ---- a8: predictModule (model_seq.py:32-54), up to 3 heads (isDR :436-440) --------- */
typedef struct {
    float* w0; /* fc.0.weight [hid,256] */
    float* b0; /* fc.0.bias   [hid] */
    float* w2; /* fc.2.weight [1,hid] */
    float* b2; /* fc.2.bias   [1] */
} amid_head_tensors;
/* probs[head][dom][b,c] = sigmoid(w2 . relu(W0 [u_dom[b] ; items[b,c]] + b0) + b2)
 * probs: [n_heads,2,B,C] contiguous. */
int amid_score_fwd(const float* u1, const float* u2, const float* items, const amid_head_tensors* heads,
                   int32_t n_heads, int32_t hid, int32_t B, int32_t C, float* probs, amid_stream_t stream);
int64_t amid_score_bwd_workspace_bytes(int32_t n_heads, int32_t hid, int32_t B, int32_t C);
/* dprobs: [n_heads,2,B,C] gradient w.r.t. the probabilities.  Outputs du1,du2 [B,128],
 * ditems [B,C,128], head gradients G[n_heads]. */
int amid_score_bwd(const float* u1, const float* u2, const float* items, const amid_head_tensors* heads,
                   int32_t n_heads, int32_t hid, int32_t B, int32_t C, const float* probs, const float* dprobs,
                   float* du1, float* du2, float* ditems, amid_head_tensors* G,
                   void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* ---- a9: losses (train_sr.py:205-212; train_sr_dr.py:212-221, 385-395) -------------- */
/* mode 0: loss_cls                      (train_sr.py:210-211)
 * mode 1: loss_cls + dr_e_w*loss_dr_e   (train_sr_dr.py:217-221)  needs 3 heads
 * mode 2: loss_dr_r                     (train_sr_dr.py:392-394)  needs 3 heads + ob_label
 * probs/dprobs [n_heads,2,B,C]; labels [B,C] fp32; domain_id, ob_label [B] int64.
 * inv_count = 1/(Bglobal*C) (the torch.mean divisor).  losses[0..2] = (cls, dr_e, dr_r)
 * means of the LOCAL batch rows scaled by inv_count.  BCE semantics = nn.BCELoss on
 * probabilities: log clamp at -100, backward divides by max(p(1-p), 1e-12). */
int amid_loss_fwd_bwd(const float* probs, int32_t n_heads, int32_t B, int32_t C, const float* labels,
                      const int64_t* domain_id, const int64_t* ob_label, int32_t mode, float dr_e_w,
                      float inv_count, float* losses, float* dprobs, amid_stream_t stream);

/* ---- a9: embedding backward = deterministic sort-by-index segmented reduction ------ */
int64_t amid_embgrad_workspace_bytes(int64_t n_rows);
/* ids [n_rows] int64, grad_rows [n_rows,128].  Produces the unique ids (ascending) in
 * uniq_ids, their summed gradient rows in uniq_grads [n_rows,128] (first n_uniq valid)
 * and n_uniq[0].  Rows with equal id are added in ascending source-row order. */
int amid_embgrad_segreduce(const int64_t* ids, const float* grad_rows, int64_t n_rows, int64_t V,
                           int64_t* uniq_ids, float* uniq_grads, int32_t* n_uniq,
                           void* workspace, int64_t workspace_bytes, amid_stream_t stream);
/* dense[uniq_ids[u],:] = uniq_grads[u,:]  (dense [V,128] must be zero-filled by the caller;
 * this is the drop-in path that feeds torch.optim.Adam a dense .grad like aten::embedding_dense_backward). */
int amid_embgrad_scatter_dense(const int64_t* uniq_ids, const float* uniq_grads, const int32_t* n_uniq,
                               int64_t max_rows, float* dense, int64_t V, amid_stream_t stream);

/* dense[ids[u] - id_offset,:] += rows[u,:] for u < n (ids unique within a call; ids outside [id_offset, id_offset + rows_dense)
 * are skipped).  Owner side of the sparse reduce-scatter of the table gradient under data parallelism (engine.py): the
 * reference has no counterpart -- it is the row-sharded form of the sum aten::embedding_dense_backward + NCCL would produce. */
int amid_embgrad_scatter_add(const int64_t* ids, const float* rows, int64_t n, int64_t id_offset, float* dense,
                             int64_t rows_dense, amid_stream_t stream);

/* ---- a9: Adam (torch.optim.Adam, betas (.9,.999), eps 1e-8, no weight decay) -------- */
/* dense: one fused pass over n elements; step = 1-based step number of this update. */
int amid_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                    float beta1, float beta2, float eps, amid_stream_t stream);
/* row-sparse with EXACT dense semantics: for each unique row, first replay the
 * zero-gradient steps last_step[row]+1 .. step-1 that a dense Adam would have applied,
 * then apply step `step` with the row gradient; last_step[row] = step. */
int amid_adam_rows_lazy(float* table, float* m, float* v, int32_t* last_step, const int64_t* uniq_ids,
                        const float* uniq_grads, const int32_t* n_uniq, int64_t max_rows, int32_t step,
                        float lr, float beta1, float beta2, float eps, amid_stream_t stream);
/* bring every row of the table up to `step` (before eval / state_dict / optimizer switch). */
int amid_adam_rows_flush(float* table, float* m, float* v, int32_t* last_step, int64_t V, int32_t step,
                         float lr, float beta1, float beta2, float eps, amid_stream_t stream);

/* ---- row-sharded table (BASELINE config 4): the lookup plan of one step, entirely on the device ------------------
 * owner = id mod G, owner-local row = id div G.  uniq_local [n]: owner-local row of every unique id of the step in bucket
 * order (by owner, ascending id inside) = the order of the step table; virtual_ids [n]: row of the step table for every
 * position; send_counts [G]: unique rows requested from each owner; flags [2] = {unique groups, out-of-range id seen}.
 * Replaces the torch.unique / argsort / bincount plan of round 1 (one sort + run-length encode + scan + 3 small kernels). */
int64_t amid_shard_plan_workspace_bytes(int64_t n);
int amid_shard_plan(const int64_t* ids, int64_t n, int64_t V, int32_t G, int64_t* uniq_local, int64_t* virtual_ids,
                    int32_t* flags, int64_t* send_counts, void* workspace, int64_t workspace_bytes, amid_stream_t stream);

/* ---- a10: eval ranking (utils.py:296-301, train_sr.py:114-115) ---------------------- */
/* scores [N,C], positive in column 0.  s0 = scores[r,0] - fix (fp32).  n_greater[r] =
 * #{c>=1 : scores[r,c] > s0}, n_equal[r] = #{c>=1 : scores[r,c] == s0}.  The rank of the
 * positive under argsort(argsort(-scores)) is n_greater when n_equal == 0. */
int amid_rank_counts(const float* scores, int64_t N, int32_t C, float fix, int32_t* n_greater,
                     int32_t* n_equal, amid_stream_t stream);

/* ---- a10 at catalogue scale (BASELINE config 5: every user against every pool item) ----
 * The reference scores 1 + neg_nums sampled candidates per user (dataset_seq.py:201, train_sr.py:56-128); the
 * full-catalogue variant scores the whole target-domain pool with the same predictModule (model_seq.py:40-54)
 * and the same ranking rule (utils.py:296-301).  hid must be 32.
 * item_proj:  Bc[i,:] = W0[:,128:] table[ids[i]] + b0      ([n_items,32]; ids NULL = rows 0..n_items-1)
 * user_proj:  A[b,dom,:] = W0[:,:128] u_dom[b]             ([B,2,32])
 * rank:       for slot s < n_users, user row r = user_rows[s], positive = Bc row pos_idx[r]:
 *             counts[s] = {#gt, #eq against s_pos ; #gt, #eq against s_pos - fix (fp32)} over Bc rows
 *             [i_lo, i_hi) except the positive itself; s_pos[s] = the positive's score.
 * scores:     the same scores written out, [n_users, i_hi - i_lo] (tie fallback and tests). */
int amid_catalogue_item_proj(const float* table, int64_t V, const int64_t* ids, int64_t n_items, const float* w0,
                             const float* b0, int32_t hid, float* Bc, amid_stream_t stream);
int amid_catalogue_user_proj(const float* u1, const float* u2, int32_t B, const float* w0, int32_t hid, float* A,
                             amid_stream_t stream);
int amid_catalogue_rank(const float* A, const int32_t* user_rows, int32_t n_users, int32_t dom, const float* Bc,
                        int32_t i_lo, int32_t i_hi, const int32_t* pos_idx, const float* w2, const float* b2, float fix,
                        int32_t* counts, float* s_pos, amid_stream_t stream);
int amid_catalogue_scores(const float* A, const int32_t* user_rows, int32_t n_users, int32_t dom, const float* Bc,
                          int32_t i_lo, int32_t i_hi, const int32_t* pos_idx, const float* w2, const float* b2,
                          float* s_pos, float* scores, amid_stream_t stream);

/* ---- a11: batch construction on the device (dataset_seq.py:177-236 __getitem__, :252-274 collate) ----
 * Histories are tokenised once into CSR arrays (amid_b200/pipeline.py): hist_dk = the history the encoder sees
 * (for the row's own domain: target and its earlier occurrences removed, :189-195), excl = the sorted unique items
 * of the row's full own-domain sequence (the set the negatives avoid, :188), pools = sorted item ids per domain
 * (:141-142, 151-158).  One launch writes a whole batch; ids are int64 throughout. */
typedef struct {
    const int64_t *hist_d1_vals, *hist_d1_offs;   /* offs: [n_rows+1] */
    const int64_t *hist_d2_vals, *hist_d2_offs;
    const int64_t *excl_vals, *excl_offs;
    const int64_t *target, *user;                 /* [n_rows] */
    const int32_t *domain, *overlap;              /* [n_rows] */
    const int64_t *pool_d1, *pool_d2;
    int64_t n_pool_d1, n_pool_d2, n_rows;
} amid_batch_source;
typedef struct {
    int64_t *seq_d1, *seq_d2;                     /* [B,L] last L items, left-padded with pad_id (:12-22) */
    int64_t *i_node, *user_node, *domain_id, *overlap_label, *long_tail_mask_d1, *long_tail_mask_d2;   /* [B] */
    int64_t *neg_samples;                         /* [B,K] or NULL (replay mode: the caller supplies negatives) */
} amid_batch_out;
/* rows[b] selects the dataset row of batch position b.  K > 0 with neg_samples != NULL draws K distinct negatives
 * per row from the target domain's pool minus the row's exclusion list (seeded, reproducible); an exhausted pool or
 * a row index out of range raises the gather error flag (amid_gather_error_host_sync). */
int amid_batch_build(const amid_batch_source* src, const int64_t* rows, int32_t B, int32_t L, int32_t K,
                     int32_t long_length, int64_t pad_id, uint64_t seed, const amid_batch_out* out, amid_stream_t stream);

/* ---- test support: the keep-mask a dropout site uses (for oracle mask injection) ---- */
/* feature site: out[r*128+c] for r<rows; attention site: out[((b*8+h)*L+i)*L+j]. */
int amid_dropout_mask_feature(const amid_dropout* drop, uint32_t site, int64_t rows, uint8_t* out, amid_stream_t stream);
int amid_dropout_mask_attn(const amid_dropout* drop, uint32_t site, int32_t B, int32_t L, uint8_t* out, amid_stream_t stream);

/* ---- tcgen05 bring-up / unit-test entry points (TF32 operands, fp32 accumulate in TMEM) ------ */
/* y[M,128] = x[M,128] w[128,128]^T + b   (what nn.Linear / Conv1d(k=1) compute on the path) */
int amid_tc_linear_test(const float* x, const float* w, const float* b, int32_t M, float* y, amid_stream_t stream);

/* BF16-operand variants of the bring-up kernels; the second one accumulates dy^T x per CTA in TMEM
 * using MN-major operand views of row-major token tiles (sum over cta of part = dy^T x). */
int amid_tc_linear16_test(const float* x, const float* w, const float* b, int32_t M, float* y, amid_stream_t stream);
int amid_tc_wgrad16_test(const float* dy, const float* x, int32_t M, float* part, int32_t n_ctas, amid_stream_t stream);

/* Split-operand ("x3") bring-up: the same linear layer at fp32-level accuracy from FP16 pair pieces (token tile in
 * tensor memory, pre-swizzled weight image fetched by one bulk copy; scratch >= 65,540 bytes), and the weight-gradient
 * kernel with BF16 triples: sum over cta of wpart[cta] = dy^T x, sum over cta of bpart[cta] = column sums of dy. */
int amid_x3_linear_test(const float* x, const float* w, const float* b, int32_t M, float* y, void* scratch,
                        amid_stream_t stream);
/* Attention kernels side by side (q, k, v, o: [B*L,128] with heads along the columns; lse: [B*8*L]).
 * impl 0 = fp32 CUDA cores, 1 = mma.sync TF32, 2 = mma.sync 3xTF32, 3 = tcgen05 FP16-pair split, first version (L <= 224),
 * 4 = round-2 tcgen05 kernels of attn_p.cuh (forward k_attn_fwd_p, 64 <= L <= 224; backward: single-pass persistent
 * k_attn_bwd_p, 64 <= L <= 256), 5 = backward only: two-pass k_attn_bwd_t2 (L <= 224; the kernel the x3 train step runs). */
int amid_attn_fwd_test(const float* q, const float* k, const float* v, float* o, float* lse, int32_t B, int32_t L,
                       const amid_dropout* drop, uint32_t site, int32_t impl, amid_stream_t stream);
/* backward: dq is the gradient with respect to q / 0.25 (the convention of amid_encoder_bwd's chain kernels) */
int amid_attn_bwd_test(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dO,
                       float* dq, float* dk, float* dv, int32_t B, int32_t L, const amid_dropout* drop, uint32_t site,
                       int32_t impl, amid_stream_t stream);
int amid_x3_wgrad_test(const float* dy, const float* x, int32_t M, float* wpart, float* bpart, int32_t n_ctas,
                       amid_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AMID_B200_H */
