/*
 * clipn.h — C ABI of libclipn.so: the B200 (sm_100a) CLIP train-step kernels.
 *
 * This is the drop-in boundary for the reference's hot path.  Every entry point replaces one
 * (group of) ATen/cuBLAS/NCCL call(s) the reference makes; the citation on each function is
 * `file:line` under /root/reference/src/open_clip.  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add and where each call slots into model.py / loss.py.
 *
 * Conventions
 *   - Plain C: raw device pointers, sizes, a CUDA stream handle.  No torch types.
 *   - The CALLER owns all memory (outputs, workspaces); the library never allocates or frees
 *     device memory and never synchronises the host with the device.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 on success, negative on error; clipn_last_error() returns a message
 *     (thread-local).  No exceptions cross the ABI.
 *   - Activations are bf16 row-major; statistics / reductions / master gradients are fp32.
 *   - Re-entrant: no global mutable state besides a read-only device-property cache.
 */
#ifndef CLIPN_H_
#define CLIPN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* clipn_stream_t; /* cudaStream_t */

#define CLIPN_OK 0
#define CLIPN_ERR_ARG (-1)
#define CLIPN_ERR_CUDA (-2)

/* ---- library ------------------------------------------------------------------------------ */
int clipn_version(void);
const char* clipn_last_error(void);
/* sm count / compute capability of the current device */
int clipn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- GEMM family (tcgen05 + TMEM + TMA) ----------------------------------------------------
 * C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *   A: a_mn_major == 0 -> stored [M,K] row-major (lda = row pitch in elements, K contiguous)
 *      a_mn_major == 1 -> stored [K,M] row-major (M contiguous)           (weight-gradient form)
 *   B: b_mn_major == 0 -> stored [N,K] row-major (torch Linear weight)    (F.linear form)
 *      b_mn_major == 1 -> stored [K,N] row-major                          (x @ W form)
 * Replaces: F.linear x3 on in_proj chunks transformer.py:195-197, out_proj :246, mlp c_fc/c_proj
 * :295-299, conv1-as-GEMM :794, `pooled @ proj` :923, `x @ text_projection` model.py:409, and their
 * autograd dgrad/wgrad (cuBLASLt in the reference).
 */
enum clipn_epilogue {
  CLIPN_EPI_STORE = 0,      /* C(bf16) = alpha*acc (+ bias[n])                                         */
  CLIPN_EPI_BIAS_GELU = 1,  /* t = bf16(acc + bias); C = t (pre-activation, saved for backward);
                               C2 = bf16(gelu_erf(t))           (c_fc + nn.GELU, transformer.py:295-299) */
  CLIPN_EPI_BIAS_RESID = 2, /* C = bf16( bf16(acc + bias) + aux[m,n] )   (out_proj/c_proj + residual
                               add, transformer.py:328-329)                                               */
  CLIPN_EPI_DGELU = 3,      /* h = aux[m,n]; C = bf16(acc * gelu'(h)); optional C2 = bf16(gelu(h))  (GELU bwd
                               fused into the c_proj dgrad; C2, when given, re-materialises the c_proj wgrad
                               operand so the forward need not keep it)                                   */
  CLIPN_EPI_ACCUM_F32 = 4,  /* C(f32)[m,n] += alpha*acc   (split-K weight gradient, red.global.add)      */
  CLIPN_EPI_STORE_F32 = 5,  /* C(f32) = alpha*acc (+ bias)                                               */
  CLIPN_EPI_LSE = 6,        /* no C.  Online row log-sum-exp partials of alpha*acc (+logit_bias):
                               part_max/part_sum[(n_tile*2+half)*M + m], and the label logit
                               (column == m + label_offset) into pos[m]  (ClipLoss fwd, loss.py:102-139)   */
  CLIPN_EPI_CLIP_DLOGITS = 7, /* C(bf16)[m,n] = gscale*( exp(s-row_lse[m]) + col_w*exp(s-col_lse[n]) - (1+col_w)/N ),
                               s = alpha*acc (+logit_bias): the softmax parts of d loss / d logits, centred on their
                               mean.  The caller restores the mean and the one-hot part
                               -(1+col_w)*gscale*[n == m+label_offset] in fp32 (bf16 would lose 2^-9 of values
                               ~1/N resp. ~2, which dominates the gradient while the features are nearly parallel).
                               Also accumulates scalar_acc[0] += gscale * sum (P_row - onehot) * acc
                               (d loss / d logit_scale)                                                    */
  CLIPN_EPI_SIGLIP = 8,     /* softplus / sigmoid epilogue for SigLipLoss (loss.py:351-367): see
                               clipn_siglip_* below                                                       */
  CLIPN_EPI_BIAS_GELU_GRAD = 9, /* t = bf16(acc + bias); C = bf16(gelu'(t)); C2 = bf16(gelu_erf(t)): the forward
                               keeps the GELU derivative instead of the pre-activation, so the backward epilogue
                               is a plain multiply (CLIPN_EPI_MUL_AUX) instead of erf + exp per element          */
  CLIPN_EPI_MUL_AUX = 10,   /* C = bf16(acc * aux[m,n])   (c_proj dgrad x saved gelu'(h); optional col_sum)      */
};

typedef struct clipn_gemm_desc {
  const void* a; int64_t lda; int32_t a_mn_major;
  const void* b; int64_t ldb; int32_t b_mn_major;
  void* c; int64_t ldc;
  void* c2; int64_t ldc2;
  const void* bias;              /* bf16 [N] or NULL */
  const void* aux; int64_t ldaux; /* bf16 [M,N] residual / pre-activation, or NULL */
  int32_t m, n, k;
  int32_t epilogue;              /* enum clipn_epilogue */
  float alpha;
  int32_t splits;                /* split-K factor, >=1 (only with CLIPN_EPI_ACCUM_F32) */
  /* loss epilogues (LSE / CLIP_DLOGITS / SIGLIP); ignored otherwise */
  const float* row_lse;          /* [M] */
  const float* col_lse;          /* [N] */
  float* part_max; float* part_sum; /* [2*ceil(N/BN) * M] each; BN from clipn_gemm_tile_n() */
  float* pos;                    /* [M] label logit (LSE: written); CLIP_DLOGITS: optional row centre (read) */
  float* scalar_acc;             /* [2] fp32 accumulators: d logit_scale (raw, pre-chain), d logit_bias */
  float logit_bias;              /* added to alpha*acc before exp (0 if none) */
  float gscale;                  /* gradient scale, e.g. 1/(2B) */
  float col_w;                   /* weight of the column-softmax term (1 with gather_with_grad, 0 w/o) */
  int32_t label_offset;          /* rank*B for local_loss (loss.py:82-83) */
  int32_t negative_only;         /* SIGLIP: no positives in this block (loss.py:344-348) */
  /* optional DEVICE scalars (fp32), so callers never sync to read logit_scale / logit_bias on the host:
     effective alpha = alpha * (*alpha_dev), effective logit_bias = logit_bias + (*logit_bias_dev) */
  const float* alpha_dev;
  const float* logit_bias_dev;
  /* optional fp32 [N] accumulator: col_sum[n] += sum_m C[m,n] (bias gradient fused into the epilogue that
     produces C; CLIPN_EPI_STORE and CLIPN_EPI_DGELU) */
  float* col_sum;
} clipn_gemm_desc;

int clipn_gemm(const clipn_gemm_desc* d, clipn_stream_t stream);
/* CUDA-core restatement of the same contract (same epilogues) used only by the tests to bisect
 * the tensor-core path. Never on the product path. */
int clipn_gemm_ref(const clipn_gemm_desc* d, clipn_stream_t stream);
/* N-tile width the tensor-core kernel will use for this N (sizes the LSE partial buffers). */
int clipn_gemm_tile_n(int n);

/* ---- LayerNorm (layers.py:11-26, eps 1e-5; fp32 statistics, bf16 in/out) --------------------- */
int clipn_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                        int64_t rows, int32_t d, float eps, clipn_stream_t stream);
/* dx_out = (dx_resid ? dx_resid : 0) + LN'(dy); dgamma/dbeta are fp32 accumulators (+=).
 * dresid_sum (optional, fp32 [d], +=): column sums of dx_resid — the bias gradient of the Linear that fed the residual
 * stream (out_proj.bias / c_proj.bias, transformer.py:328-329), fused into the pass that streams dx_resid anyway. */
int clipn_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        const void* dx_resid, void* dx_out, float* dgamma, float* dbeta, float* dresid_sum, int64_t rows,
                        int32_t d, clipn_stream_t stream);

/* ---- attention core (F.scaled_dot_product_attention, transformer.py:223-228) ------------------
 * qkv: bf16 [B*L, 3*H*64] (q | k | v, head-major inside each third, as produced by the QKV GEMM);
 * out: bf16 [B*L, H*64] (heads merged, transformer.py:244); lse: fp32 [B,H,L].
 * causal != 0 reproduces the additive -inf upper-triangular mask (transformer.py:1716-1722). */
int clipn_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t seq, int32_t heads,
                        int32_t causal, float scale, clipn_stream_t stream);
/* dbias (optional, fp32 [3*H*64], +=): column sums of dqkv == gradient of in_proj_bias, fused into the kernel.
 * Sequence limits: L <= 640 (forward and backward). L <= 384 runs one kernel with Q,K,V,dO resident in shared
 * memory and does not read `out`; 384 < L <= 640 (ViT-L/14-336: 577) runs a dQ pass with K,V resident and a
 * dK/dV pass with Q,dO resident, both taking D = rowsum(dO o O) from `out`. */
int clipn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                        float* dbias, int32_t batch, int32_t seq, int32_t heads, int32_t causal, float scale,
                        clipn_stream_t stream);

/* ---- embeddings / pooling / normalize ----------------------------------------------------------- */
/* conv1 (kernel = stride = patch, no bias; transformer.py:632-638,794-796) as im2row: image NCHW bf16
 * [B,3,H,W] -> patches bf16 [B*gh*gw, 3*P*P], column order (c, py, px) == conv1.weight.flatten(1). */
int clipn_patchify(const void* image, void* patches, int32_t batch, int32_t chans, int32_t height, int32_t width,
                   int32_t patch, clipn_stream_t stream);
/* Same im2row for any even patch size (ViT-L/14: 3*14*14 = 588 columns, whose rows are not 16-byte multiples):
 * rows are written with pitch ld_patches (multiple of 8, >= chans*patch*patch) and zero-filled beyond the data,
 * so the conv1 GEMM runs with K = ld_patches. With height == width == patch it pads a [d,3,P,P] weight the same
 * way (one "patch" per output channel). */
int clipn_patchify_padded(const void* image, void* patches, int64_t ld_patches, int32_t batch, int32_t chans,
                          int32_t height, int32_t width, int32_t patch, clipn_stream_t stream);
/* dst[r, 0:cols] += src[r, 0:cols] (fp32): folds a K-padded conv1 weight gradient back into the parameter layout */
int clipn_accum_rows_f32(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int32_t cols,
                         clipn_stream_t stream);
/* x[b,0,:] = bf16(bf16(cls)+bf16(pos[0])); x[b,1+p,:] = bf16(patch_out[b*np+p] + bf16(pos[1+p]))
 * (transformer.py:799-801) */
int clipn_vision_embed_fwd(const void* patch_out, const float* cls, const float* pos, void* x, int32_t batch,
                           int32_t npatch, int32_t d, clipn_stream_t stream);
/* dpatch_out (bf16) = dx[:,1:,:]; dcls (f32, +=) = sum_b dx[b,0]; dpos (f32, +=) = sum_b dx[b] */
int clipn_vision_embed_bwd(const void* dx, void* dpatch_out, float* dcls, float* dpos, int32_t batch, int32_t npatch,
                           int32_t d, clipn_stream_t stream);
/* x = bf16(bf16(table[ids]) + bf16(pos))   (model.py:399-401); also eot_idx[b] = argmax_l ids[b,l]
 * (transformer.py:941-944; first maximal index like torch.argmax) when eot_idx != NULL */
int clipn_text_embed_fwd(const int64_t* ids, const float* table, const float* pos, void* x, int32_t* eot_idx,
                         int32_t batch, int32_t seq, int32_t d, int32_t vocab, clipn_stream_t stream);
/* dtable (f32, +=) scatter-add of dx rows; dpos (f32, +=) = sum_b dx[b] */
int clipn_text_embed_bwd(const int64_t* ids, const void* dx, float* dtable, float* dpos, int32_t batch, int32_t seq,
                         int32_t d, int32_t vocab, clipn_stream_t stream);
/* out[b,:] = x[b*seq + idx[b], :] (idx == NULL -> row 0: CLS pooling transformer.py:787) */
int clipn_gather_rows(const void* x, const int32_t* idx, void* out, int32_t batch, int32_t seq, int32_t d,
                      clipn_stream_t stream);
/* dx (bf16 [B*seq, d]) = 0 except dx[b*seq + idx[b]] = dpooled[b] */
int clipn_scatter_rows(const void* dpooled, const int32_t* idx, void* dx, int32_t batch, int32_t seq, int32_t d,
                       clipn_stream_t stream);
/* F.normalize(x, dim=-1) (model.py:391,411): y = x / max(||x||, 1e-12); inv_norm saved (fp32 [rows]) */
int clipn_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int32_t d, clipn_stream_t stream);
/* dx = inv_norm * (dy - y * <dy, y>) ; dy fp32 or bf16 by dy_is_f32 */
int clipn_l2norm_bwd(const void* dy, int32_t dy_is_f32, const void* y, const float* inv_norm, void* dx, int64_t rows,
                     int32_t d, clipn_stream_t stream);

/* ---- small reductions / casts ------------------------------------------------------------------ */
/* out (f32 [n], +=) = column sums of x (bf16 [rows, n])  — bias gradients */
int clipn_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int32_t n, clipn_stream_t stream);
int clipn_cast_f32_to_bf16(const float* x, void* y, int64_t n, clipn_stream_t stream);

/* ---- optimizer ---------------------------------------------------------------------------------------
 * Multi-tensor AdamW (decoupled weight decay), one launch for the whole model: replaces torch.optim.AdamW as built by
 * open_clip_train/optim.py:453-454 (`optimizer.step()`, open_clip_train/train.py:182).  Arithmetic of torch's fused
 * AdamW: p -= lr*wd*p; m = lerp(m, g, 1-beta1); v = beta2*v + (1-beta2)*g*g;
 * p -= (lr / bias_correction1) * m / (sqrt(v) / bias_correction2_sqrt + eps), all in fp32, stored back in each tensor's
 * dtype (bf16 parameters carry bf16 gradients and bf16 moments, as under --precision bf16).  The caller passes the
 * bias corrections of the current step: 1 - beta1^t and sqrt(1 - beta2^t). */
#define CLIPN_ADAMW_MAX_TENSORS 512 /* per launch; longer lists are split */
typedef struct clipn_adamw_tensor {
  void* param; const void* grad; void* exp_avg; void* exp_avg_sq; /* same dtype, same numel */
  int64_t numel;
  float lr, weight_decay;
  int32_t is_bf16; /* 1: bf16 tensors, 0: fp32 tensors */
} clipn_adamw_tensor;
int clipn_adamw_multi(const clipn_adamw_tensor* tensors, int32_t n, float beta1, float beta2, float eps,
                      float bias_correction1, float bias_correction2_sqrt, clipn_stream_t stream);

/* ---- contrastive losses --------------------------------------------------------------------------
 * ClipLoss (loss.py:57-141) in its local_loss form — this rank's B rows against all N = W*B columns, both
 * directions (logits_per_image rows, logits_per_text rows; loss.py:102-104) — and SigLipLoss (loss.py:314-489).
 *
 * FUSED FORWARD.  `txt_cols` / `img_cols` list W device pointers, each bf16 [B,E]: rank r's feature buffer,
 * peer-mapped into this process (CUDA IPC / symmetric memory) for r != rank.  The call reads every peer's buffer
 * directly over NVLink — no NCCL: a P2P gather kernel (coalesced 16-byte peer loads from all SMs, each byte crossing
 * NVLink once) fills the local copies gather_txt / gather_img (bf16 [N,E], required when W > 1; the backward reads them),
 * then ONE tcgen05 launch computes both directions with the column tile stationary in shared memory and the online
 * log-sum-exp in the epilogue (logits never materialised):
 *   lse[0*B + m] = logsumexp_n( s * img[m] . txt_all[n] )      lse[1*B + m] = logsumexp_n( s * txt[m] . img_all[n] )
 *   loss_acc[0] += ( sum_m lse_img[m] - pos_img[m]  +  sum_m lse_txt[m] - pos_txt[m] ) / (2B)      (loss.py:135-139)
 * with s = scale * (*scale_dev) and pos = the label logit (column rank*B + m, loss.py:82-83).
 * (CLIPN_PEER_DIRECT=1: the GEMM's TMA producer pulls the column tiles straight from the peers and the gathered copy is
 * a by-product — kept for experiments; per-SM TMA peer reads are latency-bound, see csrc/loss.cu.)
 * Shapes: E % 64 == 0, E <= 1024, W <= 8, W*B % 8 == 0; clipn_peer_gemm_tile_n(1, W*B, E) returns 0 for shapes the
 * kernel does not take (use the generic calls below on operands gathered by the caller).
 * workspace: fp32, clipn_clip_fwd_fused_workspace(W, B, E) elements. */
int32_t clipn_peer_gemm_tile_n(int32_t world, int32_t b, int32_t e);
/* The gather step alone: every rank's [B,E] block (peer-mapped pointers) -> local gather_txt / gather_img [W*B,E]. */
int clipn_peer_gather(const void* const* txt_cols, const void* const* img_cols, int32_t world, int32_t b, int32_t e,
                      void* gather_txt, void* gather_img, clipn_stream_t stream);
int64_t clipn_clip_fwd_fused_workspace(int32_t world, int32_t b, int32_t e);
/* Measurement aid (bench.py): with clipn_stage_timing(1), every clipn_clip_fwd_fused call records CUDA events on its
 * stream between the peer gather, the GEMM and the combine kernel (up to 64 calls); clipn_stage_times writes the mean
 * milliseconds of the three stages to out[3] and returns the number of calls averaged.  Off by default. */
int clipn_stage_timing(int32_t enable);
int32_t clipn_stage_times(float* out);
int clipn_clip_fwd_fused(const void* img_rows, const void* txt_rows, const void* const* txt_cols,
                         const void* const* img_cols, int32_t world, int32_t rank, int32_t b, int32_t e, float scale,
                         const float* scale_dev, void* gather_txt, void* gather_img, float* lse, float* loss_acc,
                         float* workspace, clipn_stream_t stream);
/* SigLipLoss forward (+ d(logits) when dl_img / dl_txt are given), same operand convention.  Direction 0 (image rows
 * x all text columns) accumulates the loss value (loss.py:351-367: every other rank's text block is a negative_only
 * block, loss.py:410-487), d scale and d bias (scalar_acc[0], [1]) and writes dl_img = gscale * d loss / d logits
 * (bf16 [B, ld]); direction 1 (text rows x all image columns) only writes dl_txt.  Because a SigLIP logit's gradient
 * depends on nothing but the pair itself, rank r obtains d loss / d txt_r from ITS OWN text rows against the gathered
 * image columns: the reverse neighbour exchange of the reference (loss.py:287-307) carries no data here. */
int clipn_siglip_fwd_fused(const void* img_rows, const void* txt_rows, const void* const* txt_cols,
                           const void* const* img_cols, int32_t world, int32_t rank, int32_t b, int32_t e,
                           const float* scale_dev, const float* bias_dev, float gscale, void* gather_txt,
                           void* gather_img, float* loss_acc, float* scalar_acc, void* dl_img, void* dl_txt, int64_t ld,
                           clipn_stream_t stream);

/* GENERIC pieces (one local column operand `feats_cols` bf16 [n,E]; m local rows): the backward of both losses and
 * the forward for shapes outside the fused kernel's envelope.
 *   lse[m] = logsumexp_n( scale * rows[m] . cols[n] ),  pos[m] = scale * rows[m] . cols[label_offset+m]
 * workspace: fp32, at least clipn_clip_lse_workspace(m, n) elements.  n % 8 == 0. */
int64_t clipn_clip_lse_workspace(int32_t m, int32_t n);
int clipn_clip_lse_fwd(const void* feats_rows, const void* feats_cols, int32_t m, int32_t n, int32_t e, float scale,
                       const float* scale_dev, int32_t label_offset, float* lse, float* pos, float* workspace,
                       clipn_stream_t stream);
/* dlogits (bf16 [m, ld]) for one direction, mean-centred and WITHOUT the one-hot term, see CLIPN_EPI_CLIP_DLOGITS;
 * col_lse is the OTHER direction's global LSE vector [n] (all ranks),
 * scalar_acc[0] += sum_{m,n} (P_row - onehot) * (s / scale - row_centre[m]) * gscale  — the d logit_scale sum.
 * row_centre (optional fp32 [m]): any value close to rows[m] . cols[label_offset + m]; it cancels exactly
 * (sum_n P_row = 1) and keeps the accumulated terms small while all features are still nearly parallel. */
int clipn_clip_dlogits(const void* feats_rows, const void* feats_cols, int32_t m, int32_t n, int32_t e, float scale,
                       const float* scale_dev, int32_t label_offset, const float* row_lse, const float* col_lse,
                       float col_w, float gscale, void* dlogits, int64_t ld, float* scalar_acc, const float* row_centre,
                       clipn_stream_t stream);
/* d_rows (fp32 [m,E]) += alpha * dlogits[m,n] @ feats_cols[n,E]; split-K over n in `splits` chunks (TMA reduce-add)
 * so that the [m x E] output fills the machine.  The caller initialises d_rows: zeros, or for ClipLoss the fp32 part
 * (1+col_w) * gscale * alpha * (mean_n feats_cols[n] - feats_cols[label_offset + m]). */
int clipn_clip_dfeat(const void* dlogits, int64_t ld, const void* feats_cols, int32_t m, int32_t n, int32_t e,
                     float alpha, const float* alpha_dev, float* d_rows, int32_t splits, clipn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPN_H_ */
