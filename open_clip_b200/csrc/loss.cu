// Contrastive-loss entry points (ClipLoss loss.py:57-141, SigLipLoss loss.py:314-489) on top of the tcgen05
// GEMM: the logits are never materialised in the forward (online log-sum-exp in the GEMM epilogue), and the
// column operand is read tile-by-tile straight from every rank's (peer-mapped) feature buffer through one TMA
// tensor map per rank — the all-gather of gather_features (loss.py:29-54) is fused into the GEMM.
#include <math.h>

#include "common.cuh"
#include "gemm_internal.cuh"

namespace clipn {

// combine per-slab (max, sum) partials into the row LSE: lse[m] = log sum_n exp(s[m,n])
__global__ void __launch_bounds__(256) lse_combine_kernel(const float* __restrict__ part_max,
                                                          const float* __restrict__ part_sum, float* __restrict__ lse,
                                                          int m, int slabs) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  float mx = -INFINITY;
  for (int s = 0; s < slabs; ++s) mx = fmaxf(mx, part_max[static_cast<int64_t>(s) * m + row]);
  float sum = 0.f;
  for (int s = 0; s < slabs; ++s) {
    const float pm = part_max[static_cast<int64_t>(s) * m + row];
    if (pm != -INFINITY) sum += part_sum[static_cast<int64_t>(s) * m + row] * __expf(pm - mx);
  }
  lse[row] = mx + __logf(sum);
}

static void base_desc(clipn_gemm_desc& d, const void* rows, int b, int n, int e, float scale,
                      const float* scale_dev) {
  memset(&d, 0, sizeof(d));
  d.a = rows; d.lda = e; d.a_mn_major = 0;
  d.ldb = e; d.b_mn_major = 0;
  d.m = b; d.n = n; d.k = e;
  d.alpha = scale; d.alpha_dev = scale_dev; d.splits = 1;
}

}  // namespace clipn

using namespace clipn;

extern "C" int64_t clipn_clip_lse_workspace(int32_t b, int32_t n) {
  const int bn = gemm_tile_n(n);
  const int64_t slabs = 2 * static_cast<int64_t>((n + bn - 1) / bn);
  return 2 * slabs * b;
}

extern "C" int clipn_clip_lse_fwd(const void* feats_rows, const void* const* feats_cols, int32_t world, int32_t b,
                                  int32_t e, float scale, const float* scale_dev, int32_t label_offset, float* lse,
                                  float* pos, float* workspace, clipn_stream_t stream) {
  CLIPN_REQUIRE(feats_rows && feats_cols && lse && pos && workspace, "clip_lse_fwd: null pointer");
  CLIPN_REQUIRE(world >= 1 && world <= kMaxBMaps, "clip_lse_fwd: world must be 1..8");
  const int n = world * b;
  const int bn = gemm_tile_n(n);
  const int slabs = 2 * ((n + bn - 1) / bn);
  clipn_gemm_desc d;
  base_desc(d, feats_rows, b, n, e, scale, scale_dev);
  d.b = feats_cols[0];
  d.epilogue = CLIPN_EPI_LSE;
  d.part_max = workspace;
  d.part_sum = workspace + static_cast<int64_t>(slabs) * b;
  d.pos = pos;
  d.label_offset = label_offset;
  int rc = gemm_launch(d, feats_cols, world, b, false, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  lse_combine_kernel<<<(b + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d.part_max, d.part_sum, lse, b, slabs);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_clip_dlogits(const void* feats_rows, const void* const* feats_cols, int32_t world, int32_t b,
                                  int32_t e, float scale, const float* scale_dev, int32_t label_offset,
                                  const float* row_lse, const float* col_lse, float col_w, float gscale, void* dlogits, float* scalar_acc,
                                  clipn_stream_t stream) {
  CLIPN_REQUIRE(feats_rows && feats_cols && row_lse && dlogits, "clip_dlogits: null pointer");
  CLIPN_REQUIRE(world >= 1 && world <= kMaxBMaps, "clip_dlogits: world must be 1..8");
  const int n = world * b;
  clipn_gemm_desc d;
  base_desc(d, feats_rows, b, n, e, scale, scale_dev);
  d.b = feats_cols[0];
  d.epilogue = CLIPN_EPI_CLIP_DLOGITS;
  d.c = dlogits; d.ldc = n;
  d.row_lse = row_lse; d.col_lse = col_lse; d.col_w = col_w; d.gscale = gscale;
  d.scalar_acc = scalar_acc;
  d.label_offset = label_offset;
  return gemm_launch(d, feats_cols, world, b, false, static_cast<cudaStream_t>(stream));
}

extern "C" int clipn_clip_dfeat(const void* dlogits, const void* const* feats_cols, int32_t world, int32_t b, int32_t e,
                                float alpha, const float* alpha_dev, void* d_rows, int32_t out_is_f32,
                                clipn_stream_t stream) {
  CLIPN_REQUIRE(dlogits && feats_cols && d_rows, "clip_dfeat: null pointer");
  CLIPN_REQUIRE(world >= 1 && world <= kMaxBMaps, "clip_dfeat: world must be 1..8");
  const int n = world * b;
  clipn_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = dlogits; d.lda = n; d.a_mn_major = 0;       // [B, N], reduction over N
  d.b = feats_cols[0]; d.ldb = e; d.b_mn_major = 1;  // each rank: [B(K rows), E] -> MN-major
  d.c = d_rows; d.ldc = e;
  d.m = b; d.n = e; d.k = n;
  d.alpha = alpha; d.alpha_dev = alpha_dev; d.splits = 1;
  d.epilogue = out_is_f32 ? CLIPN_EPI_STORE_F32 : CLIPN_EPI_STORE;
  return gemm_launch(d, feats_cols, world, b, false, static_cast<cudaStream_t>(stream));
}

extern "C" int clipn_siglip_block(const void* img, const void* txt, int32_t b, int32_t e, const float* scale_dev,
                                  const float* bias_dev, int32_t negative_only, float gscale, float* loss_acc,
                                  void* dlogits, float* scalar_acc, clipn_stream_t stream) {
  CLIPN_REQUIRE(img && txt && loss_acc && scale_dev, "siglip_block: null pointer");
  clipn_gemm_desc d;
  base_desc(d, img, b, b, e, 1.0f, scale_dev);
  d.b = txt;
  d.epilogue = CLIPN_EPI_SIGLIP;
  d.c = dlogits; d.ldc = b;
  d.logit_bias = 0.f; d.logit_bias_dev = bias_dev;
  d.negative_only = negative_only;
  d.gscale = gscale;
  d.part_sum = loss_acc;
  d.scalar_acc = scalar_acc;
  const void* bp[1] = {txt};
  return gemm_launch(d, bp, 1, 0, false, static_cast<cudaStream_t>(stream));
}
