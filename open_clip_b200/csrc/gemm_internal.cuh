// Internal (non-ABI) declarations shared by gemm.cu and loss.cu.
#pragma once
#include "common.cuh"

namespace clipn {

constexpr int kMaxBMaps = 8;

struct alignas(64) TmapSet {
  CUtensorMap a;
  CUtensorMap b[kMaxBMaps];
  CUtensorMap c, c2, aux;  // epilogue staging maps (64-col x 32-row bf16 boxes)
};

struct GemmParams {
  int m, n, k;
  int tiles_m, tiles_n, kblocks, splits;
  int a_mn, b_mn;
  int b_maps, b_rows_per_map;
  void* c; int64_t ldc;
  void* c2; int64_t ldc2;
  const void* bias;
  const void* aux; int64_t ldaux;
  float alpha;
  const float* row_lse; const float* col_lse;
  float* part_max; float* part_sum; float* pos; float* scalar_acc;
  float logit_bias, gscale, col_w;
  int label_offset, negative_only;
  const float* alpha_dev; const float* logit_bias_dev;
  float* col_sum;
};

struct RefOperands {
  const void* a; int64_t lda;
  const void* b[kMaxBMaps]; int64_t ldb;
};

// b_ptrs: `b_maps` base pointers, each covering `b_rows_per_map` rows of the B operand's OUTER (row)
// dimension (N rows when K-major, K rows when MN-major). b_maps == 1 -> d.b semantics.
int gemm_launch(const clipn_gemm_desc& d, const void* const* b_ptrs, int b_maps, int64_t b_rows_per_map, bool use_ref,
                cudaStream_t stream);
int gemm_tile_n(int n);

}  // namespace clipn
