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

// Peer-streaming logits GEMM (gemm_peer.cuh): up to two directions in one launch, column operand given as one
// pointer per rank (peer-mapped), stationary in shared memory; optional local gathered copy as a by-product.
struct PeerGemmDesc {
  const void* rows[2];         // [m, e] bf16 local row operand of each direction
  const void* const* cols[2];  // per direction: `world` device pointers, each [rows_per_map, e] bf16
  void* gather[2];             // per direction: local [world * rows_per_map, e] bf16 copy of the column operand, or null
  int dirs, world, rank, m, rows_per_map, e;
  int epilogue;                // CLIPN_EPI_LSE or CLIPN_EPI_SIGLIP
  float alpha; const float* alpha_dev;
  float logit_bias; const float* logit_bias_dev;
  float gscale; int label_offset, negative_only;
  float* part_max[2]; float* part_sum[2]; float* pos[2];
  void* c[2]; int64_t ldc; float* scalar_acc[2];
};
int peer_gemm_launch(const PeerGemmDesc& d, cudaStream_t stream);
// column-tile width the peer kernel uses for embed dim e and this problem, or 0 if the shape is not supported
int peer_gemm_tile_n(int world, int rows_per_map, int e);

}  // namespace clipn
