// Contrastive-loss entry points (ClipLoss loss.py:57-141, SigLipLoss loss.py:314-489) on top of the tcgen05 GEMMs.
//
// Forward (fused, clipn_clip_fwd_fused / clipn_siglip_fwd_fused): ONE launch of the peer-streaming kernel
// (gemm_peer.cuh) computes both directions' logits against every rank's features, read tile by tile straight from
// the owning rank's peer-mapped buffer (the all-gather of gather_features loss.py:29-54 is fused into the GEMM; the
// logits are never materialised: online log-sum-exp / softplus in the epilogue) and leaves a local copy of the
// gathered operands behind.  Backward: d(logits) tiles are recomputed from the LSE vectors against the LOCAL
// gathered copy (generic kernel, one tensor map), written once in bf16 and contracted by a split-K GEMM.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_internal.cuh"

namespace clipn {

// combine per-slab (max, sum) partials into the row LSE: lse[m] = log sum_n exp(s[m,n]); optionally accumulate
// loss_acc += loss_scale * sum_m (lse[m] - pos[m])   (the cross-entropy value of this direction)
// Block = 32 consecutive rows (lanes: coalesced [slab][row] reads) x 8 slab subsets (warps); the subsets are merged
// through shared memory.  (One thread per row walking all slabs serially took 22 us at 64 slabs — latency-bound.)
__global__ void __launch_bounds__(256) lse_combine_kernel(const float* __restrict__ part_max,
                                                          const float* __restrict__ part_sum,
                                                          const float* __restrict__ pos, float* __restrict__ lse,
                                                          float* __restrict__ loss_acc, float loss_scale, int m,
                                                          int slabs, int64_t dir_stride_part, int64_t dir_stride_vec) {
  __shared__ float red[8][32];
  const int dir = blockIdx.y;
  part_max += dir * dir_stride_part;
  part_sum += dir * dir_stride_part;
  lse += dir * dir_stride_vec;
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int row = blockIdx.x * 32 + lane;
  const bool live = row < m;
  float mx = -INFINITY;
  if (live) {
#pragma unroll 4
    for (int s = sub; s < slabs; s += 8) mx = fmaxf(mx, part_max[static_cast<int64_t>(s) * m + row]);
  }
  red[sub][lane] = mx;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) mx = fmaxf(mx, red[i][lane]);
  __syncthreads();
  float sum = 0.f;
  if (live) {
#pragma unroll 4
    for (int s = sub; s < slabs; s += 8) {
      const float pm = part_max[static_cast<int64_t>(s) * m + row];
      const float ps = part_sum[static_cast<int64_t>(s) * m + row];
      if (pm != -INFINITY) sum += ps * expf(pm - mx);  // precise: the backward differentiates through exp(s - lse)
    }
  }
  red[sub][lane] = sum;
  __syncthreads();
  if (sub == 0) {
    float local = 0.f;
    if (live) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][lane];
      const float l = mx + logf(t);
      lse[row] = l;
      if (pos != nullptr) local = l - pos[dir * dir_stride_vec + row];
    }
    if (loss_acc != nullptr) {
      local = warp_sum(local);
      if (lane == 0) atomicAdd(loss_acc, local * loss_scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Peer gather: every rank's [B, E] bf16 feature block -> this rank's local [W*B, E] copies, by plain 16-byte loads from
// the peer-mapped buffers (NVLink / NVSwitch P2P) with 8 independent loads in flight per thread x 2048 threads per SM.
// Why not TMA tiles straight into the GEMM (CLIPN_PEER_DIRECT=1 keeps that path): a TMA box over a peer tensor is
// fetched as one 128-byte NVLink read per row with few requests outstanding — measured on 2 x B200, a CTA refilling
// its 128 KB column tile on demand waited ~140 us (0.9 GB/s per SM), while all SMs pulling coalesced 16-byte loads
// together run at the fabric rate.  So the operand crosses NVLink exactly once, at full rate, and the GEMM that
// follows reads L2-resident local memory.
// ---------------------------------------------------------------------------------------------------------------------
struct GatherPtrs {
  const uint4* src[2][kMaxBMaps];  // [direction][rank]
  uint4* dst[2];
  int world;
  int64_t block16;                 // 16-byte elements per rank block (B * E * 2 / 16)
};

__global__ void __launch_bounds__(512) peer_gather_kernel(const __grid_constant__ GatherPtrs g) {
  constexpr int U = 8;
  const int64_t per_dir = g.block16 * g.world;
  const int64_t total = 2 * per_dir;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        const int dir = i >= per_dir;
        const int64_t j = i - dir * per_dir;
        const int r = static_cast<int>(j / g.block16);
        v[u] = __ldg(g.src[dir][r] + (j - r * g.block16));  // peer addresses bypass the local L2 anyway
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        const int dir = i >= per_dir;
        g.dst[dir][i - dir * per_dir] = v[u];
      }
    }
  }
}

static bool peer_direct() {
  static const bool on = [] {
    const char* e = getenv("CLIPN_PEER_DIRECT");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

// Opt-in stage timing of the fused forward (bench.py / tools: clipn_stage_timing(1)): CUDA events on the launching
// stream between the gather, the GEMM and the combine kernel of each call; read back (averaged) by
// clipn_stage_times.  Off by default: no events are recorded.
namespace {
constexpr int kStageRing = 64;
struct StageTimer {
  bool on = false;
  int calls = 0;
  cudaEvent_t ev[kStageRing][4] = {};
  bool made = false;
};
StageTimer g_stage;
inline void stage_mark(int k, cudaStream_t stream) {
  if (!g_stage.on || g_stage.calls >= kStageRing) return;
  if (!g_stage.made) {
    for (auto& q : g_stage.ev)
      for (auto& e : q) cudaEventCreate(&e);
    g_stage.made = true;
  }
  cudaEventRecord(g_stage.ev[g_stage.calls][k], stream);
  if (k == 3) ++g_stage.calls;
}
}  // namespace

extern "C" int clipn_stage_timing(int32_t enable) {
  g_stage.on = enable != 0;
  g_stage.calls = 0;
  return CLIPN_OK;
}

// out[0..2] = mean milliseconds of (gather, GEMM, combine) over the calls recorded since clipn_stage_timing(1);
// returns the number of calls averaged (synchronises on the last event)
extern "C" int32_t clipn_stage_times(float* out) {
  const int n = g_stage.calls;
  out[0] = out[1] = out[2] = 0.f;
  if (n == 0) return 0;
  cudaEventSynchronize(g_stage.ev[n - 1][3]);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_stage.ev[i][k], g_stage.ev[i][k + 1]);
      out[k] += ms / n;
    }
  return n;
}

// gather every rank's columns into the local copies; afterwards the GEMM sees ONE local [W*B, E] operand per direction
static int gather_columns(const void* const* txt_cols, const void* const* img_cols, int world, int b, int e, void* gather_txt,
                          void* gather_img, cudaStream_t stream) {
  CLIPN_REQUIRE(gather_txt != nullptr && gather_img != nullptr, "fused forward: world > 1 needs the local gather buffers");
  CLIPN_REQUIRE((static_cast<int64_t>(b) * e * 2) % 16 == 0, "fused forward: B * E must be a multiple of 8");
  GatherPtrs g;
  for (int r = 0; r < world; ++r) {
    CLIPN_REQUIRE(txt_cols[r] != nullptr && img_cols[r] != nullptr, "fused forward: null column pointer");
    g.src[0][r] = reinterpret_cast<const uint4*>(txt_cols[r]);
    g.src[1][r] = reinterpret_cast<const uint4*>(img_cols[r]);
  }
  g.dst[0] = reinterpret_cast<uint4*>(gather_txt);
  g.dst[1] = reinterpret_cast<uint4*>(gather_img);
  g.world = world;
  g.block16 = static_cast<int64_t>(b) * e * 2 / 16;
  peer_gather_kernel<<<num_sms() * 4, 512, 0, stream>>>(g);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

}  // namespace clipn
extern "C" int clipn_peer_gather(const void* const* txt_cols, const void* const* img_cols, int32_t world, int32_t b,
                                 int32_t e, void* gather_txt, void* gather_img, clipn_stream_t stream) {
  CLIPN_REQUIRE(txt_cols && img_cols, "peer_gather: null pointer");
  CLIPN_REQUIRE(world >= 1 && world <= clipn::kMaxBMaps, "peer_gather: world must be 1..8");
  return clipn::gather_columns(txt_cols, img_cols, world, b, e, gather_txt, gather_img, static_cast<cudaStream_t>(stream));
}
namespace clipn {

static void base_desc(clipn_gemm_desc& d, const void* rows, const void* cols, int m, int n, int e, float scale,
                      const float* scale_dev) {
  memset(&d, 0, sizeof(d));
  d.a = rows; d.lda = e; d.a_mn_major = 0;
  d.b = cols; d.ldb = e; d.b_mn_major = 0;
  d.m = m; d.n = n; d.k = e;
  d.alpha = scale; d.alpha_dev = scale_dev; d.splits = 1;
}

}  // namespace clipn

using namespace clipn;

// ---------------------------------------------------------------------------------------------------
// fused forward (peer-streaming kernel)
// ---------------------------------------------------------------------------------------------------
extern "C" int32_t clipn_peer_gemm_tile_n(int32_t world, int32_t b, int32_t e) { return peer_gemm_tile_n(world, b, e); }

extern "C" int64_t clipn_clip_fwd_fused_workspace(int32_t world, int32_t b, int32_t e) {
  const int bn = (world > 1 && peer_direct()) ? peer_gemm_tile_n(world, b, e) : peer_gemm_tile_n(1, world * b, e);
  if (bn == 0) return 0;
  const int64_t n = static_cast<int64_t>(world) * b;
  const int64_t slabs = 2 * ((n + bn - 1) / bn);  // one per column half of a tile
  return 2 * (2 * slabs * b + b);  // per direction: part_max + part_sum [slabs, b], pos [b]
}

extern "C" int clipn_clip_fwd_fused(const void* img_rows, const void* txt_rows, const void* const* txt_cols,
                                    const void* const* img_cols, int32_t world, int32_t rank, int32_t b, int32_t e,
                                    float scale, const float* scale_dev, void* gather_txt, void* gather_img, float* lse,
                                    float* loss_acc, float* workspace, clipn_stream_t stream) {
  CLIPN_REQUIRE(img_rows && txt_rows && txt_cols && img_cols && lse && workspace, "clip_fwd_fused: null pointer");
  CLIPN_REQUIRE(world >= 1 && world <= kMaxBMaps, "clip_fwd_fused: world must be 1..8");
  CLIPN_REQUIRE((gather_txt == nullptr) == (gather_img == nullptr), "clip_fwd_fused: both gather buffers or none");
  const bool direct = world > 1 && peer_direct();  // TMA tiles straight from the peers inside the GEMM (experiment)
  const int bn = direct ? peer_gemm_tile_n(world, b, e) : peer_gemm_tile_n(1, world * b, e);
  CLIPN_REQUIRE(bn != 0, "clip_fwd_fused: unsupported shape (see clipn_peer_gemm_tile_n)");
  const int64_t n = static_cast<int64_t>(world) * b;
  const int slabs = 2 * static_cast<int>((n + bn - 1) / bn);
  const int64_t part = static_cast<int64_t>(slabs) * b;
  float* pos = workspace + 4 * part;  // [2, b]
  const void* local_txt[1] = {gather_txt};
  const void* local_img[1] = {gather_img};
  PeerGemmDesc d;
  memset(&d, 0, sizeof(d));
  d.rows[0] = img_rows; d.rows[1] = txt_rows;
  d.dirs = 2; d.m = b; d.e = e;
  stage_mark(0, static_cast<cudaStream_t>(stream));
  if (world > 1 && !direct) {
    int rcg = gather_columns(txt_cols, img_cols, world, b, e, gather_txt, gather_img, static_cast<cudaStream_t>(stream));
    if (rcg) return rcg;
    d.cols[0] = local_txt; d.cols[1] = local_img;
    d.world = 1; d.rank = 0; d.rows_per_map = static_cast<int>(n);
  } else {
    d.cols[0] = txt_cols; d.cols[1] = img_cols;
    d.gather[0] = gather_txt; d.gather[1] = gather_img;
    d.world = world; d.rank = rank; d.rows_per_map = b;
  }
  d.epilogue = CLIPN_EPI_LSE;
  d.alpha = scale; d.alpha_dev = scale_dev;
  d.label_offset = world > 1 ? rank * b : 0;
  for (int dir = 0; dir < 2; ++dir) {
    d.part_max[dir] = workspace + dir * part;
    d.part_sum[dir] = workspace + 2 * part + dir * part;
    d.pos[dir] = pos + dir * static_cast<int64_t>(b);
  }
  stage_mark(1, static_cast<cudaStream_t>(stream));
  int rc = peer_gemm_launch(d, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  stage_mark(2, static_cast<cudaStream_t>(stream));
  dim3 grid((b + 31) / 32, 2);
  lse_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(workspace, workspace + 2 * part, pos, lse,
                                                                         loss_acc, 1.0f / (2.0f * b), b, slabs, part, b);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  stage_mark(3, static_cast<cudaStream_t>(stream));
  return CLIPN_OK;
}

extern "C" int clipn_siglip_fwd_fused(const void* img_rows, const void* txt_rows, const void* const* txt_cols,
                                      const void* const* img_cols, int32_t world, int32_t rank, int32_t b, int32_t e,
                                      const float* scale_dev, const float* bias_dev, float gscale, void* gather_txt,
                                      void* gather_img, float* loss_acc, float* scalar_acc, void* dl_img, void* dl_txt,
                                      int64_t ld, clipn_stream_t stream) {
  CLIPN_REQUIRE(img_rows && txt_rows && txt_cols && img_cols && loss_acc && scale_dev, "siglip_fwd_fused: null pointer");
  CLIPN_REQUIRE((dl_img == nullptr) == (dl_txt == nullptr), "siglip_fwd_fused: both d(logits) buffers or none");
  CLIPN_REQUIRE(world >= 1 && world <= kMaxBMaps, "siglip_fwd_fused: world must be 1..8");
  const int dirs = dl_txt != nullptr ? 2 : 1;  // the text direction only produces gradients
  const bool direct = world > 1 && peer_direct();
  const int64_t n = static_cast<int64_t>(world) * b;
  const void* local_txt[1] = {gather_txt};
  const void* local_img[1] = {gather_img};
  PeerGemmDesc d;
  memset(&d, 0, sizeof(d));
  d.rows[0] = img_rows; d.rows[1] = txt_rows;
  d.dirs = dirs; d.m = b; d.e = e;
  if (world > 1 && !direct) {
    int rcg = gather_columns(txt_cols, img_cols, world, b, e, gather_txt, gather_img, static_cast<cudaStream_t>(stream));
    if (rcg) return rcg;
    d.cols[0] = local_txt; d.cols[1] = local_img;
    d.world = 1; d.rank = 0; d.rows_per_map = static_cast<int>(n);
  } else {
    d.cols[0] = txt_cols; d.cols[1] = img_cols;
    d.gather[0] = dirs == 2 ? gather_txt : nullptr;  // nothing downstream reads the copies without a backward
    d.gather[1] = dirs == 2 ? gather_img : nullptr;
    d.world = world; d.rank = rank; d.rows_per_map = b;
  }
  d.epilogue = CLIPN_EPI_SIGLIP;
  d.alpha = 1.0f; d.alpha_dev = scale_dev; d.logit_bias = 0.f; d.logit_bias_dev = bias_dev;
  d.gscale = gscale;
  d.label_offset = world > 1 ? rank * b : 0;
  d.part_sum[0] = loss_acc; d.scalar_acc[0] = scalar_acc;  // value, d scale, d bias: image direction only
  d.c[0] = dl_img; d.c[1] = dl_txt; d.ldc = ld;
  return peer_gemm_launch(d, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------
// generic (single tensor map, local operands) pieces: forward fallback for shapes the peer kernel does not take,
// and the backward
// ---------------------------------------------------------------------------------------------------
extern "C" int64_t clipn_clip_lse_workspace(int32_t m, int32_t n) {
  const int bn = gemm_tile_n(n);
  const int64_t slabs = 2 * static_cast<int64_t>((n + bn - 1) / bn);
  return 2 * slabs * m;
}

extern "C" int clipn_clip_lse_fwd(const void* feats_rows, const void* feats_cols, int32_t m, int32_t n, int32_t e,
                                  float scale, const float* scale_dev, int32_t label_offset, float* lse, float* pos,
                                  float* workspace, clipn_stream_t stream) {
  CLIPN_REQUIRE(feats_rows && feats_cols && lse && pos && workspace, "clip_lse_fwd: null pointer");
  const int bn = gemm_tile_n(n);
  const int slabs = 2 * ((n + bn - 1) / bn);
  clipn_gemm_desc d;
  base_desc(d, feats_rows, feats_cols, m, n, e, scale, scale_dev);
  d.epilogue = CLIPN_EPI_LSE;
  d.part_max = workspace;
  d.part_sum = workspace + static_cast<int64_t>(slabs) * m;
  d.pos = pos;
  d.label_offset = label_offset;
  const void* bp[1] = {feats_cols};
  int rc = gemm_launch(d, bp, 1, 0, false, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  lse_combine_kernel<<<dim3((m + 31) / 32, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d.part_max, d.part_sum, nullptr, lse, nullptr, 0.f, m, slabs, 0, 0);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_clip_dlogits(const void* feats_rows, const void* feats_cols, int32_t m, int32_t n, int32_t e,
                                  float scale, const float* scale_dev, int32_t label_offset, const float* row_lse,
                                  const float* col_lse, float col_w, float gscale, void* dlogits, int64_t ld,
                                  float* scalar_acc, const float* row_centre, clipn_stream_t stream) {
  CLIPN_REQUIRE(feats_rows && feats_cols && row_lse && dlogits, "clip_dlogits: null pointer");
  clipn_gemm_desc d;
  base_desc(d, feats_rows, feats_cols, m, n, e, scale, scale_dev);
  d.epilogue = CLIPN_EPI_CLIP_DLOGITS;
  d.c = dlogits; d.ldc = ld;
  d.row_lse = row_lse; d.col_lse = col_lse; d.col_w = col_w; d.gscale = gscale;
  d.scalar_acc = scalar_acc;
  d.pos = const_cast<float*>(row_centre);  // read-only here: per-row centre of the d logit_scale sum
  d.label_offset = label_offset;
  const void* bp[1] = {feats_cols};
  return gemm_launch(d, bp, 1, 0, false, static_cast<cudaStream_t>(stream));
}

extern "C" int clipn_clip_dfeat(const void* dlogits, int64_t ld, const void* feats_cols, int32_t m, int32_t n, int32_t e,
                                float alpha, const float* alpha_dev, float* d_rows, int32_t splits,
                                clipn_stream_t stream) {
  CLIPN_REQUIRE(dlogits && feats_cols && d_rows, "clip_dfeat: null pointer");
  clipn_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = dlogits; d.lda = ld; d.a_mn_major = 0;     // [m, n], reduction over n
  d.b = feats_cols; d.ldb = e; d.b_mn_major = 1;  // [n (K rows), e] -> MN-major
  d.c = d_rows; d.ldc = e;
  d.m = m; d.n = e; d.k = n;
  d.alpha = alpha; d.alpha_dev = alpha_dev; d.splits = splits < 1 ? 1 : splits;
  d.epilogue = CLIPN_EPI_ACCUM_F32;               // d_rows (fp32, caller-zeroed) += alpha * dlogits @ cols
  const void* bp[1] = {feats_cols};
  return gemm_launch(d, bp, 1, 0, false, static_cast<cudaStream_t>(stream));
}
