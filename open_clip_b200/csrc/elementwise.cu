// HBM-bound kernels of the CLIP towers: LayerNorm fwd/bwd, patch/token embedding, pooling gathers,
// L2-normalize, bias-gradient column sums.  All are one-pass, 16-byte vectorised, warp-per-row where a row
// reduction is needed (warp-shuffle reductions, no shared memory), grid sized in multiples of the SM count.
#include "common.cuh"

namespace clipn {

// Row kernels are templated on NCH = 16-byte chunks per lane (a lane holds NCH x 8 bf16 of a row): d <= 256*NCH.
// Dispatching on NCH keeps the per-lane register arrays exactly as large as the row needs (occupancy).
#define CLIPN_DISPATCH_NCH(d, CALL)          \
  do {                                       \
    const int nch_ = ((d) / 8 + 31) / 32;    \
    if (nch_ <= 1) { CALL(1); }              \
    else if (nch_ == 2) { CALL(2); }         \
    else if (nch_ == 3) { CALL(3); }         \
    else { CALL(4); }                        \
  } while (0)

static inline int grid_for_rows(int64_t rows, int rows_per_block) {
  int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x-mean)*rstd*gamma + beta    (layers.py:11-26)
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256, (NCH <= 2) ? 4 : 3) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int64_t rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nchunk = d >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * wpb;
  int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  // software pipeline: the next row's 16-byte loads are in flight while this row is reduced, normalised and stored
  uint4 cur[NCH], nxt[NCH];
  auto load_row = [&](int64_t r, uint4(&dst)[NCH]) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * d);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      dst[c] = (ci < nchunk) ? xr[ci] : make_uint4(0, 0, 0, 0);
    }
  };
  if (row < rows) load_row(row, cur);
  for (; row < rows; row += stride) {
    if (row + stride < rows) load_row(row + stride, nxt);
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      unpack_bf16x8(cur[c], v[c]);  // chunks past the row are zero: they do not change the sum
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
    const float mean = warp_sum(s) / d;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = v[c][j] - mean;
          sq += t * t;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / d + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + row * d);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        const float4 g0 = reinterpret_cast<const float4*>(gamma)[ci * 2], g1 = reinterpret_cast<const float4*>(gamma)[ci * 2 + 1];
        const float4 b0 = reinterpret_cast<const float4*>(beta)[ci * 2], b1 = reinterpret_cast<const float4*>(beta)[ci * 2 + 1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * g[j] + b[j];
        yr[ci] = pack_bf16x8(o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) cur[c] = nxt[c];
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dx = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma; dgamma += sum dy*xhat,
// dbeta += sum dy.  Optional residual-gradient add fused into the store.
// ------------------------------------------------------------------------------------------------
// kRSum: also accumulate the column sums of the residual gradient (dresid_sum[d] += sum_rows dx_resid) — that is the
// bias gradient of the Linear whose output the residual stream received (out_proj / c_proj, transformer.py:328-329), so
// the two stand-alone column-sum launches per block disappear into the pass that already streams dx_resid.
template <int NCH, bool kRSum>
__global__ void __launch_bounds__(256, (NCH <= 2) ? 3 : 2) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                            const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in,
                                                            const float* __restrict__ gamma,
                                                            const __nv_bfloat16* __restrict__ dx_resid,
                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dresid_sum,
                                                            int64_t rows, int d) {
  extern __shared__ float sred[];  // [2 or 3][d]
  float* s_dg = sred;
  float* s_db = sred + d;
  float* s_dr = sred + 2 * d;
  for (int i = threadIdx.x; i < (kRSum ? 3 : 2) * d; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nchunk = d >> 3;
  float acc_dg[NCH][8], acc_db[NCH][8], acc_dr[kRSum ? NCH : 1][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_dg[c][j] = acc_db[c][j] = 0.f;
  if (kRSum) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_dr[c][j] = 0.f;
  }

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * d);
    const uint4* rr = dx_resid ? reinterpret_cast<const uint4*>(dx_resid + row * d) : nullptr;
    // all three streams (x, dy, residual gradient) are requested up front: 9 independent 16-byte loads per lane
    uint4 xq[NCH], dq[NCH], rq[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        xq[c] = xr[ci];
        dq[c] = dyr[ci];
        rq[c] = rr ? rr[ci] : make_uint4(0, 0, 0, 0);
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    // the rows stay packed (bf16) in registers between the two passes; xhat and g are re-derived, not kept in fp32
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        float xv[8], dv[8];
        unpack_bf16x8(xq[c], xv);
        unpack_bf16x8(dq[c], dv);
        const float4 g0 = reinterpret_cast<const float4*>(gamma)[ci * 2], g1 = reinterpret_cast<const float4*>(gamma)[ci * 2 + 1];
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float gj = dv[j] * gm[j];
          s1 += gj;
          s2 += gj * xh;
          acc_dg[c][j] += dv[j] * xh;
          acc_db[c][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
    uint4* dxr = reinterpret_cast<uint4*>(dx + row * d);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        float xv[8], dv[8], o[8], r[8];
        unpack_bf16x8(xq[c], xv);
        unpack_bf16x8(dq[c], dv);
        unpack_bf16x8(rq[c], r);  // zeros when there is no residual gradient
        const float4 g0 = reinterpret_cast<const float4*>(gamma)[ci * 2], g1 = reinterpret_cast<const float4*>(gamma)[ci * 2 + 1];
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dv[j] * gm[j] - s1 - (xv[j] - mean) * rstd * s2) + r[j];
        dxr[ci] = pack_bf16x8(o);
        if (kRSum) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_dr[c][j] += r[j];
        }
      }
    }
  }
  // block reduce of the per-lane column partials, then one atomic per column per block
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ci = lane + c * 32;
    if (ci < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_dg[ci * 8 + j], acc_dg[c][j]);
        atomicAdd(&s_db[ci * 8 + j], acc_db[c][j]);
        if (kRSum) atomicAdd(&s_dr[ci * 8 + j], acc_dr[c][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], s_dg[i]);
    if (dbeta) atomicAdd(&dbeta[i], s_db[i]);
    if (kRSum) atomicAdd(&dresid_sum[i], s_dr[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// patchify (im2row for conv1 with kernel == stride): patches[b*np + gy*gw + gx][c*P*P + py*P + px]
// ------------------------------------------------------------------------------------------------
// Any even patch size (ViT-L/14: 14 px = 28 B per patch row, not 16-byte aligned): one bf16 pair per thread in
// OUTPUT order, so writes are coalesced; columns [chans*patch*patch, ld) are zero-filled (K padding for TMA).
__global__ void __launch_bounds__(256) patchify_pair_kernel(const __nv_bfloat16* __restrict__ img,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld, int batch,
                                                            int chans, int height, int width, int patch) {
  const int gw = width / patch, gh = height / patch;
  const int kcols = chans * patch * patch;
  const int pairs = static_cast<int>(ld / 2);
  const int64_t total = static_cast<int64_t>(batch) * gh * gw * pairs;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % pairs) * 2;
    const int64_t prow = i / pairs;
    uint32_t v = 0u;
    if (col < kcols) {
      const int px = col % patch;
      const int py = (col / patch) % patch;
      const int c = col / (patch * patch);
      const int gx = static_cast<int>(prow % gw);
      const int gy = static_cast<int>((prow / gw) % gh);
      const int64_t b = prow / (static_cast<int64_t>(gw) * gh);
      v = *reinterpret_cast<const uint32_t*>(img + ((b * chans + c) * height + gy * patch + py) * width + gx * patch + px);
    }
    *reinterpret_cast<uint32_t*>(out + prow * ld + col) = v;
  }
}

// dst[r][c] += src[r][c] for c < cols (fp32; un-pads a K-padded weight gradient into the parameter's own layout)
__global__ void __launch_bounds__(256) accum_rows_f32_kernel(float* __restrict__ dst, int64_t ld_dst,
                                                             const float* __restrict__ src, int64_t ld_src, int64_t rows,
                                                             int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const int c = static_cast<int>(i % cols);
    dst[r * ld_dst + c] += src[r * ld_src + c];
  }
}

__global__ void __launch_bounds__(256) patchify_kernel(const __nv_bfloat16* __restrict__ img,
                                                       __nv_bfloat16* __restrict__ out, int batch, int chans, int height,
                                                       int width, int patch) {
  // one 16-byte vector (8 px of one image row inside one patch) per thread-iteration
  const int gw = width / patch, gh = height / patch;
  const int vec_per_prow = patch / 8;
  const int64_t total = static_cast<int64_t>(batch) * chans * height * (width / 8);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // source order (coalesced reads): b, c, y, xvec
    const int xv = static_cast<int>(i % (width / 8));
    int64_t t = i / (width / 8);
    const int y = static_cast<int>(t % height);
    t /= height;
    const int c = static_cast<int>(t % chans);
    const int b = static_cast<int>(t / chans);
    const int gx = xv / vec_per_prow, pv = xv % vec_per_prow;
    const int gy = y / patch, py = y % patch;
    if (gx >= gw || gy >= gh) continue;
    const uint4 v = reinterpret_cast<const uint4*>(img)[i];
    const int64_t prow = (static_cast<int64_t>(b) * gh + gy) * gw + gx;
    const int64_t col = (static_cast<int64_t>(c) * patch + py) * patch + pv * 8;
    reinterpret_cast<uint4*>(out + prow * (static_cast<int64_t>(chans) * patch * patch) + col)[0] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// vision embed: cat(class_embedding, patches) + positional_embedding   (transformer.py:799-801)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vision_embed_fwd_kernel(const __nv_bfloat16* __restrict__ patch_out,
                                                               const float* __restrict__ cls,
                                                               const float* __restrict__ pos,
                                                               __nv_bfloat16* __restrict__ x, int batch, int npatch, int d) {
  const int L = npatch + 1;
  const int dv = d / 8;
  const int64_t total = static_cast<int64_t>(batch) * L * dv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % dv);
    const int64_t tok = i / dv;
    const int l = static_cast<int>(tok % L);
    const int64_t b = tok / L;
    float a[8], pz[8], o[8];
    const float4 p0 = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(l) * d)[cv * 2];
    const float4 p1 = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(l) * d)[cv * 2 + 1];
    const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) pz[j] = bf16_round(pp[j]);
    if (l == 0) {
      const float4 c0 = reinterpret_cast<const float4*>(cls)[cv * 2], c1 = reinterpret_cast<const float4*>(cls)[cv * 2 + 1];
      const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = bf16_round(cc[j]);
    } else {
      unpack_bf16x8(reinterpret_cast<const uint4*>(patch_out + (b * npatch + (l - 1)) * d)[cv], a);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] + pz[j];
    reinterpret_cast<uint4*>(x + tok * d)[cv] = pack_bf16x8(o);
  }
}

// dx [B, L, d] -> dpatch_out [B*np, d] (copy), dcls += sum_b dx[b,0], dpos[l] += sum_b dx[b,l]
// grid.x = L, each block reduces one token position over the batch.
__global__ void __launch_bounds__(256) vision_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx,
                                                               __nv_bfloat16* __restrict__ dpatch,
                                                               float* __restrict__ dcls, float* __restrict__ dpos,
                                                               int batch, int npatch, int d) {
  const int L = npatch + 1;
  const int l = blockIdx.x;
  const int bchunk = blockIdx.y, nb = gridDim.y;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int b = bchunk; b < batch; b += nb) {
      const __nv_bfloat16 v = dx[(static_cast<int64_t>(b) * L + l) * d + c];
      s += __bfloat162float(v);
      if (l > 0) dpatch[(static_cast<int64_t>(b) * npatch + (l - 1)) * d + c] = v;
    }
    if (dpos) atomicAdd(&dpos[static_cast<int64_t>(l) * d + c], s);
    if (l == 0 && dcls) atomicAdd(&dcls[c], s);
  }
}

// ------------------------------------------------------------------------------------------------
// text embed: token_embedding gather + positional add (model.py:399-401) and EOT argmax
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) text_embed_fwd_kernel(const int64_t* __restrict__ ids,
                                                             const float* __restrict__ table,
                                                             const float* __restrict__ pos, __nv_bfloat16* __restrict__ x,
                                                             int batch, int seq, int d, int vocab) {
  const int dv = d / 8;
  const int64_t total = static_cast<int64_t>(batch) * seq * dv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % dv);
    const int64_t tok = i / dv;
    const int l = static_cast<int>(tok % seq);
    int64_t id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4 t0 = reinterpret_cast<const float4*>(table + id * d)[cv * 2];
    const float4 t1 = reinterpret_cast<const float4*>(table + id * d)[cv * 2 + 1];
    const float4 p0 = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(l) * d)[cv * 2];
    const float4 p1 = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(l) * d)[cv * 2 + 1];
    const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(tt[j]) + bf16_round(pp[j]);
    reinterpret_cast<uint4*>(x + tok * d)[cv] = pack_bf16x8(o);
  }
}

__global__ void argmax_rows_kernel(const int64_t* __restrict__ ids, int32_t* __restrict__ idx, int batch, int seq) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int64_t best = ids[static_cast<int64_t>(b) * seq];
  int bi = 0;
  for (int l = 1; l < seq; ++l) {
    const int64_t v = ids[static_cast<int64_t>(b) * seq + l];
    if (v > best) {
      best = v;
      bi = l;
    }
  }
  idx[b] = bi;
}

// dtable[ids[tok]] += dx[tok] (fp32 atomics; ~B*L*d adds), dpos[l] += sum_b dx[b,l]
__global__ void __launch_bounds__(256) text_embed_bwd_kernel(const int64_t* __restrict__ ids,
                                                             const __nv_bfloat16* __restrict__ dx,
                                                             float* __restrict__ dtable, float* __restrict__ dpos,
                                                             int batch, int seq, int d, int vocab) {
  const int l = blockIdx.x;
  const int bchunk = blockIdx.y, nb = gridDim.y;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int b = bchunk; b < batch; b += nb) {
      const int64_t tok = static_cast<int64_t>(b) * seq + l;
      const float v = __bfloat162float(dx[tok * d + c]);
      s += v;
      int64_t id = ids[tok];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
      atomicAdd(&dtable[id * d + c], v);
    }
    if (dpos) atomicAdd(&dpos[static_cast<int64_t>(l) * d + c], s);
  }
}

// ------------------------------------------------------------------------------------------------
// pooling gathers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                                          int batch, int seq, int d) {
  const int dv = d / 8;
  const int64_t total = static_cast<int64_t>(batch) * dv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % dv);
    const int64_t b = i / dv;
    const int l = idx ? idx[b] : 0;
    reinterpret_cast<uint4*>(out + b * d)[cv] = reinterpret_cast<const uint4*>(x + (b * seq + l) * d)[cv];
  }
}
__global__ void __launch_bounds__(256) scatter_rows_kernel(const __nv_bfloat16* __restrict__ dpooled,
                                                           const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                                                           int batch, int seq, int d) {
  const int dv = d / 8;
  const int64_t total = static_cast<int64_t>(batch) * seq * dv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % dv);
    const int64_t tok = i / dv;
    const int l = static_cast<int>(tok % seq);
    const int64_t b = tok / seq;
    const int sel = idx ? idx[b] : 0;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (l == sel) v = reinterpret_cast<const uint4*>(dpooled + b * d)[cv];
    reinterpret_cast<uint4*>(dx + tok * d)[cv] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// F.normalize (model.py:391): warp per row
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ inv_norm, int64_t rows, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, nchunk = d >> 3;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
    float v[NCH][8];
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        unpack_bf16x8(xr[ci], v[c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += v[c][j] * v[c][j];
      }
    }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
    uint4* yr = reinterpret_cast<uint4*>(y + row * d);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[c][j] * inv;
        yr[ci] = pack_bf16x8(o);
      }
    }
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
  }
}

template <bool DY_F32, int NCH>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const void* __restrict__ dy_, const __nv_bfloat16* __restrict__ y,
                                                         const float* __restrict__ inv_norm, __nv_bfloat16* __restrict__ dx,
                                                         int64_t rows, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, nchunk = d >> 3;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const uint4* yr = reinterpret_cast<const uint4*>(y + row * d);
    float yv[NCH][8], gv[NCH][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        unpack_bf16x8(yr[ci], yv[c]);
        if constexpr (DY_F32) {
          const float4* g = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * d);
          const float4 a = g[ci * 2], b = g[ci * 2 + 1];
          gv[c][0] = a.x; gv[c][1] = a.y; gv[c][2] = a.z; gv[c][3] = a.w;
          gv[c][4] = b.x; gv[c][5] = b.y; gv[c][6] = b.z; gv[c][7] = b.w;
        } else {
          unpack_bf16x8(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + row * d)[ci], gv[c]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dot += yv[c][j] * gv[c][j];
      }
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    uint4* dxr = reinterpret_cast<uint4*>(dx + row * d);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ci = lane + c * 32;
      if (ci < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = inv * (gv[c][j] - yv[c][j] * dot);
        dxr[ci] = pack_bf16x8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] += sum_rows x[row, n].  Block = 32 x 8 threads: each warp-row
// of the block strides over rows, 8 columns (one 16-byte vector) per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                     float* __restrict__ out, int64_t rows, int n) {
  __shared__ float s_part[8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < n) {
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + ty; r < rows; r += static_cast<int64_t>(gridDim.y) * 8) {
      float v[8];
      unpack_bf16x8(*reinterpret_cast<const uint4*>(x + r * ldx + col), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_part[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.x;  // 256 columns of this block
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += s_part[k][t];
  const int c = blockIdx.x * 256 + t;
  if (c < n) atomicAdd(&out[c], s);
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

static inline int grid_for_elems(int64_t n, int per_block = 256) {
  int64_t blocks = (n + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace clipn

using namespace clipn;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" int clipn_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                   float* rstd, int64_t rows, int32_t d, float eps, clipn_stream_t stream) {
  CLIPN_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null pointer");
  CLIPN_REQUIRE(d % 8 == 0 && d <= 1024 && d > 0, "layernorm: d must be a multiple of 8 and <= 1024");
  if (rows <= 0) return CLIPN_OK;
#define CLIPN_LN_FWD(N) \
  layernorm_fwd_kernel<N><<<grid_for_rows(rows, 8), 256, 0, ST(stream)>>>(BF(x), gamma, beta, BFW(y), mean, rstd, rows, d, eps)
  CLIPN_DISPATCH_NCH(d, CLIPN_LN_FWD);
#undef CLIPN_LN_FWD
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                   const float* gamma, const void* dx_resid, void* dx_out, float* dgamma, float* dbeta,
                                   float* dresid_sum, int64_t rows, int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(dy && x && mean && rstd && gamma && dx_out, "layernorm_bwd: null pointer");
  CLIPN_REQUIRE(dresid_sum == nullptr || dx_resid != nullptr, "layernorm_bwd: dresid_sum needs dx_resid");
  CLIPN_REQUIRE(d % 8 == 0 && d <= 1024 && d > 0, "layernorm: d must be a multiple of 8 and <= 1024");
  if (rows <= 0) return CLIPN_OK;
  int64_t blocks = (rows + 7) / 8;
  const int cap = num_sms() * (d <= 512 ? 3 : 2);  // one resident wave: every block does a single dgamma/dbeta flush
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
#define CLIPN_LN_BWD(N)                                                                                                 \
  if (dresid_sum != nullptr)                                                                                            \
    layernorm_bwd_kernel<N, true><<<grid, 256, 3 * d * sizeof(float), ST(stream)>>>(                                     \
        BF(dy), BF(x), mean, rstd, gamma, BF(dx_resid), BFW(dx_out), dgamma, dbeta, dresid_sum, rows, d);                \
  else                                                                                                                  \
    layernorm_bwd_kernel<N, false><<<grid, 256, 2 * d * sizeof(float), ST(stream)>>>(                                    \
        BF(dy), BF(x), mean, rstd, gamma, BF(dx_resid), BFW(dx_out), dgamma, dbeta, nullptr, rows, d)
  CLIPN_DISPATCH_NCH(d, CLIPN_LN_BWD);
#undef CLIPN_LN_BWD
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_patchify(const void* image, void* patches, int32_t batch, int32_t chans, int32_t height,
                              int32_t width, int32_t patch, clipn_stream_t stream) {
  CLIPN_REQUIRE(image && patches, "patchify: null pointer");
  CLIPN_REQUIRE(patch % 8 == 0 && width % 8 == 0, "patchify: patch and width must be multiples of 8");
  CLIPN_REQUIRE(height % patch == 0 && width % patch == 0, "patchify: image must be a whole number of patches");
  const int64_t total = static_cast<int64_t>(batch) * chans * height * (width / 8);
  if (total <= 0) return CLIPN_OK;
  patchify_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(BF(image), BFW(patches), batch, chans, height, width, patch);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_patchify_padded(const void* image, void* patches, int64_t ld_patches, int32_t batch, int32_t chans,
                                     int32_t height, int32_t width, int32_t patch, clipn_stream_t stream) {
  CLIPN_REQUIRE(image && patches, "patchify_padded: null pointer");
  CLIPN_REQUIRE(patch > 0 && patch % 2 == 0 && width % 2 == 0, "patchify_padded: patch and width must be even");
  CLIPN_REQUIRE(height % patch == 0 && width % patch == 0, "patchify_padded: image must be a whole number of patches");
  CLIPN_REQUIRE(ld_patches % 8 == 0 && ld_patches >= static_cast<int64_t>(chans) * patch * patch,
                "patchify_padded: row pitch must be a multiple of 8 and hold chans*patch*patch columns");
  const int64_t total = static_cast<int64_t>(batch) * (height / patch) * (width / patch) * (ld_patches / 2);
  if (total <= 0) return CLIPN_OK;
  patchify_pair_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(BF(image), BFW(patches), ld_patches, batch, chans,
                                                                      height, width, patch);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_accum_rows_f32(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows,
                                    int32_t cols, clipn_stream_t stream) {
  CLIPN_REQUIRE(dst && src, "accum_rows_f32: null pointer");
  CLIPN_REQUIRE(cols > 0 && ld_dst >= cols && ld_src >= cols, "accum_rows_f32: bad pitch");
  if (rows <= 0) return CLIPN_OK;
  accum_rows_f32_kernel<<<grid_for_elems(rows * cols), 256, 0, ST(stream)>>>(dst, ld_dst, src, ld_src, rows, cols);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_vision_embed_fwd(const void* patch_out, const float* cls, const float* pos, void* x, int32_t batch,
                                      int32_t npatch, int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(patch_out && cls && pos && x, "vision_embed_fwd: null pointer");
  CLIPN_REQUIRE(d % 8 == 0, "vision_embed: d must be a multiple of 8");
  const int64_t total = static_cast<int64_t>(batch) * (npatch + 1) * (d / 8);
  if (total <= 0) return CLIPN_OK;
  vision_embed_fwd_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(BF(patch_out), cls, pos, BFW(x), batch, npatch, d);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_vision_embed_bwd(const void* dx, void* dpatch_out, float* dcls, float* dpos, int32_t batch,
                                      int32_t npatch, int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(dx && dpatch_out, "vision_embed_bwd: null pointer");
  if (batch <= 0) return CLIPN_OK;
  int nb = (num_sms() * 4 + npatch) / (npatch + 1);
  if (nb < 1) nb = 1;
  if (nb > batch) nb = batch;
  vision_embed_bwd_kernel<<<dim3(npatch + 1, nb), 256, 0, ST(stream)>>>(BF(dx), BFW(dpatch_out), dcls, dpos, batch, npatch, d);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_text_embed_fwd(const int64_t* ids, const float* table, const float* pos, void* x, int32_t* eot_idx,
                                    int32_t batch, int32_t seq, int32_t d, int32_t vocab, clipn_stream_t stream) {
  CLIPN_REQUIRE(ids && table && pos && x, "text_embed_fwd: null pointer");
  CLIPN_REQUIRE(d % 8 == 0, "text_embed: d must be a multiple of 8");
  const int64_t total = static_cast<int64_t>(batch) * seq * (d / 8);
  if (total <= 0) return CLIPN_OK;
  text_embed_fwd_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(ids, table, pos, BFW(x), batch, seq, d, vocab);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  if (eot_idx) {
    argmax_rows_kernel<<<(batch + 127) / 128, 128, 0, ST(stream)>>>(ids, eot_idx, batch, seq);
    CLIPN_CHECK_CUDA(cudaGetLastError());
  }
  return CLIPN_OK;
}

extern "C" int clipn_text_embed_bwd(const int64_t* ids, const void* dx, float* dtable, float* dpos, int32_t batch,
                                    int32_t seq, int32_t d, int32_t vocab, clipn_stream_t stream) {
  CLIPN_REQUIRE(ids && dx && dtable, "text_embed_bwd: null pointer");
  if (batch <= 0) return CLIPN_OK;
  int nb = (num_sms() * 4 + seq - 1) / seq;
  if (nb < 1) nb = 1;
  if (nb > batch) nb = batch;
  text_embed_bwd_kernel<<<dim3(seq, nb), 256, 0, ST(stream)>>>(ids, BF(dx), dtable, dpos, batch, seq, d, vocab);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_gather_rows(const void* x, const int32_t* idx, void* out, int32_t batch, int32_t seq, int32_t d,
                                 clipn_stream_t stream) {
  CLIPN_REQUIRE(x && out && d % 8 == 0, "gather_rows: bad arguments");
  const int64_t total = static_cast<int64_t>(batch) * (d / 8);
  if (total <= 0) return CLIPN_OK;
  gather_rows_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(BF(x), idx, BFW(out), batch, seq, d);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_scatter_rows(const void* dpooled, const int32_t* idx, void* dx, int32_t batch, int32_t seq,
                                  int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(dpooled && dx && d % 8 == 0, "scatter_rows: bad arguments");
  const int64_t total = static_cast<int64_t>(batch) * seq * (d / 8);
  if (total <= 0) return CLIPN_OK;
  scatter_rows_kernel<<<grid_for_elems(total), 256, 0, ST(stream)>>>(BF(dpooled), idx, BFW(dx), batch, seq, d);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(x && y && d % 8 == 0 && d <= 1024, "l2norm_fwd: bad arguments");
  if (rows <= 0) return CLIPN_OK;
#define CLIPN_L2_FWD(N) l2norm_fwd_kernel<N><<<grid_for_rows(rows, 8), 256, 0, ST(stream)>>>(BF(x), BFW(y), inv_norm, rows, d)
  CLIPN_DISPATCH_NCH(d, CLIPN_L2_FWD);
#undef CLIPN_L2_FWD
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_l2norm_bwd(const void* dy, int32_t dy_is_f32, const void* y, const float* inv_norm, void* dx,
                                int64_t rows, int32_t d, clipn_stream_t stream) {
  CLIPN_REQUIRE(dy && y && inv_norm && dx && d % 8 == 0 && d <= 1024, "l2norm_bwd: bad arguments");
  if (rows <= 0) return CLIPN_OK;
#define CLIPN_L2_BWD(N)                                                                                                 \
  if (dy_is_f32)                                                                                                        \
    l2norm_bwd_kernel<true, N><<<grid_for_rows(rows, 8), 256, 0, ST(stream)>>>(dy, BF(y), inv_norm, BFW(dx), rows, d);  \
  else                                                                                                                  \
    l2norm_bwd_kernel<false, N><<<grid_for_rows(rows, 8), 256, 0, ST(stream)>>>(dy, BF(y), inv_norm, BFW(dx), rows, d)
  CLIPN_DISPATCH_NCH(d, CLIPN_L2_BWD);
#undef CLIPN_L2_BWD
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int32_t n, clipn_stream_t stream) {
  CLIPN_REQUIRE(x && out && n % 8 == 0 && ldx % 8 == 0, "colsum: bad arguments");
  if (rows <= 0) return CLIPN_OK;
  const int gx = (n + 255) / 256;
  int gy = (num_sms() * 8 + gx - 1) / gx;
  const int64_t maxy = (rows + 7) / 8;
  if (gy > maxy) gy = static_cast<int>(maxy);
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3(gx, gy), 256, 0, ST(stream)>>>(BF(x), ldx, out, rows, n);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_cast_f32_to_bf16(const float* x, void* y, int64_t n, clipn_stream_t stream) {
  CLIPN_REQUIRE(x && y, "cast: null pointer");
  if (n <= 0) return CLIPN_OK;
  cast_f32_bf16_kernel<<<grid_for_elems(n), 256, 0, ST(stream)>>>(x, BFW(y), n);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}
