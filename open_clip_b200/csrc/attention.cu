// Attention core for the CLIP towers (F.scaled_dot_product_attention, transformer.py:223-228), head_dim 64.
//
// Sequences are short (50 / 77 / 197 tokens; 577 for ViT-L/14-336) and the op is 1.6 % of the step's FLOPs: the
// bound is HBM bytes, the practical limit is instruction issue and latency (ncu, profiles/), so the design goals
// are bytes in flight and few non-MMA instructions.  Persistent CTAs loop over (batch, head) work items with
// a 2-stage cp.async pipeline: while the warps compute item i out of shared memory, the 16-byte cp.async copies
// of item i+1 (Q, K, V and, in the backward, dO) are already in flight.  Inside an item every warp owns 16-row
// tiles; S = QK^T and PV run on tensor cores (mma.sync m16n8k16 bf16 -> fp32 — the tiles are 16x64, far below the
// 128-row tcgen05 atom), softmax is a warp-shuffle (quad) reduction in registers with online rescaling, the
// causal mask is a predicate (no mask tensor) evaluated only on key tiles that cross the diagonal or the sequence
// end.  Outputs are staged through shared memory and leave as full 128-byte rows.
// Kernels by sequence length:
//   forward   L <= 528: attention_fwd_kernel (Q,K,V resident)      528 < L <= 640: attention_fwd_long_kernel
//   backward  L <= 80 : attention_bwd_small_kernel (P, dS cached)   80 < L < 192 : attention_bwd_kernel (recompute)
//             192 <= L <= 640: attention_bwd_long_dq_kernel + attention_bwd_long_dkv_kernel (two passes)
// The single-kernel backwards recompute P from the saved log-sum-exp and use D_i = sum_j P_ij dP_ij
// (== rowsum(dO o O)), so the forward output is not re-read: phase 1 (warp = 16 queries) computes D, phase 2a dQ,
// phase 2b (warp = 16 keys, transposed tiles) dK and dV — no atomics, deterministic.  The two-pass backward takes
// D from the forward output instead.
#include <stdlib.h>

#include "common.cuh"

namespace clipn {

constexpr int HD = 64;   // head dim
constexpr int LDS = 72;  // padded smem row (bf16 elements): 144 B => conflict-free fragment loads
constexpr int kLongSeqMax = 640;  // longest sequence the K,V-resident kernels hold in 227 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// A-operand fragments (16 rows x 64 k) of a row-major smem tile starting at row r0.
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16* tile, int r0, int lane, uint32_t (&a)[4][4]) {
  const int r = r0 + (lane >> 2), c = (lane & 3) * 2;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = lds32(tile + r * LDS + kk * 16 + c);
    a[kk][1] = lds32(tile + (r + 8) * LDS + kk * 16 + c);
    a[kk][2] = lds32(tile + r * LDS + kk * 16 + c + 8);
    a[kk][3] = lds32(tile + (r + 8) * LDS + kk * 16 + c + 8);
  }
}
// acc[16 x 8] += A(16 x 64) * T[n0..n0+7][0..63]^T   (B(k,n) = T[n][k], T row-major in smem)
__device__ __forceinline__ void mma_a_tT(float (&acc)[4], const uint32_t (&a)[4][4], const __nv_bfloat16* T, int n0,
                                         int lane) {
  const __nv_bfloat16* row = T + (n0 + (lane >> 2)) * LDS + (lane & 3) * 2;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) mma_bf16(acc, a[kk], lds32(row + kk * 16), lds32(row + kk * 16 + 8));
}
// out[jn] (16 x 64) += Pa(16x16) * T[k0..k0+15][0..63]   (T row-major [k][64]; B fragments via ldmatrix.trans)
__device__ __forceinline__ void mma_p_t(float (&out)[8][4], const uint32_t (&pa)[4], const __nv_bfloat16* T, int k0,
                                        int lane) {
  const int mi = lane >> 3, ri = lane & 7;
#pragma unroll
  for (int jn = 0; jn < 8; jn += 2) {
    uint32_t b[4];
    ldmatrix_x4_trans(b, T + (k0 + (mi & 1) * 8 + ri) * LDS + (jn + (mi >> 1)) * 8);
    mma_bf16(out[jn], pa, b[0], b[1]);
    mma_bf16(out[jn + 1], pa, b[2], b[3]);
  }
}
__device__ __forceinline__ void pack_frag(uint32_t (&pa)[4], const float (&t0)[4], const float (&t1)[4]) {
  pa[0] = pack_bf16x2(t0[0], t0[1]);
  pa[1] = pack_bf16x2(t0[2], t0[3]);
  pa[2] = pack_bf16x2(t1[0], t1[1]);
  pa[3] = pack_bf16x2(t1[2], t1[3]);
}

// async copy of `rows` x 64 bf16 (128 B per row) into a padded smem tile
__device__ __forceinline__ void tile_cp_async(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t src_ld, int rows) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, v = i & 7;
    cp_async16(dst + r * LDS + v * 8, src + r * src_ld + v * 8);
  }
}
__device__ __forceinline__ void tile_zero_pad(__nv_bfloat16* dst, int rows, int rows_pad) {
  for (int i = threadIdx.x; i < (rows_pad - rows) * 8; i += blockDim.x) {
    const int r = rows + (i >> 3), v = i & 7;
    *reinterpret_cast<uint4*>(dst + r * LDS + v * 8) = make_uint4(0, 0, 0, 0);
  }
}
// write a warp's staged 16 x 64 bf16 tile (rows r0.. of `stage`, LDS pitch) to global rows as 128-byte segments
__device__ __forceinline__ void store_tile16(const __nv_bfloat16* stage, __nv_bfloat16* gdst, int64_t g_ld, int r0,
                                             int seq, int lane) {
  for (int i = lane; i < 16 * 8; i += 32) {
    const int r = i >> 3, v = i & 7;
    if (r0 + r < seq)
      *reinterpret_cast<uint4*>(gdst + static_cast<int64_t>(r0 + r) * g_ld + v * 8) =
          *reinterpret_cast<const uint4*>(stage + r * LDS + v * 8);
  }
}
// dst[c] += sum over the 16 rows of a staged 16 x 64 bf16 tile (pad rows hold zeros): in_proj_bias gradient
__device__ __forceinline__ void tile_colsum_atomic(const __nv_bfloat16* stage, float* dst, int lane) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(stage + r * LDS + lane * 2);
    const float2 f = __bfloat1622float2(v);
    s0 += f.x;
    s1 += f.y;
  }
  atomicAdd(dst + lane * 2, s0);
  atomicAdd(dst + lane * 2 + 1, s1);
}
__device__ __forceinline__ void stage_frag_tile(__nv_bfloat16* stage, const float (&o)[8][4], float s0, float s1,
                                                int lane) {
  const int r = lane >> 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j * 8 + (lane & 3) * 2;
    *reinterpret_cast<uint32_t*>(stage + r * LDS + c) = pack_bf16x2(o[j][0] * s0, o[j][1] * s0);
    *reinterpret_cast<uint32_t*>(stage + (r + 8) * LDS + c) = pack_bf16x2(o[j][2] * s1, o[j][3] * s1);
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32, (MAXW <= 5) ? 3 : 2)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ lse_out, int items, int seq, int heads, int causal, float scale, int nstages) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  const int stage_elems = 3 * Lp * LDS;
  __nv_bfloat16* sbase = reinterpret_cast<__nv_bfloat16*>(smem_att);
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float sl2 = scale * kLog2e;

  for (int s = 0; s < nstages; ++s)
    for (int t = 0; t < 3; ++t) tile_zero_pad(sbase + s * stage_elems + t * Lp * LDS, seq, Lp);

  auto prefetch = [&](int item, int s) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    __nv_bfloat16* st = sbase + s * stage_elems;
    tile_cp_async(st, base, ld, seq);
    tile_cp_async(st + Lp * LDS, base + d, ld, seq);
    tile_cp_async(st + 2 * Lp * LDS, base + 2 * d, ld, seq);
  };

  int item = blockIdx.x;
  if (nstages == 2 && item < items) prefetch(item, 0);
  cp_async_commit();
  for (int it = 0; item < items; item += gridDim.x, ++it) {
    const int s = (nstages == 2) ? (it & 1) : 0;
    const int next = item + gridDim.x;
    if (nstages == 2) {
      if (next < items) prefetch(next, s ^ 1);
    } else {
      prefetch(item, 0);  // long sequences: one stage, no lookahead
    }
    cp_async_commit();
    if (nstages == 2) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();
    __nv_bfloat16* sQ = sbase + s * stage_elems;
    const __nv_bfloat16* sK = sQ + Lp * LDS;
    const __nv_bfloat16* sV = sK + Lp * LDS;
    const int b = item / heads, h = item % heads;

    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      uint32_t qa[4][4];
      load_a_frags(sQ, r0, lane, qa);
      float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
      float o[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
      const int row_a = r0 + (lane >> 2);
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      for (int kc = 0; kc < kv_end; kc += 64) {
        float sc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[j][e] = 0.f;
          if (kc + j * 8 < kv_end) mma_a_tT(sc[j], qa, sK, kc + j * 8, lane);
        }
        float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // only key tiles that cross the sequence end or the causal diagonal need the per-element predicate
          const int k_lo = kc + j * 8;
          const bool edge = (k_lo + 8 > kv_end) || (causal && k_lo + 7 > r0);  // warp-uniform
          if (!edge) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              sc[j][e] *= sl2;
              cmax[e >> 1] = fmaxf(cmax[e >> 1], sc[j][e]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int key = k_lo + (lane & 3) * 2 + (e & 1);
              const int row = row_a + (e >> 1) * 8;
              const bool ok = key < kv_end && !(causal && key > row);
              sc[j][e] = ok ? sc[j][e] * sl2 : -INFINITY;
              cmax[e >> 1] = fmaxf(cmax[e >> 1], sc[j][e]);
            }
          }
        }
        float corr[2], mref[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 1));
          cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 2));
          const float mn = fmaxf(m_i[t], cmax[t]);
          mref[t] = (mn == -INFINITY) ? 0.f : mn;
          corr[t] = ex2_approx(m_i[t] - mref[t]);  // m_i = -inf -> 0
          m_i[t] = mn;
          l_i[t] *= corr[t];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            o[j][e] *= corr[e >> 1];
            sc[j][e] = ex2_approx(sc[j][e] - mref[e >> 1]);
            l_i[e >> 1] += sc[j][e];
          }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kc + kk * 16 < kv_end) {
            uint32_t pa[4];
            pack_frag(pa, sc[2 * kk], sc[2 * kk + 1]);
            mma_p_t(o, pa, sV, kc + kk * 16, lane);
          }
        }
      }
      float inv[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        l_i[t] += __shfl_xor_sync(0xffffffffu, l_i[t], 1);
        l_i[t] += __shfl_xor_sync(0xffffffffu, l_i[t], 2);
        inv[t] = l_i[t] > 0.f ? 1.f / l_i[t] : 0.f;
      }
      // stage the 16x64 output tile in this warp's (now dead) Q rows, then write coalesced 128-byte rows
      __syncwarp();
      stage_frag_tile(sQ + r0 * LDS, o, inv[0], inv[1], lane);
      __syncwarp();
      store_tile16(sQ + r0 * LDS, out + static_cast<int64_t>(b) * seq * d + h * HD, d, r0, seq, lane);
      if ((lane & 3) == 0 && lse_out != nullptr) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int r = row_a + t * 8;
          if (r < seq) lse_out[(static_cast<int64_t>(b) * heads + h) * seq + r] = (m_i[t] + log2f(l_i[t])) * kLn2;
        }
      }
    }
    __syncthreads();  // stage s is re-filled by the prefetch issued at the top of the next iteration
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 2)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ dqkv, int items, int seq, int heads,
                     int causal, float scale, int nstages) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  const int stage_elems = 4 * Lp * LDS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sbase = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sOut = sbase + nstages * stage_elems + warp * 16 * LDS;  // per-warp output staging tile
  float* sLse = reinterpret_cast<float*>(sbase + nstages * stage_elems + nwarps * 16 * LDS);  // log2-domain LSE
  float* sD = sLse + Lp;
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const float sl2 = scale * kLog2e;

  for (int s = 0; s < nstages; ++s)
    for (int t = 0; t < 4; ++t) tile_zero_pad(sbase + s * stage_elems + t * Lp * LDS, seq, Lp);

  auto prefetch = [&](int item, int s) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    __nv_bfloat16* st = sbase + s * stage_elems;
    tile_cp_async(st, base, ld, seq);
    tile_cp_async(st + Lp * LDS, base + d, ld, seq);
    tile_cp_async(st + 2 * Lp * LDS, base + 2 * d, ld, seq);
    tile_cp_async(st + 3 * Lp * LDS, dout + static_cast<int64_t>(b) * seq * d + h * HD, d, seq);
  };

  int item = blockIdx.x;
  if (nstages == 2 && item < items) prefetch(item, 0);
  cp_async_commit();
  for (int it = 0; item < items; item += gridDim.x, ++it) {
    const int s = (nstages == 2) ? (it & 1) : 0;
    const int next = item + gridDim.x;
    if (nstages == 2) {
      if (next < items) prefetch(next, s ^ 1);
    } else {
      prefetch(item, 0);  // long sequences: one stage, no lookahead
    }
    cp_async_commit();
    const int b = item / heads, h = item % heads;
    for (int i = threadIdx.x; i < Lp; i += blockDim.x)
      sLse[i] = (i < seq) ? lse_in[(static_cast<int64_t>(b) * heads + h) * seq + i] * kLog2e : 0.f;
    if (nstages == 2) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();
    const __nv_bfloat16* sQ = sbase + s * stage_elems;
    const __nv_bfloat16* sK = sQ + Lp * LDS;
    const __nv_bfloat16* sV = sK + Lp * LDS;
    const __nv_bfloat16* sDO = sV + Lp * LDS;
    __nv_bfloat16* dq_base = dqkv + static_cast<int64_t>(b) * seq * ld + h * HD;

    // ---------------- phase 1: D[i] = sum_j P_ij * dP_ij  (warp owns 16 queries) ----------------
    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      uint32_t qa[4][4], doa[4][4];
      load_a_frags(sQ, r0, lane, qa);
      load_a_frags(sDO, r0, lane, doa);
      const int row_a = r0 + (lane >> 2);
      const float lse_r[2] = {sLse[row_a], sLse[row_a + 8]};
      float dsum[2] = {0.f, 0.f};
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      for (int k0 = 0; k0 < kv_end; k0 += 8) {
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_a_tT(sc, qa, sK, k0, lane);
        mma_a_tT(dp, doa, sV, k0, lane);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          const bool ok = key < kv_end && row < seq && !(causal && key > row);
          if (ok) dsum[e >> 1] += ex2_approx(sc[e] * sl2 - lse_r[e >> 1]) * dp[e];
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        dsum[t] += __shfl_xor_sync(0xffffffffu, dsum[t], 1);
        dsum[t] += __shfl_xor_sync(0xffffffffu, dsum[t], 2);
      }
      if ((lane & 3) == 0) {
        sD[row_a] = dsum[0];
        sD[row_a + 8] = dsum[1];
      }
    }
    __syncthreads();

    // ---------------- phase 2a: warp owns 16 queries -> dQ ----------------
    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      uint32_t qa[4][4], doa[4][4];
      load_a_frags(sQ, r0, lane, qa);
      load_a_frags(sDO, r0, lane, doa);
      const int row_a = r0 + (lane >> 2);
      const float lse_r[2] = {sLse[row_a], sLse[row_a + 8]};
      const float d_r[2] = {sD[row_a], sD[row_a + 8]};
      float dq[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      for (int k0 = 0; k0 < kv_end; k0 += 16) {
        float ds[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          if (k0 + t * 8 < kv_end) {
            mma_a_tT(sc, qa, sK, k0 + t * 8, lane);
            mma_a_tT(dp, doa, sV, k0 + t * 8, lane);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + t * 8 + (lane & 3) * 2 + (e & 1);
            const int row = row_a + (e >> 1) * 8;
            const bool ok = key < kv_end && row < seq && !(causal && key > row);
            ds[t][e] = ok ? ex2_approx(sc[e] * sl2 - lse_r[e >> 1]) * (dp[e] - d_r[e >> 1]) : 0.f;
          }
        }
        uint32_t pa[4];
        pack_frag(pa, ds[0], ds[1]);
        mma_p_t(dq, pa, sK, k0, lane);
      }
      __syncwarp();
      stage_frag_tile(sOut, dq, scale, scale, lane);
      __syncwarp();
      store_tile16(sOut, dq_base, ld, r0, seq, lane);
    }

    // ---------------- phase 2b: warp owns 16 keys -> dK, dV (transposed tiles) ----------------
    for (int c0 = warp * 16; c0 < Lp; c0 += nwarps * 16) {
      uint32_t ka[4][4], va[4][4];
      load_a_frags(sK, c0, lane, ka);
      load_a_frags(sV, c0, lane, va);
      const int key_a = c0 + (lane >> 2);
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[j][e] = dv[j][e] = 0.f;
      const int q_begin = causal ? c0 : 0;  // queries < c0 never see these keys (c0 is a multiple of 16)
      for (int q0 = q_begin; q0 < seq; q0 += 16) {
        float pt[2][4], dst[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          const bool live = q0 + t * 8 < seq;
          if (live) {
            mma_a_tT(sc, ka, sQ, q0 + t * 8, lane);
            mma_a_tT(dp, va, sDO, q0 + t * 8, lane);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int qi = q0 + t * 8 + (lane & 3) * 2 + (e & 1);
            const int key = key_a + (e >> 1) * 8;
            const bool ok = live && qi < seq && key < seq && !(causal && key > qi);
            const float p = ok ? ex2_approx(sc[e] * sl2 - sLse[qi]) : 0.f;
            pt[t][e] = p;
            dst[t][e] = ok ? p * (dp[e] - sD[qi]) : 0.f;
          }
        }
        uint32_t pa[4], da[4];
        pack_frag(pa, pt[0], pt[1]);
        pack_frag(da, dst[0], dst[1]);
        mma_p_t(dv, pa, sDO, q0, lane);
        mma_p_t(dk, da, sQ, q0, lane);
      }
      __syncwarp();
      stage_frag_tile(sOut, dk, scale, scale, lane);
      __syncwarp();
      store_tile16(sOut, dq_base + d, ld, c0, seq, lane);
      __syncwarp();
      stage_frag_tile(sOut, dv, 1.f, 1.f, lane);
      __syncwarp();
      store_tile16(sOut, dq_base + 2 * d, ld, c0, seq, lane);
    }
    __syncthreads();  // everyone is done with stage s, sLse and sD before they are overwritten
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// backward, short sequences (L <= 80: CLIP text 77, ViT-B/32 50): P and dS are computed ONCE.
// Phase A (warp = 16 queries) keeps S/dP of its whole key range in registers (NT n-tiles of 8 keys), reduces
// D = rowsum(P o dP), forms dS, accumulates dQ and parks P and dS (bf16, [query][key]) in shared memory.
// Phase B (warp = 16 keys) reads them back TRANSPOSED with ldmatrix.trans as A-operand fragments, so
// dV = P^T dO and dK = dS^T Q need no recomputed S^T / dP^T, no exps and no masks: 5 GEMM units instead of 9.
// (ncu, round 1: the recomputing kernel is issue-bound — 139 M warp instructions, 45 % issue-active, 21 % DRAM.)
// ------------------------------------------------------------------------------------------------
// PF (experimental, CLIPN_ATTN_BWD_PREFETCH=1): K and V are dead once phase A is done, so the next item's K,V tiles
// are requested before phase B and land while it runs; only Q and dO are waited for at the top of the next item.
template <int NT, bool PF>
__global__ void __launch_bounds__(160, 2)
attention_bwd_small_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                           const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ dqkv,
                           float* __restrict__ dbias, int items, int seq, int heads, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  constexpr int Lp = NT * 8;    // 64 or 80
  constexpr int LDP = Lp + 8;   // pitch of the P / dS tiles (bf16): conflict-free ldmatrix rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sK = sQ + Lp * LDS;
  __nv_bfloat16* sV = sK + Lp * LDS;
  __nv_bfloat16* sDO = sV + Lp * LDS;
  __nv_bfloat16* sP = sDO + Lp * LDS;
  __nv_bfloat16* sdS = sP + Lp * LDP;
  __nv_bfloat16* sOut = sdS + Lp * LDP + warp * 16 * LDS;
  float* sLse = reinterpret_cast<float*>(sdS + Lp * LDP + nwarps * 16 * LDS);
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const float sl2 = scale * kLog2e;

  for (int t = 0; t < 4; ++t) tile_zero_pad(sQ + t * Lp * LDS, seq, Lp);

  auto load_kv = [&](int it) {
    const __nv_bfloat16* kb = qkv + static_cast<int64_t>(it / heads) * seq * ld + (it % heads) * HD;
    tile_cp_async(sK, kb + d, ld, seq);
    tile_cp_async(sV, kb + 2 * d, ld, seq);
  };
  if constexpr (PF) {
    if (static_cast<int>(blockIdx.x) < items) load_kv(blockIdx.x);
    cp_async_commit();
  }

  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    tile_cp_async(sQ, base, ld, seq);
    if constexpr (!PF) {
      tile_cp_async(sK, base + d, ld, seq);
      tile_cp_async(sV, base + 2 * d, ld, seq);
    }
    tile_cp_async(sDO, dout + static_cast<int64_t>(b) * seq * d + h * HD, d, seq);
    cp_async_commit();
    for (int i = threadIdx.x; i < Lp; i += blockDim.x)
      sLse[i] = (i < seq) ? lse_in[(static_cast<int64_t>(b) * heads + h) * seq + i] * kLog2e : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    __nv_bfloat16* dq_base = dqkv + static_cast<int64_t>(b) * seq * ld + h * HD;

    // ---------------- phase A: warp owns 16 queries ----------------
    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      uint32_t qa[4][4], doa[4][4];
      load_a_frags(sQ, r0, lane, qa);
      load_a_frags(sDO, r0, lane, doa);
      const int row_a = r0 + (lane >> 2);
      const float lse_r[2] = {sLse[row_a], sLse[row_a + 8]};
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      float p[NT][4], dp[NT][4];
      float dsum[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[j][e] = 0.f;
        if (j * 8 < kv_end) {
          mma_a_tT(sc, qa, sK, j * 8, lane);
          mma_a_tT(dp[j], doa, sV, j * 8, lane);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = j * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          const bool ok = key < kv_end && row < seq && !(causal && key > row);
          p[j][e] = ok ? ex2_approx(sc[e] * sl2 - lse_r[e >> 1]) : 0.f;
          dsum[e >> 1] += p[j][e] * dp[j][e];
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        dsum[t] += __shfl_xor_sync(0xffffffffu, dsum[t], 1);
        dsum[t] += __shfl_xor_sync(0xffffffffu, dsum[t], 2);
      }
      float dq[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) ds[e] = p[j][e] * (dp[j][e] - dsum[e >> 1]);
        const int c = j * 8 + (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(sP + row_a * LDP + c) = pack_bf16x2(p[j][0], p[j][1]);
        *reinterpret_cast<uint32_t*>(sP + (row_a + 8) * LDP + c) = pack_bf16x2(p[j][2], p[j][3]);
        *reinterpret_cast<uint32_t*>(sdS + row_a * LDP + c) = pack_bf16x2(ds[0], ds[1]);
        *reinterpret_cast<uint32_t*>(sdS + (row_a + 8) * LDP + c) = pack_bf16x2(ds[2], ds[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[j][e] = ds[e];  // dp now holds dS for the dQ MMAs
      }
#pragma unroll
      for (int kk = 0; kk < NT / 2; ++kk) {
        if (kk * 16 < kv_end) {
          uint32_t pa[4];
          pack_frag(pa, dp[2 * kk], dp[2 * kk + 1]);
          mma_p_t(dq, pa, sK, kk * 16, lane);
        }
      }
      __syncwarp();
      stage_frag_tile(sOut, dq, scale, scale, lane);
      __syncwarp();
      store_tile16(sOut, dq_base, ld, r0, seq, lane);
      if (dbias != nullptr) tile_colsum_atomic(sOut, dbias + h * HD, lane);
    }
    __syncthreads();
    if constexpr (PF) {  // sK / sV are not read again for this item
      if (item + static_cast<int>(gridDim.x) < items) load_kv(item + gridDim.x);
      cp_async_commit();
    }

    // ---------------- phase B: warp owns 16 keys; P^T and dS^T come from shared memory ----------------
    for (int c0 = warp * 16; c0 < Lp; c0 += nwarps * 16) {
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[j][e] = dv[j][e] = 0.f;
      const int q_begin = causal ? c0 : 0;
      const int mi = lane >> 3, ri = lane & 7;
      for (int q0 = q_begin; q0 < seq; q0 += 16) {
        uint32_t pa[4], da[4];
        const int off = (q0 + (mi >> 1) * 8 + ri) * LDP + c0 + (mi & 1) * 8;
        ldmatrix_x4_trans(pa, sP + off);
        ldmatrix_x4_trans(da, sdS + off);
        mma_p_t(dv, pa, sDO, q0, lane);
        mma_p_t(dk, da, sQ, q0, lane);
      }
      __syncwarp();
      stage_frag_tile(sOut, dk, scale, scale, lane);
      __syncwarp();
      store_tile16(sOut, dq_base + d, ld, c0, seq, lane);
      if (dbias != nullptr) tile_colsum_atomic(sOut, dbias + d + h * HD, lane);
      __syncwarp();
      stage_frag_tile(sOut, dv, 1.f, 1.f, lane);
      __syncwarp();
      store_tile16(sOut, dq_base + 2 * d, ld, c0, seq, lane);
      if (dbias != nullptr) tile_colsum_atomic(sOut, dbias + 2 * d + h * HD, lane);
    }
    __syncthreads();  // tiles, sP/sdS and sLse are rewritten by the next item
  }
  if constexpr (PF) cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Long sequences (L > 384, e.g. ViT-L/14-336: 577 tokens): the whole Q/K/V/dO set no longer fits in shared memory.
// Forward and the dQ pass keep K,V of the (batch, head) item resident and stream 16-row Q (and dO) tiles through a
// per-warp staging tile; the dK/dV pass keeps Q,dO resident and streams K,V tiles.  D = rowsum(dO o O) is taken
// from the forward output (one extra tile read) instead of a P.dP sweep.  One CTA (8 warps) per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_load(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t src_ld, int r0,
                                               int seq, int lane) {
  for (int i = lane; i < 16 * 8; i += 32) {
    const int r = i >> 3, v = i & 7;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < seq) val = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r0 + r) * src_ld + v * 8);
    *reinterpret_cast<uint4*>(dst + r * LDS + v * 8) = val;
  }
}
// D[r] = sum_c dO[r,c] * O[r,c] for the 16 rows of a staged dO tile (O read from global); result in sDw[0..15]
__device__ __forceinline__ void warp_tile_rowdot(const __nv_bfloat16* sDOt, const __nv_bfloat16* o_base, int64_t o_ld,
                                                 int r0, int seq, float* sDw, int lane) {
  const int r = lane >> 1, hf = lane & 1;
  float acc = 0.f;
  if (r0 + r < seq) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float a[8], c[8];
      unpack_bf16x8(*reinterpret_cast<const uint4*>(o_base + static_cast<int64_t>(r0 + r) * o_ld + hf * 32 + v * 8), a);
      unpack_bf16x8(*reinterpret_cast<const uint4*>(sDOt + r * LDS + hf * 32 + v * 8), c);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += a[j] * c[j];
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (hf == 0) sDw[r] = acc;
}

__global__ void __launch_bounds__(256, 1)
attention_fwd_long_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                          float* __restrict__ lse_out, int items, int seq, int heads, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sV = sK + Lp * LDS;
  __nv_bfloat16* sQt = sV + Lp * LDS + warp * 16 * LDS;  // this warp's Q tile, later its output staging tile
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const float sl2 = scale * kLog2e;
  tile_zero_pad(sK, seq, Lp);
  tile_zero_pad(sV, seq, Lp);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    tile_cp_async(sK, base + d, ld, seq);
    tile_cp_async(sV, base + 2 * d, ld, seq);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      warp_tile_load(sQt, base, ld, r0, seq, lane);
      __syncwarp();
      uint32_t qa[4][4];
      load_a_frags(sQt, 0, lane, qa);
      float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
      float o[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
      const int row_a = r0 + (lane >> 2);
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      for (int kc = 0; kc < kv_end; kc += 64) {
        float sc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[j][e] = 0.f;
          if (kc + j * 8 < kv_end) mma_a_tT(sc[j], qa, sK, kc + j * 8, lane);
        }
        float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = kc + j * 8 + (lane & 3) * 2 + (e & 1);
            const int row = row_a + (e >> 1) * 8;
            const bool ok = key < kv_end && !(causal && key > row);
            sc[j][e] = ok ? sc[j][e] * sl2 : -INFINITY;
            cmax[e >> 1] = fmaxf(cmax[e >> 1], sc[j][e]);
          }
        float corr[2], mref[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 1));
          cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 2));
          const float mn = fmaxf(m_i[t], cmax[t]);
          mref[t] = (mn == -INFINITY) ? 0.f : mn;
          corr[t] = ex2_approx(m_i[t] - mref[t]);
          m_i[t] = mn;
          l_i[t] *= corr[t];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            o[j][e] *= corr[e >> 1];
            sc[j][e] = ex2_approx(sc[j][e] - mref[e >> 1]);
            l_i[e >> 1] += sc[j][e];
          }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kc + kk * 16 < kv_end) {
            uint32_t pa[4];
            pack_frag(pa, sc[2 * kk], sc[2 * kk + 1]);
            mma_p_t(o, pa, sV, kc + kk * 16, lane);
          }
        }
      }
      float inv[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        l_i[t] += __shfl_xor_sync(0xffffffffu, l_i[t], 1);
        l_i[t] += __shfl_xor_sync(0xffffffffu, l_i[t], 2);
        inv[t] = l_i[t] > 0.f ? 1.f / l_i[t] : 0.f;
      }
      __syncwarp();
      stage_frag_tile(sQt, o, inv[0], inv[1], lane);
      __syncwarp();
      store_tile16(sQt, out + static_cast<int64_t>(b) * seq * d + h * HD, d, r0, seq, lane);
      if ((lane & 3) == 0 && lse_out != nullptr) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int r = row_a + t * 8;
          if (r < seq) lse_out[(static_cast<int64_t>(b) * heads + h) * seq + r] = (m_i[t] + log2f(l_i[t])) * kLn2;
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// dQ pass: K,V resident; per warp: Q and dO tiles staged, D from dO o O, S/dP recomputed per 16-key slab.
__global__ void __launch_bounds__(256, 1)
attention_bwd_long_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                             const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse_in,
                             __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, int items, int seq, int heads,
                             int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sV = sK + Lp * LDS;
  __nv_bfloat16* sQt = sV + Lp * LDS + warp * 32 * LDS;
  __nv_bfloat16* sDOt = sQt + 16 * LDS;
  float* sDw = reinterpret_cast<float*>(sV + Lp * LDS + nwarps * 32 * LDS) + warp * 16;
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const float sl2 = scale * kLog2e;
  tile_zero_pad(sK, seq, Lp);
  tile_zero_pad(sV, seq, Lp);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    const __nv_bfloat16* obase = out + static_cast<int64_t>(b) * seq * d + h * HD;
    const __nv_bfloat16* dobase = dout + static_cast<int64_t>(b) * seq * d + h * HD;
    const float* lse_b = lse_in + (static_cast<int64_t>(b) * heads + h) * seq;
    __nv_bfloat16* dq_base = dqkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    tile_cp_async(sK, base + d, ld, seq);
    tile_cp_async(sV, base + 2 * d, ld, seq);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int r0 = warp * 16; r0 < Lp; r0 += nwarps * 16) {
      warp_tile_load(sQt, base, ld, r0, seq, lane);
      warp_tile_load(sDOt, dobase, d, r0, seq, lane);
      __syncwarp();
      warp_tile_rowdot(sDOt, obase, d, r0, seq, sDw, lane);
      __syncwarp();
      uint32_t qa[4][4], doa[4][4];
      load_a_frags(sQt, 0, lane, qa);
      load_a_frags(sDOt, 0, lane, doa);
      const int row_a = r0 + (lane >> 2);
      const float lse_r[2] = {row_a < seq ? lse_b[row_a] * kLog2e : 0.f, row_a + 8 < seq ? lse_b[row_a + 8] * kLog2e : 0.f};
      const float d_r[2] = {sDw[lane >> 2], sDw[(lane >> 2) + 8]};
      float dq[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
      int kv_end = seq;
      if (causal && r0 + 16 < kv_end) kv_end = r0 + 16;
      for (int k0 = 0; k0 < kv_end; k0 += 16) {
        float ds[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          if (k0 + t * 8 < kv_end) {
            mma_a_tT(sc, qa, sK, k0 + t * 8, lane);
            mma_a_tT(dp, doa, sV, k0 + t * 8, lane);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + t * 8 + (lane & 3) * 2 + (e & 1);
            const int row = row_a + (e >> 1) * 8;
            const bool ok = key < kv_end && row < seq && !(causal && key > row);
            ds[t][e] = ok ? ex2_approx(sc[e] * sl2 - lse_r[e >> 1]) * (dp[e] - d_r[e >> 1]) : 0.f;
          }
        }
        uint32_t pa[4];
        pack_frag(pa, ds[0], ds[1]);
        mma_p_t(dq, pa, sK, k0, lane);
      }
      __syncwarp();
      stage_frag_tile(sQt, dq, scale, scale, lane);
      __syncwarp();
      store_tile16(sQt, dq_base, ld, r0, seq, lane);
      if (dbias != nullptr) tile_colsum_atomic(sQt, dbias + h * HD, lane);
      __syncwarp();
    }
    __syncthreads();
  }
}

// dK/dV pass: Q,dO (and D, LSE) resident; per warp: K and V tiles staged; transposed tiles as in the general kernel.
__global__ void __launch_bounds__(256, 1)
attention_bwd_long_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                              const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse_in,
                              __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, int items, int seq, int heads,
                              int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sDO = sQ + Lp * LDS;
  __nv_bfloat16* sKt = sDO + Lp * LDS + warp * 32 * LDS;
  __nv_bfloat16* sVt = sKt + 16 * LDS;
  float* sLse = reinterpret_cast<float*>(sDO + Lp * LDS + nwarps * 32 * LDS);
  float* sD = sLse + Lp;
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const float sl2 = scale * kLog2e;
  tile_zero_pad(sQ, seq, Lp);
  tile_zero_pad(sDO, seq, Lp);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / heads, h = item % heads;
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    const __nv_bfloat16* obase = out + static_cast<int64_t>(b) * seq * d + h * HD;
    const __nv_bfloat16* dobase = dout + static_cast<int64_t>(b) * seq * d + h * HD;
    __nv_bfloat16* dq_base = dqkv + static_cast<int64_t>(b) * seq * ld + h * HD;
    tile_cp_async(sQ, base, ld, seq);
    tile_cp_async(sDO, dobase, d, seq);
    cp_async_commit();
    for (int i = threadIdx.x; i < Lp; i += blockDim.x)
      sLse[i] = (i < seq) ? lse_in[(static_cast<int64_t>(b) * heads + h) * seq + i] * kLog2e : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    for (int i = threadIdx.x; i < Lp * 8; i += blockDim.x) {  // D[r] = sum_c dO[r,c] O[r,c]  (8 lanes per row)
      const int r = i >> 3, v = i & 7;
      float acc = 0.f;
      if (r < seq) {
        float a[8], c[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(obase + static_cast<int64_t>(r) * d + v * 8), a);
        unpack_bf16x8(*reinterpret_cast<const uint4*>(sDO + r * LDS + v * 8), c);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += a[j] * c[j];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (v == 0) sD[r] = acc;
    }
    __syncthreads();
    for (int c0 = warp * 16; c0 < Lp; c0 += nwarps * 16) {
      warp_tile_load(sKt, base + d, ld, c0, seq, lane);
      warp_tile_load(sVt, base + 2 * d, ld, c0, seq, lane);
      __syncwarp();
      uint32_t ka[4][4], va[4][4];
      load_a_frags(sKt, 0, lane, ka);
      load_a_frags(sVt, 0, lane, va);
      const int key_a = c0 + (lane >> 2);
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[j][e] = dv[j][e] = 0.f;
      const int q_begin = causal ? c0 : 0;
      for (int q0 = q_begin; q0 < seq; q0 += 16) {
        float pt[2][4], dst[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          const bool live = q0 + t * 8 < seq;
          if (live) {
            mma_a_tT(sc, ka, sQ, q0 + t * 8, lane);
            mma_a_tT(dp, va, sDO, q0 + t * 8, lane);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int qi = q0 + t * 8 + (lane & 3) * 2 + (e & 1);
            const int key = key_a + (e >> 1) * 8;
            const bool ok = live && qi < seq && key < seq && !(causal && key > qi);
            const float p = ok ? ex2_approx(sc[e] * sl2 - sLse[qi]) : 0.f;
            pt[t][e] = p;
            dst[t][e] = ok ? p * (dp[e] - sD[qi]) : 0.f;
          }
        }
        uint32_t pa[4], da[4];
        pack_frag(pa, pt[0], pt[1]);
        pack_frag(da, dst[0], dst[1]);
        mma_p_t(dv, pa, sDO, q0, lane);
        mma_p_t(dk, da, sQ, q0, lane);
      }
      __syncwarp();
      stage_frag_tile(sKt, dk, scale, scale, lane);
      stage_frag_tile(sVt, dv, 1.f, 1.f, lane);
      __syncwarp();
      store_tile16(sKt, dq_base + d, ld, c0, seq, lane);
      store_tile16(sVt, dq_base + 2 * d, ld, c0, seq, lane);
      if (dbias != nullptr) {
        tile_colsum_atomic(sKt, dbias + d + h * HD, lane);
        tile_colsum_atomic(sVt, dbias + 2 * d + h * HD, lane);
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

static int pick_warps(int seq) {
  int tiles = (seq + 15) / 16;
  return tiles < 8 ? tiles : 8;
}

}  // namespace clipn

using namespace clipn;

namespace clipn {
// attention_tc.cu: tcgen05 / TMEM / TMA forward (default); CLIPN_ATTN_TC=0 selects the mma.sync kernels below
bool attention_tc_enabled();
int attention_tc_fwd(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int causal, float scale,
                     cudaStream_t stream);
int attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias,
                     int batch, int seq, int heads, int causal, float scale, cudaStream_t stream);
}  // namespace clipn

extern "C" int clipn_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t seq, int32_t heads,
                                   int32_t causal, float scale, clipn_stream_t stream) {
  CLIPN_REQUIRE(qkv && out, "attention_fwd: null pointer");
  CLIPN_REQUIRE(seq > 0 && heads > 0, "attention_fwd: bad dims");
  if (batch <= 0) return CLIPN_OK;
  if (clipn::attention_tc_enabled())
    return clipn::attention_tc_fwd(qkv, out, lse, batch, seq, heads, causal, scale, static_cast<cudaStream_t>(stream));
  const int Lp = (seq + 15) & ~15;
  const size_t stage_bytes = static_cast<size_t>(3) * Lp * LDS * 2;
  const int nst = (2 * stage_bytes <= 113 * 1024) ? 2 : 1;  // double-buffer when two CTAs still fit per SM
  const size_t smem = nst * stage_bytes;
  const int items = batch * heads;
  if (smem > 227 * 1024) {  // K,V resident, Q streamed per warp
    const size_t smem_long = (static_cast<size_t>(2) * Lp + 8 * 16) * LDS * 2;
    CLIPN_REQUIRE(smem_long <= 227 * 1024 && seq <= kLongSeqMax, "attention_fwd: sequence too long (L <= 640)");
    int grid_long = num_sms() < items ? num_sms() : items;
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_fwd_long_kernel<<<grid_long, 256, smem_long, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse, items, seq, heads, causal,
        scale);
    CLIPN_CHECK_CUDA(cudaGetLastError());
    return CLIPN_OK;
  }
  const int nw = pick_warps(seq);
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  // residency is register-bound (128 regs/thread): 4 CTAs of 4 warps (L <= 64), 3 of 5, 2 of 8
  const int cap = nw <= 4 ? 4 : (nw <= 5 ? 3 : 2);
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > items) grid = items;
  auto st = static_cast<cudaStream_t>(stream);
  if (nw <= 5) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_fwd_kernel<5><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<__nv_bfloat16*>(out), lse, items, seq, heads,
                                                         causal, scale, nst);
  } else {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_fwd_kernel<8><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<__nv_bfloat16*>(out), lse, items, seq, heads,
                                                         causal, scale, nst);
  }
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                   float* dbias, int32_t batch, int32_t seq, int32_t heads, int32_t causal, float scale,
                                   clipn_stream_t stream) {
  // L <= 384: D = rowsum(dO o O) is recomputed as sum_j P_ij dP_ij and `out` is not read; longer sequences read it
  CLIPN_REQUIRE(qkv && dout && lse && dqkv, "attention_bwd: null pointer");
  CLIPN_REQUIRE(seq > 0 && heads > 0, "attention_bwd: bad dims");
  if (batch <= 0) return CLIPN_OK;
  if (clipn::attention_tc_enabled() && out != nullptr)
    return clipn::attention_tc_bwd(qkv, out, dout, lse, dqkv, dbias, batch, seq, heads, causal, scale,
                                   static_cast<cudaStream_t>(stream));
  const int Lp = (seq + 15) & ~15;
  const int nw = pick_warps(seq);
  static const bool small_enabled = [] {
    const char* e = getenv("CLIPN_ATTN_BWD_SMALL");
    return !(e != nullptr && e[0] == '0');
  }();
  if (seq <= 80 && small_enabled) {
    const int lp = seq <= 64 ? 64 : 80;
    const size_t smem = (static_cast<size_t>(4) * lp * LDS + static_cast<size_t>(2) * lp * (lp + 8) +
                         static_cast<size_t>(nw) * 16 * LDS) * 2 + static_cast<size_t>(lp) * sizeof(float);
    const int items = batch * heads;
    int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
    const int reg_cap = nw <= 4 ? 3 : 2;  // ~168 regs/thread: three 4-warp CTAs (L <= 64) or two 5-warp CTAs fit 64K regs
    if (per_sm > reg_cap) per_sm = reg_cap;
    int grid = num_sms() * per_sm;
    if (grid > items) grid = items;
    auto st = static_cast<cudaStream_t>(stream);
    static const bool prefetch_kv = [] {  // experimental, off by default (not yet measured on hardware)
      const char* e = getenv("CLIPN_ATTN_BWD_PREFETCH");
      return e != nullptr && e[0] == '1';
    }();
    using SmallKernel = void (*)(const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, float*, int, int,
                                 int, int, float);
    SmallKernel kern;
    if (lp == 64) kern = prefetch_kv ? attention_bwd_small_kernel<8, true> : attention_bwd_small_kernel<8, false>;
    else kern = prefetch_kv ? attention_bwd_small_kernel<10, true> : attention_bwd_small_kernel<10, false>;
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, nw * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                      reinterpret_cast<const __nv_bfloat16*>(dout), lse,
                                      reinterpret_cast<__nv_bfloat16*>(dqkv), dbias, items, seq, heads, causal, scale);
    CLIPN_CHECK_CUDA(cudaGetLastError());
    return CLIPN_OK;
  }
  const size_t stage_bytes = static_cast<size_t>(4) * Lp * LDS * 2;
  const size_t extra = static_cast<size_t>(nw) * 16 * LDS * 2 + static_cast<size_t>(2) * Lp * sizeof(float);
  const int nst = (2 * stage_bytes + extra <= 113 * 1024) ? 2 : 1;
  const size_t smem = nst * stage_bytes + extra;
  const int items = batch * heads;
  // The two-pass path is also the faster one well below the smem limit (B200, B=256 L=197 H=12: 1344 us vs 1659 us
  // for the single-kernel recompute schedule), so it is the default from L = 192 up. CLIPN_ATTN_BWD_TWO_PASS_MIN moves it.
  static const int two_pass_min_seq = [] {
    const char* e = getenv("CLIPN_ATTN_BWD_TWO_PASS_MIN");
    return e != nullptr ? atoi(e) : 192;
  }();
  if (smem > 227 * 1024 || (seq >= two_pass_min_seq && out != nullptr)) {
    // two passes: dQ with K,V resident, then dK/dV with Q,dO resident; D from dO o O
    CLIPN_REQUIRE(out != nullptr, "attention_bwd: the long-sequence path needs the forward output");
    const size_t smem_dq = (static_cast<size_t>(2) * Lp + 8 * 32) * LDS * 2 + 8 * 16 * sizeof(float);
    const size_t smem_dkv = (static_cast<size_t>(2) * Lp + 8 * 32) * LDS * 2 + static_cast<size_t>(2) * Lp * sizeof(float);
    CLIPN_REQUIRE(smem_dq <= 227 * 1024 && smem_dkv <= 227 * 1024 && seq <= kLongSeqMax,
                  "attention_bwd: sequence too long (L <= 640)");
    const int grid_long = num_sms() < items ? num_sms() : items;
    auto stl = static_cast<cudaStream_t>(stream);
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_long_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_long_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_bwd_long_dq_kernel<<<grid_long, 256, smem_dq, stl>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
        reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), nullptr, items, seq,
        heads, causal, scale);
    attention_bwd_long_dkv_kernel<<<grid_long, 256, smem_dkv, stl>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
        reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), nullptr, items, seq,
        heads, causal, scale);
    CLIPN_CHECK_CUDA(cudaGetLastError());
    if (dbias != nullptr)
      return clipn_colsum(dqkv, 3 * static_cast<int64_t>(heads) * HD, dbias, static_cast<int64_t>(batch) * seq,
                          3 * heads * HD, stream);
    return CLIPN_OK;
  }
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  if (grid > items) grid = items;
  auto st = static_cast<cudaStream_t>(stream);
  if (nw <= 5) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_bwd_kernel<5><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<const __nv_bfloat16*>(dout), lse,
                                                         reinterpret_cast<__nv_bfloat16*>(dqkv), items, seq, heads, causal,
                                                         scale, nst);
  } else {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attention_bwd_kernel<8><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<const __nv_bfloat16*>(dout), lse,
                                                         reinterpret_cast<__nv_bfloat16*>(dqkv), items, seq, heads, causal,
                                                         scale, nst);
  }
  CLIPN_CHECK_CUDA(cudaGetLastError());
  // the general (long-sequence) kernel does not fuse the in_proj_bias gradient: one extra pass over dqkv
  if (dbias != nullptr)
    return clipn_colsum(dqkv, 3 * static_cast<int64_t>(heads) * HD, dbias, static_cast<int64_t>(batch) * seq,
                        3 * heads * HD, stream);
  return CLIPN_OK;
}
