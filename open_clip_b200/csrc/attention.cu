// Attention core for the CLIP towers (F.scaled_dot_product_attention, transformer.py:223-228), head_dim 64.
//
// Sequences are short (50 / 77 / 197 tokens) so one CTA owns one (batch, head): Q, K, V (and dO in the
// backward) live in shared memory for the whole CTA, each warp owns 16-row tiles, S = QK^T and PV run on
// tensor cores (mma.sync m16n8k16 bf16 -> fp32; the tiles are far below the 128-row tcgen05 atom and the
// op is 1.6 % of the step's FLOPs — it is HBM-bound), softmax is a warp-shuffle (quad)
// reduction in registers with online rescaling, the causal mask is a predicate (no mask tensor).
// Backward recomputes P from the saved log-sum-exp; pass A (warp = 16 queries) produces dQ, pass B
// (warp = 16 keys, transposed tiles) produces dK and dV, so there are no atomics and the result is
// deterministic.
#include "common.cuh"

namespace clipn {

constexpr int HD = 64;       // head dim
constexpr int LDS = 72;      // padded smem row (bf16 elements): 144 B => conflict-free fragment loads
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
// For a 16-deep k-slab starting at smem row k0 of T (row-major [k][64]): out[jn] += Pa(16x16) * T[k0..k0+15][jn*8..]
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

__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t src_ld, int rows,
                                          int rows_pad) {
  for (int i = threadIdx.x; i < rows_pad * 8; i += blockDim.x) {
    const int r = i >> 3, v = i & 7;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows) val = *reinterpret_cast<const uint4*>(src + r * src_ld + v * 8);
    *reinterpret_cast<uint4*>(dst + r * LDS + v * 8) = val;
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out,
                                                            int seq, int heads, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sK = sQ + Lp * LDS;
  __nv_bfloat16* sV = sK + Lp * LDS;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
  load_tile(sQ, base, ld, seq, Lp);
  load_tile(sK, base + d, ld, seq, Lp);
  load_tile(sV, base + 2 * d, ld, seq, Lp);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float sl2 = scale * kLog2e;
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
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
        if (kc + j * 8 < kv_end) mma_a_tT(s[j], qa, sK, kc + j * 8, lane);
      }
      float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kc + j * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          const bool ok = key < kv_end && key < seq && !(causal && key > row);
          s[j][e] = ok ? s[j][e] * sl2 : -INFINITY;
          cmax[e >> 1] = fmaxf(cmax[e >> 1], s[j][e]);
        }
      float corr[2], mref[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 1));
        cmax[t] = fmaxf(cmax[t], __shfl_xor_sync(0xffffffffu, cmax[t], 2));
        const float mn = fmaxf(m_i[t], cmax[t]);
        mref[t] = (mn == -INFINITY) ? 0.f : mn;
        corr[t] = exp2f(m_i[t] - mref[t]);  // m_i = -inf -> 0
        m_i[t] = mn;
        l_i[t] *= corr[t];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[j][e] *= corr[e >> 1];
          s[j][e] = exp2f(s[j][e] - mref[e >> 1]);
          l_i[e >> 1] += s[j][e];
        }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kc + kk * 16 < kv_end) {
          uint32_t pa[4];
          pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
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
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + (lane & 3) * 2;
      *reinterpret_cast<uint32_t*>(sQ + row_a * LDS + c) = pack_bf16x2(o[j][0] * inv[0], o[j][1] * inv[0]);
      *reinterpret_cast<uint32_t*>(sQ + (row_a + 8) * LDS + c) = pack_bf16x2(o[j][2] * inv[1], o[j][3] * inv[1]);
    }
    __syncwarp();
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = r0 + (i >> 3), v = i & 7;
      if (r < seq)
        *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * seq + r) * d + h * HD + v * 8) =
            *reinterpret_cast<const uint4*>(sQ + r * LDS + v * 8);
    }
    if ((lane & 3) == 0 && lse_out != nullptr) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r = row_a + t * 8;
        if (r < seq)
          lse_out[(static_cast<int64_t>(b) * heads + h) * seq + r] = (m_i[t] + log2f(l_i[t])) * kLn2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ out,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            const float* __restrict__ lse_in,
                                                            __nv_bfloat16* __restrict__ dqkv, int seq, int heads,
                                                            int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int Lp = (seq + 15) & ~15;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);
  __nv_bfloat16* sK = sQ + Lp * LDS;
  __nv_bfloat16* sV = sK + Lp * LDS;
  __nv_bfloat16* sDO = sV + Lp * LDS;
  float* sLse = reinterpret_cast<float*>(sDO + Lp * LDS);  // log2-domain LSE
  float* sD = sLse + Lp;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int d = heads * HD;
  const int64_t ld = 3 * static_cast<int64_t>(d);
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * seq * ld + h * HD;
  const __nv_bfloat16* obase = out + static_cast<int64_t>(b) * seq * d + h * HD;
  const __nv_bfloat16* dobase = dout + static_cast<int64_t>(b) * seq * d + h * HD;
  load_tile(sQ, base, ld, seq, Lp);
  load_tile(sK, base + d, ld, seq, Lp);
  load_tile(sV, base + 2 * d, ld, seq, Lp);
  load_tile(sDO, dobase, d, seq, Lp);
  for (int i = threadIdx.x; i < Lp; i += blockDim.x)
    sLse[i] = (i < seq) ? lse_in[(static_cast<int64_t>(b) * heads + h) * seq + i] * kLog2e : 0.f;
  __syncthreads();
  // D[r] = sum_c dO[r,c] * O[r,c]  (8 lanes per row)
  for (int i = threadIdx.x; i < Lp * 8; i += blockDim.x) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float sl2 = scale * kLog2e;
  __nv_bfloat16* dq_base = dqkv + static_cast<int64_t>(b) * seq * ld + h * HD;

  // ---------------- pass A: warp owns 16 queries -> dQ ----------------
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
    for (int kc = 0; kc < kv_end; kc += 64) {
      float ds[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        if (kc + j * 8 < kv_end) {
          mma_a_tT(s, qa, sK, kc + j * 8, lane);
          mma_a_tT(dp, doa, sV, kc + j * 8, lane);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kc + j * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          const bool ok = key < kv_end && key < seq && row < seq && !(causal && key > row);
          const float p = ok ? exp2f(s[e] * sl2 - lse_r[e >> 1]) : 0.f;
          ds[j][e] = p * (dp[e] - d_r[e >> 1]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kc + kk * 16 < kv_end) {
          uint32_t pa[4];
          pa[0] = pack_bf16x2(ds[2 * kk][0], ds[2 * kk][1]);
          pa[1] = pack_bf16x2(ds[2 * kk][2], ds[2 * kk][3]);
          pa[2] = pack_bf16x2(ds[2 * kk + 1][0], ds[2 * kk + 1][1]);
          pa[3] = pack_bf16x2(ds[2 * kk + 1][2], ds[2 * kk + 1][3]);
          mma_p_t(dq, pa, sK, kc + kk * 16, lane);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + (lane & 3) * 2;
      if (row_a < seq)
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(row_a) * ld + c) =
            pack_bf16x2(dq[j][0] * scale, dq[j][1] * scale);
      if (row_a + 8 < seq)
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(row_a + 8) * ld + c) =
            pack_bf16x2(dq[j][2] * scale, dq[j][3] * scale);
    }
  }

  // ---------------- pass B: warp owns 16 keys -> dK, dV (transposed tiles) ----------------
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
    const int q_begin = causal ? (c0 & ~63) : 0;  // queries < c0 never see these keys
    for (int qc = q_begin; qc < seq; qc += 64) {
      float pt[8][4], dst[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        const bool live = (qc + j * 8 < seq) && !(causal && qc + j * 8 + 7 < c0);
        if (live) {
          mma_a_tT(s, ka, sQ, qc + j * 8, lane);
          mma_a_tT(dp, va, sDO, qc + j * 8, lane);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = qc + j * 8 + (lane & 3) * 2 + (e & 1);
          const int key = key_a + (e >> 1) * 8;
          const bool ok = live && qi < seq && key < seq && !(causal && key > qi);
          const float p = ok ? exp2f(s[e] * sl2 - sLse[qi < Lp ? qi : 0]) : 0.f;
          pt[j][e] = p;
          dst[j][e] = ok ? p * (dp[e] - sD[qi]) : 0.f;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (qc + kk * 16 < seq) {
          uint32_t pa[4], da[4];
          pa[0] = pack_bf16x2(pt[2 * kk][0], pt[2 * kk][1]);
          pa[1] = pack_bf16x2(pt[2 * kk][2], pt[2 * kk][3]);
          pa[2] = pack_bf16x2(pt[2 * kk + 1][0], pt[2 * kk + 1][1]);
          pa[3] = pack_bf16x2(pt[2 * kk + 1][2], pt[2 * kk + 1][3]);
          da[0] = pack_bf16x2(dst[2 * kk][0], dst[2 * kk][1]);
          da[1] = pack_bf16x2(dst[2 * kk][2], dst[2 * kk][3]);
          da[2] = pack_bf16x2(dst[2 * kk + 1][0], dst[2 * kk + 1][1]);
          da[3] = pack_bf16x2(dst[2 * kk + 1][2], dst[2 * kk + 1][3]);
          mma_p_t(dv, pa, sDO, qc + kk * 16, lane);
          mma_p_t(dk, da, sQ, qc + kk * 16, lane);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + (lane & 3) * 2;
      if (key_a < seq) {
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(key_a) * ld + d + c) =
            pack_bf16x2(dk[j][0] * scale, dk[j][1] * scale);
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(key_a) * ld + 2 * d + c) = pack_bf16x2(dv[j][0], dv[j][1]);
      }
      if (key_a + 8 < seq) {
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(key_a + 8) * ld + d + c) =
            pack_bf16x2(dk[j][2] * scale, dk[j][3] * scale);
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<int64_t>(key_a + 8) * ld + 2 * d + c) =
            pack_bf16x2(dv[j][2], dv[j][3]);
      }
    }
  }
}

static int pick_warps(int seq) {
  int tiles = (seq + 15) / 16;
  return tiles < 8 ? tiles : 8;
}

}  // namespace clipn

using namespace clipn;

extern "C" int clipn_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t seq, int32_t heads,
                                   int32_t causal, float scale, clipn_stream_t stream) {
  CLIPN_REQUIRE(qkv && out, "attention_fwd: null pointer");
  CLIPN_REQUIRE(seq > 0 && heads > 0, "attention_fwd: bad dims");
  if (batch <= 0) return CLIPN_OK;
  const int Lp = (seq + 15) & ~15;
  const size_t smem = static_cast<size_t>(3) * Lp * LDS * 2;
  CLIPN_REQUIRE(smem <= 227 * 1024, "attention_fwd: sequence too long for the single-CTA kernel (L <= 512)");
  CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attention_fwd_kernel<<<batch * heads, pick_warps(seq) * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse, seq, heads, causal, scale);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

extern "C" int clipn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                   int32_t batch, int32_t seq, int32_t heads, int32_t causal, float scale,
                                   clipn_stream_t stream) {
  CLIPN_REQUIRE(qkv && out && dout && lse && dqkv, "attention_bwd: null pointer");
  CLIPN_REQUIRE(seq > 0 && heads > 0, "attention_bwd: bad dims");
  if (batch <= 0) return CLIPN_OK;
  const int Lp = (seq + 15) & ~15;
  const size_t smem = static_cast<size_t>(4) * Lp * LDS * 2 + static_cast<size_t>(2) * Lp * sizeof(float);
  CLIPN_REQUIRE(smem <= 227 * 1024, "attention_bwd: sequence too long for the single-CTA kernel (L <= 384)");
  CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attention_bwd_kernel<<<batch * heads, pick_warps(seq) * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), seq, heads, causal, scale);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}
