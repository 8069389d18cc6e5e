// tcgen05 / TMEM / TMA GEMM family for the CLIP towers and the contrastive-loss logits.
//
// One persistent, warp-specialised kernel (1 CTA per SM, 384 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, SWIZZLE_128B, 3-6 stage mbarrier ring)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16,
//                                 fp32 accumulators double-buffered in TMEM: 2 x BN columns)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b.x32 -> registers -> fused epilogue -> per-warp swizzled smem
//                                 staging tile -> TMA store; residual / pre-activation operands arrive by TMA load
//                                 into the same staging tile, so all global traffic of the epilogue is coalesced
//                                 bulk copies and no cross-warp synchronisation exists in the epilogue)
// Operands may be K-major (torch Linear layout) or MN-major (weight-gradient / `x @ W` layouts); the
// MN-major case uses the canonical ((8,n),(8,k)) SW128 layout with LBO = 8 KiB between 64-wide MN groups.
// The B operand may be split across up to 8 tensor maps (one per rank's peer-mapped feature buffer), which
// is how the all-gather of gather_features (loss.py:29-54) is fused into the logits GEMM.
//
// Replaces (reference, cuBLASLt via ATen): transformer.py:195-197,246,295-299,794,923; model.py:409;
// loss.py:102-110 and the autograd dgrad/wgrad of each.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_internal.cuh"

namespace clipn {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int EPI_BUF_BYTES = 32 * 128;  // one warp's staging tile: 32 rows x 64 bf16 (128 B, SW128)
constexpr float kLog2e = 1.4426950408889634f;

// what each epilogue reads / writes through the per-warp TMA staging tiles
template <int EPI>
struct EpiTraits {
  static constexpr bool kOutTma = EPI == CLIPN_EPI_STORE || EPI == CLIPN_EPI_BIAS_GELU || EPI == CLIPN_EPI_BIAS_RESID ||
                                  EPI == CLIPN_EPI_DGELU || EPI == CLIPN_EPI_CLIP_DLOGITS || EPI == CLIPN_EPI_SIGLIP ||
                                  EPI == CLIPN_EPI_BIAS_GELU_GRAD || EPI == CLIPN_EPI_MUL_AUX;
  static constexpr int kNumOut =
      (EPI == CLIPN_EPI_BIAS_GELU || EPI == CLIPN_EPI_DGELU || EPI == CLIPN_EPI_BIAS_GELU_GRAD) ? 2 : 1;
  static constexpr bool kAux = EPI == CLIPN_EPI_BIAS_RESID || EPI == CLIPN_EPI_DGELU || EPI == CLIPN_EPI_MUL_AUX;
  static constexpr bool kColSum = EPI == CLIPN_EPI_STORE || EPI == CLIPN_EPI_DGELU || EPI == CLIPN_EPI_MUL_AUX;
  static constexpr bool kRedF32 = EPI == CLIPN_EPI_ACCUM_F32;  // fp32 tile reduce-added by TMA (cp.reduce.async.bulk)
  static constexpr int kBufs = kOutTma ? kNumOut : (kRedF32 ? 1 : 0);
};

// CTAS = 1: one CTA computes a 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, cta_group::2) computes 256 x BN;
// each CTA stages its own 128 rows of A and BN/2 rows of B, so a stage is 16 KB + BN*64 B and the ring gets deeper.
template <int BN, int EPI, int CTAS>
struct TileCfg {
  static constexpr int B_STAGE_BYTES = (BN / CTAS) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // Pair kernel, GELU fwd / bwd epilogues (two outputs, or one output x one operand): per warp 2 KB tiles
  // (32 rows x 32 cols, SW64) — each output and the aux operand double-buffered — so the TMA stores of piece i and the
  // aux TMA load of piece i+1 overlap the math of piece i.  Otherwise: kBufs x 4 KB (32 rows x 64 cols, SW128).
  static constexpr bool kPipedEpi = CTAS == 2 && (EpiTraits<EPI>::kNumOut == 2 || EPI == CLIPN_EPI_MUL_AUX);
  static constexpr int PIPE_TILES = 2 * EpiTraits<EPI>::kNumOut + (EpiTraits<EPI>::kAux ? 2 : 0);  // 4 or 6
  static constexpr int EPI_BYTES =
      kPipedEpi ? kEpiWarps * PIPE_TILES * 2048 : kEpiWarps * EpiTraits<EPI>::kBufs * EPI_BUF_BYTES;  // 0 / 32 / 64 / 96 KB
  static constexpr int BUDGET = 227 * 1024 - 1024 - 256;
  static constexpr int STAGES_FIT = (BUDGET - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 or 256 (power of two)
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 2 * kEpiWarps) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

// ---------------------------------------------------------------------------------------------------
// Epilogue math. Each epilogue thread owns one output row; it receives 32 consecutive columns at a time.
// ---------------------------------------------------------------------------------------------------
struct EpiState {
  float run_max, run_sum, pos, acc0, acc1;
  bool has_pos;
};

__device__ __forceinline__ void epi_begin(EpiState& st) {
  st.run_max = -INFINITY;
  st.run_sum = 0.f;
  st.pos = 0.f;
  st.has_pos = false;
  st.acc0 = 0.f;
  st.acc1 = 0.f;
}

// nvalid = number of valid columns in this 32-wide chunk (multiple of 8; N-tail support)
__device__ __forceinline__ void store_bf16_row32(void* base, int64_t ld, int row, int col, const float (&v)[32],
                                                 int nvalid) {
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + static_cast<int64_t>(row) * ld + col);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[i * 8 + j];
    if (i * 8 < nvalid) dst[i] = pack_bf16x8(t);
  }
}
__device__ __forceinline__ void load_bf16_row32(const void* base, int64_t ld, int row, int col, float (&v)[32],
                                                int nvalid) {
  const uint4* src =
      reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + static_cast<int64_t>(row) * ld + col);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack_bf16x8((i * 8 < nvalid) ? __ldg(src + i) : make_uint4(0, 0, 0, 0), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i * 8 + j] = t[j];
  }
}
__device__ __forceinline__ void load_bias32(const void* bias, int col, float (&b)[32], int nvalid) {
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(bias) + col);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack_bf16x8((i * 8 < nvalid) ? __ldg(src + i) : make_uint4(0, 0, 0, 0), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[i * 8 + j] = t[j];
  }
}

// Bias gradient fused into the epilogue: col_sum[col + l] += sum over this warp's 32 rows of v[l].
// Register transpose-reduce: at step `off` a lane keeps the half of its column range selected by its own bit
// `off` and receives the partner's partial sums for it; after 5 steps lane l holds column l's total
// (31 shuffles + 31 adds per 32x32 piece).  Warp-convergent; destroys v.
__device__ __forceinline__ void epi_col_sum(const GemmParams& p, int row, int col, float (&v)[32]) {
  const int lane = lane_id();
  const bool row_ok = row < p.m;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = row_ok ? v[i] : 0.f;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  if (col + lane < p.n) atomicAdd(p.col_sum + col + lane, v[0]);
}

// Staging tile access: row r (0..31) of a 1024-aligned 32x128B SW128 tile, 16-byte chunk c (0..7).
__device__ __forceinline__ uint4* stage_chunk(uint8_t* buf, int r, int c) {
  return reinterpret_cast<uint4*>(buf + r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void stage_write32(uint8_t* buf, int r, int half, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[i * 8 + j];
    *stage_chunk(buf, r, half * 4 + i) = pack_bf16x8(t);
  }
}
__device__ __forceinline__ void stage_read32(uint8_t* buf, int r, int half, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack_bf16x8(*stage_chunk(buf, r, half * 4 + i), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i * 8 + j] = t[j];
  }
}

// v: accumulator chunk in, primary bf16 output out (when the epilogue has one); aux: residual / pre-activation
// values (EpiTraits::kAux); o1: second output (kNumOut == 2).  fp32 outputs / atomics / reductions are issued here.
// Caller guarantees col < N (warp-uniform); row may be >= M.
template <int EPI>
__device__ __forceinline__ void epi_compute(const GemmParams& p, int row, int col, float (&v)[32],
                                            const float (&aux)[32], float (&o1)[32], EpiState& st) {
  const bool row_ok = row < p.m;
  const int nvalid = (p.n - col) < 32 ? (p.n - col) : 32;
  if constexpr (EPI == CLIPN_EPI_STORE || EPI == CLIPN_EPI_STORE_F32) {
    const f32x2 al = f2_splat(p.alpha);
    if (p.bias != nullptr) {
      float b[32];
      load_bias32(p.bias, col, b, nvalid);
#pragma unroll
      for (int i = 0; i < 32; i += 2) f2_unpack(f2_fma(f2_pack(v[i], v[i + 1]), al, f2_pack(b[i], b[i + 1])), v[i], v[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) f2_unpack(f2_mul(f2_pack(v[i], v[i + 1]), al), v[i], v[i + 1]);
    }
    if constexpr (EPI == CLIPN_EPI_STORE_F32) {
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.c) + static_cast<int64_t>(row) * p.ldc + col);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i * 4 < nvalid) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  } else if constexpr (EPI == CLIPN_EPI_BIAS_GELU) {
    float b[32];
    load_bias32(p.bias, col, b, nvalid);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      f2_unpack(f2_add(f2_pack(v[i], v[i + 1]), f2_pack(b[i], b[i + 1])), v[i], v[i + 1]);
      bf16_round2(v[i], v[i + 1]);
      float u0, u1;
      gelu_pair<false>(v[i], v[i + 1], o1[i], o1[i + 1], u0, u1);
    }
  } else if constexpr (EPI == CLIPN_EPI_BIAS_GELU_GRAD) {
    float b[32];
    load_bias32(p.bias, col, b, nvalid);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float t0, t1;
      f2_unpack(f2_add(f2_pack(v[i], v[i + 1]), f2_pack(b[i], b[i + 1])), t0, t1);
      bf16_round2(t0, t1);
      gelu_pair<true>(t0, t1, o1[i], o1[i + 1], v[i], v[i + 1]);  // C2 = gelu(t), C = gelu'(t)
    }
  } else if constexpr (EPI == CLIPN_EPI_MUL_AUX) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) f2_unpack(f2_mul(f2_pack(v[i], v[i + 1]), f2_pack(aux[i], aux[i + 1])), v[i], v[i + 1]);
  } else if constexpr (EPI == CLIPN_EPI_BIAS_RESID) {
    float b[32];
    load_bias32(p.bias, col, b, nvalid);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      f2_unpack(f2_add(f2_pack(v[i], v[i + 1]), f2_pack(b[i], b[i + 1])), v[i], v[i + 1]);
      bf16_round2(v[i], v[i + 1]);
      f2_unpack(f2_add(f2_pack(v[i], v[i + 1]), f2_pack(aux[i], aux[i + 1])), v[i], v[i + 1]);
    }
  } else if constexpr (EPI == CLIPN_EPI_DGELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float g0, g1;
      gelu_pair<true>(aux[i], aux[i + 1], o1[i], o1[i + 1], g0, g1);
      f2_unpack(f2_mul(f2_pack(v[i], v[i + 1]), f2_pack(g0, g1)), v[i], v[i + 1]);
    }
  } else if constexpr (EPI == CLIPN_EPI_ACCUM_F32) {
    const f32x2 al = f2_splat(p.alpha);
#pragma unroll
    for (int i = 0; i < 32; i += 2) f2_unpack(f2_mul(f2_pack(v[i], v[i + 1]), al), v[i], v[i + 1]);  // reduce-added to C by the caller
  } else if constexpr (EPI == CLIPN_EPI_LSE) {
    // online log-sum-exp in the log2 domain: t = (alpha*acc + bias) * log2(e); one FFMA + FMNMX + FADD + MUFU.EX2 +
    // FADD per element (raw ex2.approx: the range-checked exp2f() costs ~6 more issue slots per element and this
    // epilogue competes with a K = 512 mainloop).  run_max / pos are kept in the log2 domain until epi_finish.
    const float a2 = p.alpha * kLog2e, b2 = p.logit_bias * kLog2e;
    float cmax = -INFINITY;
    if (nvalid == 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = fmaf(v[i], a2, b2);
        cmax = fmaxf(cmax, v[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = (i < nvalid) ? fmaf(v[i], a2, b2) : -INFINITY;
        cmax = fmaxf(cmax, v[i]);
      }
    }
    // the label column of this warp's 32 rows lies in [row0 + off, row0 + off + 32): test the chunk once per warp
    const int lab0 = row - static_cast<int>(lane_id()) + p.label_offset;
    if (col < lab0 + 32 && col + 32 > lab0) {
      const int label = row + p.label_offset;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col + i == label) {
          st.pos = v[i];
          st.has_pos = true;
        }
    }
    const float nmax = fmaxf(st.run_max, cmax);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      s0 += ex2_approx(v[i] - nmax);
      s1 += ex2_approx(v[i + 1] - nmax);
    }
    st.run_sum = st.run_sum * ex2_approx(st.run_max - nmax) + (s0 + s1);
    st.run_max = nmax;
  } else if constexpr (EPI == CLIPN_EPI_CLIP_DLOGITS) {
    const int label = row + p.label_offset;
    const float a2 = p.alpha * kLog2e, b2 = p.logit_bias * kLog2e;
    const float rl2 = row_ok ? __ldg(p.row_lse + row) * kLog2e : 0.f;
    const int lab0 = row - static_cast<int>(lane_id()) + p.label_offset;
    const bool has_label = col < lab0 + 32 && col + 32 > lab0;  // warp-uniform
    float cl2[32];
    if (p.col_w != 0.f) {
      // column LSE vector: 32 consecutive floats, the same for every lane (broadcast loads through L1)
      if (nvalid == 32 && (reinterpret_cast<uintptr_t>(p.col_lse + col) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(p.col_lse + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldg(src + i);
          cl2[4 * i] = t.x * kLog2e; cl2[4 * i + 1] = t.y * kLog2e; cl2[4 * i + 2] = t.z * kLog2e; cl2[4 * i + 3] = t.w * kLog2e;
        }
      } else {  // ragged tail / a vector that is not 16-byte aligned (batch not a multiple of 4)
#pragma unroll
        for (int i = 0; i < 32; ++i) cl2[i] = (i < nvalid) ? __ldg(p.col_lse + col + i) * kLog2e : 0.f;
      }
    }
    const float shift = (1.f + p.col_w) / static_cast<float>(p.n);
    // d logit_scale = sum (P_row - onehot) * dot: with nearly parallel features (fresh model) every dot is ~1 and the sum
    // is the small difference of two O(B) totals — accumulated naively (one fp32 atomic per warp and tile) it lost all
    // but two digits at N = 32768.  Subtracting a per-row centre c ~ dot[row, label] (optional vector `pos`) from every
    // dot changes the exact value by c * (sum_n P_row - 1) = 0 and leaves only small terms to accumulate.
    const float centre = (p.pos != nullptr && row_ok) ? __ldg(p.pos + row) : 0.f;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float dot = v[i];
      const float s2 = fmaf(dot, a2, b2);
      float pr = ex2_approx(s2 - rl2);
      float pc = (p.col_w != 0.f) ? p.col_w * ex2_approx(s2 - cl2[i]) : 0.f;
      if (i >= nvalid) { pr = 0.f; pc = 0.f; }
      // C holds the softmax parts CENTRED on their mean (1 + col_w) / N, without the one-hot of the label column: bf16
      // then rounds the deviation from the uniform distribution instead of values ~1/N (and never a value ~2 next to
      // them), which is what keeps d(features) accurate while the features are still nearly parallel (fresh model).
      // The caller restores both parts in fp32: (1 + col_w) * gscale * alpha * (mean_n cols[n] - cols[label]).
      v[i] = p.gscale * (pr + pc - shift);
      if (has_label && col + i == label) pr -= 1.f;
      a0 = fmaf(pr, dot - centre, a0);
      a1 += pr;
    }
    if (row_ok) {
      st.acc0 += a0;
      st.acc1 += a1;
    }
  } else if constexpr (EPI == CLIPN_EPI_SIGLIP) {
    // -logsigmoid(yz) = max(-yz, 0) + log1p(e), e = exp(-|yz|); d/dz = -y * sigmoid(-yz).  Raw MUFU ops
    // (ex2 / rcp / lg2 .approx): 3 per element with the loss value, 2 without (p.part_sum == nullptr).
    const int lab0 = row - static_cast<int>(lane_id()) + p.label_offset;
    const bool has_label = !p.negative_only && col < lab0 + 32 && col + 32 > lab0;  // warp-uniform
    const int label = row + p.label_offset;
    const bool want_loss = p.part_sum != nullptr;
    float ls = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float dot = v[i];
      const float z = fmaf(dot, p.alpha, p.logit_bias);
      const float y = (has_label && col + i == label) ? 1.f : -1.f;
      const float yz = y * z;
      const float e = ex2_approx(-fabsf(yz) * kLog2e);
      const float r = rcp_approx(1.f + e);                 // 1 / (1 + e)
      const float sig = (yz >= 0.f) ? e * r : r;            // sigmoid(-yz)
      const float dz = (i < nvalid) ? -y * sig : 0.f;
      v[i] = p.gscale * dz;
      a0 = fmaf(dz, dot, a0);
      a1 += dz;
      if (want_loss && i < nvalid) ls += fmaxf(-yz, 0.f) - lg2_approx(r) * 0.69314718055994531f;  // log1p(e) = -ln r
    }
    if (row_ok) {
      st.run_sum += ls;
      st.acc0 += a0;
      st.acc1 += a1;
    }
  }
}

// called once per (row, n-slab) after all chunks; `slab` = n_tile*2 + half. Warp-convergent.
template <int EPI>
__device__ __forceinline__ void epi_finish(const GemmParams& p, int row, int slab, EpiState& st) {
  if constexpr (EPI == CLIPN_EPI_LSE) {
    if (row < p.m) {
      constexpr float kLn2 = 0.69314718055994531f;
      p.part_max[static_cast<int64_t>(slab) * p.m + row] = st.run_max * kLn2;  // log2 domain -> natural
      p.part_sum[static_cast<int64_t>(slab) * p.m + row] = st.run_sum;
      if (st.has_pos) p.pos[row] = st.pos * kLn2;
    }
  } else if constexpr (EPI == CLIPN_EPI_CLIP_DLOGITS) {
    const float a0 = warp_sum(st.acc0), a1 = warp_sum(st.acc1);
    if (lane_id() == 0 && p.scalar_acc != nullptr) {
      atomicAdd(p.scalar_acc + 0, a0 * p.gscale);
      atomicAdd(p.scalar_acc + 1, a1 * p.gscale);
    }
  } else if constexpr (EPI == CLIPN_EPI_SIGLIP) {
    const float l = warp_sum(st.run_sum), a0 = warp_sum(st.acc0), a1 = warp_sum(st.acc1);
    if (lane_id() == 0) {
      if (p.part_sum != nullptr) atomicAdd(p.part_sum, l * p.gscale);
      if (p.scalar_acc != nullptr) {
        atomicAdd(p.scalar_acc + 0, a0 * p.gscale);
        atomicAdd(p.scalar_acc + 1, a1 * p.gscale);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// The tensor-core kernel
// ---------------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ TmapSet tm, const GemmParams p_in) {
  using Cfg = TileCfg<BN, EPI, 1>;
  using Tr = EpiTraits<EPI>;
  GemmParams p = p_in;
  if (p.alpha_dev != nullptr) p.alpha *= __ldg(p.alpha_dev);
  if (p.logit_bias_dev != nullptr) p.logit_bias += __ldg(p.logit_bias_dev);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_bar = tmem_empty + 2;  // one per epilogue warp
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aux_bar + kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.a);
    for (int i = 0; i < p.b_maps; ++i) tma_prefetch_desc(&tm.b[i]);
    if (Tr::kOutTma || Tr::kRedF32) tma_prefetch_desc(&tm.c);
    if (Tr::kNumOut == 2 && p.c2 != nullptr) tma_prefetch_desc(&tm.c2);
    if (Tr::kAux) tma_prefetch_desc(&tm.aux);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_work = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.splits;
        const int tile = w / p.splits;
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int kb0 = static_cast<int>((static_cast<int64_t>(split) * p.kblocks) / p.splits);
        const int kb1 = static_cast<int>((static_cast<int64_t>(split + 1) * p.kblocks) / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d(sa, &tm.a, &full_bar[stage], kb * BK, m0);
          } else {
            tma_load_2d(sa, &tm.a, &full_bar[stage], m0, kb * BK);
            tma_load_2d(sa + 8192, &tm.a, &full_bar[stage], m0 + 64, kb * BK);
          }
          if (!p.b_mn) {
            const int map = n0 / p.b_rows_per_map;
            tma_load_2d(sb, &tm.b[map], &full_bar[stage], kb * BK, n0 - map * p.b_rows_per_map);
          } else {
            const int map = (kb * BK) / p.b_rows_per_map;
            const int krow = kb * BK - map * p.b_rows_per_map;
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * 8192, &tm.b[map], &full_bar[stage], n0 + 64 * i, krow);
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.splits;
        const int kb0 = static_cast<int>((static_cast<int64_t>(split) * p.kblocks) / p.splits);
        const int kb1 = static_cast<int>((static_cast<int64_t>(split + 1) * p.kblocks) / p.splits);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = p.a_mn ? umma_smem_desc(sa + k * 2048, 8192, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = p.b_mn ? umma_smem_desc(sb + k * 2048, 8192, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int e = warp - 4;
    const int q = e & 3;   // TMEM lane quarter == warp % 4
    const int h = e >> 2;  // column half of the tile
    uint8_t* buf0 = epi_smem + e * (Tr::kBufs * EPI_BUF_BYTES);
    uint8_t* buf1 = buf0 + EPI_BUF_BYTES;
    const bool store_c = Tr::kOutTma && (EPI != CLIPN_EPI_SIGLIP || p.c != nullptr);
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.splits;
      const int tn = tile % p.tiles_n;
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = tn * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      EpiState st;
      epi_begin(st);
#pragma unroll 1
      for (int jc = 0; jc < BN / 128; ++jc) {
        const int cl0 = h * (BN / 2) + jc * 64;  // first column (within the tile) of this 64-wide chunk
        const bool chunk_live = n0 + cl0 < p.n;  // warp-uniform
        if (Tr::kOutTma && chunk_live) {
          // staging tiles must be free: the TMA stores of the previous chunk have finished reading them
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          if (Tr::kAux) {
            if (lane == 0) {
              mbar_expect_tx(&aux_bar[e], EPI_BUF_BYTES);
              tma_load_2d(buf0, &tm.aux, &aux_bar[e], n0 + cl0, row0);
            }
            mbar_wait(&aux_bar[e], aux_phase);
            aux_phase ^= 1;
          }
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int cl = cl0 + half * 32;
          float v[32], aux[32], o1[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + cl, v);
          if (n0 + cl < p.n) {
            if (Tr::kAux) stage_read32(buf0, lane, half, aux);
            epi_compute<EPI>(p, row, n0 + cl, v, aux, o1, st);
            if (store_c) stage_write32(buf0, lane, half, v);
            if (Tr::kNumOut == 2 && p.c2 != nullptr) stage_write32(buf1, lane, half, o1);
            if (Tr::kColSum && p.col_sum != nullptr)
              epi_col_sum(p, row, n0 + cl, v);
            if (Tr::kRedF32) {
              // fp32 32x32 tile -> swizzled staging -> cp.reduce.async.bulk.tensor (.add) into C
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(stage_chunk(buf0, lane, i)) =
                    make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_reduce_add_2d(&tm.c, buf0, n0 + cl, row0);
                tma_store_commit();
              }
            }
          }
        }
        if (Tr::kOutTma && chunk_live && store_c) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm.c, buf0, n0 + cl0, row0);
            if (Tr::kNumOut == 2 && p.c2 != nullptr) tma_store_2d(&tm.c2, buf1, n0 + cl0, row0);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      epi_finish<EPI>(p, row, tn * 2 + h, st);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if ((Tr::kOutTma || Tr::kRedF32) && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// 32 rows x 32 bf16 (64-byte rows) SWIZZLE_64B staging tile: 16-byte chunk c (0..3) of row r
__device__ __forceinline__ uint4* stage64_chunk(uint8_t* buf, int r, int c) {
  return reinterpret_cast<uint4*>(buf + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}
__device__ __forceinline__ void stage64_write32(uint8_t* buf, int r, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[i * 8 + j];
    *stage64_chunk(buf, r, i) = pack_bf16x8(t);
  }
}
__device__ __forceinline__ void stage64_read32(uint8_t* buf, int r, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack_bf16x8(*stage64_chunk(buf, r, i), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i * 8 + j] = t[j];
  }
}

#include "gemm_pair.cuh"
#include "gemm_peer.cuh"

// ---------------------------------------------------------------------------------------------------
// CUDA-core restatement (tests only): identical slab structure and epilogue math, no tensor cores, plain
// global loads/stores.  grid = (ceil(M/128), tiles_n*2), block = 128 threads (thread == row).
// ---------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_ref_kernel(const GemmParams p_in, RefOperands ops, int bn) {
  using Tr = EpiTraits<EPI>;
  GemmParams p = p_in;
  if (p.alpha_dev != nullptr) p.alpha *= __ldg(p.alpha_dev);
  if (p.logit_bias_dev != nullptr) p.logit_bias += __ldg(p.logit_bias_dev);
  const int row = blockIdx.x * 128 + threadIdx.x;
  const int slab = blockIdx.y;
  const int n_begin = (slab >> 1) * bn + (slab & 1) * (bn / 2);
  const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(ops.a);
  const bool store_c = Tr::kOutTma && (EPI != CLIPN_EPI_SIGLIP || p.c != nullptr);
  EpiState st;
  epi_begin(st);
  for (int cl = 0; cl < bn / 2; cl += 32) {
    const int col = n_begin + cl;
    if (col >= p.n) break;
    const int nvalid = (p.n - col) < 32 ? (p.n - col) : 32;
    float v[32], aux[32], o1[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = aux[i] = o1[i] = 0.f;
    if (row < p.m) {
      for (int k = 0; k < p.k; ++k) {
        const float a = __bfloat162float(p.a_mn ? A[static_cast<int64_t>(k) * ops.lda + row]
                                                  : A[static_cast<int64_t>(row) * ops.lda + k]);
        for (int i = 0; i < nvalid; ++i) {
          const int n = col + i;
          float b;
          if (!p.b_mn) {
            const int map = n / p.b_rows_per_map;
            b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(
                ops.b[map])[static_cast<int64_t>(n - map * p.b_rows_per_map) * ops.ldb + k]);
          } else {
            const int map = k / p.b_rows_per_map;
            b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(
                ops.b[map])[static_cast<int64_t>(k - map * p.b_rows_per_map) * ops.ldb + n]);
          }
          v[i] += a * b;
        }
      }
      if (Tr::kAux) load_bf16_row32(p.aux, p.ldaux, row, col, aux, nvalid);
    }
    epi_compute<EPI>(p, row, col, v, aux, o1, st);
    if (row < p.m) {
      if (Tr::kRedF32) {
        float* dst = reinterpret_cast<float*>(p.c) + static_cast<int64_t>(row) * p.ldc + col;
        for (int i = 0; i < nvalid; ++i) atomicAdd(dst + i, v[i]);
      }
      if (store_c) store_bf16_row32(p.c, p.ldc, row, col, v, nvalid);
      if (Tr::kNumOut == 2 && p.c2 != nullptr) store_bf16_row32(p.c2, p.ldc2, row, col, o1, nvalid);
    }
    if (Tr::kColSum && p.col_sum != nullptr) epi_col_sum(p, row, col, v);
  }
  epi_finish<EPI>(p, row, slab, st);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int gemm_tile_n(int n) {
  // CLIPN_GEMM_TILE_N=128 forces the 128-wide tile everywhere (tuning experiments: wave quantisation of the
  // N = 768 / 512 outputs); the default picks 256 whenever it divides N.
  static const int forced = [] {
    const char* e = getenv("CLIPN_GEMM_TILE_N");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced == 128) return 128;
  return (n % 256 == 0) ? 256 : 128;
}

template <int BN, int EPI>
static int launch_tc(const TmapSet& tm, const GemmParams& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN, EPI, 1>;
  static bool configured = false;  // benign race: idempotent attribute set
  if (!configured) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES));
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<BN, EPI><<<grid, kThreads, Cfg::SMEM_BYTES, stream>>>(tm, p);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

template <int BN, int EPI>
static int launch_tc2(const TmapSet& tm, const GemmParams& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN, EPI, 2>;
  static bool configured = false;
  if (!configured) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES));
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;  // pair tiles
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? total : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CLIPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<BN, EPI>, tm, p));
  return CLIPN_OK;
}

template <int EPI>
static int launch_ref(const GemmParams& p, const RefOperands& ops, int bn, cudaStream_t stream) {
  dim3 grid((p.m + 127) / 128, p.tiles_n * 2);
  gemm_ref_kernel<EPI><<<grid, 128, 0, stream>>>(p, ops, bn);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

int peer_gemm_tile_n(int world, int rows_per_map, int e) {
  if (e <= 0 || e % BK != 0 || e > 1024 || world < 1 || world > kMaxBMaps || rows_per_map <= 0) return 0;
  const int bn = e <= 512 ? 256 : 128;
  if (world > 1 && rows_per_map % bn != 0) return 0;  // a column tile must not straddle two ranks' buffers
  return bn;
}

template <int BN, int EPI>
static int launch_peer(const PeerTmaps& tm, const PeerParams& p, cudaStream_t stream) {
  using Cfg = PeerCfg<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_peer_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES));
    configured = true;
  }
  const int64_t total = static_cast<int64_t>(p.dirs) * p.tiles_n * p.tiles_m;
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? static_cast<int>(total) : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CLIPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_peer_kernel<BN, EPI>, tm, p));
  return CLIPN_OK;
}

int peer_gemm_launch(const PeerGemmDesc& d, cudaStream_t stream) {
  CLIPN_REQUIRE(d.dirs == 1 || d.dirs == 2, "peer gemm: 1 or 2 directions");
  CLIPN_REQUIRE(d.epilogue == CLIPN_EPI_LSE || d.epilogue == CLIPN_EPI_SIGLIP, "peer gemm: LSE or SIGLIP epilogue");
  CLIPN_REQUIRE(d.m > 0, "peer gemm: empty problem");
  const int bn = peer_gemm_tile_n(d.world, d.rows_per_map, d.e);
  CLIPN_REQUIRE(bn != 0,
                "peer gemm: needs embed dim % 64 == 0 and <= 1024, world <= 8, per-rank rows % 256 == 0 (128 for embed dim > "
                "512) when world > 1");
  int cc_major = 0, sms = 0, cc_minor = 0;
  clipn_device_info(&sms, &cc_major, &cc_minor);
  CLIPN_REQUIRE(cc_major == 10, "peer gemm: the tcgen05 kernels require an sm_100 (B200) device");
  const int n = d.world * d.rows_per_map;
  PeerParams p;
  memset(&p, 0, sizeof(p));
  p.m = d.m; p.n = n; p.kblocks = d.e / BK;
  p.rank = d.world > 1 ? d.rank : 0;
  p.rows_per_map = d.rows_per_map;
  p.tiles_m = (d.m + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (n + bn - 1) / bn;
  p.dirs = d.dirs;
  p.label_offset = d.label_offset; p.negative_only = d.negative_only;
  p.alpha = d.alpha; p.alpha_dev = d.alpha_dev; p.logit_bias = d.logit_bias; p.logit_bias_dev = d.logit_bias_dev;
  p.gscale = d.gscale;
  p.gather = 0;
  PeerTmaps tm;
  int rc;
  for (int dir = 0; dir < d.dirs; ++dir) {
    CLIPN_REQUIRE(d.rows[dir] != nullptr && d.cols[dir] != nullptr, "peer gemm: null operand");
    if (d.epilogue == CLIPN_EPI_LSE)
      CLIPN_REQUIRE(d.part_max[dir] && d.part_sum[dir] && d.pos[dir], "peer gemm: LSE buffers required");
    p.part_max[dir] = d.part_max[dir]; p.part_sum[dir] = d.part_sum[dir]; p.pos[dir] = d.pos[dir];
    p.c[dir] = d.c[dir]; p.scalar_acc[dir] = d.scalar_acc[dir];
    rc = make_tmap_2d(&tm.a[dir], d.rows[dir], 2, d.e, d.m, static_cast<uint64_t>(d.e) * 2, BK, BM, 128);
    if (rc) return rc;
    for (int i = 0; i < d.world; ++i) {
      CLIPN_REQUIRE(d.cols[dir][i] != nullptr, "peer gemm: null column pointer");
      rc = make_tmap_2d(&tm.b[dir][i], d.cols[dir][i], 2, d.e, d.rows_per_map, static_cast<uint64_t>(d.e) * 2, BK, bn / 2,
                        128);
      if (rc) return rc;
    }
    if (d.gather[dir] != nullptr) {
      p.gather = 1;
      rc = make_tmap_2d(&tm.g[dir], d.gather[dir], 2, d.e, n, static_cast<uint64_t>(d.e) * 2, BK, bn / 2, 128);
      if (rc) return rc;
    }
    if (d.epilogue == CLIPN_EPI_SIGLIP && d.c[dir] != nullptr) {
      CLIPN_REQUIRE(d.ldc % 8 == 0 && d.ldc >= n, "peer gemm: ldc must be a multiple of 8 and >= n");
      rc = make_tmap_2d(&tm.c[dir], d.c[dir], 2, n, d.m, static_cast<uint64_t>(d.ldc) * 2, 64, 32, 128);
      if (rc) return rc;
    }
  }
  if (p.gather)
    for (int dir = 0; dir < d.dirs; ++dir)
      CLIPN_REQUIRE(d.gather[dir] != nullptr, "peer gemm: gather buffers must be given for every direction or none");
  if (d.epilogue == CLIPN_EPI_LSE)
    return bn == 256 ? launch_peer<256, CLIPN_EPI_LSE>(tm, p, stream) : launch_peer<128, CLIPN_EPI_LSE>(tm, p, stream);
  return bn == 256 ? launch_peer<256, CLIPN_EPI_SIGLIP>(tm, p, stream) : launch_peer<128, CLIPN_EPI_SIGLIP>(tm, p, stream);
}

#define CLIPN_DISPATCH_EPI(EPIVAR, MACRO)                                   \
  switch (EPIVAR) {                                                         \
    case CLIPN_EPI_STORE: MACRO(CLIPN_EPI_STORE); break;                   \
    case CLIPN_EPI_BIAS_GELU: MACRO(CLIPN_EPI_BIAS_GELU); break;           \
    case CLIPN_EPI_BIAS_RESID: MACRO(CLIPN_EPI_BIAS_RESID); break;         \
    case CLIPN_EPI_DGELU: MACRO(CLIPN_EPI_DGELU); break;                   \
    case CLIPN_EPI_ACCUM_F32: MACRO(CLIPN_EPI_ACCUM_F32); break;           \
    case CLIPN_EPI_STORE_F32: MACRO(CLIPN_EPI_STORE_F32); break;           \
    case CLIPN_EPI_LSE: MACRO(CLIPN_EPI_LSE); break;                       \
    case CLIPN_EPI_CLIP_DLOGITS: MACRO(CLIPN_EPI_CLIP_DLOGITS); break;     \
    case CLIPN_EPI_SIGLIP: MACRO(CLIPN_EPI_SIGLIP); break;                 \
    case CLIPN_EPI_BIAS_GELU_GRAD: MACRO(CLIPN_EPI_BIAS_GELU_GRAD); break; \
    case CLIPN_EPI_MUL_AUX: MACRO(CLIPN_EPI_MUL_AUX); break;               \
    default: return set_error_arg("unknown epilogue", __FILE__, __LINE__); \
  }

int gemm_launch(const clipn_gemm_desc& d, const void* const* b_ptrs, int b_maps, int64_t b_rows_per_map, bool use_ref,
                cudaStream_t stream) {
  CLIPN_REQUIRE(d.m > 0 && d.n > 0 && d.k > 0, "gemm: empty problem");
  // the loss epilogues handle any N (per-element masks, TMA-clipped stores): a contrastive batch need not be a
  // multiple of 8 (reference ClipLoss / SigLipLoss take any batch); the others read bias / aux vectors 8 at a time
  const bool any_n = d.epilogue == CLIPN_EPI_LSE || d.epilogue == CLIPN_EPI_CLIP_DLOGITS || d.epilogue == CLIPN_EPI_SIGLIP;
  CLIPN_REQUIRE(any_n || d.n % 8 == 0, "gemm: N must be a multiple of 8");
  // row pitches must be 16-byte multiples for the TMA maps; the contraction length itself is free (TMA zero-fills)
  CLIPN_REQUIRE(d.lda % 8 == 0 && d.ldb % 8 == 0, "gemm: leading dims must be multiples of 8");
  CLIPN_REQUIRE(b_maps >= 1 && b_maps <= kMaxBMaps, "gemm: 1..8 B maps");
  const int ep = d.epilogue;
  const bool needs_c = !(ep == CLIPN_EPI_LSE || (ep == CLIPN_EPI_SIGLIP && d.c == nullptr));
  if (needs_c) CLIPN_REQUIRE(d.c != nullptr && d.ldc % 8 == 0, "gemm: C missing or ldc not a multiple of 8");
  const bool two_out = ep == CLIPN_EPI_BIAS_GELU || ep == CLIPN_EPI_DGELU || ep == CLIPN_EPI_BIAS_GELU_GRAD;
  const bool has_aux = ep == CLIPN_EPI_BIAS_RESID || ep == CLIPN_EPI_DGELU || ep == CLIPN_EPI_MUL_AUX;
  if (ep == CLIPN_EPI_BIAS_GELU || ep == CLIPN_EPI_BIAS_GELU_GRAD) CLIPN_REQUIRE(d.c2 != nullptr, "gemm: C2 required");
  if (d.c2 != nullptr) CLIPN_REQUIRE(d.ldc2 % 8 == 0, "gemm: ldc2 must be a multiple of 8");  // DGELU: C2 optional
  if (ep == CLIPN_EPI_BIAS_GELU || ep == CLIPN_EPI_BIAS_RESID || ep == CLIPN_EPI_BIAS_GELU_GRAD)
    CLIPN_REQUIRE(d.bias != nullptr, "gemm: bias required");
  if (has_aux) CLIPN_REQUIRE(d.aux != nullptr && d.ldaux % 8 == 0, "gemm: aux required");
  if (ep == CLIPN_EPI_LSE) CLIPN_REQUIRE(d.part_max && d.part_sum && d.pos, "gemm: LSE buffers required");
  if (ep == CLIPN_EPI_CLIP_DLOGITS) CLIPN_REQUIRE(d.row_lse && (d.col_w == 0.f || d.col_lse), "gemm: lse vectors");

  GemmParams p;
  p.m = d.m; p.n = d.n; p.k = d.k;
  p.a_mn = d.a_mn_major ? 1 : 0;
  p.b_mn = d.b_mn_major ? 1 : 0;
  const int bn = gemm_tile_n(d.n);
  p.tiles_m = (d.m + BM - 1) / BM;
  p.tiles_n = (d.n + bn - 1) / bn;
  p.kblocks = (d.k + BK - 1) / BK;
  p.splits = (ep == CLIPN_EPI_ACCUM_F32 && d.splits > 1) ? d.splits : 1;
  if (p.splits > p.kblocks) p.splits = p.kblocks;
  p.b_maps = b_maps;
  p.b_rows_per_map = b_maps > 1 ? static_cast<int>(b_rows_per_map) : 0x40000000;
  if (b_maps > 1) {
    CLIPN_REQUIRE(b_rows_per_map % bn == 0 || p.b_mn, "gemm: per-rank rows must be a multiple of the N tile");
    CLIPN_REQUIRE(b_rows_per_map % BK == 0 || !p.b_mn, "gemm: per-rank rows must be a multiple of 64");
  }
  p.c = d.c; p.ldc = d.ldc; p.c2 = d.c2; p.ldc2 = d.ldc2;
  p.bias = d.bias; p.aux = d.aux; p.ldaux = d.ldaux;
  p.alpha = d.alpha;
  p.row_lse = d.row_lse; p.col_lse = d.col_lse; p.part_max = d.part_max; p.part_sum = d.part_sum; p.pos = d.pos;
  p.scalar_acc = d.scalar_acc; p.logit_bias = d.logit_bias; p.gscale = d.gscale; p.col_w = d.col_w;
  p.label_offset = d.label_offset; p.negative_only = d.negative_only;
  p.alpha_dev = d.alpha_dev; p.logit_bias_dev = d.logit_bias_dev;
  p.col_sum = d.col_sum;
  if (d.col_sum != nullptr)
    CLIPN_REQUIRE(ep == CLIPN_EPI_STORE || ep == CLIPN_EPI_DGELU || ep == CLIPN_EPI_MUL_AUX,
                  "gemm: col_sum only with STORE / DGELU / MUL_AUX epilogues");

  if (use_ref) {
    RefOperands ops;
    ops.a = d.a; ops.lda = d.lda; ops.ldb = d.ldb;
    for (int i = 0; i < kMaxBMaps; ++i) ops.b[i] = b_ptrs[i < b_maps ? i : 0];
#define CLIPN_REF_CASE(E) return launch_ref<E>(p, ops, bn, stream)
    CLIPN_DISPATCH_EPI(ep, CLIPN_REF_CASE)
#undef CLIPN_REF_CASE
    return CLIPN_OK;
  }

  int cc_major = 0, sms = 0, cc_minor = 0;
  clipn_device_info(&sms, &cc_major, &cc_minor);
  CLIPN_REQUIRE(cc_major == 10, "gemm: the tcgen05 kernels require an sm_100 (B200) device");

  // CTA-pair kernel (256-row tiles, cta_group::2) whenever there are at least two 128-row tiles of work;
  // CLIPN_GEMM_PAIR=0 forces the single-CTA kernel (A/B testing, bisecting).
  static const bool pair_enabled = [] {
    const char* e = getenv("CLIPN_GEMM_PAIR");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool use_pair = pair_enabled && d.m > BM;
  const int b_box_rows = use_pair ? bn / 2 : bn;  // each CTA of a pair stages half of the B tile

  TmapSet tm;
  int rc;
  if (!p.a_mn) rc = make_tmap_2d(&tm.a, d.a, 2, d.k, d.m, d.lda * 2, BK, BM, 128);
  else rc = make_tmap_2d(&tm.a, d.a, 2, d.m, d.k, d.lda * 2, 64, BK, 128);
  if (rc) return rc;
  for (int i = 0; i < b_maps; ++i) {
    if (!p.b_mn) {
      const uint64_t rows = b_maps > 1 ? static_cast<uint64_t>(b_rows_per_map) : static_cast<uint64_t>(d.n);
      rc = make_tmap_2d(&tm.b[i], b_ptrs[i], 2, d.k, rows, d.ldb * 2, BK, b_box_rows, 128);
    } else {
      const uint64_t rows = b_maps > 1 ? static_cast<uint64_t>(b_rows_per_map) : static_cast<uint64_t>(d.k);
      rc = make_tmap_2d(&tm.b[i], b_ptrs[i], 2, d.n, rows, d.ldb * 2, 64, BK, 128);
    }
    if (rc) return rc;
  }
  // epilogue staging maps: 64-column x 32-row bf16 boxes (one warp's tile), SW128
  const bool out_tma = ep == CLIPN_EPI_STORE || ep == CLIPN_EPI_BIAS_GELU || ep == CLIPN_EPI_BIAS_RESID ||
                       ep == CLIPN_EPI_DGELU || ep == CLIPN_EPI_CLIP_DLOGITS || (ep == CLIPN_EPI_SIGLIP && d.c != nullptr) ||
                       ep == CLIPN_EPI_BIAS_GELU_GRAD || ep == CLIPN_EPI_MUL_AUX;
  // pair kernel + two-output epilogue: 32-column pieces (64-byte rows, SW64), see TileCfg::kPipedEpi
  const bool piped = use_pair && (two_out || ep == CLIPN_EPI_MUL_AUX);
  const uint32_t ebox = piped ? 32 : 64;
  const int esw = piped ? 64 : 128;
  if (out_tma) {
    rc = make_tmap_2d(&tm.c, d.c, 2, d.n, d.m, d.ldc * 2, ebox, 32, esw);
    if (rc) return rc;
  }
  if (ep == CLIPN_EPI_ACCUM_F32) {
    rc = make_tmap_2d(&tm.c, d.c, 4, d.n, d.m, d.ldc * 4, 32, 32, 128);
    if (rc) return rc;
  }
  if (two_out && d.c2 != nullptr) {
    rc = make_tmap_2d(&tm.c2, d.c2, 2, d.n, d.m, d.ldc2 * 2, ebox, 32, esw);
    if (rc) return rc;
  }
  if (has_aux) {
    rc = make_tmap_2d(&tm.aux, d.aux, 2, d.n, d.m, d.ldaux * 2, ebox, 32, esw);
    if (rc) return rc;
  }
  if (use_pair) {
    p.tiles_m = (d.m + 2 * BM - 1) / (2 * BM);
#define CLIPN_TC2_CASE(E) \
  return (bn == 256) ? launch_tc2<256, E>(tm, p, stream) : launch_tc2<128, E>(tm, p, stream)
    CLIPN_DISPATCH_EPI(ep, CLIPN_TC2_CASE)
#undef CLIPN_TC2_CASE
  }
#define CLIPN_TC_CASE(E) \
  return (bn == 256) ? launch_tc<256, E>(tm, p, stream) : launch_tc<128, E>(tm, p, stream)
  CLIPN_DISPATCH_EPI(ep, CLIPN_TC_CASE)
#undef CLIPN_TC_CASE
  return CLIPN_OK;
}

}  // namespace clipn

extern "C" int clipn_gemm(const clipn_gemm_desc* d, clipn_stream_t stream) {
  if (d == nullptr) return clipn::set_error_arg("gemm: null descriptor", __FILE__, __LINE__);
  const void* b[1] = {d->b};
  return clipn::gemm_launch(*d, b, 1, 0, false, static_cast<cudaStream_t>(stream));
}
extern "C" int clipn_gemm_ref(const clipn_gemm_desc* d, clipn_stream_t stream) {
  if (d == nullptr) return clipn::set_error_arg("gemm: null descriptor", __FILE__, __LINE__);
  const void* b[1] = {d->b};
  return clipn::gemm_launch(*d, b, 1, 0, true, static_cast<cudaStream_t>(stream));
}
extern "C" int clipn_gemm_tile_n(int n) { return clipn::gemm_tile_n(n); }
