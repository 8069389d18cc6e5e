// Attention core on the 5th-generation tensor cores (F.scaled_dot_product_attention, reference transformer.py:223-228),
// forward.  tcgen05.mma with the S = Q K^T and O = P V accumulators in TMEM, operands staged by 3-D TMA boxes
// ([column, token, sequence]: rows past the end of a sequence are zero-filled on load and clipped on store, so no
// padding copies exist), softmax on the TMEM rows (one thread per query row, warp-shuffle-free: tcgen05.ld gives every
// thread its whole row), P handed to the second MMA through shared memory in the canonical SW128 K-major layout.
//
// Tile = 128 query rows x 128 keys, head_dim 64:
//   L <= 64   : TWO sequences share a tile (rows 0-63 / 64-127, block-diagonal mask) — ViT-B/32's 50 tokens fill 78 %
//               of the rows instead of 39 %
//   L <= 128  : one sequence, one key tile, softmax in one sweep (text tower: 77 tokens, causal by predicate)
//   L  > 128  : ceil(L/128) query tiles x key tiles, TWO PASSES over the key tiles: pass 1 only takes the row maximum
//               (S recomputed: QK^T is 1/3 of the tensor work and the tensor pipe idles behind the MUFU anyway), pass 2
//               exponentiates against the final maximum and accumulates P V — the O accumulator in TMEM is never
//               rescaled (ViT-L/14-336: 577 tokens = 5 x 5 tiles)
//
// Warp roles (384 threads, persistent, 1 CTA / SM): warp 0 TMA producer (Q tiles + a 5-stage K/V ring), warp 1 MMA
// issuer, warp 2 TMEM allocator, warps 4-7 / 8-11 two softmax groups working on two different items: while one group
// exponentiates, the tensor core runs the other group's QK^T / PV (FA-style ping-pong; TMEM: 2 x (128 S + 64 O) columns).
// Producer and MMA issuer walk the same deterministic schedule (alternating one "slot" per group: [PV of the previous
// step] + [QK^T of the next step]), so the K/V ring is consumed in exactly the order it is filled.
#include <stdlib.h>

#include "common.cuh"

#ifndef AT_PIPELINED_LD
#define AT_PIPELINED_LD 0
#endif

namespace clipn {

constexpr int kAtThreads = 384;
constexpr int AT_TILE = 128 * 64 * 2;  // one [128 x 64] bf16 tile, SWIZZLE_128B
constexpr int AT_P = 128 * 128 * 2;    // P tile: two 64-wide K atoms
constexpr int AT_STAGES = 5;
constexpr int AT_TMEM_GROUP = 192;
constexpr float kLog2eAt = 1.4426950408889634f;     // per group: S at +0 (128 columns), O at +128 (64 columns)
constexpr int AT_SMEM = 2 * AT_TILE + 2 * AT_P + 2 * AT_TILE + AT_STAGES * AT_TILE + 256 + 1024;
static_assert(AT_SMEM <= 227 * 1024, "attention smem budget");

struct AttnFwdParams {
  int L, B, H, D;      // tokens, sequences, heads, H * 64
  int G, RB;           // sequences per tile (2 when L <= 64), rows per sequence in the tile (128 / G)
  int nq, nkv, causal; // query / key tiles per sequence
  int items;           // ceil(B / G) * H * nq
  float scale_log2;    // softmax scale * log2(e)
  float* lse;          // [B, H, L]
};
struct alignas(64) AttnMaps {
  CUtensorMap qkv;  // [B][L][3D], box (64, RB, G)
  CUtensorMap out;  // [B][L][D],  box (64, RB, G)
};

struct AtItem {
  int b0, h, qt, kvn, T;
};
__device__ __forceinline__ AtItem at_decode(const AttnFwdParams& p, int idx) {
  AtItem it;
  it.qt = idx % p.nq;
  const int r = idx / p.nq;
  it.h = r % p.H;
  it.b0 = (r / p.H) * p.G;
  it.kvn = p.causal ? (it.qt + 1 < p.nkv ? it.qt + 1 : p.nkv) : p.nkv;  // causal: key tiles above the diagonal are skipped
  it.T = it.kvn == 1 ? 1 : 2 * it.kvn;
  return it;
}
__device__ __forceinline__ void at_step(const AtItem& it, int t, int& kt, bool& do_max, bool& do_exp) {
  if (it.kvn == 1) { kt = 0; do_max = true; do_exp = true; }
  else if (t < it.kvn) { kt = t; do_max = true; do_exp = false; }
  else { kt = t - it.kvn; do_max = false; do_exp = true; }
}

// One group's position in the shared schedule (walked identically by the producer and the MMA issuer).
struct AtStream {
  int idx;           // current item
  AtItem it;
  int t;             // next step whose QK^T is to be issued
  bool active;       // a QK^T remains for the current item
  bool pv_pending;   // the PV of the last issued step is still to be issued
  int pv_kt, pv_h, pv_b0;  // key tile / head / first sequence of the step that owes it
  bool pv_first, pv_last;
};
__device__ __forceinline__ void at_stream_init(AtStream& s, const AttnFwdParams& p, int g) {
  s.idx = blockIdx.x + g * gridDim.x;
  s.active = s.idx < p.items;
  s.pv_pending = false;
  s.t = 0;
  if (s.active) s.it = at_decode(p, s.idx);
}
// after the QK^T of step t has been issued: records the PV this step owes (if any) and moves on — to the next step, or to
// the group's next item when this was the item's last QK^T (the owed PV is issued later from the recorded copy)
__device__ __forceinline__ void at_stream_after_s(AtStream& s, const AttnFwdParams& p) {
  int kt;
  bool do_max, do_exp;
  at_step(s.it, s.t, kt, do_max, do_exp);
  const bool last = s.t == s.it.T - 1;
  s.pv_pending = do_exp;
  if (do_exp) {
    s.pv_kt = kt;
    s.pv_h = s.it.h;
    s.pv_b0 = s.it.b0;
    s.pv_first = kt == 0;
    s.pv_last = last;
  }
  ++s.t;
  if (last) {
    s.idx += 2 * gridDim.x;
    s.t = 0;
    s.active = s.idx < p.items;
    if (s.active) s.it = at_decode(p, s.idx);
  }
}

__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_fwd_kernel(const __grid_constant__ AttnMaps tm, const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                       // [2] Q tile of each group
  uint8_t* p_s = q_s + 2 * AT_TILE;          // [2] P tile of each group
  uint8_t* o_s = p_s + 2 * AT_P;             // [2] O staging (TMA store source)
  uint8_t* ring = o_s + 2 * AT_TILE;         // [AT_STAGES] K / V tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + AT_STAGES * AT_TILE);
  uint64_t* q_full = bars;                   // [2]
  uint64_t* q_empty = q_full + 2;            // [2]
  uint64_t* kv_full = q_empty + 2;           // [AT_STAGES]
  uint64_t* kv_empty = kv_full + AT_STAGES;  // [AT_STAGES]
  uint64_t* s_full = kv_empty + AT_STAGES;   // [2] MMA -> softmax: S ready
  uint64_t* s_free = s_full + 2;             // [2] softmax -> MMA: S read
  uint64_t* p_full = s_free + 2;             // [2] softmax -> MMA: P written
  uint64_t* pv_done = p_full + 2;            // [2] MMA -> softmax: PV retired (P reusable, O valid after the last one)
  uint64_t* o_free = pv_done + 2;            // [2] softmax -> MMA: O read
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.qkv);
    tma_prefetch_desc(&tm.out);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_free[i], 4);
    }
    for (int i = 0; i < AT_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp >= 4) {
    // the P tiles start as zeros: columns a row never writes (the other sequence's block, keys past the tile) stay 0
    const int g = (warp - 4) >> 2;
    const int r = ((warp - 4) & 3) * 32 + lane;
    uint4* row0 = reinterpret_cast<uint4*>(p_s + g * AT_P + r * 128);
    uint4* row1 = reinterpret_cast<uint4*>(p_s + g * AT_P + 16384 + r * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      row0[i] = make_uint4(0, 0, 0, 0);
      row1[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== producer: Q tiles + K/V ring, in the MMA issuer's order =====================
      AtStream st[2];
      at_stream_init(st[0], p, 0);
      at_stream_init(st[1], p, 1);
      uint32_t n_q[2] = {0, 0};
      int stage = 0;
      uint32_t phase = 0;
      auto ring_load = [&](const AtItem& it, int part, int kt) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], AT_TILE);
        tma_load_3d(ring + stage * AT_TILE, &tm.qkv, &kv_full[stage], part * p.D + it.h * 64, p.G == 1 ? kt * 128 : 0, it.b0);
        if (++stage == AT_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      // slot of a group = [QK^T of its next step] then [PV its previous step owes]: the next S is in flight while the
      // softmax of the previous one is still writing P
      while (st[0].active || st[0].pv_pending || st[1].active || st[1].pv_pending) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          AtStream& s = st[g];
          const bool owe = s.pv_pending;
          const int owe_kt = s.pv_kt;
          AtItem owe_it;               // the item of the step that owes the PV (s.it may already be the next item)
          owe_it.h = s.pv_h;
          owe_it.b0 = s.pv_b0;
          s.pv_pending = false;
          if (s.active) {
            int kt;
            bool dm, de;
            at_step(s.it, s.t, kt, dm, de);
            if (s.t == 0) {
              if (n_q[g] > 0) mbar_wait(&q_empty[g], (n_q[g] - 1) & 1);
              mbar_expect_tx(&q_full[g], AT_TILE);
              tma_load_3d(q_s + g * AT_TILE, &tm.qkv, &q_full[g], s.it.h * 64, p.G == 1 ? s.it.qt * 128 : 0, s.it.b0);
              ++n_q[g];
            }
            ring_load(s.it, 1, kt);  // K tile
            at_stream_after_s(s, p);
          }
          if (owe) ring_load(owe_it, 2, owe_kt);  // V tile of the step whose PV is owed
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S[128 x 128] = Q (K-major) x K^T (K-major)
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);  // O[128 x 64] += P (K-major) x V (MN-major)
      AtStream st[2];
      at_stream_init(st[0], p, 0);
      at_stream_init(st[1], p, 1);
      uint32_t n_s[2] = {0, 0}, n_p[2] = {0, 0}, n_q[2] = {0, 0}, n_o[2] = {0, 0};
      int stage = 0;
      uint32_t phase = 0;
      while (st[0].active || st[0].pv_pending || st[1].active || st[1].pv_pending) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          AtStream& s = st[g];
          const uint32_t t_s = tmem_base + g * AT_TMEM_GROUP;
          const bool owe = s.pv_pending;
          const int owe_kt = s.pv_kt;
          const bool owe_first = s.pv_first, owe_last = s.pv_last;
          s.pv_pending = false;
          if (s.active) {
            // ---- QK^T of the next step: needs only the S columns back (all rows loaded them), not the P tile
            if (s.t == 0) mbar_wait(&q_full[g], n_q[g] & 1);
            if (n_s[g] > 0) mbar_wait(&s_free[g], (n_s[g] - 1) & 1);
            mbar_wait(&kv_full[stage], phase);
            tc_fence_after();
            const uint32_t sq = smem_u32(q_s + g * AT_TILE);
            const uint32_t sk = smem_u32(ring + stage * AT_TILE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(t_s, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), idesc_s,
                        k > 0 ? 1u : 0u);
            umma_commit(&s_full[g]);
            umma_commit(&kv_empty[stage]);
            if (++stage == AT_STAGES) {
              stage = 0;
              phase ^= 1;
            }
            if (s.t == s.it.T - 1) {
              umma_commit(&q_empty[g]);  // the item's last QK^T: its Q tile may be replaced
              ++n_q[g];
            }
            ++n_s[g];
            at_stream_after_s(s, p);
          }
          if (owe) {
            // ---- PV owed by the previous step
            mbar_wait(&p_full[g], n_p[g] & 1);
            if (owe_first && n_o[g] > 0) mbar_wait(&o_free[g], (n_o[g] - 1) & 1);
            mbar_wait(&kv_full[stage], phase);
            tc_fence_after();
            const int kv_valid = p.G == 2 ? 128 : (p.L - owe_kt * 128 < 128 ? p.L - owe_kt * 128 : 128);
            const int nkk = (kv_valid + 15) >> 4;
            const uint32_t sp = smem_u32(p_s + g * AT_P);
            const uint32_t sv = smem_u32(ring + stage * AT_TILE);
            for (int kk = 0; kk < nkk; ++kk)
              umma_bf16(t_s + 128, umma_smem_desc(sp + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_smem_desc(sv + kk * 2048, 8192, 1024), idesc_pv, (owe_first && kk == 0) ? 0u : 1u);
            umma_commit(&pv_done[g]);
            umma_commit(&kv_empty[stage]);
            if (++stage == AT_STAGES) {
              stage = 0;
              phase ^= 1;
            }
            ++n_p[g];
            if (owe_last) ++n_o[g];
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax groups: one thread per query row =====================
    const int g = (warp - 4) >> 2;
    const int quarter = (warp - 4) & 3;
    const int r = quarter * 32 + lane;
    const bool leader = quarter == 0 && lane == 0;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * AT_TMEM_GROUP;
    uint8_t* pbuf = p_s + g * AT_P;
    uint8_t* obuf = o_s + g * AT_TILE;
    uint32_t n_s = 0, n_p = 0;
    for (int idx = blockIdx.x + g * gridDim.x; idx < p.items; idx += 2 * gridDim.x) {
      const AtItem it = at_decode(p, idx);
      const int seq = p.G == 2 ? (r >> 6) : 0;
      const int qi = p.G == 2 ? (r & 63) : it.qt * 128 + r;
      const int b = it.b0 + seq;
      const int cbase = seq * 64;  // this row's key block inside the tile (G == 2: block-diagonal)
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < it.T; ++t) {
        int kt;
        bool do_max, do_exp;
        at_step(it, t, kt, do_max, do_exp);
        const int kv0 = p.G == 2 ? 0 : kt * 128;
        const int kv_valid = p.G == 2 ? p.L : (p.L - kv0 < 128 ? p.L - kv0 : 128);
        int kmax = kv_valid;  // keys [0, kmax) of the block are visible to this row
        if (p.causal) {
          const int c = qi - kv0 + 1;
          kmax = c < kmax ? c : kmax;
          kmax = kmax < 1 ? 1 : kmax;  // rows past the sequence end are never stored; keep them finite
        }
        const int nchunk = (kv_valid + 31) >> 5;  // warp-uniform
        mbar_wait(&s_full[g], n_s & 1);
        tc_fence_after();
        // TMEM reads are software-pipelined: chunk j+1 is requested before the math of chunk j (two register sets,
        // the loop is written two chunks at a time so that the sets are never selected dynamically)
        uint32_t ra[32], rb[32];
        // `release`: this is the step's last sweep over S — the moment the last chunk is in registers the S columns are
        // handed back to the MMA issuer, so the next QK^T runs while this step's math / P stores are still going on
        auto hand_back_s = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[g]);
        };
        auto pipelined = [&](bool release, auto&& process) {
#if !AT_PIPELINED_LD
          // (measured: on this kernel the two-set pipelining costs more in register moves than the overlap returns)
          for (int j = 0; j < nchunk; ++j) {
            tmem_ld_32x32_issue(t_row + cbase + j * 32, ra);
            tmem_ld_wait(ra);
            if (release && j + 1 == nchunk) hand_back_s();
            process(ra, j);
          }
          (void)rb;
          return;
#endif
          tmem_ld_32x32_issue(t_row + cbase, ra);
          tmem_ld_wait(ra);
          for (int j = 0; j < nchunk; j += 2) {
            if (j + 1 < nchunk) tmem_ld_32x32_issue(t_row + cbase + (j + 1) * 32, rb);
            else if (release) hand_back_s();
            process(ra, j);
            if (j + 1 < nchunk) {
              tmem_ld_wait(rb);
              if (j + 2 < nchunk) tmem_ld_32x32_issue(t_row + cbase + (j + 2) * 32, ra);
              else if (release) hand_back_s();
              process(rb, j + 1);
              if (j + 2 < nchunk) tmem_ld_wait(ra);
            }
          }
        };
        if (do_max) {
          pipelined(!do_exp, [&](uint32_t (&cur)[32], int j) {
            if (__all_sync(0xffffffffu, (j + 1) * 32 <= kmax)) {  // every key of the chunk visible to every row of the warp
              float m0 = __uint_as_float(cur[0]), m1 = __uint_as_float(cur[1]), m2 = __uint_as_float(cur[2]),
                    m3 = __uint_as_float(cur[3]);
#pragma unroll
              for (int i = 4; i < 32; i += 4) {
                m0 = fmaxf(m0, __uint_as_float(cur[i]));
                m1 = fmaxf(m1, __uint_as_float(cur[i + 1]));
                m2 = fmaxf(m2, __uint_as_float(cur[i + 2]));
                m3 = fmaxf(m3, __uint_as_float(cur[i + 3]));
              }
              m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * p.scale_log2);  // scale > 0: max commutes with it
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                m = fmaxf(m, (j * 32 + i < kmax) ? __uint_as_float(cur[i]) * p.scale_log2 : -INFINITY);
            }
          });
        }
        if (do_exp) {
          if (n_p > 0) mbar_wait(&pv_done[g], (n_p - 1) & 1);  // the previous PV has finished reading the P tile
          const f32x2 sc2 = f2_splat(p.scale_log2), nm2 = f2_splat(-m);
          pipelined(true, [&](uint32_t (&cur)[32], int j) {
            float v[32];
            f32x2 acc2 = f2_splat(0.f);
            if (__all_sync(0xffffffffu, (j + 1) * 32 <= kmax)) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float a0, a1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, nm2), a0, a1);
                v[i] = ex2_approx(a0);
                v[i + 1] = ex2_approx(a1);
                acc2 = f2_add(acc2, f2_pack(v[i], v[i + 1]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float a0, a1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, nm2), a0, a1);
                v[i] = (j * 32 + i < kmax) ? ex2_approx(a0) : 0.f;
                v[i + 1] = (j * 32 + i + 1 < kmax) ? ex2_approx(a1) : 0.f;
                acc2 = f2_add(acc2, f2_pack(v[i], v[i + 1]));
              }
            }
            float s0, s1;
            f2_unpack(acc2, s0, s1);
            l += s0 + s1;
            const int c0 = cbase + j * 32;  // first column of this chunk inside the P tile
            uint8_t* rowp = pbuf + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float t8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) t8[i] = v[q4 * 8 + i];
              const int ch = ((c0 & 63) >> 3) + q4;
              *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) = pack_bf16x8(t8);
            }
          });
          fence_proxy_async_smem();
        }
        if (do_exp) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[g]);
        }
        ++n_s;
        if (do_exp) ++n_p;
      }
      // ---- O epilogue: O / l -> bf16 -> swizzled staging tile -> one TMA store (rows past L / B are clipped)
      mbar_wait(&pv_done[g], (n_p - 1) & 1);
      tc_fence_after();
      if (leader) tma_store_wait_read<0>();  // the previous item's store has read the staging tile
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      const float inv = 1.f / l;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[32];
        tmem_ld_32x32(t_row + 128 + j * 32, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = v[q4 * 8 + i] * inv;
          const int ch = j * 4 + q4;
          *reinterpret_cast<uint4*>(obuf + r * 128 + ((ch ^ (r & 7)) << 4)) = pack_bf16x8(t8);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (leader) {
        tma_store_3d(&tm.out, obuf, it.h * 64, p.G == 1 ? it.qt * 128 : 0, it.b0);
        tma_store_commit();
      }
      if (qi < p.L && b < p.B)
        p.lse[(static_cast<int64_t>(b) * p.H + it.h) * p.L + qi] = (m + lg2_approx(l)) * 0.69314718055994531f;
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================================
// Backward, sequences of at most 128 tokens (both CLIP ViT-B towers: 50 / 77; ViT-L text: 77) — one tile per item,
// everything in one pass:
//   S = Q K^T, dP = dO V^T               (TMEM, 2 x 128 columns)
//   P = exp(S - lse), dS = P o (dP - D) * scale                          (8 warps: TMEM lane quarter x column half)
//       D = rowsum(dO o O) = sum_key P dP  (O = P V; the whole key range of a row is in this tile), so the O tile is
//       never loaded: the two column halves of a row exchange their fp32 partial sums through shared memory
//   dV = P^T dO, dK = dS^T Q, dQ = dS K  (TMEM, 3 x 64 columns; dV on its own commit, drained while dK / dQ run)
// P and dS are written once to shared memory in the SW128 layout that is at the same time the K-major operand
// [q][key] (dQ = dS K) and the MN-major operand [key][q]^T (dV = P^T dO, dK = dS^T Q): no transposition anywhere.
// Likewise every input tile [rows x 64] serves as K-major operand (contraction over head_dim) and MN-major operand
// (contraction over rows) in its natural TMA layout.  Outputs leave by three TMA stores from staging tiles (their own
// 48 KB when two sequences share a tile, else the input stage's Q / K / V tiles); the in_proj bias gradient (column sums
// of the staged dQ / dV; the K third is identically zero) is summed by warps 2 / 3 in registers and flushed with one
// atomicAdd per column when the CTA's contiguous item range changes head.
// ===================================================================================================
// Shared-memory plan (bytes): two input stages [Q, K, V, dO] = 128 K, then
//   G == 2 (L <= 64, two sequences per tile): output staging [dQ, dK, dV] 48 K | P 24 K | dS 24 K — P and dS are block
//           diagonal, stored as [seq-0 block | 8 K of zeros | seq-1 block] with the two 64-column panels OVERLAPPING on
//           the zero block (panel stride 8 K instead of 16 K): the input stage is free the moment the output MMAs
//           retire (tcgen05.commit arrives `in_empty`), the stores drain from their own tiles
//   G == 1 (64 < L <= 128): P 32 K | dS 32 K, outputs staged over the stage's own Q / K / V tiles
constexpr int AT_BWD_SMEM = 2 * 4 * AT_TILE + 3 * AT_TILE + 2 * (AT_P - 8192) + 2 * 128 * 4 + 256 + 1024;
static_assert(AT_BWD_SMEM <= 227 * 1024, "attention bwd smem budget");

struct AttnBwdParams {
  int L, B, H, D, G, RB, causal;
  int items, nb;       // nb = ceil(B / G); item idx = h * nb + bi
  float scale, scale_log2;
  const float* lse;    // [B, H, L]
  const __nv_bfloat16* out;  // [B*L, D]
  float* dbias;        // [3D] or null
};
struct alignas(64) AttnBwdMaps {
  CUtensorMap qkv;   // [B][L][3D] loads
  CUtensorMap dout;  // [B][L][D]  loads
  CUtensorMap out;   // [B][L][D]  loads (forward output)
  CUtensorMap dqkv;  // [B][L][3D] stores
};

// lane l receives the sum over the warp's 32 rows of column l (v is destroyed) — see epi_col_sum in gemm.cu
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32]) {
  const int lane = lane_id();
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
  return v[0];
}

__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_bwd_kernel(const __grid_constant__ AttnBwdMaps tm, const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const bool two = p.G == 2;
  const int pan = two ? 8192 : 16384;       // stride between the two 64-column panels of P / dS
  const int pbytes = two ? AT_P - 8192 : AT_P;
  uint8_t* in_s = smem;                     // [2 stages][Q, K, V, dO]
  uint8_t* out_s = in_s + 2 * 4 * AT_TILE;  // G == 2 only: [dQ, dK, dV] staging
  uint8_t* p_s = two ? out_s + 3 * AT_TILE : out_s;  // P tile
  uint8_t* ds_s = p_s + pbytes;             // dS tile
  float* xch = reinterpret_cast<float*>(smem + 2 * 4 * AT_TILE + 3 * AT_TILE + 2 * (AT_P - 8192));  // [2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 256);
  uint64_t* in_full = bars;       // [2]
  uint64_t* in_empty = bars + 2;  // [2]
  uint64_t* sdp_full = bars + 4;  // S and dP in TMEM
  uint64_t* pds_full = bars + 5;  // P and dS in shared memory (8 warps)
  uint64_t* out_full = bars + 6;  // dV, dK, dQ in TMEM
  uint64_t* acc_free = bars + 7;  // TMEM read by the epilogue (8 warps)
  // `staged` and `sums_done` have TWO barriers each, alternating by item: with G == 1 nothing stops the epilogue warps
  // (resp. warp 2) from completing item n+1's phase while a consumer is still waiting for item n's — on a single barrier
  // the parity would then alias and the waiter hang (seen on 8 GPUs under NCCL traffic).  A consumer never lags by
  // more than one item (item n+2 needs the input stage that item n's consumers release), so two barriers suffice.
  uint64_t* staged = bars + 8;    // [2] dQ / dK / dV staged in the stage's tiles (8 warps) -> warps 2 and 3
  uint64_t* sums_done = bars + 10; // [2] warp 2 -> warp 3
  uint64_t* stg_free = bars + 12; // G == 2: the staging tiles have been read by the stores and the sums (warp 3)
  uint64_t* p_full = bars + 13;   // P in shared memory (8 warps): dV's MMAs start while dS is still being computed
  uint64_t* dv_full = bars + 14;  // dV in TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // contiguous item range of this CTA (items are head-major: a CTA sees at most a couple of heads)
  const int i0 = static_cast<int>((static_cast<int64_t>(p.items) * blockIdx.x) / gridDim.x);
  const int i1 = static_cast<int>((static_cast<int64_t>(p.items) * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.qkv);
    tma_prefetch_desc(&tm.dout);
    tma_prefetch_desc(&tm.dqkv);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(out_full, 1);
    mbar_init(acc_free, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&staged[i], 8);
      mbar_init(&sums_done[i], 1);
    }
    mbar_init(stg_free, 1);
    mbar_init(p_full, 8);
    mbar_init(dv_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp >= 4) {
    // P / dS start as zeros and are never used as staging: the zero block between the two sequences' blocks (G == 2)
    // and the key columns no row writes stay zero for the whole kernel
    const int t = threadIdx.x - 128;  // 0..255
    uint4* z = reinterpret_cast<uint4*>(p_s);
    for (int i = t; i < 2 * pbytes / 16; i += 256) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int nk = p.G == 2 ? 8 : (p.L + 15) >> 4;  // contraction steps over rows (queries or keys): 16 rows each

  if (warp == 0) {
    if (elect_one()) {
      // ===================== producer =====================
      for (int idx = i0; idx < i1; ++idx) {
        const int n = idx - i0, stage = n & 1;
        const int h = idx / p.nb, b0 = (idx % p.nb) * p.G;
        if (n >= 2) mbar_wait(&in_empty[stage], ((n >> 1) - 1) & 1);
        uint8_t* st = in_s + stage * 4 * AT_TILE;
        mbar_expect_tx(&in_full[stage], 4 * AT_TILE);
        tma_load_3d(st, &tm.qkv, &in_full[stage], h * 64, 0, b0);
        tma_load_3d(st + AT_TILE, &tm.qkv, &in_full[stage], p.D + h * 64, 0, b0);
        tma_load_3d(st + 2 * AT_TILE, &tm.qkv, &in_full[stage], 2 * p.D + h * 64, 0, b0);
        tma_load_3d(st + 3 * AT_TILE, &tm.dout, &in_full[stage], h * 64, 0, b0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);   // [128 x 128] = A (K-major) x B (K-major), K = head_dim
      const uint32_t id_t = umma_idesc_bf16(128, 64, 1, 1);    // [keys x 64] = A^T (MN-major) x B (MN-major), K = queries
      const uint32_t id_q = umma_idesc_bf16(128, 64, 0, 1);    // [queries x 64] = A (K-major) x B (MN-major), K = keys
      const uint32_t sp = smem_u32(p_s), sds = smem_u32(ds_s);
      // S / dP of item n+1 are issued right behind the output MMAs of item n (their TMEM columns are free as soon as
      // item n's P / dS are in shared memory), so the softmax-gradient warps find them ready after item n's epilogue
      auto issue_sdp = [&](int n) {
        const int stage = n & 1;
        const uint32_t sq = smem_u32(in_s + stage * 4 * AT_TILE), sk = sq + AT_TILE, sv = sq + 2 * AT_TILE, sdo = sq + 3 * AT_TILE;
        mbar_wait(&in_full[stage], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + 128, umma_smem_desc(sdo + k * 32, 16, 1024), umma_smem_desc(sv + k * 32, 16, 1024), id_s,
                    k > 0);
        umma_commit(sdp_full);
      };
      if (i0 < i1) issue_sdp(0);
      for (int idx = i0; idx < i1; ++idx) {
        const int n = idx - i0, stage = n & 1;
        const uint32_t sq = smem_u32(in_s + stage * 4 * AT_TILE), sk = sq + AT_TILE, sdo = sq + 3 * AT_TILE;
        mbar_wait(p_full, n & 1);
        if (n > 0) mbar_wait(acc_free, (n - 1) & 1);  // the previous item's dV / dK / dQ have been read out of TMEM
        tc_fence_after();
        for (int kk = 0; kk < nk; ++kk)  // dV[key, d] = sum_q P[q, key] dO[q, d]
          umma_bf16(tmem_base + 256, umma_smem_desc(sp + kk * 2048, pan, 1024), umma_smem_desc(sdo + kk * 2048, 8192, 1024),
                    id_t, kk > 0);
        umma_commit(dv_full);  // drained by the epilogue warps while dK / dQ are still in the pipe
        mbar_wait(pds_full, n & 1);
        tc_fence_after();
        for (int kk = 0; kk < nk; ++kk)  // dK[key, d] = sum_q dS[q, key] Q[q, d]
          umma_bf16(tmem_base + 320, umma_smem_desc(sds + kk * 2048, pan, 1024), umma_smem_desc(sq + kk * 2048, 8192, 1024),
                    id_t, kk > 0);
        for (int kk = 0; kk < nk; ++kk)  // dQ[q, d] = sum_key dS[q, key] K[key, d]
          umma_bf16(tmem_base + 384, umma_smem_desc(sds + (kk >> 2) * pan + (kk & 3) * 32, 16, 1024),
                    umma_smem_desc(sk + kk * 2048, 8192, 1024), id_q, kk > 0);
        umma_commit(out_full);
        if (two) umma_commit(&in_empty[stage]);  // nothing reads the stage's tiles any more: the producer may refill it
        if (idx + 1 < i1) issue_sdp(n + 1);
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== warps 2 / 3: TMA stores + the in_proj bias gradient, off the compute warps' critical path ====
    // Column sums of the STAGED (bf16-rounded) tiles — warp 2: dQ, warp 3: dV (+ the three TMA stores).  Lane c owns
    // columns 2c, 2c+1 and walks the 128 rows: a row's 32 lanes read its 128 bytes exactly once (conflict-free in the
    // SW128 layout).  The K third is not summed: sum_key dS[q, key] = scale * (sum_key P dP - D sum_key P) = 0 — a key
    // bias has no gradient (softmax is invariant to a common shift of the keys).
    const int tile = warp == 2 ? 0 : 2;  // Q slot holds dQ, V slot holds dV
    float a0 = 0.f, a1 = 0.f;
    int cur_h = -1;
    auto flush = [&]() {
      if (p.dbias != nullptr && cur_h >= 0) {
        float* d = p.dbias + tile * p.D + cur_h * 64 + 2 * lane;
        atomicAdd(d, a0);
        atomicAdd(d + 1, a1);
      }
      a0 = a1 = 0.f;
    };
    for (int idx = i0; idx < i1; ++idx) {
      const int n = idx - i0, stage = n & 1;
      const int h = idx / p.nb, b0 = (idx % p.nb) * p.G;
      uint8_t* st = two ? out_s : in_s + stage * 4 * AT_TILE;  // where the epilogue staged dQ | dK | dV
      if (h != cur_h) {
        flush();
        cur_h = h;
      }
      mbar_wait(&staged[n & 1], (n >> 1) & 1);
      if (warp == 3 && lane == 0) {
        tma_store_3d(&tm.dqkv, st, h * 64, 0, b0);
        tma_store_3d(&tm.dqkv, st + AT_TILE, p.D + h * 64, 0, b0);
        tma_store_3d(&tm.dqkv, st + 2 * AT_TILE, 2 * p.D + h * 64, 0, b0);
        tma_store_commit();
      }
      if (p.dbias != nullptr) {
        const uint8_t* base = st + tile * AT_TILE + (lane & 3) * 4;
        float b0s = 0.f, b1s = 0.f, c0s = 0.f, c1s = 0.f;  // four independent chains
#pragma unroll 8
        for (int r = 0; r < 128; r += 2) {
          const uint32_t w0 = *reinterpret_cast<const uint32_t*>(base + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
          const uint32_t w1 = *reinterpret_cast<const uint32_t*>(base + (r + 1) * 128 + (((lane >> 2) ^ ((r + 1) & 7)) << 4));
          b0s += __uint_as_float(w0 << 16);
          b1s += __uint_as_float(w0 & 0xffff0000u);
          c0s += __uint_as_float(w1 << 16);
          c1s += __uint_as_float(w1 & 0xffff0000u);
        }
        a0 += b0s + c0s;
        a1 += b1s + c1s;
      }
      __syncwarp();
      if (warp == 2) {
        if (lane == 0) mbar_arrive(&sums_done[n & 1]);  // the dQ tile has been summed
      } else if (lane == 0) {
        tma_store_wait_read<0>();       // the staged tiles have been read by the stores ...
        mbar_wait(&sums_done[n & 1], (n >> 1) & 1);    // ... and by warp 2: they may be overwritten
        mbar_arrive(two ? stg_free : &in_empty[stage]);
      }
      __syncwarp();
    }
    flush();
    if (warp == 3 && lane == 0) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ===================== softmax-gradient + epilogue warps: (TMEM lane quarter, column half) =====================
    const int quarter = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int seq = p.G == 2 ? (r >> 6) : 0;
    const int qi = p.G == 2 ? (r & 63) : r;
    const int cbase = seq * 64;
    int kmax = p.L;
    if (p.causal) kmax = qi + 1 < kmax ? qi + 1 : kmax;
    const int nchunk = p.G == 2 ? 2 : (p.L + 31) >> 5;   // 32-column chunks holding keys of this row's block
    const int per = p.G == 2 ? 1 : 2;                    // chunks per column half
    // Row constants: the log-sum-exp is fetched one item AHEAD; D = rowsum(dO o O) is never read from memory — with
    // the whole key range of a row in one tile it equals sum_key P dP (O = P V), which the two column halves of a row
    // accumulate in fp32 from the accumulators they hold anyway and exchange through shared memory.
    // (returns the raw value: the conversion to log2 units happens where it is consumed, one item later, so the load
    // has a whole item to land)
    auto row_lse = [&](int idx, bool& valid) -> float {
      const int h = idx / p.nb, b = (idx % p.nb) * p.G + seq;
      valid = qi < p.L && b < p.B;
      return valid ? __ldg(p.lse + (static_cast<int64_t>(b) * p.H + h) * p.L + qi) : 0.f;
    };
    bool valid = false, nvalid_row = false;
    float lse_raw = 0.f, nlse_raw = 0.f;
    if (i0 < i1) lse_raw = row_lse(i0, valid);
    const int prow = two ? (r & 63) : r;                 // row inside the row's own P / dS block
    const int pblk = two ? seq * 16384 : 0;              // G == 2: [seq-0 block | zeros | seq-1 block]
    for (int idx = i0; idx < i1; ++idx) {
      const int n = idx - i0, stage = n & 1;
      uint8_t* st = two ? out_s : in_s + stage * 4 * AT_TILE;
      if (idx + 1 < i1) nlse_raw = row_lse(idx + 1, nvalid_row);
      mbar_wait(sdp_full, n & 1);
      tc_fence_after();
      const float lse2 = lse_raw * kLog2eAt;
      // pass 1: P (stored, and kept as packed bf16) and this half's share of D
      uint32_t pk[2][16];
      float part = 0.f;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = half * per + jj;
        if (jj < per && j < nchunk) {
          float sv[32], dp[32];
          tmem_ld_32x32(t_row + cbase + j * 32, sv);
          tmem_ld_32x32(t_row + 128 + cbase + j * 32, dp);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = valid && (j * 32 + i < kmax);
            const float pr = ok ? ex2_approx(fmaf(sv[i], p.scale_log2, -lse2)) : 0.f;
            sv[i] = pr;
            part = fmaf(pr, dp[i], part);
          }
          const int c0 = j * 32;  // first key of the chunk inside the row's block
          uint8_t* rowp = p_s + pblk + (two ? 0 : (c0 >> 6) * 16384) + prow * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float t8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = sv[q4 * 8 + i];
            const uint4 w = pack_bf16x8(t8);
            pk[jj][q4 * 4] = w.x;
            pk[jj][q4 * 4 + 1] = w.y;
            pk[jj][q4 * 4 + 2] = w.z;
            pk[jj][q4 * 4 + 3] = w.w;
            *reinterpret_cast<uint4*>(rowp + (((((c0 & 63) >> 3) + q4) ^ (r & 7)) << 4)) = w;
          }
        }
      }
      xch[half * 128 + r] = part;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the row's other column half (warp +-4)
      const float drow = part + xch[(1 - half) * 128 + r];
      // pass 2: dS = P (dP - D) scale
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = half * per + jj;
        if (jj < per && j < nchunk) {
          float dp[32];
          tmem_ld_32x32(t_row + 128 + cbase + j * 32, dp);
          const int c0 = j * 32;
          uint8_t* rowp = ds_s + pblk + (two ? 0 : (c0 >> 6) * 16384) + prow * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float u8[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w = pk[jj][q4 * 4 + i];
              u8[2 * i] = __uint_as_float(w << 16) * (dp[q4 * 8 + 2 * i] - drow) * p.scale;
              u8[2 * i + 1] = __uint_as_float(w & 0xffff0000u) * (dp[q4 * 8 + 2 * i + 1] - drow) * p.scale;
            }
            *reinterpret_cast<uint4*>(rowp + (((((c0 & 63) >> 3) + q4) ^ (r & 7)) << 4)) = pack_bf16x8(u8);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // ---- epilogue: dV | dK | dQ columns [half*32, half*32+32) of row r -> bf16 -> staging tile; dV first (its MMAs
      // ran during pass 2), dK / dQ when theirs retire
      mbar_wait(dv_full, n & 1);
      if (two && n > 0) mbar_wait(stg_free, (n - 1) & 1);  // the previous item's stores / sums are done with the tiles
      tc_fence_after();
#pragma unroll 1
      for (int k3 = 0; k3 < 3; ++k3) {  // tile 2: dV (TMEM 256, V tile), 1: dK (320, K tile), 0: dQ (384, Q tile)
        const int part_i = 2 - k3;
        if (k3 == 1) {
          mbar_wait(out_full, n & 1);
          tc_fence_after();
        }
        float v[32];
        tmem_ld_32x32(t_row + 384 - part_i * 64 + half * 32, v);
        uint8_t* dst = st + part_i * AT_TILE + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = v[q4 * 8 + i];
          *reinterpret_cast<uint4*>(dst + (((half * 4 + q4) ^ (r & 7)) << 4)) = pack_bf16x8(t8);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_free);
        mbar_arrive(&staged[n & 1]);
      }
      valid = nvalid_row;
      lse_raw = nlse_raw;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================================
// Backward, sequences longer than 128 tokens (ViT-B/16: 197, ViT-L/14-336: 577): two launches of one kernel.
//   kDKDV : item = (sequence, head, KEY tile j): K_j, V_j stay in shared memory, the query tiles i stream past
//           them; dV_j += P_ij^T dO_i and dK_j += dS_ij^T Q_i accumulate in TMEM over i and are stored once
//   kDQ   : item = (sequence, head, QUERY tile i): Q_i, dO_i stay, the key tiles stream; dQ_i += dS_ij K_j
// S and dP are recomputed in both launches (4 of the 7-9 MMAs per tile pair: the tensor pipe waits for the MUFU
// anyway) so that no gradient is ever accumulated through global memory.  Same P / dS shared-memory tiles, same
// operand-layout reuse and the same 8-warp softmax-gradient stage as the single-tile kernel above.
// ===================================================================================================
constexpr int AT_BWDL_SMEM = 2 * 2 * AT_TILE + 2 * 2 * AT_TILE + 2 * AT_P + 192 * 4 + 256 + 1024;
static_assert(AT_BWDL_SMEM <= 227 * 1024, "attention bwd (long) smem budget");

struct AttnBwdLongParams {
  int L, B, H, D, nt, causal;  // nt = ceil(L / 128) tiles per sequence
  int items;                   // B * H * nt, item idx = (h * B + b) * nt + s  (head-major: few heads per CTA)
  float scale, scale_log2;
  const float* lse;
  const __nv_bfloat16* out;
  float* dbias;
};

template <bool kDKDV>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_bwd_long_kernel(const __grid_constant__ AttnBwdMaps tm, const __grid_constant__ AttnBwdLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stat_base = smem;                 // [2 items] stationary pair: kDKDV ? (K_j, V_j) : (Q_i, dO_i)
  uint8_t* str_s = stat_base + 2 * 2 * AT_TILE;  // [2 stages] streamed pair: kDKDV ? (Q_i, dO_i) : (K_j, V_j)
  uint8_t* p_s = str_s + 2 * 2 * AT_TILE;    // P tile
  uint8_t* ds_s = p_s + AT_P;                // dS tile
  float* bias_s = reinterpret_cast<float*>(ds_s + AT_P);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 192);
  uint64_t* stat_full = bars;      // [2] stationary pair of item n & 1 landed
  uint64_t* stat_empty = bars + 2; // [2] its tiles (and the output staging over them) have been read
  uint64_t* str_full = bars + 4;   // [2]
  uint64_t* str_empty = bars + 6;  // [2]
  uint64_t* sdp_full = bars + 8;
  uint64_t* pds_full = bars + 9;   // 8 warps
  uint64_t* out_full = bars + 10;
  uint64_t* acc_free = bars + 11;  // 8 warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int i0 = static_cast<int>((static_cast<int64_t>(p.items) * blockIdx.x) / gridDim.x);
  const int i1 = static_cast<int>((static_cast<int64_t>(p.items) * (blockIdx.x + 1)) / gridDim.x);
  // inner range of an item with stationary tile s: the other side's tiles [t_lo, t_hi)
  auto inner = [&](int s, int& t_lo, int& t_hi) {
    if (kDKDV) { t_lo = p.causal ? s : 0; t_hi = p.nt; }       // query tiles that see key tile s
    else { t_lo = 0; t_hi = p.causal ? s + 1 : p.nt; }         // key tiles visible to query tile s
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.qkv);
    tma_prefetch_desc(&tm.dout);
    tma_prefetch_desc(&tm.dqkv);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&stat_full[i], 1);
      mbar_init(&stat_empty[i], 1);
      mbar_init(&str_full[i], 1);
      mbar_init(&str_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(out_full, 1);
    mbar_init(acc_free, 8);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp >= 4) {
    const int t = threadIdx.x - 128;
    uint4* z = reinterpret_cast<uint4*>(p_s);
    for (int i = t; i < 2 * AT_P / 16; i += 256) z[i] = make_uint4(0, 0, 0, 0);
    if (t < 192) bias_s[t] = 0.f;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== producer =====================
      uint32_t n_str = 0;
      for (int idx = i0; idx < i1; ++idx) {
        const int n = idx - i0, slot = n & 1;
        const int s = idx % p.nt, b = (idx / p.nt) % p.B, h = idx / (p.nt * p.B);
        uint8_t* stat_s = stat_base + slot * 2 * AT_TILE;
        if (n >= 2) mbar_wait(&stat_empty[slot], ((n >> 1) - 1) & 1);
        mbar_expect_tx(&stat_full[slot], 2 * AT_TILE);
        if (kDKDV) {
          tma_load_3d(stat_s, &tm.qkv, &stat_full[slot], p.D + h * 64, s * 128, b);               // K_s
          tma_load_3d(stat_s + AT_TILE, &tm.qkv, &stat_full[slot], 2 * p.D + h * 64, s * 128, b);  // V_s
        } else {
          tma_load_3d(stat_s, &tm.qkv, &stat_full[slot], h * 64, s * 128, b);                      // Q_s
          tma_load_3d(stat_s + AT_TILE, &tm.dout, &stat_full[slot], h * 64, s * 128, b);           // dO_s
        }
        int t_lo, t_hi;
        inner(s, t_lo, t_hi);
        for (int t = t_lo; t < t_hi; ++t, ++n_str) {
          const int stage = n_str & 1;
          if (n_str >= 2) mbar_wait(&str_empty[stage], ((n_str >> 1) - 1) & 1);
          uint8_t* st = str_s + stage * 2 * AT_TILE;
          mbar_expect_tx(&str_full[stage], 2 * AT_TILE);
          if (kDKDV) {
            tma_load_3d(st, &tm.qkv, &str_full[stage], h * 64, t * 128, b);                 // Q_t
            tma_load_3d(st + AT_TILE, &tm.dout, &str_full[stage], h * 64, t * 128, b);      // dO_t
          } else {
            tma_load_3d(st, &tm.qkv, &str_full[stage], p.D + h * 64, t * 128, b);           // K_t
            tma_load_3d(st + AT_TILE, &tm.qkv, &str_full[stage], 2 * p.D + h * 64, t * 128, b);  // V_t
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t id_t = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t id_q = umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t sp = smem_u32(p_s), sds = smem_u32(ds_s);
      uint32_t n_str = 0;
      for (int idx = i0; idx < i1; ++idx) {
        const int n = idx - i0, slot = n & 1;
        const int s = idx % p.nt;
        const uint32_t s0 = smem_u32(stat_base + slot * 2 * AT_TILE), s1 = s0 + AT_TILE;
        int t_lo, t_hi;
        inner(s, t_lo, t_hi);
        mbar_wait(&stat_full[slot], (n >> 1) & 1);
        if (n > 0) mbar_wait(acc_free, (n - 1) & 1);
        for (int t = t_lo; t < t_hi; ++t, ++n_str) {
          const int stage = n_str & 1;
          const uint32_t r0 = smem_u32(str_s + stage * 2 * AT_TILE), r1 = r0 + AT_TILE;
          // operand roles: Q / dO tiles index queries, K / V tiles index keys
          const uint32_t sq = kDKDV ? r0 : s0, sdo = kDKDV ? r1 : s1, sk = kDKDV ? s0 : r0, sv = kDKDV ? s1 : r1;
          const int qt = kDKDV ? t : s, kt = kDKDV ? s : t;
          mbar_wait(&str_full[stage], (n_str >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), id_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + 128, umma_smem_desc(sdo + k * 32, 16, 1024), umma_smem_desc(sv + k * 32, 16, 1024), id_s,
                      k > 0);
          umma_commit(sdp_full);
          mbar_wait(pds_full, n_str & 1);
          tc_fence_after();
          const int q_valid = p.L - qt * 128 < 128 ? p.L - qt * 128 : 128;
          const int k_valid = p.L - kt * 128 < 128 ? p.L - kt * 128 : 128;
          const uint32_t acc = (t > t_lo) ? 1u : 0u;
          if (kDKDV) {
            const int nk = (q_valid + 15) >> 4;  // contraction over the queries of tile t
            for (int kk = 0; kk < nk; ++kk)
              umma_bf16(tmem_base + 256, umma_smem_desc(sp + kk * 2048, 16384, 1024),
                        umma_smem_desc(sdo + kk * 2048, 8192, 1024), id_t, (kk > 0) ? 1u : acc);
            for (int kk = 0; kk < nk; ++kk)
              umma_bf16(tmem_base + 320, umma_smem_desc(sds + kk * 2048, 16384, 1024),
                        umma_smem_desc(sq + kk * 2048, 8192, 1024), id_t, (kk > 0) ? 1u : acc);
          } else {
            const int nk = (k_valid + 15) >> 4;  // contraction over the keys of tile t
            for (int kk = 0; kk < nk; ++kk)
              umma_bf16(tmem_base + 256, umma_smem_desc(sds + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_smem_desc(sk + kk * 2048, 8192, 1024), id_q, (kk > 0) ? 1u : acc);
          }
          umma_commit(&str_empty[stage]);
        }
        umma_commit(out_full);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax-gradient + epilogue warps =====================
    const int quarter = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const bool leader = warp == 4 && lane == 0;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int cur_h = -1;
    uint32_t n_str = 0;
    auto flush_bias = [&]() {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int t = threadIdx.x - 128;
      if (t < 192 && p.dbias != nullptr && cur_h >= 0) {
        atomicAdd(p.dbias + (t >> 6) * p.D + cur_h * 64 + (t & 63), bias_s[t]);
        bias_s[t] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int idx = i0; idx < i1; ++idx) {
      const int n = idx - i0, slot = n & 1;
      const int s = idx % p.nt, b = (idx / p.nt) % p.B, h = idx / (p.nt * p.B);
      uint8_t* stat_s = stat_base + slot * 2 * AT_TILE;
      if (h != cur_h) {
        if (cur_h >= 0) flush_bias();
        cur_h = h;
      }
      int t_lo, t_hi;
      inner(s, t_lo, t_hi);
      mbar_wait(&stat_full[slot], (n >> 1) & 1);
      for (int t = t_lo; t < t_hi; ++t, ++n_str) {
        const int stage = n_str & 1;
        const int qt = kDKDV ? t : s, kt = kDKDV ? s : t;
        const uint8_t* do_tile = kDKDV ? str_s + stage * 2 * AT_TILE + AT_TILE : stat_s + AT_TILE;
        const int qi = qt * 128 + r;
        const bool valid = qi < p.L;
        const int kv0 = kt * 128;
        const int k_valid = p.L - kv0 < 128 ? p.L - kv0 : 128;
        int kmax = k_valid;
        if (p.causal) {
          const int c = qi - kv0 + 1;
          kmax = c < kmax ? c : kmax;
        }
        mbar_wait(&str_full[stage], (n_str >> 1) & 1);
        float lse2 = 0.f, drow = 0.f;
        if (valid) {
          lse2 = __ldg(p.lse + (static_cast<int64_t>(b) * p.H + h) * p.L + qi) * kLog2eAt;
          const uint4* orow = reinterpret_cast<const uint4*>(p.out + (static_cast<int64_t>(b) * p.L + qi) * p.D + h * 64);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float a[8], d8[8];
            unpack_bf16x8(__ldg(orow + c), a);
            unpack_bf16x8(*reinterpret_cast<const uint4*>(do_tile + r * 128 + ((c ^ (r & 7)) << 4)), d8);
#pragma unroll
            for (int i = 0; i < 8; ++i) drow = fmaf(a[i], d8[i], drow);
          }
        }
        mbar_wait(sdp_full, n_str & 1);
        tc_fence_after();
        const int nchunk = (k_valid + 31) >> 5;
        for (int jj = 0; jj < 2; ++jj) {
          const int j = half * 2 + jj;
          if (j >= nchunk) break;
          float sv[32], dp[32];
          tmem_ld_32x32(t_row + j * 32, sv);
          tmem_ld_32x32(t_row + 128 + j * 32, dp);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = valid && (j * 32 + i < kmax);
            const float pr = ok ? ex2_approx(fmaf(sv[i], p.scale_log2, -lse2)) : 0.f;
            sv[i] = pr;
            dp[i] = pr * (dp[i] - drow) * p.scale;
          }
          const int c0 = j * 32;
          const int off = (c0 >> 6) * 16384 + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float t8[8], u8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              t8[i] = sv[q4 * 8 + i];
              u8[i] = dp[q4 * 8 + i];
            }
            const int ch = ((((c0 & 63) >> 3) + q4) ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(p_s + off + ch) = pack_bf16x8(t8);
            *reinterpret_cast<uint4*>(ds_s + off + ch) = pack_bf16x8(u8);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
      }
      // ---- item epilogue: accumulators -> bf16 -> the stationary pair's tiles -> TMA store
      mbar_wait(out_full, n & 1);
      tc_fence_after();
      constexpr int kParts = kDKDV ? 2 : 1;
#pragma unroll 1
      for (int pi = 0; pi < kParts; ++pi) {
        // kDKDV: pi 0 = dV (TMEM 256) staged over V (stat tile 1), pi 1 = dK (TMEM 320) over K (stat tile 0); kDQ: dQ over Q
        const int part = kDKDV ? (pi == 0 ? 2 : 1) : 0;  // column third of dqkv: 0 q, 1 k, 2 v
        float v[32];
        tmem_ld_32x32(t_row + 256 + pi * 64 + half * 32, v);
        uint8_t* dst = stat_s + (kDKDV ? (pi == 0 ? AT_TILE : 0) : 0) + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = v[q4 * 8 + i];
          const uint4 pk = pack_bf16x8(t8);
          *reinterpret_cast<uint4*>(dst + (((half * 4 + q4) ^ (r & 7)) << 4)) = pk;
          unpack_bf16x8(pk, t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[q4 * 8 + i] = (s * 128 + r < p.L) ? t8[i] : 0.f;  // rows past L are clipped by the store
        }
        if (p.dbias != nullptr) {
          const float cs = warp_transpose_sum(v);
          atomicAdd(bias_s + part * 64 + half * 32 + lane, cs);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (leader) {
        if (kDKDV) {
          tma_store_3d(&tm.dqkv, stat_s, p.D + h * 64, s * 128, b);                 // dK_s
          tma_store_3d(&tm.dqkv, stat_s + AT_TILE, 2 * p.D + h * 64, s * 128, b);   // dV_s
        } else {
          tma_store_3d(&tm.dqkv, stat_s, h * 64, s * 128, b);                       // dQ_s
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&stat_empty[slot]);
      }
    }
    if (cur_h >= 0) flush_bias();
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
bool attention_tc_enabled() {
  static const bool on = [] {
    const char* e = getenv("CLIPN_ATTN_TC");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int attention_tc_fwd(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int causal, float scale,
                     cudaStream_t stream) {
  int cc_major = 0, sms = 0, cc_minor = 0;
  clipn_device_info(&sms, &cc_major, &cc_minor);
  CLIPN_REQUIRE(cc_major == 10, "attention: the tcgen05 kernels require an sm_100 (B200) device");
  CLIPN_REQUIRE(lse != nullptr, "attention_fwd: lse buffer required");
  AttnFwdParams p;
  p.L = seq; p.B = batch; p.H = heads; p.D = heads * 64;
  p.G = seq <= 64 ? 2 : 1;
  p.RB = 128 / p.G;
  p.nq = p.G == 2 ? 1 : (seq + 127) / 128;
  p.nkv = p.nq;
  p.causal = causal ? 1 : 0;
  p.items = ((batch + p.G - 1) / p.G) * heads * p.nq;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  AttnMaps tm;
  const uint64_t d3 = static_cast<uint64_t>(3) * p.D, d1 = static_cast<uint64_t>(p.D);
  int rc = make_tmap_3d(&tm.qkv, qkv, 2, d3, seq, batch, d3 * 2, d3 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&tm.out, out, 2, d1, seq, batch, d1 * 2, d1 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    configured = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attention_tc_fwd_kernel<<<grid, kAtThreads, AT_SMEM, stream>>>(tm, p);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

int attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias,
                     int batch, int seq, int heads, int causal, float scale, cudaStream_t stream) {
  CLIPN_REQUIRE(out != nullptr, "attention_bwd: the forward output is required (D = rowsum(dO o O))");
  if (seq > 128) {
    AttnBwdLongParams q;
    q.L = seq; q.B = batch; q.H = heads; q.D = heads * 64;
    q.nt = (seq + 127) / 128;
    q.causal = causal ? 1 : 0;
    q.items = batch * heads * q.nt;
    q.scale = scale; q.scale_log2 = scale * 1.4426950408889634f;
    q.lse = lse;
    q.out = reinterpret_cast<const __nv_bfloat16*>(out);
    q.dbias = dbias;
    AttnBwdMaps tl;
    const uint64_t e3 = static_cast<uint64_t>(3) * q.D, e1 = static_cast<uint64_t>(q.D);
    int rl = make_tmap_3d(&tl.qkv, qkv, 2, e3, seq, batch, e3 * 2, e3 * 2 * seq, 64, 128, 1, 128);
    if (rl) return rl;
    rl = make_tmap_3d(&tl.dout, dout, 2, e1, seq, batch, e1 * 2, e1 * 2 * seq, 64, 128, 1, 128);
    if (rl) return rl;
    rl = make_tmap_3d(&tl.dqkv, dqkv, 2, e3, seq, batch, e3 * 2, e3 * 2 * seq, 64, 128, 1, 128);
    if (rl) return rl;
    static bool configured_long = false;
    if (!configured_long) {
      CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_bwd_long_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            AT_BWDL_SMEM));
      CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_bwd_long_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            AT_BWDL_SMEM));
      configured_long = true;
    }
    const int gl = q.items < num_sms() ? q.items : num_sms();
    attention_tc_bwd_long_kernel<true><<<gl, kAtThreads, AT_BWDL_SMEM, stream>>>(tl, q);   // dK, dV
    CLIPN_CHECK_CUDA(cudaGetLastError());
    attention_tc_bwd_long_kernel<false><<<gl, kAtThreads, AT_BWDL_SMEM, stream>>>(tl, q);  // dQ
    CLIPN_CHECK_CUDA(cudaGetLastError());
    return CLIPN_OK;
  }
  AttnBwdParams p;
  p.L = seq; p.B = batch; p.H = heads; p.D = heads * 64;
  p.G = seq <= 64 ? 2 : 1;
  p.RB = 128 / p.G;
  p.causal = causal ? 1 : 0;
  p.nb = (batch + p.G - 1) / p.G;
  p.items = p.nb * heads;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.out = reinterpret_cast<const __nv_bfloat16*>(out);
  p.dbias = dbias;
  AttnBwdMaps tm;
  const uint64_t d3 = static_cast<uint64_t>(3) * p.D, d1 = static_cast<uint64_t>(p.D);
  int rc = make_tmap_3d(&tm.qkv, qkv, 2, d3, seq, batch, d3 * 2, d3 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&tm.dout, dout, 2, d1, seq, batch, d1 * 2, d1 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&tm.out, out, 2, d1, seq, batch, d1 * 2, d1 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&tm.dqkv, dqkv, 2, d3, seq, batch, d3 * 2, d3 * 2 * seq, 64, p.RB, p.G, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    CLIPN_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BWD_SMEM));
    configured = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attention_tc_bwd_kernel<<<grid, kAtThreads, AT_BWD_SMEM, stream>>>(tm, p);
  CLIPN_CHECK_CUDA(cudaGetLastError());
  return CLIPN_OK;
}

}  // namespace clipn
