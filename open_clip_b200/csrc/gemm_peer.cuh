// Peer-streaming logits GEMM: the all-gather of gather_features (reference loss.py:29-54) fused into the similarity
// GEMM of ClipLoss.get_logits (loss.py:91-116) / SigLipLoss (loss.py:351-367, 406-489), both directions in ONE launch.
// Included by gemm.cu inside namespace clipn, after the epilogue helpers and gemm_pair.cuh.
//
//   S_d[m, c] = rows_d[m, :] . cols_d[c, :]      d = 0: local image rows x every rank's text rows
//                                                d = 1: local text rows  x every rank's image rows
//
// Why a dedicated kernel: with K = embed_dim <= 512 the whole column tile (BN = 256 rows x K: 128 KB per CTA of the
// pair) fits in shared memory, so the COLUMN operand — the one that lives in the other ranks' HBM — is the stationary
// operand: a CTA pair pulls a [BN x K] tile from the owning rank's symmetric buffer through that rank's TMA tensor
// map exactly once (NVLink / NVSwitch P2P) and streams the local row operand past it out of L2 (4 MB, resident).
// The tile is single-buffered but refilled PIECE BY PIECE (one 64-wide k-block at a time): on the last M tile of a
// column tile every k-block's MMAs release their piece, the next tile's piece is requested at once and has a whole
// unit of tensor work (4096 cycles) to arrive.  BN = 256 (not 128 with two buffers, the first version): the row
// operand then needs 32 B/cycle/SM from L2 instead of 64 — the 128-wide version was bound by the latency-bandwidth
// product of its 96 KB ring (ncu: tensor pipe 52 %, the MMA thread spinning on `full`; profiles/r02_ncu_peer_v1.txt).
// Every peer byte therefore crosses NVLink once per launch (x ~1.1 for tiles shared by two neighbouring work
// ranges) instead of once per 256-row M tile as in the M-major walk of gemm_tc2_kernel (16x at local batch 4096).
// The resident tile is also written to a LOCAL [N, K] copy by TMA store as a by-product (the materialised
// all-gather), which is what the backward's d-logits / d-feature GEMMs read — the backward touches no peer memory.
//
// Work decomposition: units (direction, column tile, 256-row M tile), M fastest; the unit list is cut into equal
// contiguous ranges, one per cluster (stream-K style, no tail wave).  Column tiles are visited starting at this
// rank's own block so the first tiles come from local HBM while the first peer loads are in flight.
//
// Warp roles (384 threads, cluster of 2 CTAs, cta_group::2, M = 256):
//   warp 0 : TMA producer of the row operand  (A ring, 16 KB stages, both CTAs load their own 128 rows)
//   warp 1 : MMA issuer (leader CTA)           tcgen05.mma.cta_group::2 kind::f16, D = 256 x BN fp32 in TMEM, x2
//   warp 2 : TMEM allocator, then relay: as each piece of a column tile lands it TMA-stores it to the gathered
//            copy and reports it to the leader's `piece_full` barrier
//   warp 3 : TMA producer of the column-tile pieces (peer reads)
//   warps 4-11 : epilogue, all 8 warps on every unit (TMEM lane quarter x column half): online LSE (CLIP) or
//                softplus/sigmoid (SigLIP).  (Splitting the warps between the two accumulators doubled the latency of one
//                unit's epilogue past the 4096-cycle mainloop and stalled the MMA on tmem_empty: 59 % tensor pipe.)
#pragma once

struct alignas(64) PeerTmaps {
  CUtensorMap a[2];             // row operands [m, K]
  CUtensorMap b[2][kMaxBMaps];  // column operands, one map per rank: [rows_per_map, K] (peer-mapped)
  CUtensorMap g[2];             // local gathered copies [n, K] (TMA store targets)
  CUtensorMap c[2];             // SIGLIP: d(logits) outputs [m, n] bf16
};

struct PeerParams {
  int m, n, kblocks;          // local rows, total columns (world * rows_per_map), K / 64
  int rank, rows_per_map;     // this rank; columns owned by each rank
  int tiles_m, tiles_n, dirs; // 256-row M tiles, BN-column tiles, 1 or 2 directions
  int label_offset, gather, negative_only;
  float alpha, logit_bias, gscale;
  const float* alpha_dev;
  const float* logit_bias_dev;
  float* part_max[2];
  float* part_sum[2];         // LSE: [2 * tiles_n, m] partial sums; SIGLIP: loss accumulator (or null)
  float* pos[2];
  void* c[2];
  float* scalar_acc[2];
};

template <int BN, int EPI>
struct PeerCfg {
  static constexpr int BNH = BN / 2;                   // column-tile rows staged by each CTA of the pair
  static constexpr int KB_MAX = 8 * 256 / BN;          // 8 k-blocks (K <= 512) at BN = 256, 16 (K <= 1024) at BN = 128
  static constexpr int B_PIECE = BNH * BK * 2;         // one k-block of the resident tile: 16 KB / 8 KB
  static constexpr int B_BYTES = KB_MAX * B_PIECE;     // resident column tile (this CTA's half): 128 KB
  static constexpr int EPI_BYTES = (EPI == CLIPN_EPI_SIGLIP) ? kEpiWarps * EPI_BUF_BYTES : 0;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 640;
  static constexpr int STAGES_FIT = (BUDGET - B_BYTES - EPI_BYTES) / A_STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int NUM_BARS = 2 * STAGES + 3 * KB_MAX + 5;
  static constexpr int BAR_BYTES = NUM_BARS * 8 + 16;
  static constexpr int SMEM_BYTES = B_BYTES + STAGES * A_STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static_assert(STAGES >= 3, "row-operand ring too shallow");
  static_assert(BAR_BYTES <= 640, "barrier area");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  static_assert(B_PIECE % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
  static_assert(TMEM_COLS <= 512, "TMEM columns");
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_peer_kernel(const __grid_constant__ PeerTmaps tm, const __grid_constant__ PeerParams p) {
  using Cfg = PeerCfg<BN, EPI>;
  constexpr int BNH = Cfg::BNH;
  constexpr bool kStore = EPI == CLIPN_EPI_SIGLIP;
  // p stays in the constant bank (its per-direction arrays are indexed dynamically: a local copy would live on the stack)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres = smem;                      // [KB_MAX] pieces of BNH rows x 64 k (SW128, K-major)
  uint8_t* aring = smem + Cfg::B_BYTES;      // [STAGES] 128 rows x 64 k
  uint8_t* epi_smem = aring + Cfg::STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* b_land = empty_bar + Cfg::STAGES;    // own CTA: this piece's bytes have landed
  uint64_t* piece_full = b_land + Cfg::KB_MAX;   // leader: both CTAs' halves of this piece are in place
  uint64_t* piece_empty = piece_full + Cfg::KB_MAX;  // both CTAs: the last MMAs reading this piece have retired
  uint64_t* st_done = piece_empty + Cfg::KB_MAX; // own CTA: the gathered-copy stores of the tile have left smem
  uint64_t* tmem_full = st_done + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    for (int d = 0; d < p.dirs; ++d) {
      tma_prefetch_desc(&tm.a[d]);
      if (p.gather) tma_prefetch_desc(&tm.g[d]);
      if (kStore && p.c[d] != nullptr) tma_prefetch_desc(&tm.c[d]);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader: one arrive.expect_tx covering both CTAs' bytes (see gemm_pair.cuh)
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < Cfg::KB_MAX; ++i) {
      mbar_init(&b_land[i], 1);       // arrive.expect_tx of this CTA's loader
      mbar_init(&piece_full[i], 2);   // leader: one arrive per CTA (relay threads)
      mbar_init(&piece_empty[i], 1);  // multicast tcgen05.commit
    }
    mbar_init(st_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);          // multicast tcgen05.commit
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);  // leader: the 8 epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // ---- this cluster's contiguous unit range; unit = (direction, column tile, M tile), M fastest
  const int64_t total_units = static_cast<int64_t>(p.dirs) * p.tiles_n * p.tiles_m;
  const int u0 = static_cast<int>((total_units * cluster_id) / num_clusters);
  const int u1 = static_cast<int>((total_units * (cluster_id + 1)) / num_clusters);
  const int t_first = u0 < u1 ? u0 / p.tiles_m : 0;        // tile id = d * tiles_n + nt
  const int t_last = u0 < u1 ? (u1 - 1) / p.tiles_m : -1;
  const int rot = p.rank * (p.rows_per_map / BN);  // column tiles are visited starting at this rank's own block
  auto tile_col0 = [&](int nt) {
    int t = nt + rot;
    if (t >= p.tiles_n) t -= p.tiles_n;
    return t * BN;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===================== row-operand producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u0; u < u1; ++u) {
        const int mt = u % p.tiles_m;
        const int d = (u / p.tiles_m) / p.tiles_n;
        const int m0 = mt * (2 * BM) + static_cast<int>(rank) * BM;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * A_STAGE_BYTES);
          tma_load_2d_2sm(aring + stage * A_STAGE_BYTES, &tm.a[d], bar, kb * BK, m0);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ===================== column-tile loader: peer reads, one k-block piece at a time (both CTAs) =====================
      for (int t = t_first; t <= t_last; ++t) {
        const int s = t - t_first;
        const int d = t / p.tiles_n;
        const int n0 = tile_col0(t % p.tiles_n);
        const int map = n0 / p.rows_per_map;
        const int r0 = n0 - map * p.rows_per_map + static_cast<int>(rank) * BNH;
        if (s > 0 && p.gather) mbar_wait(st_done, (s - 1) & 1);  // the previous tile's gathered-copy stores have read smem
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (s > 0) mbar_wait(&piece_empty[kb], (s - 1) & 1);   // the previous tile's last MMAs on this piece retired
          mbar_expect_tx(&b_land[kb], Cfg::B_PIECE);
          tma_load_2d(bres + kb * Cfg::B_PIECE, &tm.b[d][map], &b_land[kb], kb * BK, r0);
        }
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      // ===================== relay: landed piece -> gathered copy (TMA store) -> leader's piece_full =====================
      for (int t = t_first; t <= t_last; ++t) {
        const int s = t - t_first;
        const int d = t / p.tiles_n;
        const int n0 = tile_col0(t % p.tiles_n) + static_cast<int>(rank) * BNH;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&b_land[kb], s & 1);
          if (p.gather) {
            tma_store_2d(&tm.g[d], bres + kb * Cfg::B_PIECE, kb * BK, n0);
            tma_store_commit();
          }
          if (rank == 0) mbar_arrive(&piece_full[kb]);
          else mbar_arrive_cluster(mapa_u32(smem_u32(&piece_full[kb]), 0));
        }
        if (p.gather) {
          tma_store_wait_read<0>();
          mbar_arrive(st_done);
        }
      }
      if (p.gather) tma_store_wait_all<0>();
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ===================== MMA issuer (leader CTA) =====================
      const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t sb0 = smem_u32(bres);
      for (int u = u0; u < u1; ++u) {
        const int it = u - u0;
        const int tile = u / p.tiles_m;
        const bool first_of_tile = u == u0 || (u - 1) / p.tiles_m != tile;
        const bool last_of_tile = u + 1 == u1 || (u + 1) / p.tiles_m != tile;
        const uint32_t tile_parity = static_cast<uint32_t>(tile - t_first) & 1;
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (first_of_tile) mbar_wait(&piece_full[kb], tile_parity);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(aring + stage * A_STAGE_BYTES);
          const uint32_t sb = sb0 + kb * Cfg::B_PIECE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_2sm(d_tmem, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (last_of_tile) umma_commit_2sm(&piece_empty[kb]);  // this piece may be refilled with the next column tile
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps per unit, warp = (TMEM lane quarter q, column half h) =====================
    const int e = warp - 4;
    const int q = e & 3;   // TMEM lane quarter == warp % 4
    const int h = e >> 2;  // column half of the tile
    uint8_t* buf0 = epi_smem + e * EPI_BUF_BYTES;
    GemmParams gp;
    gp.m = p.m; gp.n = p.n; gp.gscale = p.gscale;
    gp.alpha = p.alpha_dev != nullptr ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    gp.logit_bias = p.logit_bias_dev != nullptr ? p.logit_bias + __ldg(p.logit_bias_dev) : p.logit_bias;
    gp.label_offset = p.label_offset; gp.negative_only = p.negative_only; gp.col_w = 0.f;
    gp.row_lse = nullptr; gp.col_lse = nullptr; gp.col_sum = nullptr; gp.bias = nullptr;
    const int nunits = u1 - u0;
    for (int it = 0; it < nunits; ++it) {
      const int u = u0 + it;
      const int acc = it & 1;
      const int mt = u % p.tiles_m;
      const int tile = u / p.tiles_m;
      const int d = tile / p.tiles_n;
      const int n0 = tile_col0(tile % p.tiles_n);
      const int row0 = mt * (2 * BM) + static_cast<int>(rank) * BM + q * 32;
      const int row = row0 + lane;
      gp.part_max = p.part_max[d]; gp.part_sum = p.part_sum[d]; gp.pos = p.pos[d];
      gp.c = p.c[d]; gp.scalar_acc = p.scalar_acc[d];
      const bool store_c = kStore && p.c[d] != nullptr;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      EpiState st;
      epi_begin(st);
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int cl = h * (BN / 2) + c * 32;
        if (store_c && (c & 1) == 0) {
          if (lane == 0) tma_store_wait_read<0>();  // the previous store has finished reading the staging tile
          __syncwarp();
        }
        float v[32], aux[32], o1[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + cl, v);
        if (n0 + cl < p.n) {
          epi_compute<EPI>(gp, row, n0 + cl, v, aux, o1, st);
          if (store_c) stage_write32(buf0, lane, c & 1, v);
        }
        if (store_c && (c & 1) == 1 && n0 + cl - 32 < p.n) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm.c[d], buf0, n0 + cl - 32, row0);  // boxes past N / M are clipped by the TMA unit
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      }
      epi_finish<EPI>(gp, row, (n0 / BN) * 2 + h, st);  // LSE partials: one slab per column half of a tile
    }
    if (kStore && lane == 0) tma_store_wait_all<0>();
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}
