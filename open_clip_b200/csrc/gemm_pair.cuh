// CTA-pair (cta_group::2) variant of the GEMM kernel.  Included by gemm.cu inside namespace clipn, after the
// epilogue helpers.  Same warp roles, same epilogues, same operand layouts as gemm_tc_kernel; differences:
//   * launched as clusters of 2 CTAs (one TPC); the pair owns a 256 x BN output tile, CTA r the rows r*128..
//   * every CTA TMA-loads its own 128 rows of A and its BN/2 rows of B (stage = 16 KB + BN*64 B  -> deeper ring,
//     1/3 fewer operand bytes per MAC through L2 -> SMEM)
//   * CTA 0 issues tcgen05.mma.cta_group::2 (M = 256); completion is multicast to both CTAs' mbarriers
//   * the leader's `full` barrier collects the transaction bytes of BOTH CTAs' loads (+ one remote arrive)
//   * the leader's `tmem_empty` barrier collects the epilogue arrivals of both CTAs
#pragma once

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ TmapSet tm, const GemmParams p_in) {
  using Cfg = TileCfg<BN, EPI, 2>;
  using Tr = EpiTraits<EPI>;
  constexpr int BNH = BN / 2;  // B rows staged by each CTA
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
  uint64_t* aux_bar = tmem_empty + 2;  // two per epilogue warp (double-buffered aux tiles)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aux_bar + 2 * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.a);
    for (int i = 0; i < p.b_maps; ++i) tma_prefetch_desc(&tm.b[i]);
    if (Tr::kOutTma || Tr::kRedF32) tma_prefetch_desc(&tm.c);
    if (Tr::kNumOut == 2 && p.c2 != nullptr) tma_prefetch_desc(&tm.c2);
    if (Tr::kAux) tma_prefetch_desc(&tm.aux);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      // leader: one arrive.expect_tx covering BOTH CTAs' bytes.  The peer producer does not arrive: a remote
      // mbarrier.arrive.release.cluster compiles to MEMBAR.ALL.GPU and drains the producer's in-flight TMA loads
      // every stage (measured: tensor pipe 33 % active).  It cannot run ahead of the leader anyway: its empty
      // barrier only fires after the leader's MMAs consumed the stage.
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);               // multicast tcgen05.commit
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);  // leader: epilogue warps of both CTAs
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_work = p.tiles_m * p.tiles_n * p.splits;  // tiles_m counts 256-row pair tiles here

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int split = w % p.splits;
        const int tile = w / p.splits;
        const int m0 = (tile / p.tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = (tile % p.tiles_n) * BN + static_cast<int>(rank) * BNH;
        const int kb0 = static_cast<int>((static_cast<int64_t>(split) * p.kblocks) / p.splits);
        const int kb1 = static_cast<int>((static_cast<int64_t>(split + 1) * p.kblocks) / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);  // the LEADER's barrier
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tm.a, bar, kb * BK, m0);
          } else {
            tma_load_2d_2sm(sa, &tm.a, bar, m0, kb * BK);
            tma_load_2d_2sm(sa + 8192, &tm.a, bar, m0 + 64, kb * BK);
          }
          if (!p.b_mn) {
            const int map = n0 / p.b_rows_per_map;
            tma_load_2d_2sm(sb, &tm.b[map], bar, kb * BK, n0 - map * p.b_rows_per_map);
          } else {
            const int map = (kb * BK) / p.b_rows_per_map;
            const int krow = kb * BK - map * p.b_rows_per_map;
#pragma unroll
            for (int i = 0; i < BNH / 64; ++i) tma_load_2d_2sm(sb + i * 8192, &tm.b[map], bar, n0 + 64 * i, krow);
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
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
            umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[acc]);  // accumulators ready for both CTAs' epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int e = warp - 4;
    const int q = e & 3;   // TMEM lane quarter == warp % 4
    const int h = e >> 2;  // column half of the tile
    if constexpr (Cfg::kPipedEpi) {
      // ---- software-pipelined epilogue for the two-output GELU variants: 32-column pieces, double-buffered
      // staging; piece i's TMA stores and piece i+1's aux TMA load are in flight while piece i is computed ----
      constexpr int NP = BN / 64;  // pieces per warp per tile
      constexpr int OUT1 = 4096;                  // second output's tiles (two-output epilogues)
      constexpr int AUX = Tr::kNumOut * 4096;     // aux operand's tiles
      uint8_t* wb = epi_smem + e * (Cfg::PIPE_TILES * 2048);
      uint64_t* abar = aux_bar + 2 * e;
      auto coords = [&](int w, int i, int& col, int& row0) {
        const int tile = w / p.splits;
        col = (tile % p.tiles_n) * BN + h * (BN / 2) + i * 32;
        row0 = (tile / p.tiles_n) * (2 * BM) + static_cast<int>(rank) * BM + q * 32;
      };
      uint32_t pc = 0;  // running piece counter: buffer = pc & 1, aux phase = (pc >> 1) & 1
      if (Tr::kAux && lane == 0 && cluster_id < total_work) {
        int col, row0;
        coords(cluster_id, 0, col, row0);
        mbar_expect_tx(&abar[0], 2048);
        tma_load_2d(wb + AUX, &tm.aux, &abar[0], col, row0);
      }
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        EpiState st;
        epi_begin(st);
#pragma unroll 1
        for (int i = 0; i < NP; ++i, ++pc) {
          const int b = pc & 1;
          int col, row0;
          coords(w, i, col, row0);
          if (lane == 0) tma_store_wait_read<1>();  // the stores that used buffers `b` two pieces ago have read them
          __syncwarp();
          if (Tr::kAux) {
            if (lane == 0) {
              int nw = w, ni = i + 1;
              if (ni == NP) {
                ni = 0;
                nw = w + num_clusters;
              }
              if (nw < total_work) {
                int ncol, nrow0;
                coords(nw, ni, ncol, nrow0);
                mbar_expect_tx(&abar[b ^ 1], 2048);
                tma_load_2d(wb + AUX + (b ^ 1) * 2048, &tm.aux, &abar[b ^ 1], ncol, nrow0);
              }
            }
            mbar_wait(&abar[b], (pc >> 1) & 1);
          }
          float v[32], aux[32], o1[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + h * (BN / 2) + i * 32, v);
          if (Tr::kAux) stage64_read32(wb + AUX + b * 2048, lane, aux);
          if (col < p.n) epi_compute<EPI>(p, row0 + lane, col, v, aux, o1, st);
          stage64_write32(wb + b * 2048, lane, v);
          if (Tr::kNumOut == 2 && p.c2 != nullptr) stage64_write32(wb + OUT1 + b * 2048, lane, o1);
          if (Tr::kColSum && p.col_sum != nullptr && col < p.n) epi_col_sum(p, row0 + lane, col, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm.c, wb + b * 2048, col, row0);  // boxes past N / M are clipped by the TMA unit
            if (Tr::kNumOut == 2 && p.c2 != nullptr) tma_store_2d(&tm.c2, wb + OUT1 + b * 2048, col, row0);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[acc]), 0));  // publishes no memory: no MEMBAR
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (lane == 0) tma_store_wait_all<0>();
    } else {
    uint8_t* buf0 = epi_smem + e * (Tr::kBufs * EPI_BUF_BYTES);
    uint8_t* buf1 = buf0 + EPI_BUF_BYTES;
    const bool store_c = Tr::kOutTma && (EPI != CLIPN_EPI_SIGLIP || p.c != nullptr);
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int tile = w / p.splits;
      const int tn = tile % p.tiles_n;
      const int m0 = (tile / p.tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = tn * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      EpiState st;
      epi_begin(st);
#pragma unroll 1
      for (int jc = 0; jc < BN / 128; ++jc) {
        const int cl0 = h * (BN / 2) + jc * 64;
        const bool chunk_live = n0 + cl0 < p.n;
        if (Tr::kOutTma && chunk_live) {
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
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[acc]), 0));  // publishes no memory: no MEMBAR
      }
      // LSE partial slabs are indexed by 128-column halves of the N tile, exactly like the single-CTA kernel
      epi_finish<EPI>(p, row, tn * 2 + h, st);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if ((Tr::kOutTma || Tr::kRedF32) && lane == 0) tma_store_wait_all<0>();
    }  // !kPipedEpi
  }

  // neither CTA may exit (or free TMEM) while its peer can still read its smem / signal its barriers
  __syncwarp();  // re-converge the single-lane producer / MMA warps before the .aligned cluster barrier
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}
