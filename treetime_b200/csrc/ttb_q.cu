// ttb_q.cu -- instantiates the kernels for ONE alphabet size (-DTTB_Q=<q>) and exports
// their launchers as a TtbQOps table.
#include "ttb_qops.h"
#if TTB_Q > 8
#include "ttb_mma.cuh"
#endif
#include <algorithm>
#include <cstdlib>

#ifndef TTB_Q
#error "compile with -DTTB_Q=<n_states>"
#endif

namespace {
constexpr int Q = TTB_Q;

constexpr bool HAS_SYM = (Q <= TTB_SS_REG_MAXQ);   // symmetric register-resident site-specific kernels
size_t level_smem(int rows, const TtbDev& d, bool ss, bool sym = false) {
  if (ss && sym) return Pipe<Q, TTB_SS_SYM_STAGES>::smem_bytes(rows, stage_pq<Q, true>(d.pq), stage_tu<Q, true>(d.tu_stride), true);
  return ss ? Pipe<Q>::smem_bytes(rows, stage_pq<Q, true>(d.pq), stage_tu<Q, true>(d.tu_stride), true)
            : Pipe<Q>::smem_bytes(rows, d.pq, d.tu_stride, false);
}
size_t post_smem(const TtbDev& d, bool ss = false, bool sym = false) { return level_smem(Pipe<Q>::CB * Q, d, ss, sym); }
size_t pre_smem(const TtbDev& d, bool ss = false, bool sym = false) { return level_smem(Q + Pipe<Q>::CB * Q, d, ss, sym); }

// Launch with the programmatic-stream-serialization attribute (see pdl_wait() in ttb_kernels.cuh).
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr bool HAS_SS = (Q <= 8);   // site-specific models: nucleotide-sized alphabets only
constexpr bool HAS_F32 = (Q <= 8);  // float message storage (ttb_set_message_storage): the HBM-bound alphabet sizes

template <bool SS>
int prepare_t(const TtbDev& d) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(post_level_kernel<Q, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, SS))) != cudaSuccess) return (int)e;
  if (!SS && (e = cudaFuncSetAttribute(post_level_kernel<Q, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, false))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
  // masked (ARG mode) variants
  if ((e = cudaFuncSetAttribute(post_level_kernel<Q, SS, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, SS))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, SS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, SS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
  if constexpr (SS && HAS_SYM) {
    if ((e = cudaFuncSetAttribute(post_level_kernel<Q, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, true, true))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, true, true))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, true, true))) != cudaSuccess) return (int)e;
  }
  if constexpr (!SS) {       // merged-level launches (opt-in): single model, double storage, no masks
    if ((e = cudaFuncSetAttribute(post_level_kernel<Q, false, false, false, false, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, false))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(post_level_kernel<Q, false, true, false, false, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, false))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, false, false, false, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, false))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, false, false, false, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, false))) != cudaSuccess) return (int)e;
  }
  if constexpr (HAS_F32) {   // float message storage: same stages (the rows just hold floats)
    if ((e = cudaFuncSetAttribute(post_level_kernel<Q, SS, false, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, SS))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, SS, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, SS, false, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, SS))) != cudaSuccess) return (int)e;
    if constexpr (SS && HAS_SYM) {
      if ((e = cudaFuncSetAttribute(post_level_kernel<Q, true, false, true, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem(d, true, true))) != cudaSuccess) return (int)e;
      if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, false, true, true, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, true, true))) != cudaSuccess) return (int)e;
      if ((e = cudaFuncSetAttribute(pre_level_kernel<Q, true, true, true, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre_smem(d, true, true))) != cudaSuccess) return (int)e;
    }
  }
  return 0;
}

#if TTB_Q > 8
// Tensor-pipe level kernels (ttb_mma.cuh): NW pattern warps per block, chosen at run time (TTB_MMA_NW, measurement knob).
// Pattern warps per block (measurement knobs TTB_MMA_NW_POST / TTB_MMA_NW_PRE, TTB_MMA_NW for both).  Defaults from the
// cfg4 sweep (profiles/R2p_mma_sweep.txt): postorder 8 warps x 2 blocks per SM, preorder 16 warps x 1 block.
int mma_nw(bool pre) {   // read at enqueue time (graph capture), so a test can select a variant per engine
  const char* e = getenv(pre ? "TTB_MMA_NW_PRE" : "TTB_MMA_NW_POST");
  if (!e) e = getenv("TTB_MMA_NW");
  const int dflt = pre ? 16 : 8, v = e ? atoi(e) : dflt;
  return (v == 4 || v == 8 || v == 16) ? v : dflt;
}
template <int NW>
size_t mma_post_smem(const TtbDev& d) { return MmaCfg<Q, NW>::PipeT::smem_bytes(Q, MmaQ<Q>::PFQ, d.tu_stride); }
template <int NW>
size_t mma_pre_smem(const TtbDev& d, bool tips) { return MmaCfg<Q, NW>::PipeT::smem_bytes(2 * Q, 2 * MmaQ<Q>::PFQ, tips ? d.tu_stride : 0); }
template <int NW>
int prepare_mma(const TtbDev& d) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(post_level_mma_kernel<Q, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_post_smem<NW>(d))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_mma_kernel<Q, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_pre_smem<NW>(d, false))) != cudaSuccess) return (int)e;
  if ((e = cudaFuncSetAttribute(pre_level_mma_kernel<Q, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_pre_smem<NW>(d, true))) != cudaSuccess) return (int)e;
  // two resident blocks per SM need the full shared-memory carve-out
  cudaFuncSetAttribute(post_level_mma_kernel<Q, NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(pre_level_mma_kernel<Q, NW, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(pre_level_mma_kernel<Q, NW, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return 0;
}
template <int NW>
void launch_post_mma(const TtbPassPlan& pl, const TtbLevelLaunch& L, int fbase, cudaStream_t s) {
  launch_pdl(post_level_mma_kernel<Q, NW>, (unsigned)((long long)L.n_groups * pl.tiles), MmaCfg<Q, NW>::THREADS, mma_post_smem<NW>(pl.d), s, pl.d,
             pl.d_post_chunks, pl.d_post_group_ptr + L.group_off, pl.tiles, fbase);
}
template <int NW>
void launch_pre_mma(const TtbPassPlan& pl, const TtbLevelLaunch& L, cudaStream_t s) {
  const unsigned grid = (unsigned)((long long)L.n_groups * pl.tiles);
  if (pl.tips)
    launch_pdl(pre_level_mma_kernel<Q, NW, true>, grid, MmaCfg<Q, NW>::THREADS, mma_pre_smem<NW>(pl.d, true), s, pl.d, pl.d_pre_chunks,
               pl.d_pre_group_ptr + L.group_off, pl.tiles, pl.count_diff);
  else
    launch_pdl(pre_level_mma_kernel<Q, NW, false>, grid, MmaCfg<Q, NW>::THREADS, mma_pre_smem<NW>(pl.d, false), s, pl.d, pl.d_pre_chunks,
               pl.d_pre_group_ptr + L.group_off, pl.tiles, pl.count_diff);
}
#endif

int prepare_q(const TtbDev& d) {
  if (int e = prepare_t<false>(d)) return e;
#if TTB_Q > 8
  if (int e = prepare_mma<4>(d)) return e;
  if (int e = prepare_mma<8>(d)) return e;
  if (int e = prepare_mma<16>(d)) return e;
#endif
  if constexpr (HAS_SS) {
    if (int e = prepare_t<true>(d)) return e;
  }
  return 0;
}

template <bool SS, bool SYM = false, bool MASK = false, typename ST = double>
int enqueue_pass_t(const TtbPassPlan& pl, cudaStream_t s, cudaEvent_t* ev, int* pk) {
  const TtbDev& d = pl.d;
  const int tiles = pl.tiles;
  int nk = 0;
  if (ev) cudaEventRecord(ev[0], s);
  bool use_mma = false;   // large alphabets, single model, no masks: level kernels on the fp64 tensor pipe
#if TTB_Q > 8
  if constexpr (!SS && !MASK && sizeof(ST) == 8) use_mma = d.Pf != nullptr;
#endif
  if (!SS) {
#if TTB_Q > 8
    if (use_mma) {   // exp(Qt) in both layouts from one kernel
      expqt_frag_kernel<Q><<<(d.n_nodes + TTB_EXPQT_WARPS * TTB_EXPQT_NODES - 1) / (TTB_EXPQT_WARPS * TTB_EXPQT_NODES), TTB_EXPQT_WARPS * 32, 0, s>>>(d, d.Pf);
    } else
#endif
    {
      const int nthr = d.n_nodes * Q;
      expqt_kernel<Q><<<(nthr + 127) / 128, 128, 0, s>>>(d);
    }
    const long long ntab = (long long)d.n_tips * d.n_codes * Q;
    tip_table_kernel<Q><<<(unsigned)((ntab + 255) / 256), 256, 0, s>>>(d, pl.d_tip_nodes);
    nk += 2;
  }
  if (ev) { cudaEventRecord(ev[1], s); pk[0] = nk; }
  const size_t psm = post_smem(d, SS, SYM);
  int l0 = 0;
  if (!SS) {
    // level 1 (all children are tips) is a pure write stream: dedicated kernel, no pipeline
    const TtbLevelLaunch& L = pl.post_levels[0];
    const double* pairs = nullptr;
    if constexpr (Q <= 8 && !MASK) {
      if (pl.d_leaf_pairs && pl.n_leaf_chunks > 0) {      // cherry tables: one block per level-1 chunk
        leaf_pair_table_kernel<Q><<<pl.n_leaf_chunks, 128, 0, s>>>(d, pl.d_post_chunks, pl.d_leaf_pairs);
        pairs = pl.d_leaf_pairs;
        ++nk;
      }
    }
#if TTB_Q > 8
    if (use_mma)
      launch_pdl(post_leaf_mma_kernel<Q>, (unsigned)((long long)L.n_groups * tiles), 512, 0, s, d, pl.d_post_chunks,
                 pl.d_post_group_ptr + L.group_off, tiles, 0);
    else
#endif
    launch_pdl(post_leaf_level_kernel<Q, false, ST>, (unsigned)((long long)L.n_groups * tiles), TTB_BLOCK, 0, s, d, pl.d_post_chunks,
               pl.d_post_group_ptr + L.group_off, tiles, 0, pairs);
    ++nk;
    l0 = 1;
  }
  int fbase = l0 ? pl.post_levels[0].n_groups : 0;
  for (int l = l0; l < pl.n_post_levels; ++l) {
    const TtbLevelLaunch& L = pl.post_levels[l];
    constexpr bool CAN_DEP = !SS && !MASK && sizeof(ST) == 8;     // build_groups merges levels only for these
#if TTB_Q > 8
    if (use_mma && !L.dep) {
      switch (mma_nw(false)) {
        case 4: launch_post_mma<4>(pl, L, fbase, s); break;
        case 16: launch_post_mma<16>(pl, L, fbase, s); break;
        default: launch_post_mma<8>(pl, L, fbase, s); break;
      }
      fbase += L.n_groups;
      ++nk;
      continue;
    }
#endif
    if constexpr (CAN_DEP) {
      if (L.dep) {
        launch_pdl(post_level_kernel<Q, false, false, false, false, double, true>, (unsigned)((long long)L.n_groups * tiles), TTB_LEVEL_THREADS, psm, s, d,
                   pl.d_post_chunks, pl.d_post_group_ptr + L.group_off, tiles, fbase, pl.d_post_dep);
        fbase += L.n_groups;
        ++nk;
        continue;
      }
    }
    launch_pdl(post_level_kernel<Q, SS, false, SYM, MASK, ST>, (unsigned)((long long)L.n_groups * tiles), TTB_LEVEL_THREADS, psm, s, d, pl.d_post_chunks,
               pl.d_post_group_ptr + L.group_off, tiles, fbase, nullptr);
    fbase += L.n_groups;
    ++nk;
  }
  if (ev) { cudaEventRecord(ev[2], s); pk[1] = nk - pk[0]; }
  fsum_kernel<<<dim3(tiles, TTB_FLANES), TTB_BLOCK, 0, s>>>(d);
  root_kernel<Q, SS, ST><<<tiles, TTB_BLOCK, 0, s>>>(d, pl.lh_only ? 1 : 0);
  nk += 2;
  if (!pl.lh_only) {
    zero_slots_kernel<<<4, 256, 0, s>>>(d);
    ++nk;
  }
  if (ev) { cudaEventRecord(ev[3], s); pk[2] = nk - pk[0] - pk[1]; }
  if (!pl.lh_only) {
    const size_t rsm = pre_smem(d, SS, SYM);
    for (int l = 0; l < pl.n_pre_levels; ++l) {
      const TtbLevelLaunch& L = pl.pre_levels[l];
      const unsigned grid = (unsigned)((long long)L.n_groups * tiles);
      constexpr bool CAN_DEP = !SS && !MASK && sizeof(ST) == 8;
#if TTB_Q > 8
      if (use_mma && !L.dep) {
        switch (mma_nw(true)) {
          case 4: launch_pre_mma<4>(pl, L, s); break;
          case 16: launch_pre_mma<16>(pl, L, s); break;
          default: launch_pre_mma<8>(pl, L, s); break;
        }
        ++nk;
        continue;
      }
#endif
      if constexpr (CAN_DEP) {
        if (L.dep) {
          if (pl.tips)
            launch_pdl(pre_level_kernel<Q, true, false, false, false, double, true>, grid, TTB_LEVEL_THREADS, rsm, s, d, pl.d_pre_chunks,
                       pl.d_pre_group_ptr + L.group_off, tiles, pl.count_diff, pl.d_pre_dep);
          else
            launch_pdl(pre_level_kernel<Q, false, false, false, false, double, true>, grid, TTB_LEVEL_THREADS, rsm, s, d, pl.d_pre_chunks,
                       pl.d_pre_group_ptr + L.group_off, tiles, pl.count_diff, pl.d_pre_dep);
          ++nk;
          continue;
        }
      }
      if (pl.tips)
        launch_pdl(pre_level_kernel<Q, true, SS, SYM, MASK, ST>, grid, TTB_LEVEL_THREADS, rsm, s, d, pl.d_pre_chunks, pl.d_pre_group_ptr + L.group_off, tiles,
                   pl.count_diff, nullptr);
      else
        launch_pdl(pre_level_kernel<Q, false, SS, SYM, MASK, ST>, grid, TTB_LEVEL_THREADS, rsm, s, d, pl.d_pre_chunks, pl.d_pre_group_ptr + L.group_off, tiles,
                   pl.count_diff, nullptr);
      ++nk;
    }
  }
  if (ev) { cudaEventRecord(ev[4], s); pk[3] = nk - pk[0] - pk[1] - pk[2]; }
  finish_kernel<<<1, 256, 0, s>>>(d, tiles);
  ++nk;
  if (ev) { cudaEventRecord(ev[5], s); pk[2] += 1; }
  return nk;
}

int enqueue_pass_q(const TtbPassPlan& pl, cudaStream_t s, cudaEvent_t* ev, int* pk) {
  if (pl.d.mask_id) {   // per-branch masks: the general kernels with the mask test compiled in
    if constexpr (HAS_SS) {
      if (pl.d.site_specific) return enqueue_pass_t<true, false, true>(pl, s, ev, pk);
    }
    return enqueue_pass_t<false, false, true>(pl, s, ev, pk);
  }
  if constexpr (HAS_F32) {
    if (pl.d.f32) {      // float message storage (not combined with masks: ttb_marginal refuses that)
      if constexpr (HAS_SYM) {
        if (pl.d.site_specific && pl.d.ss_sym) return enqueue_pass_t<true, true, false, float>(pl, s, ev, pk);
      }
      if (pl.d.site_specific) return enqueue_pass_t<true, false, false, float>(pl, s, ev, pk);
      return enqueue_pass_t<false, false, false, float>(pl, s, ev, pk);
    }
  }
  if constexpr (HAS_SYM) {
    if (pl.d.site_specific && pl.d.ss_sym) return enqueue_pass_t<true, true>(pl, s, ev, pk);
  }
  if constexpr (HAS_SS) {
    if (pl.d.site_specific) return enqueue_pass_t<true>(pl, s, ev, pk);
  }
  return enqueue_pass_t<false>(pl, s, ev, pk);
}

int enqueue_joint_q(const TtbPassPlan& pl, cudaStream_t s, int trace) {
  const TtbDev& d = pl.d;
  if (d.site_specific) return 0;
  const int tiles = pl.tiles;
  int nk = 0;
  const int nthr = d.n_nodes * Q;
  expqt_kernel<Q><<<(nthr + 127) / 128, 128, 0, s>>>(d);
  const long long nt = std::max((long long)d.n_nodes * Q * Q, (long long)d.n_tips * d.n_codes * Q);
  joint_tables_kernel<Q><<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(d, pl.d_tip_nodes);
  nk += 2;
  const size_t psm = post_smem(d, false);
  int l0 = 0;
  if (pl.n_post_leaf_nodes) {
    const TtbLevelLaunch& L = pl.post_levels[0];
    post_leaf_level_kernel<Q, true><<<(unsigned)((long long)L.n_groups * tiles), TTB_BLOCK, 0, s>>>(d, pl.d_post_chunks,
                                                                                                pl.d_post_group_ptr + L.group_off, tiles, 0, nullptr);
    ++nk;
    l0 = 1;
  }
  for (int l = l0; l < pl.n_post_levels; ++l) {
    const TtbLevelLaunch& L = pl.post_levels[l];
    if (L.dep)
      post_level_kernel<Q, false, true, false, false, double, true><<<(unsigned)((long long)L.n_groups * tiles), TTB_LEVEL_THREADS, psm, s>>>(
          d, pl.d_post_chunks, pl.d_post_group_ptr + L.group_off, tiles, 0, pl.d_post_dep);
    else
      post_level_kernel<Q, false, true><<<(unsigned)((long long)L.n_groups * tiles), TTB_LEVEL_THREADS, psm, s>>>(d, pl.d_post_chunks,
                                                                                                        pl.d_post_group_ptr + L.group_off, tiles, 0, nullptr);
    ++nk;
  }
  joint_root_kernel<Q><<<tiles, TTB_BLOCK, 0, s>>>(d);
  zero_slots_kernel<<<4, 256, 0, s>>>(d);
  nk += 2;
  for (int l = 0; trace && l < pl.n_jpre_levels; ++l) {
    const TtbLevelLaunch& L = pl.jpre_levels[l];
    joint_pre_level_kernel<Q><<<(unsigned)((long long)L.n_groups * tiles), TTB_BLOCK, 0, s>>>(d, pl.d_jpre_nodes + L.group_off, tiles,
                                                                                            pl.count_diff);
    ++nk;
  }
  finish_kernel<<<1, 256, 0, s>>>(d, tiles);
  return nk + 1;
}

int enqueue_joint_retrace_q(const TtbPassPlan& pl, const uint8_t* d_root_idx, cudaStream_t s) {
  const TtbDev& d = pl.d;
  const int tiles = pl.tiles;
  joint_root_override_kernel<Q><<<tiles, TTB_BLOCK, 0, s>>>(d, d_root_idx);
  zero_slots_kernel<<<4, 256, 0, s>>>(d);
  int nk = 2;
  for (int l = 0; l < pl.n_jpre_levels; ++l) {
    const TtbLevelLaunch& L = pl.jpre_levels[l];
    joint_pre_level_kernel<Q><<<(unsigned)((long long)L.n_groups * tiles), TTB_BLOCK, 0, s>>>(d, pl.d_jpre_nodes + L.group_off, tiles,
                                                                                            pl.count_diff);
    ++nk;
  }
  finish_kernel<<<1, 256, 0, s>>>(d, tiles);
  return nk + 1;
}

void fetch_node_q(const TtbDev& d, int tiles, int node, int which, double* out, cudaStream_t s) {
  if constexpr (HAS_SS) {
    if (d.site_specific) {
      fetch_node_kernel<Q, true><<<tiles, TTB_BLOCK, 0, s>>>(d, node, which, out);
      return;
    }
  }
  fetch_node_kernel<Q, false><<<tiles, TTB_BLOCK, 0, s>>>(d, node, which, out);
}

void branch_eval_q(const TtbDev& d, int n_eval, int nb, const int* nodes, const int* kinds, const double* ts, int mode,
                   double* partial, double* out, cudaStream_t s) {
  bool done = false;
  if constexpr (HAS_SS) {
    if (d.site_specific) {
      branch_eval_kernel<Q, true><<<dim3(n_eval, nb), TTB_BLOCK, 0, s>>>(d, nodes, kinds, ts, mode, partial);
      done = true;
    }
  }
  if (!done) branch_eval_kernel<Q, false><<<dim3(n_eval, nb), TTB_BLOCK, 0, s>>>(d, nodes, kinds, ts, mode, partial);
  branch_reduce_kernel<<<(n_eval + 127) / 128, 128, 0, s>>>(partial, n_eval, nb, out);
}

void counts_q(const TtbDev& d, int tiles, int chunks, int chunk, double* partial, double* out, cudaStream_t s) {
  counts_kernel<Q><<<dim3(tiles, chunks), TTB_BLOCK, 0, s>>>(d, chunk, partial);
  const int width = Q * Q + Q;
  counts_reduce_kernel<<<(width + 127) / 128, 128, 0, s>>>(partial, chunks * tiles, width, out);
}
int seqgen_q(const TtbDev& d, int tiles, unsigned long long seed, const uint8_t* root_idx, const double* uniforms, uint8_t* states,
             cudaStream_t s) {
  int nk = 0;
  if constexpr (HAS_SS) {
    if (d.site_specific) {
      seqgen_kernel<Q, true><<<tiles, TTB_BLOCK, 0, s>>>(d, seed, root_idx, uniforms, states);
      return 1;
    }
  }
  const int nthr = d.n_nodes * Q;
  expqt_kernel<Q><<<(nthr + 127) / 128, 128, 0, s>>>(d);
  ++nk;
  seqgen_kernel<Q, false><<<tiles, TTB_BLOCK, 0, s>>>(d, seed, root_idx, uniforms, states);
  return nk + 1;
}
void site_counts_q(const TtbDev& d, int tiles, int chunks, int chunk, double* partial, double* out, cudaStream_t s) {
  bool done = false;
  if constexpr (HAS_SS) {
    if (d.site_specific) {
      site_counts_kernel<Q, true><<<dim3(tiles, chunks), TTB_BLOCK, 0, s>>>(d, chunk, partial);
      done = true;
    }
  }
  if (!done) site_counts_kernel<Q, false><<<dim3(tiles, chunks), TTB_BLOCK, 0, s>>>(d, chunk, partial);
  int dev = 0, n_sm = 0;   // grid-stride reduction: four blocks per multiprocessor of the current device
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  site_counts_reduce_kernel<<<std::max(1, n_sm) * 4, 256, 0, s>>>(partial, chunks, (long long)(Q * Q + Q) * d.ld, out);
}
void sample_states_q(const TtbDev& d, int tiles, int n, const int* nodes, const double* uniforms, const uint8_t* prev_idx,
                     const uint8_t* prev_idxtip, unsigned long long* counts, cudaStream_t s) {
  sample_states_kernel<Q><<<dim3(tiles, n), TTB_BLOCK, 0, s>>>(d, nodes, uniforms, prev_idx, prev_idxtip, counts);
}
}  // namespace

#define TTB_CAT2(a, b) a##b
#define TTB_CAT(a, b) TTB_CAT2(a, b)
extern const TtbQOps TTB_CAT(ttb_qops_, TTB_Q) = {prepare_q, enqueue_pass_q, enqueue_joint_q, enqueue_joint_retrace_q, fetch_node_q, branch_eval_q, counts_q, seqgen_q, site_counts_q, sample_states_q};
