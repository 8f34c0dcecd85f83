// ttb_qops.h -- per-alphabet-size launch table.  Every n_states value has its own
// translation unit (ttb_q.cu compiled with -DTTB_Q=<q>) so the template instantiations
// build in parallel; ttb_api.cu only sees this table.
#pragma once
#include <cuda_runtime.h>
#include "ttb_kernels.cuh"

// One level launch: blocks = n_groups x tiles; group g covers chunks
// [group_ptr[group_off + g], group_ptr[group_off + g + 1]).
struct TtbLevelLaunch {
  int group_off, n_groups;
  int dep;   // 1: the launch merges several consecutive small levels (one group per tile; the kernel honours the chunks' `dep`)
};

struct TtbPassPlan {
  TtbDev d;
  int tiles;
  const int* d_tip_nodes;          // tip row -> node id
  const TtbChunk* d_post_chunks;
  const int* d_post_node_chunk;    // first chunk of every scheduled postorder node (+ sentinel)
  int n_post_leaf_nodes;           // nodes of postorder level 1 (all children are tips)
  double* d_leaf_pairs;            // cherry tables of postorder level 1 ([n_leaf_chunks][n_codes^2][stride]) or null
  int n_leaf_chunks;               // chunks of postorder level 1
  const int* d_post_group_ptr;
  const int* d_post_dep;           // per chunk: global index of the chunk that wrote what it reads (merged-level launches), -1 = none
  const TtbLevelLaunch* post_levels;
  int n_post_levels;
  const TtbChunk* d_pre_chunks;    // schedule matching `tips`
  const int* d_pre_group_ptr;
  const int* d_pre_dep;
  const TtbLevelLaunch* pre_levels;
  int n_pre_levels;
  bool lh_only, tips;
  int count_diff;
  // joint reconstruction: nodes to back-trace, grouped by depth (matching `tips`)
  const int* d_jpre_nodes;
  const TtbLevelLaunch* jpre_levels;   // group_off = first node, n_groups = node count
  int n_jpre_levels;
};

struct TtbQOps {
  // one-time per process: opt in to the dynamic shared memory the level kernels need; returns 0 or a cudaError
  int (*prepare)(const TtbDev& d);
  // enqueue every kernel of one pass; optional events ev[6] bracket the phases; returns #kernels
  int (*enqueue_pass)(const TtbPassPlan& plan, cudaStream_t s, cudaEvent_t* ev, int* phase_kernels);
  // joint (max-product) reconstruction, same schedule; returns #kernels, 0 if unsupported for this model
  int (*enqueue_joint)(const TtbPassPlan& plan, cudaStream_t s, int trace);
  // root states chosen by the caller (root sampling) + back-trace
  int (*enqueue_joint_retrace)(const TtbPassPlan& plan, const uint8_t* d_root_idx, cudaStream_t s);
  void (*fetch_node)(const TtbDev& d, int tiles, int node, int which, double* out, cudaStream_t s);
  void (*branch_eval)(const TtbDev& d, int n_eval, int nb, const int* nodes, const int* kinds, const double* ts,
                      int mode, double* partial, double* out, cudaStream_t s);
  void (*counts)(const TtbDev& d, int tiles, int chunks, int chunk, double* partial, double* out, cudaStream_t s);
  // N4: evolve sequences down the tree; states[n_nodes][ld]; returns #kernels
  int (*seqgen)(const TtbDev& d, int tiles, unsigned long long seed, const uint8_t* root_idx, const double* uniforms, uint8_t* states,
                cudaStream_t s);
  // A10 per pattern: out[q*q+q][ld] (partial: [chunks][q*q+q][ld])
  void (*site_counts)(const TtbDev& d, int tiles, int chunks, int chunk, double* partial, double* out, cudaStream_t s);
  // sample_from_profile=True: draw the states of n nodes from their profiles (device arrays throughout)
  void (*sample_states)(const TtbDev& d, int tiles, int n, const int* nodes, const double* uniforms, const uint8_t* prev_idx,
                        const uint8_t* prev_idxtip, unsigned long long* counts, cudaStream_t s);
};

const TtbQOps* ttb_qops(int q);
