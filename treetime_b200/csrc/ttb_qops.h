// ttb_qops.h -- per-alphabet-size launch table.  Every n_states value has its own
// translation unit (ttb_q.cu compiled with -DTTB_Q=<q>) so the template instantiations
// build in parallel; ttb_api.cu only sees this table.
#pragma once
#include <cuda_runtime.h>
#include "ttb_kernels.cuh"

struct TtbLevel {
  int begin, count;
};

struct TtbPassPlan {
  TtbDev d;
  int tiles;
  const int* d_post_nodes;
  const TtbLevel* post_levels;
  int n_post_levels;
  const int* d_pre_nodes;      // parents list matching `tips`
  const TtbLevel* pre_levels;
  int n_pre_levels;
  bool lh_only, tips;
  int count_diff;
};

struct TtbQOps {
  // enqueue every kernel of one pass; optional events ev[6] bracket the phases; returns #kernels
  int (*enqueue_pass)(const TtbPassPlan& plan, cudaStream_t s, cudaEvent_t* ev, int* phase_kernels);
  void (*fetch_node)(const TtbDev& d, int tiles, int node, int which, double* out, cudaStream_t s);
  void (*branch_eval)(const TtbDev& d, int n_eval, int nb, const int* nodes, const int* kinds, const double* ts,
                      int mode, double* partial, double* out, cudaStream_t s);
  void (*counts)(const TtbDev& d, int tiles, int chunks, int chunk, double* partial, double* out, cudaStream_t s);
};

const TtbQOps* ttb_qops(int q);
