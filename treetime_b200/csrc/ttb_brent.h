// ttb_brent.h -- device-resident state of the lock-step Brent minimiser (see ttb_brent.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct TtbBrent {
  // bracket and Brent state, one entry per branch (names as in scipy's Brent.optimize)
  double *xa, *xb, *xc, *fa, *fb, *fc;
  double *x, *w, *v, *fx, *fw, *fv, *a, *b, *deltax, *rat, *u;
  double* ts;        // trial branch length u^2 handed to branch_eval_kernel; < 0: branch finished, skip it
  const double* f;   // objective values of the last evaluation (all-reduced by the caller when sharded)
  int *nit, *nfev;
  uint8_t* active;
  int* flags;        // [0] branches still active after this step, [1] bracket errors (1: ordering, 2: values)
};

// one step for all n branches: stage -1 start, 0..2 bracket points evaluated, >= 3 trial point evaluated
void ttb_brent_step(const TtbBrent& B, int n, int stage, double tol, int maxiter, cudaStream_t s);
