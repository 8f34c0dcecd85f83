// ttb_kernels.cuh -- sm_100a kernels of the marginal ancestral-reconstruction engine.
//
// Data layout in HBM (DESIGN.md "Layout"): every message array is STATE-PLANAR,
//   A[slot][state][pattern]   (pattern stride = ld, a multiple of 32 doubles = 256 B)
// so a warp that owns 32 consecutive patterns reads/writes q fully coalesced 256-byte
// rows per node and no byte of padding ever crosses HBM (the 5->8 padded AoS layout would
// move 60 % more bytes).  One thread = one alignment pattern; the q-state vectors live in
// registers; the per-branch exp(Qt) matrices are staged in shared memory and read as
// warp-wide broadcasts.
//
// Arithmetic contract (reference lines cited per kernel): products
// are taken in linear space with exact power-of-two rescaling instead of the reference's
// sum of logs; one log per (internal node, pattern) survives.  This is the same function
// up to fp64 rounding (measured: |dLH|/|LH| ~ 1e-15, profiles ~1e-16).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TTB_TINY 1e-12        // ttconf.TINY_NUMBER  (treetime/config.py:4)
#define TTB_SUPERTINY 1e-24   // ttconf.SUPERTINY_NUMBER (treetime/config.py:5)
#define TTB_BLOCK 128
#define TTB_CB 4              // children whose exp(Qt) are staged per shared-memory batch

struct TtbDev {
  int q;
  long long Lp;       // patterns in this shard
  long long ld;       // padded pattern stride (multiple of 32)
  int n_nodes, n_int, n_tips, n_codes;
  int gap_index;
  // tree
  const int* parent;
  const int* child_ptr;
  const int* child_idx;
  const int* tip_row;   // node -> row of codes (tips) or -1
  const int* int_slot;  // node -> slot in S/F/M (internal) or -1
  // alignment
  const uint8_t* codes;       // [n_tips][ld]
  const double* code_prof;    // [n_codes][q]
  const uint32_t* code_mask;  // bit i set <=> code_prof[code][i] != 0
  const double* mult;         // [ld]
  // model
  const double* t;       // [n_nodes]
  const double* eig;     // [q]
  const double* v;       // [q][q]
  const double* vinv;    // [q][q]
  const double* Pi;      // [q]
  const double* mu;      // [1] (device scalar so that a new rate does not invalidate the graph)
  // state
  double* P;     // [n_nodes][q*q]  exp(Q t_c), P[i*q+j] = Prob(child=i | parent=j)
  double* S;     // [n_int][q][ld]  marginal_subtree_LH
  double* F;     // [n_int][ld]     marginal_subtree_LH_prefactor
  double* M;     // [n_int][q][ld]  marginal_profile
  double* Mtip;  // [n_tips][q][ld] marginal_profile of tips (reconstruct_tip_states) or null
  uint8_t* idx;     // [n_int][ld]  argmax state
  uint8_t* idxtip;  // [n_tips][ld] or null
  double* LH;       // [ld] tree.sequence_LH
  double* lh_partial;             // [tiles]
  unsigned long long* nd_slots;   // [1024]
  double* results;                // {total_lh, n_diff}
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Deterministic block sum (fixed shuffle tree + fixed order over warps); result valid in thread 0.
template <int BLOCK>
__device__ __forceinline__ double block_sum(double x, double* sred) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sred[w] = x;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < BLOCK / 32; ++i) r += sred[i];
  }
  return r;
}

// ---------------------------------------------------------------------------------------
// A1: batched exp(Qt) for every branch.  Reference: GTR._exp_lt / GTR.expQt, gtr.py:1027-1067:
//   expQt = max(0, v . diag(exp(mu t lambda)) . v_inv).  One thread per (branch, row i).
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void expqt_kernel(TtbDev p) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= p.n_nodes * Q) return;
  const int n = gid / Q, i = gid % Q;
  const double mt = p.mu[0] * p.t[n];
  double ev[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) ev[k] = p.v[i * Q + k];
  double e[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) e[k] = exp(mt * p.eig[k]);
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) acc = fma(ev[k], e[k] * p.vinv[k * Q + j], acc);
    p.P[(size_t)n * Q * Q + i * Q + j] = fmax(0.0, acc);
  }
}

// Shared-memory carve-up used by the level kernels.
template <int Q>
struct Smem {
  double* sP;        // [TTB_CB][Q*Q]
  double* sprof;     // [n_codes][Q]
  uint32_t* smask;   // [n_codes]
  double* sred;      // [BLOCK/32]
  __device__ Smem(unsigned char* base, int n_codes) {
    sP = reinterpret_cast<double*>(base);
    sprof = sP + TTB_CB * Q * Q;
    sred = sprof + n_codes * Q;
    smask = reinterpret_cast<uint32_t*>(sred + TTB_BLOCK / 32);
  }
  static size_t bytes(int n_codes) {
    return sizeof(double) * (TTB_CB * Q * Q + (size_t)n_codes * Q + TTB_BLOCK / 32) + sizeof(uint32_t) * n_codes;
  }
};

template <int Q>
__device__ __forceinline__ void load_code_tables(const TtbDev& p, Smem<Q>& sm) {
  for (int k = threadIdx.x; k < p.n_codes * Q; k += blockDim.x) sm.sprof[k] = p.code_prof[k];
  for (int k = threadIdx.x; k < p.n_codes; k += blockDim.x) sm.smask[k] = p.code_mask[k];
}

// Stage exp(Qt) of children [c0, c0+nb) of a node into shared memory.
template <int Q>
__device__ __forceinline__ void stage_P(const TtbDev& p, Smem<Q>& sm, int c0, int nb) {
  __syncthreads();
  for (int k = threadIdx.x; k < nb * Q * Q; k += blockDim.x) {
    const int c = p.child_idx[c0 + k / (Q * Q)];
    sm.sP[k] = p.P[(size_t)c * Q * Q + (k % (Q * Q))];
  }
  __syncthreads();
}

// Child -> parent message U[j] = sum_i S_c[i] P[i][j]  (gtr.propagate_profile, gtr.py:965-995,
// without the log) and the child's subtree profile S_c (tips: the 0/1 ambiguity profile of the
// pattern character, seq2prof, seq_utils.py:207-229; zero entries are skipped, which is exact).
template <int Q, bool WANT_S>
__device__ __forceinline__ void child_message(const TtbDev& p, const Smem<Q>& sm, const double* __restrict__ Pc,
                                              int c, long long a, double (&U)[Q], double (&Sc)[Q]) {
  const int row = p.tip_row[c];
  if (row >= 0) {
    const int code = p.codes[(size_t)row * p.ld + a];
    uint32_t m = sm.smask[code];
#pragma unroll
    for (int j = 0; j < Q; ++j) U[j] = 0.0;
    if (WANT_S) {
#pragma unroll
      for (int i = 0; i < Q; ++i) Sc[i] = sm.sprof[code * Q + i];
    }
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      const double w = sm.sprof[code * Q + i];
#pragma unroll
      for (int j = 0; j < Q; ++j) U[j] = fma(w, Pc[i * Q + j], U[j]);
    }
  } else {
    const double* __restrict__ s = p.S + (size_t)p.int_slot[c] * Q * p.ld + a;
    double sc[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) sc[i] = __ldg(s + (size_t)i * p.ld);
#pragma unroll
    for (int j = 0; j < Q; ++j) U[j] = sc[0] * Pc[j];
#pragma unroll
    for (int i = 1; i < Q; ++i) {
#pragma unroll
      for (int j = 0; j < Q; ++j) U[j] = fma(sc[i], Pc[i * Q + j], U[j]);
    }
    if (WANT_S) {
#pragma unroll
      for (int i = 0; i < Q; ++i) Sc[i] = sc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// A3-A5: one postorder level.  Reference: postorder_traversal_marginal, treeanc.py:857-878 +
// normalize_profile(log=True), seq_utils.py:279-307.  Block = (internal node, 128-pattern tile).
//   X[j] = prod_c U_c[j];  Z = sum_j X[j];  S_n = X/Z;  F_n = sum_c F_c + log Z.
// Nodes with many children are rescaled by exact powers of two so the product cannot underflow.
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) post_level_kernel(TtbDev p, const int* __restrict__ level_nodes, int tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<Q> sm(smem_raw, p.n_codes);
  const int node = level_nodes[blockIdx.x / tiles];
  const long long a = (long long)(blockIdx.x % tiles) * TTB_BLOCK + threadIdx.x;
  const bool act = a < p.Lp;
  load_code_tables<Q>(p, sm);
  double X[Q];
#pragma unroll
  for (int j = 0; j < Q; ++j) X[j] = 1.0;
  double F = 0.0;
  int scale = 0;  // X holds the true product times 2^(256*scale)
  const int cb = p.child_ptr[node], ce = p.child_ptr[node + 1];
  int seen = 0;
  for (int c0 = cb; c0 < ce; c0 += TTB_CB) {
    const int nb = min(TTB_CB, ce - c0);
    stage_P<Q>(p, sm, c0, nb);
    if (act) {
      for (int b = 0; b < nb; ++b) {
        const int c = p.child_idx[c0 + b];
        double U[Q], dummy[Q];
        child_message<Q, false>(p, sm, sm.sP + b * Q * Q, c, a, U, dummy);
        const int slot = p.int_slot[c];
        if (slot >= 0) F += __ldg(p.F + (size_t)slot * p.ld + a);
#pragma unroll
        for (int j = 0; j < Q; ++j) X[j] *= U[j];
        if (++seen > 2) {  // polytomy: keep the running product in range (exact scaling)
          double mx = X[0];
#pragma unroll
          for (int j = 1; j < Q; ++j) mx = fmax(mx, X[j]);
          if (mx < 0x1p-256 && mx > 0.0) {
#pragma unroll
            for (int j = 0; j < Q; ++j) X[j] *= 0x1p+256;
            ++scale;
          }
        }
      }
    }
  }
  if (act) {
    double Z = X[0];
#pragma unroll
    for (int j = 1; j < Q; ++j) Z += X[j];
    const double inv = 1.0 / Z;
    const int slot = p.int_slot[node];
    double* __restrict__ s = p.S + (size_t)slot * Q * p.ld + a;
#pragma unroll
    for (int j = 0; j < Q; ++j) s[(size_t)j * p.ld] = X[j] * inv;
    p.F[(size_t)slot * p.ld + a] = F + (log(Z) - scale * (256.0 * 0.693147180559945309417232121458));
  }
}

__device__ __forceinline__ int argmax_first(const double* x, int q) {
  int best = 0;
  double bv = x[0];
  for (int i = 1; i < q; ++i)
    if (x[i] > bv) { bv = x[i]; best = i; }
  return best;
}

// ---------------------------------------------------------------------------------------
// A5': root.  Reference: total_LH_and_root_sequence, treeanc.py:814-838.
//   profile_r = normalize(Pi * S_r);  LH_a = F_r + log Z_r;  partial sums of LH_a * m_a.
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) root_kernel(TtbDev p, int lh_only) {
  __shared__ double sred[TTB_BLOCK / 32];
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  double contrib = 0.0;
  if (a < p.Lp) {
    const int slot = p.int_slot[0];
    const double* __restrict__ s = p.S + (size_t)slot * Q * p.ld + a;
    double R[Q];
    double Z = 0.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      R[j] = p.Pi[j] * s[(size_t)j * p.ld];
      Z += R[j];
    }
    const double lh = p.F[(size_t)slot * p.ld + a] + log(Z);
    p.LH[a] = lh;
    contrib = lh * p.mult[a];
    if (!lh_only) {
      const double inv = 1.0 / Z;
      double* __restrict__ m = p.M + (size_t)slot * Q * p.ld + a;
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        R[j] *= inv;
        m[(size_t)j * p.ld] = R[j];
      }
      int best = 0;
      double bv = R[0];
#pragma unroll
      for (int i = 1; i < Q; ++i)
        if (R[i] > bv) { bv = R[i]; best = i; }
      p.idx[(size_t)slot * p.ld + a] = (uint8_t)best;
    }
  }
  const double bs = block_sum<TTB_BLOCK>(contrib, sred);
  if (threadIdx.x == 0) p.lh_partial[blockIdx.x] = bs;
}

// Final deterministic reduction: total LH over tiles, N_diff over slots.
static __global__ void __launch_bounds__(256) finish_kernel(TtbDev p, int tiles) {
  __shared__ double sred[256 / 32];
  double x = 0.0;
  for (int i = threadIdx.x; i < tiles; i += 256) x += p.lh_partial[i];
  const double tot = block_sum<256>(x, sred);
  __shared__ unsigned long long snd[256];
  unsigned long long nd = 0;
  for (int i = threadIdx.x; i < 1024; i += 256) nd += p.nd_slots[i];
  snd[threadIdx.x] = nd;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int i = 0; i < 256; ++i) s += snd[i];
    p.results[0] = tot;
    p.results[1] = (double)s;
  }
}

static __global__ void zero_slots_kernel(TtbDev p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1024) p.nd_slots[i] = 0ull;
}

// Outside message of child c given the (clamped) parent profile Mp and the child's
// up-message U:  O ~ max(1e-12, profile_p) / U_c, normalised.  Reference:
// treeanc.py:895-899 (log(max(TINY, up.marginal_profile)) - marginal_log_Lx, normalize(log=True)).
// For small alphabets the quotient is formed division-free as Mp[j] * prod_{k != j} U[k].
template <int Q>
__device__ __forceinline__ void outgroup_message(const double (&Mp)[Q], const double (&U)[Q], double (&O)[Q]) {
  double z = 0.0;
  if (Q <= 8) {
    double pre[Q], suf[Q];
    pre[0] = 1.0;
#pragma unroll
    for (int j = 1; j < Q; ++j) pre[j] = pre[j - 1] * U[j - 1];
    suf[Q - 1] = 1.0;
#pragma unroll
    for (int j = Q - 2; j >= 0; --j) suf[j] = suf[j + 1] * U[j + 1];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      O[j] = Mp[j] * (pre[j] * suf[j]);
      z += O[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      O[j] = Mp[j] / U[j];
      z += O[j];
    }
  }
  const double inv = 1.0 / z;
#pragma unroll
  for (int j = 0; j < Q; ++j) O[j] *= inv;
}

// ---------------------------------------------------------------------------------------
// A6-A7: one preorder level.  Reference: preorder_traversal_marginal, treeanc.py:887-930 +
// GTR.evolve (gtr.py:997-1025) + prof2seq argmax (seq_utils.py:271).
// Block = (parent p, 128-pattern tile); the parent's marginal profile is read once and
// reused for all its children.  Per child c:
//   O_c   ~ max(1e-12, profile_p) / U_c                     (outside message, not stored:
//                                                            it is recomputed on demand by
//                                                            fetch / branch kernels)
//   msg_i = sum_j O_c[j] P_c[i][j];  profile_c = normalize(S_c * msg);  state = argmax.
// Tips are skipped unless TIPS (reconstruct_tip_states).
// ---------------------------------------------------------------------------------------
template <int Q, bool TIPS>
__global__ void __launch_bounds__(TTB_BLOCK) pre_level_kernel(TtbDev p, const int* __restrict__ level_parents, int tiles,
                                                             int count_diff) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<Q> sm(smem_raw, p.n_codes);
  const int pn = level_parents[blockIdx.x / tiles];
  const long long a = (long long)(blockIdx.x % tiles) * TTB_BLOCK + threadIdx.x;
  const bool act = a < p.Lp;
  load_code_tables<Q>(p, sm);
  double Mp[Q];
  if (act) {
    const double* __restrict__ m = p.M + (size_t)p.int_slot[pn] * Q * p.ld + a;
#pragma unroll
    for (int j = 0; j < Q; ++j) Mp[j] = fmax(TTB_TINY, __ldg(m + (size_t)j * p.ld));
  }
  unsigned int ndiff = 0;
  const int cb = p.child_ptr[pn], ce = p.child_ptr[pn + 1];
  for (int c0 = cb; c0 < ce; c0 += TTB_CB) {
    const int nb = min(TTB_CB, ce - c0);
    stage_P<Q>(p, sm, c0, nb);
    if (act) {
      for (int b = 0; b < nb; ++b) {
        const int c = p.child_idx[c0 + b];
        const int row = p.tip_row[c];
        if (!TIPS && row >= 0) continue;
        const double* Pc = sm.sP + b * Q * Q;
        double U[Q], Sc[Q], O[Q];
        child_message<Q, true>(p, sm, Pc, c, a, U, Sc);
        outgroup_message<Q>(Mp, U, O);
        double prof[Q];
        double z = 0.0;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          double msg = O[0] * Pc[i * Q];
#pragma unroll
          for (int j = 1; j < Q; ++j) msg = fma(O[j], Pc[i * Q + j], msg);
          prof[i] = Sc[i] * msg;
          z += prof[i];
        }
        const double inv = 1.0 / z;
        double* __restrict__ out;
        uint8_t* ip;
        if (row >= 0) {
          out = p.Mtip + (size_t)row * Q * p.ld + a;
          ip = p.idxtip + (size_t)row * p.ld + a;
        } else {
          const int slot = p.int_slot[c];
          out = p.M + (size_t)slot * Q * p.ld + a;
          ip = p.idx + (size_t)slot * p.ld + a;
        }
        int best = 0;
        double bv = -1.0;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const double x = prof[i] * inv;
          out[(size_t)i * p.ld] = x;
          if (x > bv) { bv = x; best = i; }
        }
        if (count_diff) ndiff += (*ip != (uint8_t)best);
        *ip = (uint8_t)best;
      }
    }
  }
  if (count_diff) {
    ndiff = __reduce_add_sync(0xffffffffu, ndiff);
    if ((threadIdx.x & 31) == 0 && ndiff) atomicAdd(p.nd_slots + (blockIdx.x & 1023), (unsigned long long)ndiff);
  }
}

// ---------------------------------------------------------------------------------------
// Per-node fetch in the reference's (L', q) row-major layout (TreeAnc node attributes).
// ---------------------------------------------------------------------------------------
template <int Q>
__device__ __forceinline__ void node_subtree(const TtbDev& p, int n, long long a, double (&Sc)[Q]) {
  const int row = p.tip_row[n];
  if (row >= 0) {
    const int code = p.codes[(size_t)row * p.ld + a];
#pragma unroll
    for (int i = 0; i < Q; ++i) Sc[i] = p.code_prof[code * Q + i];
  } else {
    const double* s = p.S + (size_t)p.int_slot[n] * Q * p.ld + a;
#pragma unroll
    for (int i = 0; i < Q; ++i) Sc[i] = s[(size_t)i * p.ld];
  }
}

// (pp, pc) of the branch above node n from resident messages: pc = S_n, pp = O_n
// (marginal_branch_profile, treeanc.py:1122-1146).  kind 1 = merged root branch
// (treeanc.py:1317-1326): n = n1, pp = normalize(S_n2 * Pi).
template <int Q>
__device__ __forceinline__ void branch_profiles(const TtbDev& p, int n, int kind, long long a, double (&pp)[Q], double (&pc)[Q]) {
  node_subtree<Q>(p, n, a, pc);
  if (kind == 1) {
    const int c0 = p.child_ptr[0];
    const int n1 = p.child_idx[c0], n2 = p.child_idx[c0 + 1];
    const int other = (n == n1) ? n2 : n1;
    double s2[Q];
    node_subtree<Q>(p, other, a, s2);
    double z = 0.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      pp[j] = s2[j] * p.Pi[j];
      z += pp[j];
    }
    const double inv = 1.0 / z;
#pragma unroll
    for (int j = 0; j < Q; ++j) pp[j] *= inv;
    return;
  }
  const int up = p.parent[n];
  double Mp[Q], U[Q];
  const double* m = p.M + (size_t)p.int_slot[up] * Q * p.ld + a;
#pragma unroll
  for (int j = 0; j < Q; ++j) Mp[j] = fmax(TTB_TINY, m[(size_t)j * p.ld]);
  const double* Pc = p.P + (size_t)n * Q * Q;
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    double u = 0.0;
#pragma unroll
    for (int i = 0; i < Q; ++i) u = fma(pc[i], Pc[i * Q + j], u);
    U[j] = u;
  }
  outgroup_message<Q>(Mp, U, pp);
}

template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) fetch_node_kernel(TtbDev p, int node, int which, double* __restrict__ out) {
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  if (a >= p.Lp) return;
  double x[Q], y[Q];
  if (which == 0) {
    node_subtree<Q>(p, node, a, x);
  } else if (which == 1) {
    if (node == 0) {
#pragma unroll
      for (int j = 0; j < Q; ++j) x[j] = p.Pi[j];
    } else {
      branch_profiles<Q>(p, node, 0, a, x, y);
    }
  } else {
    const int row = p.tip_row[node];
    const double* m = (row >= 0) ? p.Mtip + (size_t)row * Q * p.ld + a : p.M + (size_t)p.int_slot[node] * Q * p.ld + a;
#pragma unroll
    for (int j = 0; j < Q; ++j) x[j] = m[(size_t)j * p.ld];
  }
#pragma unroll
  for (int j = 0; j < Q; ++j) out[(size_t)a * Q + j] = x[j];
}

// ---------------------------------------------------------------------------------------
// A8: branch-length likelihood surface.  Reference: GTR.prob_t_profiles, gtr.py:922-963:
//   f(t) = sum_a m_a log(sum_ij pc[a,i] expQt(t)[i,j] pp[a,j] + 1e-24) (1-pp[a,gap])(1-pc[a,gap])
// grid = (n_eval, NB): block (e, b) forms expQt(t_e) in shared memory and strides over the
// patterns; partial[e][b] is reduced in fixed order by branch_reduce_kernel.
// mode 0: objective; mode 1: sum_a m_a (pp_a . pc_a)  (hamming numerator, gtr.py:871-874).
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) branch_eval_kernel(TtbDev p, const int* __restrict__ nodes, const int* __restrict__ kinds,
                                                               const double* __restrict__ ts, int mode, double* __restrict__ partial) {
  __shared__ double sPt[Q * Q];
  __shared__ double sred[TTB_BLOCK / 32];
  const int e = blockIdx.x;
  const int node = nodes[e];
  const int kind = kinds ? kinds[e] : 0;
  if (mode == 0) {
    const double mt = p.mu[0] * ts[e];
    for (int k = threadIdx.x; k < Q * Q; k += TTB_BLOCK) {
      const int i = k / Q, j = k % Q;
      double acc = 0.0;
      for (int m = 0; m < Q; ++m) acc = fma(p.v[i * Q + m], exp(mt * p.eig[m]) * p.vinv[m * Q + j], acc);
      sPt[k] = fmax(0.0, acc);
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (long long a = (long long)blockIdx.y * TTB_BLOCK + threadIdx.x; a < p.Lp; a += (long long)gridDim.y * TTB_BLOCK) {
    double pp[Q], pc[Q];
    branch_profiles<Q>(p, node, kind, a, pp, pc);
    if (mode == 0) {
      double g = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < Q; ++j) w = fma(sPt[i * Q + j], pp[j], w);
        g = fma(pc[i], w, g);
      }
      double val = p.mult[a] * log(g + TTB_SUPERTINY);
      if (p.gap_index >= 0) val *= (1.0 - pp[p.gap_index]) * (1.0 - pc[p.gap_index]);
      acc += val;
    } else {
      double d = 0.0;
#pragma unroll
      for (int j = 0; j < Q; ++j) d = fma(pp[j], pc[j], d);
      acc += p.mult[a] * d;
    }
  }
  const double bs = block_sum<TTB_BLOCK>(acc, sred);
  if (threadIdx.x == 0) partial[(size_t)e * gridDim.y + blockIdx.y] = bs;
}

static __global__ void branch_reduce_kernel(const double* __restrict__ partial, int n_eval, int nb, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_eval) return;
  double s = 0.0;
  for (int b = 0; b < nb; ++b) s += partial[(size_t)e * nb + b];
  out[e] = s;
}

// ---------------------------------------------------------------------------------------
// A10: expected substitution statistics.  Reference: get_branch_mutation_matrix
// (treeanc.py:1085-1120) accumulated as in infer_gtr(marginal=True) (treeanc.py:1556-1572):
//   M_a[i][j] = pc[a,i] pp[a,j] (expQt[i][j] + 1e-24) / sum_ij(...)
//   n_ij += M_a m_a;   T_k += 0.5 t m_a (sum_i M_a[i][k] + sum_j M_a[k][j])
// grid = (tiles, branch chunks); partial[(chunk*tiles+tile)][Q*Q+Q] reduced in fixed order.
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) counts_kernel(TtbDev p, int chunk, double* __restrict__ partial) {
  __shared__ double sred[TTB_BLOCK / 32];
  __shared__ double sPc[Q * Q];
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  const bool act = a < p.Lp;
  double nij[Q * Q], Ti[Q];
#pragma unroll
  for (int k = 0; k < Q * Q; ++k) nij[k] = 0.0;
#pragma unroll
  for (int k = 0; k < Q; ++k) Ti[k] = 0.0;
  const int n0 = 1 + blockIdx.y * chunk;
  const int n1 = min(p.n_nodes, n0 + chunk);
  const double m = act ? p.mult[a] : 0.0;
  for (int n = n0; n < n1; ++n) {
    __syncthreads();
    for (int k = threadIdx.x; k < Q * Q; k += TTB_BLOCK) sPc[k] = p.P[(size_t)n * Q * Q + k] + TTB_SUPERTINY;
    __syncthreads();
    if (!act) continue;
    double pp[Q], pc[Q];
    branch_profiles<Q>(p, n, 0, a, pp, pc);
    double tot = 0.0;
    double mm[Q * Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        mm[i * Q + j] = pc[i] * pp[j] * sPc[i * Q + j];
        tot += mm[i * Q + j];
      }
    const double w = m / tot;
    const double ht = 0.5 * p.t[n];
#pragma unroll
    for (int i = 0; i < Q; ++i)
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const double x = mm[i * Q + j] * w;
        nij[i * Q + j] += x;
        Ti[i] += ht * x;
        Ti[j] += ht * x;
      }
  }
  double* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (Q * Q + Q);
  for (int k = 0; k < Q * Q; ++k) {
    const double bs = block_sum<TTB_BLOCK>(nij[k], sred);
    if (threadIdx.x == 0) out[k] = bs;
  }
  for (int k = 0; k < Q; ++k) {
    const double bs = block_sum<TTB_BLOCK>(Ti[k], sred);
    if (threadIdx.x == 0) out[Q * Q + k] = bs;
  }
}

static __global__ void counts_reduce_kernel(const double* __restrict__ partial, int n_part, int width, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= width) return;
  double s = 0.0;
  for (int b = 0; b < n_part; ++b) s += partial[(size_t)b * width + k];
  out[k] = s;
}
