// ttb_kernels.cuh -- sm_100a kernels of the marginal ancestral-reconstruction engine.
//
// Data layout in HBM (DESIGN.md "Layout"): every message array is STATE-PLANAR,
//   A[slot][state][pattern]   (pattern stride = ld, a multiple of 32 doubles = 256 B)
// so the q values of 128 consecutive patterns of one node are q contiguous 1 KB rows and no
// byte of padding ever crosses HBM (a 5->8 padded AoS layout would move 60 % more bytes).
//
// Level kernels (postorder / preorder) are TMA pipelines: a block owns one 128-pattern tile
// and a run of nodes of the level; one warp issues `cp.async.bulk` (1-D TMA) copies of the
// child rows, the child's exp(Qt) and the tip tables into a 3-stage shared-memory ring and
// signals a "full" mbarrier per stage (expect_tx / complete_tx); all four warps consume a stage
// with one thread per pattern and the q-state vectors in registers, then every warp releases the
// stage on an "empty" mbarrier the producer waits on -- no block-wide barrier in steady state.
// Bytes in flight are decoupled from registers/occupancy, which is what an HBM-bound fp64
// kernel with ~100 registers of state needs.
//
// Arithmetic contract (reference lines cited per kernel): products are
// taken in linear space with exact power-of-two rescaling instead of the reference's sum of
// logs; one log per (internal node, pattern) survives.  Same function up to fp64 rounding
// (measured: |dLH|/|LH| ~ 1e-16, profiles ~1e-13).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TTB_TINY 1e-12        // ttconf.TINY_NUMBER  (treetime/config.py:4)
#define TTB_SUPERTINY 1e-24   // ttconf.SUPERTINY_NUMBER (treetime/config.py:5)
#define TTB_BLOCK 128
// the streaming level kernels add one producer warp (bulk-copy issue only) to the TTB_BLOCK pattern threads
#define TTB_LEVEL_THREADS (TTB_BLOCK + 32)
#define TTB_TILE 128          // patterns per tile = threads per block: one 1 KB row per state
#define TTB_CB 2              // children per pipeline chunk (binary nodes = one chunk)
#define TTB_FLANES 16         // lanes of the two-stage reduction of the per-run log-prefactor sums
#define TTB_PF_STRIDE (2 * 18 * 32)   // doubles per branch of the fragment-ordered exp(Qt) (ttb_mma.cuh: 2 products x 18 fragments x 32 lanes)

// One pipeline chunk of a level kernel: up to TTB_CB children of one node (32 bytes).
//   postorder: out = slot of the node being computed; src[b] = slot of internal child b or
//              -1 - tip_row for a tip child; cnode[b] = child node id (row of P).
//   preorder : out = slot of the PARENT (profile source); src/cnode = children to reconstruct.
struct TtbChunk {
  int out;
  int flags;  // bit0 first chunk of the node, bit1 last chunk, bits 8.. number of children
  int src[TTB_CB];
  int cnode[TTB_CB];
  int lo[TTB_CB];  // site-specific models: interpolation bracket of child b's branch (patched by ss_patch_chunks_kernel)
};

struct TtbDev {
  int q;
  long long Lp;       // patterns in this shard
  long long ld;       // padded pattern stride of the byte / scalar rows (multiple of 32)
  int tiles;          // number of 128-pattern tiles = ceil(Lp / 128)
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
  const double* mult;         // [ld]
  // model
  const double* t;       // [n_nodes]
  const double* eig;     // [q]
  const double* v;       // [q][q]
  const double* vinv;    // [q][q]
  const double* Pi;      // [q]
  const double* mu;      // [1] (device scalar so that a new rate does not invalidate the graph)
  // site-specific model (gtr_site_specific.py): per-pattern eigen-systems, pattern-contiguous planes
  int site_specific;
  const double* ss_eig;   // [q][ld]      eigenvalue k of pattern a at k*ld + a
  const double* ss_mu;    // [ld]
  const double* ss_V;     // [q*q][ld]    V_a[i][k]    at (i*q+k)*ld + a
  const double* ss_Vinv;  // [q*q][ld]    Vinv_a[k][j] at (k*q+j)*ld + a
  const double* ss_Pi;    // [q][ld]
  const int* ss_lo;       // [n_nodes] lower grid index of the interpolation bracket of every branch length
  const double* ss_w;     // [n_nodes] (t - t_lo)/(t_hi - t_lo), or < 0: evaluate exp(Qt) exactly
  const double2* ss_rec;  // [n_nodes] {ss_w, t}: 16-byte records the level kernels stage with one bulk copy per child
  const double* ss_E;     // [tiles][ss_ngrid][q][128] exp(t_g mu_a lambda_k(a)) on the grid, tile-blocked like the messages:
                          // the two grid rows of a branch are one contiguous 2q KB block per tile (no exp in the level kernels)
  // symmetric form (reversible models, gtr.py:612-629): Vinv_a[k][j] = V_a[j][k] * c_k / Pi_a[j] with c_k the squared
  // 1-norm of eigenvector k.  When the uploaded model has this structure (checked at upload) the level kernels keep only
  // V_a and 1/Pi_a in registers (q^2 + q instead of 2 q^2 doubles) and read the eigen-factors pre-multiplied by c_k.
  int ss_sym;
  const double* ss_c;     // [q][ld]      c_k of pattern a at k*ld + a
  const double* ss_Ec;    // like ss_E, every entry times c_k(a)
  const double* ss_grid;  // [ss_ngrid] the grid itself (branch objective at trial lengths)
  int ss_ngrid;
  double ss_tmax;         // interpolate while t < ss_tmax (= 10 / rate_scale), 0 = never
  // per-branch masks (ARG mode, arg.py:128-133): mask_id[node] = row of `masks` or -1, masks[n_masks][ld] in {0,1};
  // both null when no node has a mask.  A masked (branch, pattern) carries no information: its up-message is 1
  // (treeanc.py:867-872), the child keeps its subtree profile (:914-917) and the pattern drops out of the
  // branch's multiplicities (:1294,1332,1564-1572).
  const int* mask_id;
  const uint8_t* masks;
  // state
  int pq;        // stride of one exp(Qt) matrix in doubles (q*q rounded up to even: 16-byte multiple for TMA)
  int tu_stride; // stride of one tip table in doubles (n_codes*q rounded up to even)
  double* TU;    // [n_tips][tu_stride]  tip message table: TU[code*q+j] = sum_i prof[code][i] P[i][j]
  double* P;     // [n_nodes][pq]  exp(Q t_c), P[i*q+j] = Prob(child=i | parent=j)
  double* Pf;    // [n_nodes][TTB_PF_STRIDE] the same matrices in mma-fragment order (q > 8, ttb_mma.cuh) or null
  // message arrays are TILE-BLOCKED state-planar: [slot][tile][state][128]: the q rows of one
  // (node, tile) are one contiguous q KB block = one TMA copy, and 128 consecutive patterns of a
  // state are one coalesced 1 KB row
  // storage type of S / M / Mtip: 0 = double, 1 = float (ttb_set_message_storage: the level kernels move half the bytes,
  // all arithmetic stays fp64; the pointers below then address float arrays of the same shape)
  int f32;
  double* S;     // [n_int][tiles][q][128]  marginal_subtree_LH
  double* Fpart; // [n_fgroups][ld] per-pattern sums of log-normalisers over the nodes of one postorder block run
  int n_fgroups; //                 (sum over all runs = marginal_subtree_LH_prefactor of the root)
  double* Fred;  // [TTB_FLANES][ld] first reduction stage of Fpart
  double* M;     // [n_int][tiles][q][128]  marginal_profile
  double* Mtip;  // [n_tips][tiles][q][128] marginal_profile of tips (reconstruct_tip_states) or null
  uint8_t* idx;     // [n_int][ld]  argmax state
  uint8_t* idxtip;  // [n_tips][ld] or null
  // joint (max-product) reconstruction, treeanc.py:934-1080 (N2): reuses S for the summed child
  // messages; LP = log(max(1e-12, exp(Qt))), TL / TC = per tip-branch tables of the leaf message and
  // its argmax, Cx = best child state for every parent state (the back-pointers)
  double* LP;       // [n_nodes][pq]
  double* TL;       // [n_tips][tu_stride]
  uint8_t* TC;      // [n_tips][tu_stride]
  uint8_t* Cx;      // [n_int][tiles][q][128]
  double* LH;       // [ld] tree.sequence_LH
  double* lh_partial;             // [tiles]
  unsigned long long* nd_slots;   // [1024]
  double* results;                // {total_lh, n_diff, n_diff of tips}
  unsigned long long* trace;   // measurement only (TTB_TRACE=<file>, ttb_profile_marginal): per-block clock64 timelines of the level kernels, or null
  int dbg;   // measurement only (TTB_DBG): bit0: the tensor-pipe level kernels skip the arithmetic, bit1: they skip the message copies (ttb_mma.cuh)
};

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (global -> shared).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (level kernels are launched with the programmatic-stream-serialization
// attribute): pdl_launch_dependents() lets the next level's blocks become resident as soon as every block of
// this level has started, so that their prologue (barrier init, per-pattern model into registers, descriptor
// prefetch) overlaps this level's tail; pdl_wait() blocks until the previous grid has completed and its
// writes are visible.  Every block that touches node data calls pdl_wait(), hence completion of level l
// implies completion of all earlier levels.  Both are no-ops for an ordinary launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Orders this thread's generic-proxy shared-memory writes before later async-proxy (TMA) writes.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Orders this thread's generic-proxy global-memory writes before later async-proxy (TMA) reads of them (merged-level
// launches: a block's bulk copies re-read message tiles its own pattern threads wrote a few chunks earlier).
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void st_release_cta_shared(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// cp.async.bulk: bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Element (slot, state 0, pattern a) of a tile-blocked message array; consecutive states are
// TTB_TILE doubles apart.
// true if pattern a of the branch above `node` is masked out
__device__ __forceinline__ bool masked_out(const TtbDev& p, int node, long long a) {
  if (!p.mask_id) return false;
  const int m = __ldg(p.mask_id + node);
  return m >= 0 && __ldg(p.masks + (size_t)m * p.ld + a) == 0;
}
// multiplicity(mask=node.mask)[a] (sequence_data.py:302-306) for the branch above `node`; kind 1 = merged root branch:
// mask(n1) * mask(n2) when both children of the root carry one, no mask otherwise (treeanc.py:1326-1333)
__device__ __forceinline__ double branch_weight(const TtbDev& p, int node, int kind, long long a) {
  const double m = p.mult[a];
  if (!p.mask_id) return m;
  if (kind == 1) {
    const int c0 = p.child_ptr[0];
    const int m1 = __ldg(p.mask_id + p.child_idx[c0]), m2 = __ldg(p.mask_id + p.child_idx[c0 + 1]);
    if (m1 < 0 || m2 < 0) return m;
    return (__ldg(p.masks + (size_t)m1 * p.ld + a) && __ldg(p.masks + (size_t)m2 * p.ld + a)) ? m : 0.0;
  }
  return masked_out(p, node, a) ? 0.0 : m;
}

template <int Q>
__device__ __forceinline__ size_t msg_off(const TtbDev& p, int slot, long long a) {
  return ((size_t)slot * p.tiles + (size_t)(a / TTB_TILE)) * (size_t)(Q * TTB_TILE) + (size_t)(a % TTB_TILE);
}

// S / M / Mtip as arrays of their storage type ST (double, or float with ttb_set_message_storage)
template <typename ST>
__device__ __forceinline__ ST* msg_base(double* p) { return reinterpret_cast<ST*>(p); }
template <typename ST>
__device__ __forceinline__ const ST* msg_base(const double* p) { return reinterpret_cast<const ST*>(p); }
// run-time typed read (kernels outside the pass: fetch, branch objective, counts, sampling)
__device__ __forceinline__ double msg_ld(const double* base, size_t off, int f32) {
  return f32 ? (double)reinterpret_cast<const float*>(base)[off] : base[off];
}

// 1/a without the IEEE slow path: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps -> within 1 ulp for normal a;
// 0 -> NaN like the exact quotient's inf * 0 in the reference's log-space form, so a vanished message surfaces the same way.
__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}
// The same with ONE Newton step: relative error ~1e-12 (2^-20 seed squared).  Only where the consumer's tolerance is far
// above that and nothing accumulates: the outside message of the large-alphabet preorder (profiles: 1e-6).
__device__ __forceinline__ double fast_rcp1(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  return fma(r, fma(-a, r, 1.0), r);
}
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
  double ev[Q], e[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) ev[k] = p.v[i * Q + k];
#pragma unroll
  for (int k = 0; k < Q; ++k) e[k] = exp(mt * p.eig[k]);
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) acc = fma(ev[k], e[k] * p.vinv[k * Q + j], acc);
    p.P[(size_t)n * p.pq + i * Q + j] = fmax(0.0, acc);
  }
}

// Tip message tables: TU[row][code][j] = sum_i prof[code][i] * P_tip[i][j]  (seq2prof +
// propagate_profile for a leaf, treeanc.py:846-853 + gtr.py:965-995).  One thread per
// (tip, code, j); terms with prof == 0 are skipped, which is exact.
template <int Q>
__global__ void tip_table_kernel(TtbDev p, const int* __restrict__ tip_nodes) {
  const int per = p.n_codes * Q;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_tips * per) return;
  const int row = (int)(gid / per), r = (int)(gid % per);
  const int code = r / Q, j = r % Q;
  const double* P = p.P + (size_t)tip_nodes[row] * p.pq;
  double u = 0.0;
  for (int i = 0; i < Q; ++i) {
    const double w = p.code_prof[code * Q + i];
    if (w != 0.0) u = fma(w, P[i * Q + j], u);
  }
  p.TU[(size_t)row * p.tu_stride + r] = u;
}

// ---------------------------------------------------------------------------------------
// A9: site-specific GTR.  Reference: GTR_site_specific._expQt / expQt / propagate_profile /
// evolve (gtr_site_specific.py:350-437).  Every pattern a has its own eigen-system, so
//   expQt_a(t) = V_a diag(e_k) Vinv_a,  e_k = exp(lambda_k mu_a t)
// and for t * rate_scale < 10 the reference interpolates the stacked matrices linearly on a
// 61-point t-grid (:331-348,367-371); by linearity that is the same matrix with
//   e_k = e_k(t_lo) + (e_k(t_hi) - e_k(t_lo)) * (t - t_lo)/(t_hi - t_lo).
// The up-message is clamped at 1e-12 (:406), the matrix itself is not clamped.
// Per-thread view: this thread's pattern column of the pattern-contiguous model planes.
// ---------------------------------------------------------------------------------------
// max(lo, x) for lo >= 0 and non-NaN x through the integer order of the bit patterns: there is no fp64
// min/max instruction, fmax() (and `x > lo ? x : lo`, which the compiler turns into it) costs ~10
// instructions for its NaN handling; this is a 64-bit integer compare + select.  Negative x (sign bit set)
// compares below every non-negative lo.
__device__ __forceinline__ double at_least(double x, double lo) {
  const long long xb = __double_as_longlong(x), lb = __double_as_longlong(lo);
  return __longlong_as_double(xb > lb ? xb : lb);
}

template <int Q, bool REG = false, bool SYM = false>
struct SiteModel {
  const double* V;    // V_a[i][k] at V[(i*Q+k)*vs]
  const double* Vi;   // Vinv_a[k][j] at Vi[(k*Q+j)*vs]
  long long vs;       // plane stride of V / Vi: ld in global memory, TTB_TILE for a staged tile
  const double* lam;
  double mu;
  long long ld;
  const double* E;    // this pattern's column of the grid table
  // REG: this pattern's eigen-system held in registers for the whole block (2*Q*Q doubles; Q <= 5).
  // The level kernels handle one pattern per thread for a whole run of nodes, so the 2*Q*Q shared-memory
  // loads per matvec pair of the staged variant (the measured limiter at q = 5) disappear.
  // SYM (with REG): only V_a and r_j = 1/Pi_a[j]; Vinv_a[k][j] = V_a[j][k] c_k r_j with c_k folded into the eigen-factors
  // (grid table ss_Ec; the exact path multiplies by c_k from the planes at cpl).
  double Vr[REG ? Q * Q : 1], Vir[(REG && !SYM) ? Q * Q : 1], rr[(REG && SYM) ? Q : 1];
  const double* cpl;
  __device__ SiteModel() {}
  // model planes read from global memory (fetch / branch kernels)
  __device__ SiteModel(const TtbDev& p, long long a)
      : V(p.ss_V + a), Vi(p.ss_Vinv + a), vs(p.ld), lam(p.ss_eig + a), mu(p.ss_mu[a]), ld(p.ld), E(e_column(p, a)) {}
  // V / Vinv of this block's pattern tile staged in shared memory (level kernels, Q > 5)
  __device__ SiteModel(const TtbDev& p, long long a, const double* smem_model, int tid)
      : V(smem_model + tid), Vi(smem_model + Q * Q * TTB_TILE + tid), vs(TTB_TILE), lam(p.ss_eig + a), mu(p.ss_mu[a]), ld(p.ld),
        E(e_column(p, a)) {}
  // this pattern's column of the tile-blocked grid table: row (g, k) at E[(g*Q + k) * TTB_TILE]
  __device__ static __forceinline__ const double* e_column(const TtbDev& p, long long a) {
    return (SYM ? p.ss_Ec : p.ss_E) + (size_t)(a / TTB_TILE) * ((size_t)p.ss_ngrid * Q * TTB_TILE) + (size_t)(a % TTB_TILE);
  }
  // level kernels: registers (REG) or the staged tile
  // `pre`: preorder kernel of the symmetric form -- rr[] then holds the clamp thresholds TINY * Pi_a[j] instead of 1 / Pi_a[j]
  // (see up_pre / down_pre)
  __device__ __forceinline__ void init_level(const TtbDev& p, long long a, bool act, const double* smem_model, int tid, bool pre = false) {
    lam = p.ss_eig + a; ld = p.ld; E = e_column(p, a);
    mu = act ? p.ss_mu[a] : 0.0;
    if constexpr (REG && SYM) {
      V = Vi = nullptr; vs = 0;
      cpl = p.ss_c + a;
#pragma unroll
      for (int r = 0; r < Q * Q; ++r) Vr[r] = act ? __ldg(p.ss_V + (size_t)r * p.ld + a) : 0.0;
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const double pi = act ? __ldg(p.ss_Pi + (size_t)j * p.ld + a) : 1.0;
        rr[j] = pre ? TTB_TINY * pi : (act ? 1.0 / pi : 0.0);
      }
    } else if constexpr (REG) {
      V = Vi = nullptr; vs = 0;
#pragma unroll
      for (int r = 0; r < Q * Q; ++r) {
        Vr[r] = act ? __ldg(p.ss_V + (size_t)r * p.ld + a) : 0.0;
        Vir[r] = act ? __ldg(p.ss_Vinv + (size_t)r * p.ld + a) : 0.0;
      }
    } else {
      V = smem_model + tid; Vi = smem_model + Q * Q * TTB_TILE + tid; vs = TTB_TILE;
    }
  }
  __device__ __forceinline__ double v(int r) const { if constexpr (REG) return Vr[r]; else return V[(size_t)r * vs]; }
  __device__ __forceinline__ double vi(int r) const { if constexpr (REG) return Vir[r]; else return Vi[(size_t)r * vs]; }
  // exact eigen-factors (t * rate_scale >= 10 or approximate=False): rare, kept out of line so the
  // inlined exp() bodies do not inflate the level kernels' register pressure
  __device__ __noinline__ static void efac_exact(double tmu, const double* lam, long long ld, double* e) {
    for (int k = 0; k < Q; ++k) e[k] = exp(tmu * __ldg(lam + (size_t)k * ld));
  }
  __device__ __noinline__ static void efac_exact_sym(double tmu, const double* lam, const double* c, long long ld, double* e) {
    for (int k = 0; k < Q; ++k) e[k] = __ldg(c + (size_t)k * ld) * exp(tmu * __ldg(lam + (size_t)k * ld));
  }
  __device__ __forceinline__ void efac_at(double t, int lo, double w, double (&e)[Q]) const {
    if (w < 0.0) {
      double ex[Q];   // only this array lives in local memory; e[] stays in registers on the hot path
      efac_exact(t * mu, lam, ld, ex);
#pragma unroll
      for (int k = 0; k < Q; ++k) e[k] = ex[k];
    } else {
      const double* Elo = E + (size_t)lo * Q * TTB_TILE;
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const double elo = __ldg(Elo + k * TTB_TILE), ehi = __ldg(Elo + (Q + k) * TTB_TILE);
        e[k] = elo + (ehi - elo) * w;
      }
    }
  }
  // the same from the two grid rows staged in shared memory (rows = this thread's column of [2][Q][128])
  __device__ __forceinline__ void efac_staged(double t, double w, const double* rows, double (&e)[Q]) const {
    if (w < 0.0) {
      double ex[Q];   // only this array lives in local memory; e[] stays in registers on the hot path
      if constexpr (SYM) efac_exact_sym(t * mu, lam, cpl, ld, ex);
      else efac_exact(t * mu, lam, ld, ex);
#pragma unroll
      for (int k = 0; k < Q; ++k) e[k] = ex[k];
    } else {
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const double elo = rows[k * TTB_TILE], ehi = rows[(Q + k) * TTB_TILE];
        e[k] = elo + (ehi - elo) * w;
      }
    }
  }
  __device__ __forceinline__ void efac(const TtbDev& p, int node, double (&e)[Q]) const {
    efac_at(p.t[node], p.ss_lo[node], p.ss_w[node], e);
  }
  // child -> parent: U[j] = max(1e-12, sum_i S[i] P[i][j])
  __device__ __forceinline__ void up(const double (&S)[Q], const double (&e)[Q], double (&U)[Q], bool clamp = true) const {
    double wk[Q];
    if constexpr (REG && SYM) {   // U[j] = r_j sum_k V[j][k] (c_k e_k) sum_i S[i] V[i][k]
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < Q; ++i) acc = fma(S[i], Vr[i * Q + k], acc);
        wk[k] = acc * e[k];
      }
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) acc = fma(wk[k], Vr[j * Q + k], acc);
        acc *= rr[j];
        U[j] = clamp ? at_least(acc, TTB_TINY) : acc;
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) acc = fma(S[i], v(i * Q + k), acc);
      wk[k] = acc * e[k];
    }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) acc = fma(wk[k], vi(k * Q + j), acc);
      U[j] = clamp ? at_least(acc, TTB_TINY) : acc;
    }
  }
  // Symmetric form in the PREORDER: the factors r_j = 1 / Pi_a[j] cancel.  The up-message is U[j] = r_j u[j] with
  // u = V (c e * (V^T S)), clamped: max(TINY, r_j u[j]) = r_j max(TINY Pi_j, u[j]); the outside message enters the way down
  // as y[j] = r_j O[j] with O[j] = Mp[j] prod_{k != j} U[k], so y[j] = (prod_k r_k) Mp[j] prod_{k != j} max(TINY Pi_k, u[k]) --
  // the constant prod_k r_k drops out in the final normalisation of the profile.  up_pre returns the clamped u (rr[] holds
  // TINY Pi_j, see init_level), down_pre takes y: 2 Q multiplications fewer per child and pattern than up + down.
  __device__ __forceinline__ void up_pre(const double (&S)[Q], const double (&e)[Q], double (&U)[Q]) const {
    static_assert(REG && SYM, "symmetric register-resident form only");
    double wk[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) acc = fma(S[i], Vr[i * Q + k], acc);
      wk[k] = acc * e[k];
    }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) acc = fma(wk[k], Vr[j * Q + k], acc);
      U[j] = at_least(acc, rr[j]);
    }
  }
  __device__ __forceinline__ void down_pre(const double (&y)[Q], const double (&e)[Q], double (&msg)[Q]) const {
    static_assert(REG && SYM, "symmetric register-resident form only");
    double wk[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < Q; ++j) acc = fma(Vr[j * Q + k], y[j], acc);
      wk[k] = acc * e[k];
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) acc = fma(Vr[i * Q + k], wk[k], acc);
      msg[i] = acc;
    }
  }
  // parent -> child: msg[i] = sum_j P[i][j] O[j]
  __device__ __forceinline__ void down(const double (&O)[Q], const double (&e)[Q], double (&msg)[Q]) const {
    double wk[Q];
    if constexpr (REG && SYM) {   // msg[i] = sum_k V[i][k] (c_k e_k) sum_j V[j][k] r_j O[j]
      double y[Q];
#pragma unroll
      for (int j = 0; j < Q; ++j) y[j] = rr[j] * O[j];
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < Q; ++j) acc = fma(Vr[j * Q + k], y[j], acc);
        wk[k] = acc * e[k];
      }
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) acc = fma(Vr[i * Q + k], wk[k], acc);
        msg[i] = acc;
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < Q; ++j) acc = fma(vi(k * Q + j), O[j], acc);
      wk[k] = acc * e[k];
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) acc = fma(v(i * Q + k), wk[k], acc);
      msg[i] = acc;
    }
  }
};
// Large alphabets: ptxas takes 192 registers for the postorder kernel when left alone (one block per SM: ncu shows 5
// resident warps, issue slots 27 %); it needs only 126 without spilling, which lets three blocks share an SM.
#ifndef TTB_POST_LARGEQ_REGS
#define TTB_POST_LARGEQ_REGS 128
#endif
// site-specific level kernels keep the eigen-system in registers up to this alphabet size
#define TTB_SS_REG_MAXQ 5
// symmetric variant: register cap and ring depth chosen so that three blocks (12 pattern warps) fit one SM
#ifndef TTB_SS_SYM_REGS
#define TTB_SS_SYM_REGS 128
#endif
#ifndef TTB_SS_SYM_STAGES
#define TTB_SS_SYM_STAGES 2
#endif
// ... and then the level kernels also stage, per child, the two grid rows of its branch (2q KB per tile, in the
// stage's P area) and its {w, t} record (16 bytes, in the TU area): no global load is left on the consumers' path.
template <int Q, bool SS>
__host__ __device__ constexpr bool ss_staged() { return SS && Q <= TTB_SS_REG_MAXQ; }
template <int Q, bool SS>
__host__ __device__ inline int stage_pq(int pq) { return ss_staged<Q, SS>() ? 2 * Q * TTB_TILE : pq; }
template <int Q, bool SS>
__host__ __device__ inline int stage_tu(int tu_stride) { return ss_staged<Q, SS>() ? 2 : tu_stride; }

// ---------------------------------------------------------------------------------------
// N2: joint ML reconstruction.  Reference: TreeAnc._ml_anc_joint, treeanc.py:934-1080.
//   log_transitions = log(max(1e-12, expQt(t)))                         (:967)
//   leaf message    = log(max(profile, 1e-12))                          (:968-976)
//   Lx_c[j] = max_i (log_transitions_c[i][j] + msg_c[i]),  Cx_c[j] = argmax_i   (:986-1000)
//   msg_n[i] = sum_c Lx_c[i]                                            (:978)
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void joint_tables_kernel(TtbDev p, const int* __restrict__ tip_nodes) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nlp = (long long)p.n_nodes * Q * Q;
  if (gid < nlp) {
    const int n = (int)(gid / (Q * Q)), k = (int)(gid % (Q * Q));
    p.LP[(size_t)n * p.pq + k] = log(fmax(TTB_TINY, p.P[(size_t)n * p.pq + k]));
  }
  const int per = p.n_codes * Q;
  if (gid < (long long)p.n_tips * per) {
    const int row = (int)(gid / per), r = (int)(gid % per);
    const int code = r / Q, j = r % Q;
    const double* P = p.P + (size_t)tip_nodes[row] * p.pq;
    double best = -1e300;
    int arg = 0;
    for (int i = 0; i < Q; ++i) {
      const double v = log(fmax(TTB_TINY, P[i * Q + j])) + log(fmax(p.code_prof[code * Q + i], TTB_TINY));
      if (v > best) { best = v; arg = i; }
    }
    p.TL[(size_t)row * p.tu_stride + r] = best;
    p.TC[(size_t)row * p.tu_stride + r] = (uint8_t)arg;
  }
}

// Root of the joint pass (treeanc.py:1003-1023): Lx_r = msg_r + log Pi; state = first argmax of
// exp(Lx_r - max) (what prof2seq sees); sequence_LH = Lx_r[state].
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) joint_root_kernel(TtbDev p) {
  __shared__ double sred[TTB_BLOCK / 32];
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  double contrib = 0.0;
  if (a < p.Lp) {
    const int slot = p.int_slot[0];
    double* __restrict__ s = p.S + msg_off<Q>(p, slot, a);
    double R[Q];
    double mx = -1e300;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      R[i] = s[i * TTB_TILE] + log(p.Pi[i]);
      mx = fmax(mx, R[i]);
    }
    int best = 0;
    double bv = -1.0;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      s[i * TTB_TILE] = R[i];   // root.joint_Lx (fetchable for root sampling on the host)
      const double e = exp(R[i] - mx);
      if (e > bv) { bv = e; best = i; }
    }
    p.idx[(size_t)slot * p.ld + a] = (uint8_t)best;
    const double lh = R[best];
    p.LH[a] = lh;
    contrib = lh * p.mult[a];
  }
  const double bs = block_sum<TTB_BLOCK>(contrib, sred);
  if (threadIdx.x == 0) p.lh_partial[blockIdx.x] = bs;
}

// Root of the joint pass with a caller-chosen root state (sample_from_profile='root',
// treeanc.py:1008-1023: the root is sampled on the host from exp(joint_Lx - max)).
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) joint_root_override_kernel(TtbDev p, const uint8_t* __restrict__ root_idx) {
  __shared__ double sred[TTB_BLOCK / 32];
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  double contrib = 0.0;
  if (a < p.Lp) {
    const int slot = p.int_slot[0];
    const int st = root_idx[a];
    const double lh = p.S[msg_off<Q>(p, slot, a) + (size_t)st * TTB_TILE];   // joint_Lx of the root
    p.idx[(size_t)slot * p.ld + a] = (uint8_t)st;
    p.LH[a] = lh;
    contrib = lh * p.mult[a];
  }
  const double bs = block_sum<TTB_BLOCK>(contrib, sred);
  if (threadIdx.x == 0) p.lh_partial[blockIdx.x] = bs;
}

// Backtrace of one depth level (treeanc.py:1034-1048): state_c = Cx_c[state_parent]; tips read the
// per-branch table instead.  One thread per (node, pattern).
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) joint_pre_level_kernel(TtbDev p, const int* __restrict__ nodes, int tiles, int count_diff) {
  const int n = nodes[blockIdx.x / tiles];
  const long long a = (long long)(blockIdx.x % tiles) * TTB_TILE + threadIdx.x;
  unsigned int nd = 0;
  if (a < p.Lp) {
    const int ps = p.idx[(size_t)p.int_slot[p.parent[n]] * p.ld + a];
    const int row = p.tip_row[n];
    int st;
    uint8_t* ip;
    if (row >= 0) {
      const int code = p.codes[(size_t)row * p.ld + a];
      st = p.TC[(size_t)row * p.tu_stride + code * Q + ps];
      ip = p.idxtip + (size_t)row * p.ld + a;
    } else {
      const int slot = p.int_slot[n];
      st = p.Cx[((size_t)slot * p.tiles + (size_t)(a / TTB_TILE)) * (size_t)(Q * TTB_TILE) + (size_t)ps * TTB_TILE + (size_t)(a % TTB_TILE)];
      ip = p.idx + (size_t)slot * p.ld + a;
    }
    if (count_diff) nd = (*ip != (uint8_t)st);
    *ip = (uint8_t)st;
  }
  if (count_diff) {
    nd = __reduce_add_sync(0xffffffffu, nd);   // a block handles one node: all tips or all internal
    if ((threadIdx.x & 31) == 0 && nd)
      atomicAdd(p.nd_slots + (p.tip_row[n] >= 0 ? 512 : 0) + (blockIdx.x & 511), (unsigned long long)nd);
  }
}

// E[g][k][a] = exp(t_g * mu_a * lambda_k(a)): the eigen-factor of gtr_site_specific._expQt (:363) on the
// interpolation grid (:336-344).  One thread per (g, k, a).
static __global__ void ss_grid_table_kernel(TtbDev p, double* __restrict__ E, const double* __restrict__ c = nullptr) {
  const long long n = (long long)p.tiles * p.ss_ngrid * p.q * TTB_TILE;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(i % TTB_TILE);
    const int k = (int)((i / TTB_TILE) % p.q), g = (int)((i / ((long long)TTB_TILE * p.q)) % p.ss_ngrid);
    const long long a = (i / ((long long)TTB_TILE * p.q * p.ss_ngrid)) * TTB_TILE + lane;
    E[i] = (a < p.Lp) ? (c ? c[(size_t)k * p.ld + a] : 1.0) * exp(p.ss_grid[g] * p.ss_mu[a] * p.ss_eig[(size_t)k * p.ld + a]) : 1.0;
  }
}

// Site-specific models: write every child's interpolation bracket into its chunk descriptor, so the producer
// warp knows which grid rows to stage without a dependent global load (run when branch lengths change).
static __global__ void ss_patch_chunks_kernel(TtbChunk* __restrict__ chunks, int n_chunks, const int* __restrict__ ss_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int nch = chunks[i].flags >> 8;
  for (int b = 0; b < TTB_CB; ++b) chunks[i].lo[b] = b < nch ? ss_lo[chunks[i].cnode[b]] : 0;
}

// ---------------------------------------------------------------------------------------
// Shared-memory ring of the level kernels.
// ---------------------------------------------------------------------------------------
template <int Q, int NST = ((Q <= 8) ? 3 : 2), int NCONS = TTB_BLOCK>
struct Pipe {
  static constexpr int STAGES = NST;
  static constexpr int CB = (Q <= 8) ? TTB_CB : 1;   // children per chunk (the host schedule uses the same rule)
  // per-stage byte offsets (all multiples of 16)
  int off_P, off_TU, off_codes, off_oidx, off_desc, stage_bytes;
  unsigned char* base;
  uint64_t* full;   // [STAGES] producer -> consumers (transaction barrier)
  uint64_t* empty;  // [STAGES] consumers -> producer (one arrival per thread)
  uint64_t* mbar;   // site-specific models: "model tile loaded" barrier
  uint32_t* done;   // [TTB_BLOCK / 32] merged-level launches: chunks completed by every pattern warp (see wait_done)
  double* model;    // site-specific models: V and Vinv planes of this block's pattern tile [2*Q*Q][TILE]
  __host__ __device__ static int stage_size(int rows, int pq, int tu_stride) {
    return rows * TTB_TILE * 8 + CB * pq * 8 + CB * tu_stride * 8 + 2 * CB * TTB_TILE + 32;
  }
  __host__ static size_t smem_bytes(int rows, int pq, int tu_stride, bool site_specific = false) {
    return 128 + (size_t)STAGES * stage_size(rows, pq, tu_stride) + ((site_specific && Q > TTB_SS_REG_MAXQ) ? (size_t)2 * Q * Q * TTB_TILE * 8 : 0);
  }
  __device__ Pipe(unsigned char* smem, int rows, int pq, int tu_stride) {
    full = reinterpret_cast<uint64_t*>(smem);
    empty = full + STAGES;
    mbar = empty + STAGES;
    done = reinterpret_cast<uint32_t*>(smem + 96);
    base = smem + 128;
    off_P = rows * TTB_TILE * 8;
    off_TU = off_P + CB * pq * 8;
    off_codes = off_TU + CB * tu_stride * 8;
    off_oidx = off_codes + CB * TTB_TILE;
    off_desc = off_oidx + CB * TTB_TILE;
    stage_bytes = off_desc + 32;
    model = reinterpret_cast<double*>(base + (size_t)STAGES * stage_bytes);
  }
  // Stage the per-pattern eigen-systems of this tile once per block (warp 0), wait with wait_model().
  __device__ void load_model(const TtbDev& p, long long a0, int cols, int lane) const {
    if (lane == 0) mbar_arrive_expect_tx(mbar, (uint32_t)(2 * Q * Q * cols * 8));
    __syncwarp();
    for (int r = lane; r < 2 * Q * Q; r += 32) {
      const double* src = (r < Q * Q ? p.ss_V + (size_t)r * p.ld : p.ss_Vinv + (size_t)(r - Q * Q) * p.ld) + a0;
      tma_load_1d(model + (size_t)r * TTB_TILE, src, cols * 8, mbar);
    }
  }
  __device__ void wait_model() const { mbar_wait(mbar, 0); }
  __device__ void init() const {  // one thread
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NCONS);   // one arrival per consumer thread
    }
    mbar_init(mbar, 1);
    for (int w = 0; w < TTB_BLOCK / 32; ++w) done[w] = 0u;
    mbar_fence_init();
  }
  // Merged-level launches (several consecutive small levels in ONE launch, a block walks them all for its pattern tile):
  // a chunk may read message tiles that this block's own pattern threads wrote a few chunks earlier.  Patterns are
  // independent, so the dependency never leaves the block: every pattern warp publishes the number of chunks it has
  // completed (its global writes ordered before by fence.proxy.async + release), and the producer warp waits until all
  // warps have passed the chunk that wrote the tile before it issues the bulk copy that reads it.
  __device__ __forceinline__ void publish_done(int warp, int lane, uint32_t n_completed) const {
    fence_proxy_async_global();
    __syncwarp();
    if (lane == 0) st_release_cta_shared(done + warp, n_completed);
  }
  __device__ __forceinline__ void wait_done(int lane, uint32_t need) const {
    if (lane < TTB_BLOCK / 32) {
      while (ld_acquire_cta_shared(done + lane) < need) {
      }
    }
    __syncwarp();
    fence_proxy_async_global();
  }
  __device__ double* rows(int s) const { return reinterpret_cast<double*>(base + (size_t)s * stage_bytes); }
  __device__ double* P(int s) const { return reinterpret_cast<double*>(base + (size_t)s * stage_bytes + off_P); }
  __device__ double* TU(int s) const { return reinterpret_cast<double*>(base + (size_t)s * stage_bytes + off_TU); }
  __device__ uint8_t* codes(int s) const { return base + (size_t)s * stage_bytes + off_codes; }
  __device__ uint8_t* oidx(int s) const { return base + (size_t)s * stage_bytes + off_oidx; }
  __device__ const int4* desc(int s) const { return reinterpret_cast<const int4*>(base + (size_t)s * stage_bytes + off_desc); }
  // Producer side: the stage used by chunk number u (0-based within the block) is free once all
  // consumer warps released its previous use.
  // Ring cursor: stage s of round r (phase = r & 1); advance() walks it without a modulo per chunk.
  struct Cursor {
    int s = 0, phase = 0;
    __device__ __forceinline__ void advance() {
      if (++s == STAGES) { s = 0; phase ^= 1; }
    }
  };
  __device__ void producer_acquire(const Cursor& c, int u) const {
    if (u >= STAGES) mbar_wait(empty + c.s, c.phase ^ 1);
  }
  __device__ void consumer_wait(const Cursor& c) const { mbar_wait(full + c.s, c.phase); }
  __device__ void consumer_release(const Cursor& c) const {
    // every thread releases for itself: same speed as a warp-elected arrive (measured) and
    // compute-sanitizer racecheck can follow it
    mbar_arrive(empty + c.s);
  }
};

// Chunk descriptor held in registers as scalars (no local-memory arrays).
struct Chunk {
  int out, flags, src0, src1, cnode0, cnode1, lo0, lo1;
  __device__ __forceinline__ int nch() const { return flags >> 8; }
  __device__ __forceinline__ int lo(int b) const { return b ? lo1 : lo0; }
  // staged grid rows of child b: children of one chunk with the same bracket share one copy
  __device__ __forceinline__ int erow(int b) const { return (b && lo1 != lo0) ? 1 : 0; }
  __device__ __forceinline__ int src(int b) const { return b ? src1 : src0; }
  __device__ __forceinline__ int cnode(int b) const { return b ? cnode1 : cnode0; }
};
__device__ __forceinline__ Chunk chunk_from(const int4 a, const int4 b) {
  Chunk c;
  c.out = a.x; c.flags = a.y; c.src0 = a.z; c.src1 = a.w; c.cnode0 = b.x; c.cnode1 = b.y; c.lo0 = b.z; c.lo1 = b.w;
  return c;
}
__device__ __forceinline__ Chunk load_chunk_global(const TtbChunk* __restrict__ c) {
  const int4* q = reinterpret_cast<const int4*>(c);
  return chunk_from(__ldg(q), __ldg(q + 1));
}
__device__ __forceinline__ Chunk load_chunk_smem(const int4* q) { return chunk_from(q[0], q[1]); }

// ---------------------------------------------------------------------------------------
// A3-A5: one postorder level.  Reference: postorder_traversal_marginal, treeanc.py:857-878 +
// normalize_profile(log=True), seq_utils.py:279-307.
//   X[j] = prod_c U_c[j],  U_c[j] = sum_i S_c[i] P_c[i][j]  (tips: table lookup)
//   Z = sum_j X[j];  S_n = X/Z;  F_n = sum_c F_c + log Z.
// Nodes with many children are rescaled by exact powers of two so the product cannot underflow.
// Block = (run of nodes of the level given by group_ptr, one 128-pattern tile).
// Stage rows: child b -> rows [b*(Q+1), b*(Q+1)+Q) = S_c, row b*(Q+1)+Q = F_c.
// ---------------------------------------------------------------------------------------
template <int Q, bool SS, bool JOINT = false, bool SYM = false, bool MASK = false, typename ST = double, bool DEP = false>
__global__ void __launch_bounds__(TTB_LEVEL_THREADS) __maxnreg__((SS && Q <= TTB_SS_REG_MAXQ) ? (SYM ? TTB_SS_SYM_REGS : 168) : (Q > 8 ? (JOINT ? 168 : TTB_POST_LARGEQ_REGS) : 255)) post_level_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                              const int* __restrict__ group_ptr, int tiles, int fbase,
                                                              const int* __restrict__ dep) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int RPC = Q;  // rows per child (the log-prefactors never travel: see Fpart)
  constexpr bool EST = ss_staged<Q, SS>();   // site-specific, grid rows + branch record staged per child
  constexpr int EROWS = 2 * Q * TTB_TILE;
  static_assert(!SYM || (SS && Q <= TTB_SS_REG_MAXQ && !JOINT), "SYM is a variant of the register-resident site-specific kernels");
  static_assert(sizeof(ST) == 8 || !JOINT, "the joint pass keeps its log-space sums in double");
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * sizeof(ST);   // one (node, tile) message block
  using PipeT = Pipe<Q, SYM ? TTB_SS_SYM_STAGES : Pipe<Q>::STAGES>;
  PipeT pipe(smem_raw, PipeT::CB * RPC, stage_pq<Q, SS>(p.pq), stage_tu<Q, SS>(p.tu_stride));
  const int g = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int k0 = group_ptr[g], k1 = group_ptr[g + 1];
  const int n_chunks = k1 - k0;
  const long long a0 = (long long)tile * TTB_TILE;
  const int cols = (int)min((long long)TTB_TILE, p.ld - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long a = a0 + tid;
  const bool act = a < p.Lp;
  pdl_launch_dependents();
  if (tid == 0) pipe.init();
  __syncthreads();

  typename PipeT::Cursor cur;
  auto issue = [&](int u, const Chunk& c) {  // executed by the producer warp: fill the stage of chunk number u
    const int s = cur.s;
    pipe.producer_acquire(cur, u);
    uint64_t* bar = pipe.full + s;
    const int nch = c.nch();
    uint32_t bytes = 32;
    for (int b = 0; b < nch; ++b) {
      if (SS) {  // site-specific: the transition matrices are per pattern; per branch only the grid rows + record
        bytes += (c.src(b) >= 0) ? MSG_BYTES : (uint32_t)cols;
        if (EST) bytes += 16u + ((b == 0 || c.lo1 != c.lo0) ? (uint32_t)(EROWS * 8) : 0u);
      } else
        bytes += (c.src(b) >= 0) ? (uint32_t)(MSG_BYTES + p.pq * 8) : (uint32_t)(cols + p.tu_stride * 8);
    }
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      tma_load_1d((void*)pipe.desc(s), chunks + k0 + u, 32, bar);
    }
    __syncwarp();
    for (int job = lane; job < nch * 3; job += 32) {
      const int b = job / 3, r = job % 3;
      const int src = c.src(b);
      if (r == 2) {
        if (EST) {
          tma_load_1d(pipe.TU(s) + b * 2, p.ss_rec + c.cnode(b), 16, bar);
          if (b == 0 || c.lo1 != c.lo0)
            tma_load_1d(pipe.P(s) + c.erow(b) * EROWS, (SYM ? p.ss_Ec : p.ss_E) + ((size_t)tile * p.ss_ngrid + c.lo(b)) * (Q * TTB_TILE), EROWS * 8, bar);
        }
      } else if (src >= 0) {
        if (r == 0)   // the child's q rows of this tile are one contiguous block
          tma_load_1d(reinterpret_cast<ST*>(pipe.rows(s)) + (b * RPC) * TTB_TILE, msg_base<ST>(p.S) + msg_off<Q>(p, src, a0), MSG_BYTES, bar);
        else if (r == 1 && !SS)
          tma_load_1d(pipe.P(s) + b * p.pq, (JOINT ? p.LP : p.P) + (size_t)c.cnode(b) * p.pq, p.pq * 8, bar);
      } else {
        const int row = -1 - src;
        if (r == 0)
          tma_load_1d(pipe.codes(s) + b * TTB_TILE, p.codes + (size_t)row * p.ld + a0, cols, bar);
        else if (r == 1 && !SS)
          tma_load_1d(pipe.TU(s) + b * p.tu_stride, (JOINT ? p.TL : p.TU) + (size_t)row * p.tu_stride, p.tu_stride * 8, bar);
      }
    }
  };

  constexpr bool MREG = Q <= TTB_SS_REG_MAXQ;
  if (SS && !MREG && warp == 0) pipe.load_model(p, a0, cols, lane);
  if (warp == TTB_BLOCK / 32) {
    // Producer warp: runs ahead of the pattern threads by up to STAGES chunks (bounded by the empty
    // barriers); the descriptor of the next chunk is fetched while the current one is issued, so no
    // global-memory latency sits on the consumers' path.
    Chunk c = load_chunk_global(chunks + k0);
    int d = DEP ? __ldg(dep + k0) : -1;   // merged-level launch (DEP): the chunk (global index) that wrote what this one reads
    pdl_wait();   // everything the bulk copies read was written by earlier levels
    for (int u = 0; u < n_chunks; ++u) {
      const Chunk cn = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
      const int dn = DEP ? __ldg(dep + k0 + min(u + 1, n_chunks - 1)) : -1;
      if (DEP && d >= k0) pipe.wait_done(lane, (uint32_t)(d - k0 + 1));
      issue(u, c);
      cur.advance();
      c = cn;
      d = dn;
    }
    return;
  }
  if (SS && !MREG) pipe.wait_model();
  SiteModel<Q, MREG, SYM> sm;
  if constexpr (SS) sm.init_level(p, a, act, pipe.model, tid);
  pdl_wait();

  double X[Q];
  // sum of log-normalisers of this block's nodes (for this thread's pattern), kept as Facc + log(Zprod):
  // the normalisers are multiplied up and one log is taken whenever the product leaves [1e-150, 1e150]
  double Facc = 0.0, Zprod = 1.0;
  int scale = 0, seen = 0;
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    pipe.consumer_wait(cur);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    if (c.flags & 1) {
#pragma unroll
      for (int j = 0; j < Q; ++j) X[j] = JOINT ? 0.0 : 1.0;
      scale = 0;
      seen = 0;
    }
    const int nch = c.nch();
    if (act) {
#pragma unroll
      for (int b = 0; b < Pipe<Q>::CB; ++b) {
        if (b >= nch) break;
        double U[Q];
        if constexpr (SS) {
          double sc[Q], e[Q];
          if (c.src(b) < 0) {
            const int code = pipe.codes(s)[b * TTB_TILE + tid];
#pragma unroll
            for (int i = 0; i < Q; ++i) sc[i] = __ldg(p.code_prof + code * Q + i);
          } else {
            const ST* rows = reinterpret_cast<const ST*>(pipe.rows(s)) + (b * RPC) * TTB_TILE + tid;
#pragma unroll
            for (int i = 0; i < Q; ++i) sc[i] = (double)rows[i * TTB_TILE];
            }
          if constexpr (EST) {
            const double2 rec = reinterpret_cast<const double2*>(pipe.TU(s))[b];
            sm.efac_staged(rec.y, rec.x, pipe.P(s) + c.erow(b) * EROWS + tid, e);
          } else
            sm.efac(p, c.cnode(b), e);
          sm.up(sc, e, U);
        } else if (c.src(b) < 0) {
          const int code = pipe.codes(s)[b * TTB_TILE + tid];
          const double* tu = pipe.TU(s) + b * p.tu_stride + code * Q;
#pragma unroll
          for (int j = 0; j < Q; ++j) U[j] = tu[j];
        } else {
          const ST* rows = reinterpret_cast<const ST*>(pipe.rows(s)) + (b * RPC) * TTB_TILE + tid;
          const double* Pc = pipe.P(s) + b * p.pq;
          double sc[Q];
#pragma unroll
          for (int i = 0; i < Q; ++i) sc[i] = (double)rows[i * TTB_TILE];
          if constexpr (JOINT) {
            // max-plus "matvec" with back-pointers: Lx_c[j] = max_i (logP[i][j] + msg_c[i]), first maximum
            uint8_t* cx = p.Cx + ((size_t)c.src(b) * p.tiles + (size_t)tile) * (size_t)(Q * TTB_TILE) + tid;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
              double best = Pc[j] + sc[0];
              int arg = 0;
#pragma unroll
              for (int i = 1; i < Q; ++i) {
                const double v = Pc[i * Q + j] + sc[i];
                if (v > best) { best = v; arg = i; }
              }
              U[j] = best;
              cx[j * TTB_TILE] = (uint8_t)arg;
            }
          } else {
#pragma unroll
            for (int j = 0; j < Q; ++j) U[j] = sc[0] * Pc[j];
#pragma unroll
            for (int i = 1; i < Q; ++i)
#pragma unroll
              for (int j = 0; j < Q; ++j) U[j] = fma(sc[i], Pc[i * Q + j], U[j]);
          }
        }
        if constexpr (JOINT) {
#pragma unroll
          for (int j = 0; j < Q; ++j) X[j] += U[j];
          continue;
        }
        if constexpr (MASK) {
          if (masked_out(p, c.cnode(b), a)) continue;   // log_Lx * 0: the child says nothing about this pattern
        }
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
    pipe.consumer_release(cur);  // all smem reads of this stage are done
    cur.advance();
    if (JOINT) {
      if (act && (c.flags & 2)) {
        ST* __restrict__ so = msg_base<ST>(p.S) + msg_off<Q>(p, c.out, a);
#pragma unroll
        for (int j = 0; j < Q; ++j) so[j * TTB_TILE] = (ST)X[j];
      }
    } else if (act && (c.flags & 2)) {
      double Z = X[0];
#pragma unroll
      for (int j = 1; j < Q; ++j) Z += X[j];
      const double inv = fast_rcp(Z);
      ST* __restrict__ so = msg_base<ST>(p.S) + msg_off<Q>(p, c.out, a);
#pragma unroll
      for (int j = 0; j < Q; ++j) so[j * TTB_TILE] = (ST)(X[j] * inv);
      if (scale) Facc -= scale * (256.0 * 0.693147180559945309417232121458);
      if (Z < 1e-150 || Z > 1e150) {
        Facc += log(Z);
      } else {
        Zprod *= Z;
        if (Zprod < 1e-150 || Zprod > 1e150) {
          Facc += log(Zprod);
          Zprod = 1.0;
        }
      }
    }
    if constexpr (DEP) pipe.publish_done(warp, lane, (uint32_t)(u + 1));
  }
  if (!JOINT && act) p.Fpart[(size_t)(fbase + g) * p.ld + a] = Facc + log(Zprod);
}

// Cherry tables for postorder level 1.  The subtree profile of a node whose two children are tips depends on the
// pattern only through the pair of tip characters (c0, c1):
//   S[c0][c1][j] = TU_0[c0][j] TU_1[c1][j] / Z,   Z = sum_j TU_0[c0][j] TU_1[c1][j]
// -- n_codes^2 entries per node (324 for nucleotides) against tens of thousands of patterns.  One block per level-1
// chunk forms the table once per pass (same operations as the per-pattern path, so S is bit-identical; log Z is taken
// per entry); the level kernel then only looks entries up.  Entry = TTB_PAIR_STRIDE(Q) doubles: S[0..Q), log Z.
// Chunks that are not a complete two-tip node (polytomies of tips) keep the per-pattern path.
#define TTB_PAIR_STRIDE(Q) (((Q) + 2) / 2 * 2)
template <int Q>
__global__ void leaf_pair_table_kernel(TtbDev p, const TtbChunk* __restrict__ chunks, double* __restrict__ table) {
  const Chunk c = load_chunk_global(chunks + blockIdx.x);
  if ((c.flags & 3) != 3 || c.nch() != 2) return;
  const int nc = p.n_codes;
  const double* t0 = p.TU + (size_t)(-1 - c.src0) * p.tu_stride;
  const double* t1 = p.TU + (size_t)(-1 - c.src1) * p.tu_stride;
  double* out = table + (size_t)blockIdx.x * nc * nc * TTB_PAIR_STRIDE(Q);
  for (int e = threadIdx.x; e < nc * nc; e += blockDim.x) {
    const int c0 = e / nc, c1 = e % nc;
    double X[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) X[j] = 1.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) X[j] *= __ldg(t0 + c0 * Q + j);
#pragma unroll
    for (int j = 0; j < Q; ++j) X[j] *= __ldg(t1 + c1 * Q + j);
    double Z = X[0];
#pragma unroll
    for (int j = 1; j < Q; ++j) Z += X[j];
    const double inv = fast_rcp(Z);
#pragma unroll
    for (int j = 0; j < Q; ++j) out[(size_t)e * TTB_PAIR_STRIDE(Q) + j] = X[j] * inv;
    out[(size_t)e * TTB_PAIR_STRIDE(Q) + Q] = log(Z);
  }
}

// Postorder level 1: every child is a tip, so there is nothing to stream in but one code byte
// per (tip, pattern); the kernel is a pure write stream of q doubles per (node, pattern).
// Block = (run of nodes, tile); one thread per pattern; two-tip nodes through the cherry tables above
// (pair_table != null), everything else from the tip tables through L1.
template <int Q, bool JOINT = false, typename ST = double>
__global__ void __launch_bounds__(TTB_BLOCK) post_leaf_level_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                                   const int* __restrict__ group_ptr, int tiles, int fbase,
                                                                   const double* __restrict__ pair_table) {
  const int g = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const long long a = (long long)tile * TTB_TILE + threadIdx.x;
  pdl_launch_dependents();
  if (a >= p.Lp) return;
  const int k0 = group_ptr[g], k1 = group_ptr[g + 1];
  pdl_wait();   // the tip tables come from the preceding kernel
  double X[Q];
  double Facc = 0.0, Zprod = 1.0;   // log-normalisers as Facc + log(Zprod), see post_level_kernel
  int scale = 0, seen = 0;
  // Software pipeline: the descriptor and the code bytes of the NEXT chunk are requested before the
  // current one is processed -- the code byte comes from DRAM and everything else depends on it.
  Chunk c = load_chunk_global(chunks + k0);
  int code[Pipe<Q>::CB];
#pragma unroll
  for (int b = 0; b < Pipe<Q>::CB; ++b) code[b] = b < c.nch() ? __ldg(p.codes + (size_t)(-1 - c.src(b)) * p.ld + a) : 0;
  for (int k = k0; k < k1; ++k) {
    Chunk cn = c;
    int ncode[Pipe<Q>::CB];
#pragma unroll
    for (int b = 0; b < Pipe<Q>::CB; ++b) ncode[b] = 0;
    if (k + 1 < k1) {
      cn = load_chunk_global(chunks + k + 1);
#pragma unroll
      for (int b = 0; b < Pipe<Q>::CB; ++b)
        if (b < cn.nch()) ncode[b] = __ldg(p.codes + (size_t)(-1 - cn.src(b)) * p.ld + a);
    }
    const int nch = c.nch();
    if constexpr (!JOINT && Q <= 8) {
      if (pair_table && (c.flags & 3) == 3 && nch == 2) {     // a complete two-tip node: look the result up
        const double* e = pair_table + ((size_t)k * p.n_codes * p.n_codes + (size_t)code[0] * p.n_codes + code[1]) * TTB_PAIR_STRIDE(Q);
        ST* __restrict__ so = msg_base<ST>(p.S) + msg_off<Q>(p, c.out, a);
        double v[TTB_PAIR_STRIDE(Q)];
#pragma unroll
        for (int j = 0; j < TTB_PAIR_STRIDE(Q); j += 2) {       // 16-byte loads: entries are 16-byte aligned
          const double2 t = __ldg(reinterpret_cast<const double2*>(e + j));
          v[j] = t.x; v[j + 1] = t.y;
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) so[j * TTB_TILE] = (ST)v[j];
        Facc += v[Q];
        c = cn;
#pragma unroll
        for (int b = 0; b < Pipe<Q>::CB; ++b) code[b] = ncode[b];
        continue;
      }
    }
    if (c.flags & 1) {
#pragma unroll
      for (int j = 0; j < Q; ++j) X[j] = JOINT ? 0.0 : 1.0;
      scale = 0;
      seen = 0;
    }
#pragma unroll
    for (int b = 0; b < Pipe<Q>::CB; ++b) {
      if (b >= nch) break;
      const int row = -1 - c.src(b);
      const double* tu = (JOINT ? p.TL : p.TU) + (size_t)row * p.tu_stride + code[b] * Q;
      if constexpr (JOINT) {
#pragma unroll
        for (int j = 0; j < Q; ++j) X[j] += __ldg(tu + j);
        continue;
      }
      if (!JOINT && masked_out(p, c.cnode(b), a)) continue;
#pragma unroll
      for (int j = 0; j < Q; ++j) X[j] *= __ldg(tu + j);
      if (++seen > 2) {
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
    if (JOINT) {
      if (c.flags & 2) {
        ST* __restrict__ so = msg_base<ST>(p.S) + msg_off<Q>(p, c.out, a);
#pragma unroll
        for (int j = 0; j < Q; ++j) so[j * TTB_TILE] = (ST)X[j];
      }
    } else if (c.flags & 2) {
      double Z = X[0];
#pragma unroll
      for (int j = 1; j < Q; ++j) Z += X[j];
      const double inv = fast_rcp(Z);
      ST* __restrict__ so = msg_base<ST>(p.S) + msg_off<Q>(p, c.out, a);
#pragma unroll
      for (int j = 0; j < Q; ++j) so[j * TTB_TILE] = (ST)(X[j] * inv);
      if (scale) Facc -= scale * (256.0 * 0.693147180559945309417232121458);
      if (Z < 1e-150 || Z > 1e150) {
        Facc += log(Z);
      } else {
        Zprod *= Z;
        if (Zprod < 1e-150 || Zprod > 1e150) {
          Facc += log(Zprod);
          Zprod = 1.0;
        }
      }
    }
    c = cn;
#pragma unroll
    for (int b = 0; b < Pipe<Q>::CB; ++b) code[b] = ncode[b];
  }
  if (!JOINT) p.Fpart[(size_t)(fbase + g) * p.ld + a] = Facc + log(Zprod);
}

// First stage of the log-prefactor reduction: Fred[c][a] = sum of Fpart[g][a] over g = c mod TTB_FLANES
// (grid = (tiles, TTB_FLANES); fixed order => deterministic).
static __global__ void __launch_bounds__(TTB_BLOCK) fsum_kernel(TtbDev p) {
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  if (a >= p.Lp) return;
  double acc = 0.0;
#pragma unroll 4
  for (int g = blockIdx.y; g < p.n_fgroups; g += TTB_FLANES) acc += p.Fpart[(size_t)g * p.ld + a];
  p.Fred[(size_t)blockIdx.y * p.ld + a] = acc;
}

// ---------------------------------------------------------------------------------------
// A5': root.  Reference: total_LH_and_root_sequence, treeanc.py:814-838.
//   profile_r = normalize(Pi * S_r);  LH_a = F_r + log Z_r;  partial sums of LH_a * m_a.
// ---------------------------------------------------------------------------------------
template <int Q, bool SS, typename ST = double>
__global__ void __launch_bounds__(TTB_BLOCK) root_kernel(TtbDev p, int lh_only) {
  __shared__ double sred[TTB_BLOCK / 32];
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  double contrib = 0.0;
  if (a < p.Lp) {
    const int slot = p.int_slot[0];
    const ST* __restrict__ s = msg_base<ST>(p.S) + msg_off<Q>(p, slot, a);
    double R[Q];
    double Z = 0.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      R[j] = (SS ? p.ss_Pi[(size_t)j * p.ld + a] : p.Pi[j]) * (double)s[j * TTB_TILE];   // Pi.T at the root, treeanc.py:817-820
      Z += R[j];
    }
    double F = 0.0;   // fixed summation order (fsum_kernel lanes, then here): deterministic
#pragma unroll
    for (int c = 0; c < TTB_FLANES; ++c) F += p.Fred[(size_t)c * p.ld + a];
    const double lh = F + log(Z);
    p.LH[a] = lh;
    contrib = lh * p.mult[a];
    if (!lh_only) {
      const double inv = 1.0 / Z;
      ST* __restrict__ m = msg_base<ST>(p.M) + msg_off<Q>(p, slot, a);
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        R[j] *= inv;
        m[j * TTB_TILE] = (ST)R[j];
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
  __shared__ unsigned long long snd[256], snt[256];
  unsigned long long nd = 0, nt = 0;
  for (int i = threadIdx.x; i < 512; i += 256) { nd += p.nd_slots[i]; nt += p.nd_slots[512 + i]; }
  snd[threadIdx.x] = nd;
  snt[threadIdx.x] = nt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0, st = 0;
    for (int i = 0; i < 256; ++i) { s += snd[i]; st += snt[i]; }
    p.results[0] = tot;
    p.results[1] = (double)(s + st);   // N_diff
    p.results[2] = (double)st;         // ... of which tips
  }
}

// Re-pitch byte matrices between the host's packed [rows][Lp] layout and the device's padded
// [rows][ld] layout, so that host<->device transfers are single contiguous copies.
static __global__ void pitch_bytes_kernel(const uint8_t* __restrict__ src, long long src_ld, uint8_t* __restrict__ dst,
                                          long long dst_ld, long long cols, long long rows, uint8_t fill) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint8_t* __restrict__ s = src + r * src_ld;
    uint8_t* __restrict__ d = dst + r * dst_ld;
    for (long long c = threadIdx.x; c < dst_ld; c += blockDim.x) d[c] = (c < cols) ? s[c] : fill;
  }
}

// ---------------------------------------------------------------------------------------
// N3: pattern compression on the device.  Reference: SequenceData.make_compressed_alignment
// (sequence_data.py:325-464) + seq2array's overhang filling (seq_utils.py:196-202).
// The raw ASCII alignment [n_seq][L] stays resident; per column the extrema over the
// non-ambiguous characters decide whether the column is constant (possibly after replacing the
// ambiguous character by the single other letter, :386-392); the host numbers the patterns
// (a 30k-element job) and the device gathers the first-occurrence columns into the padded
// code matrix through the character -> code table.
// ---------------------------------------------------------------------------------------
// one warp per sequence: leading / trailing gaps become `fill`
static __global__ void fill_overhangs_kernel(uint8_t* __restrict__ aln, long long L, long long n_seq, uint8_t gap, uint8_t fill) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_seq) return;
  const int lane = threadIdx.x & 31;
  uint8_t* r = aln + row * L;
  long long first = L, last = -1;
  for (long long c = lane; c < L; c += 32)
    if (r[c] != gap) { first = min(first, c); last = max(last, c); }
  for (int o = 16; o > 0; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  for (long long c = lane; c < L; c += 32)
    if (c < first || c > last) r[c] = fill;
}

// one thread per column, rows streamed (coalesced across columns)
static __global__ void column_stats_kernel(const uint8_t* __restrict__ aln, long long L, long long n_seq, int amb,
                                           uint8_t* __restrict__ lo, uint8_t* __restrict__ hi, uint8_t* __restrict__ all_amb) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L) return;
  int mn = 255, mx = 0, any = 0;
  for (long long r = 0; r < n_seq; ++r) {
    const int v = aln[r * L + c];
    if (v != amb) { mn = min(mn, v); mx = max(mx, v); any = 1; }
  }
  lo[c] = (uint8_t)mn;
  hi[c] = (uint8_t)mx;
  all_amb[c] = any ? 0 : 1;
}

// codes[tip][p] = lut[const_letter[p] ? const_letter[p] : aln[seq_row[tip]][first_pos[p]]]; missing tips = missing_code
static __global__ void gather_patterns_kernel(const uint8_t* __restrict__ aln, long long L, const long long* __restrict__ first_pos,
                                              const uint8_t* __restrict__ const_letter, const int* __restrict__ seq_row,
                                              const uint8_t* __restrict__ lut, int missing_code, long long Lp, long long ld,
                                              long long n_tips, uint8_t* __restrict__ codes, int* __restrict__ bad) {
  for (long long t = blockIdx.x; t < n_tips; t += gridDim.x) {
    const int sr = seq_row[t];
    const uint8_t* r = aln + (long long)sr * L;
    uint8_t* d = codes + t * ld;
    for (long long pidx = threadIdx.x; pidx < ld; pidx += blockDim.x) {
      uint8_t code = 0;
      if (pidx < Lp) {
        if (sr < 0) {
          code = (uint8_t)missing_code;
        } else {
          const uint8_t ch = const_letter[pidx] ? const_letter[pidx] : r[first_pos[pidx]];
          code = lut[ch];
          if (code == 255) { atomicExch(bad, (int)ch + 1); code = (uint8_t)missing_code; }
        }
      }
      d[pidx] = code;
    }
  }
}

// Sparse alignment input (the analogue of TreeTime's VCF / dict-of-differences alignments,
// sequence_data.py:363-383): every tip row starts as the reference row, then the listed
// (tip row, pattern, code) differences are scattered in.
static __global__ void fill_ref_codes_kernel(const uint8_t* __restrict__ ref, uint8_t* __restrict__ codes, long long ld,
                                             long long Lp, long long rows) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    uint8_t* __restrict__ d = codes + r * ld;
    for (long long c = threadIdx.x; c < ld; c += blockDim.x) d[c] = (c < Lp) ? ref[c] : 0;
  }
}
static __global__ void scatter_codes_kernel(const int* __restrict__ row, const int* __restrict__ pos, const uint8_t* __restrict__ code,
                                            long long n, uint8_t* __restrict__ codes, long long ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) codes[(long long)row[i] * ld + pos[i]] = code[i];
}

// Sparse result: the reconstructed sequences as (node, pattern, state) wherever an internal node's
// state differs from its parent's (the content of `node.mutations`, treeanc.py:27-42, on compressed
// patterns) -- together with the root row this determines every sequence.  grid = (tiles, node chunks).
// Sparse form of the reconstructed sequences, ORDERED by (node, position) without a sort: one block per node counts
// the positions whose state differs from the parent's (mut_count_kernel), a single-block scan turns the counts into
// offsets (mut_scan_kernel), and the same walk writes every node's entries in position order (mut_write_kernel:
// ballot-free block scan of the per-thread counts of 16 positions).  State rows are read as 16-byte words.
__device__ __forceinline__ unsigned int mut_diff_mask(const uint4& x, const uint4& y, long long a, long long Lp) {
  // bit i set: position a + i differs (positions >= Lp are padding)
  const unsigned int w[4] = {__vcmpne4(x.x, y.x), __vcmpne4(x.y, y.y), __vcmpne4(x.z, y.z), __vcmpne4(x.w, y.w)};
  unsigned int m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int b = 0; b < 4; ++b) m |= ((w[i] >> (8 * b)) & 1u) << (4 * i + b);
  const long long left = Lp - a;
  return left >= 16 ? m : (left <= 0 ? 0u : (m & ((1u << left) - 1u)));
}
static __global__ void __launch_bounds__(256) mut_count_kernel(TtbDev p, long long* __restrict__ counts) {
  __shared__ int sred[8];
  const int n = blockIdx.x + 1;
  const int slot = p.int_slot[n];
  if (n == 1 && threadIdx.x == 0) counts[0] = 0;   // the root has no parent
  if (slot < 0) {
    if (threadIdx.x == 0) counts[n] = 0;
    return;
  }
  const uint4* row = reinterpret_cast<const uint4*>(p.idx + (size_t)slot * p.ld);
  const uint4* prow = reinterpret_cast<const uint4*>(p.idx + (size_t)p.int_slot[p.parent[n]] * p.ld);
  int c = 0;
  for (long long a = 16LL * threadIdx.x; a < p.Lp; a += 16LL * 256) c += __popc(mut_diff_mask(row[a >> 4], prow[a >> 4], a, p.Lp));
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sred[w];
    counts[n] = t;
  }
}
// exclusive scan of counts[0..n) in place; total -> *total.  One block of 1024 threads.
static __global__ void __launch_bounds__(1024) mut_scan_kernel(long long* __restrict__ counts, int n, unsigned long long* __restrict__ total) {
  __shared__ long long part[1024];
  const int per = (n + 1023) / 1024, lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const long long v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = part[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) {
    const long long c = counts[i];
    counts[i] = run;
    run += c;
  }
  if (threadIdx.x == 1023) *total = (unsigned long long)part[1023];
}
static __global__ void __launch_bounds__(256) mut_write_kernel(TtbDev p, const long long* __restrict__ offsets, long long max_n,
                                                               int* __restrict__ out_node, int* __restrict__ out_pos, uint8_t* __restrict__ out_state) {
  __shared__ int wsum[8];
  const int n = blockIdx.x + 1;
  const int slot = p.int_slot[n];
  if (slot < 0) return;
  const uint8_t* rowb = p.idx + (size_t)slot * p.ld;
  const uint4* row = reinterpret_cast<const uint4*>(rowb);
  const uint4* prow = reinterpret_cast<const uint4*>(p.idx + (size_t)p.int_slot[p.parent[n]] * p.ld);
  long long base = offsets[n];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long a0 = 0; a0 < p.Lp; a0 += 16LL * 256) {
    const long long a = a0 + 16LL * threadIdx.x;
    uint4 x = make_uint4(0, 0, 0, 0);
    unsigned int m = 0;
    if (a < p.Lp) {
      x = row[a >> 4];
      m = mut_diff_mask(x, prow[a >> 4], a, p.Lp);
    }
    const int c = __popc(m);
    int inc = c;   // inclusive scan over the block, in thread (= position) order
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      before += w < warp ? wsum[w] : 0;
      all += wsum[w];
    }
    long long k = base + before + inc - c;
    const unsigned int xs[4] = {x.x, x.y, x.z, x.w};
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      if (k < max_n) {
        out_node[k] = n;
        out_pos[k] = (int)(a + i);
        out_state[k] = (uint8_t)(xs[i >> 2] >> (8 * (i & 3)));
      }
      ++k;
    }
    base += all;
    __syncthreads();
  }
}

// N2: parent/child state-pair counts of a batch of branches -- the sufficient statistics of the joint
// branch-length optimisation (TreeAnc.add_branch_state, treeanc.py:1148-1163; GTR.state_pair, gtr.py:631-705).
// counts[b][i][c] = sum of multiplicities of the patterns where the parent of nodes[b] is in state i and the
// child shows c (c = reconstructed state index for internal nodes and, with tip_states, for tips; otherwise
// the tip's alignment code, so the host can treat ambiguous characters the way the reference does);
// first[b][i][c] = first such pattern (the reference's large-alphabet path lists pairs in order of first
// occurrence), 0x7fffffff if none.  One block per branch; W = row width of the tables.
static __global__ void pair_counts_kernel(TtbDev p, const int* __restrict__ nodes, int tip_states, int W,
                                          double* __restrict__ counts, int* __restrict__ first) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  const int bins = p.q * W;
  double* sc = reinterpret_cast<double*>(pc_smem);
  int* sf = reinterpret_cast<int*>(sc + bins);
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    sc[k] = 0.0;
    sf[k] = 0x7fffffff;
  }
  __syncthreads();
  const int n = nodes[blockIdx.x];
  const uint8_t* pr = p.idx + (size_t)p.int_slot[p.parent[n]] * p.ld;
  const int row = p.tip_row[n];
  const uint8_t* ch = row < 0 ? p.idx + (size_t)p.int_slot[n] * p.ld
                              : (tip_states ? p.idxtip : p.codes) + (size_t)row * p.ld;
  for (long long a = threadIdx.x; a < p.Lp; a += blockDim.x) {
    const int i = pr[a], c = ch[a];
    if (i >= p.q || c >= W) continue;   // 0xff: state not set
    atomicAdd(sc + i * W + c, p.mult[a]);
    atomicMin(sf + i * W + c, (int)a);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    counts[(size_t)blockIdx.x * bins + k] = sc[k];
    first[(size_t)blockIdx.x * bins + k] = sf[k];
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
// NORM = false leaves the message unnormalised (its consumer normalises the product anyway; only used where
// nothing is clamped in between).
template <int Q, bool NORM = true>
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
  if (!NORM) return;
  const double inv = 1.0 / z;
#pragma unroll
  for (int j = 0; j < Q; ++j) O[j] *= inv;
}

// ---------------------------------------------------------------------------------------
// A6-A7: one preorder level.  Reference: preorder_traversal_marginal, treeanc.py:887-930 +
// GTR.evolve (gtr.py:997-1025) + prof2seq argmax (seq_utils.py:271).
// Block = (run of parents of the level, one 128-pattern tile); the parent's marginal profile is
// fetched once (first chunk) and reused for all its children.  Per child c:
//   O_c   ~ max(1e-12, profile_p) / U_c     (outside message; NOT stored: recomputed on demand
//                                            by the fetch / branch kernels from resident data)
//   msg_i = sum_j O_c[j] P_c[i][j];  profile_c = normalize(S_c * msg);  state = argmax.
// Tips take part only with TIPS (reconstruct_tip_states).
// Stage rows: [0, Q) parent profile, child b -> rows [Q + b*Q, Q + (b+1)*Q) = S_c.
// ---------------------------------------------------------------------------------------
template <int Q, bool TIPS, bool SS, bool SYM = false, bool MASK = false, typename ST = double, bool DEP = false>
__global__ void __launch_bounds__(TTB_LEVEL_THREADS) __maxnreg__((SS && Q <= TTB_SS_REG_MAXQ) ? (SYM ? TTB_SS_SYM_REGS : 168) : (Q > 8 ? 168 : 255)) pre_level_kernel(TtbDev p, const TtbChunk* __restrict__ chunks,
                                                             const int* __restrict__ group_ptr, int tiles, int count_diff,
                                                             const int* __restrict__ dep) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool EST = ss_staged<Q, SS>();
  constexpr int EROWS = 2 * Q * TTB_TILE;
  static_assert(sizeof(ST) == 8 || Q <= 8, "float message storage: small alphabets only (the q >= 20 kernel parks doubles in its stage)");
  constexpr uint32_t MSG_BYTES = Q * TTB_TILE * sizeof(ST);   // one (node, tile) message block
  using PipeT = Pipe<Q, SYM ? TTB_SS_SYM_STAGES : Pipe<Q>::STAGES>;
  PipeT pipe(smem_raw, Q + PipeT::CB * Q, stage_pq<Q, SS>(p.pq), stage_tu<Q, SS>(p.tu_stride));
  const int g = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int k0 = group_ptr[g], k1 = group_ptr[g + 1];
  const int n_chunks = k1 - k0;
  const long long a0 = (long long)tile * TTB_TILE;
  const int cols = (int)min((long long)TTB_TILE, p.ld - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long a = a0 + tid;
  const bool act = a < p.Lp;
  pdl_launch_dependents();
  if (tid == 0) pipe.init();
  __syncthreads();

  typename PipeT::Cursor cur;
  auto issue = [&](int u, const Chunk& c) {  // executed by the producer warp
    const int s = cur.s;
    pipe.producer_acquire(cur, u);
    uint64_t* bar = pipe.full + s;
    const int nch = c.nch();
    const bool first = c.flags & 1;
    uint32_t bytes = 32 + (first ? MSG_BYTES : 0u);
    for (int b = 0; b < nch; ++b) {
      if (SS) {
        bytes += (c.src(b) >= 0) ? (uint32_t)(MSG_BYTES + cols) : (uint32_t)(2 * cols);
        if (EST) bytes += 16u + ((b == 0 || c.lo1 != c.lo0) ? (uint32_t)(EROWS * 8) : 0u);
      } else
        bytes += (c.src(b) >= 0) ? (uint32_t)(MSG_BYTES + cols + p.pq * 8) : (uint32_t)(2 * cols + p.pq * 8 + p.tu_stride * 8);
    }
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      tma_load_1d((void*)pipe.desc(s), chunks + k0 + u, 32, bar);
    }
    __syncwarp();
    if (first && lane == 31)   // the parent's profile tile: one contiguous q KB block
      tma_load_1d(pipe.rows(s), msg_base<ST>(p.M) + msg_off<Q>(p, c.out, a0), MSG_BYTES, bar);
    // jobs of child b: 0 profile block / codes | 1 old states / tip table | 2 tip old states | 3 exp(Qt)
    for (int job = lane; job < nch * 4; job += 32) {
      const int b = job >> 2, r = job & 3;
      const int src = c.src(b);
      if (r == 3) {
        if (!SS)
          tma_load_1d(pipe.P(s) + b * p.pq, p.P + (size_t)c.cnode(b) * p.pq, p.pq * 8, bar);
        else if (EST) {
          tma_load_1d(pipe.TU(s) + b * 2, p.ss_rec + c.cnode(b), 16, bar);
          if (b == 0 || c.lo1 != c.lo0)
            tma_load_1d(pipe.P(s) + c.erow(b) * EROWS, (SYM ? p.ss_Ec : p.ss_E) + ((size_t)tile * p.ss_ngrid + c.lo(b)) * (Q * TTB_TILE), EROWS * 8, bar);
        }
      } else if (src >= 0) {
        if (r == 0)
          tma_load_1d(reinterpret_cast<ST*>(pipe.rows(s)) + (Q + b * Q) * TTB_TILE, msg_base<ST>(p.S) + msg_off<Q>(p, src, a0), MSG_BYTES, bar);
        else if (r == 1)
          tma_load_1d(pipe.oidx(s) + b * TTB_TILE, p.idx + (size_t)src * p.ld + a0, cols, bar);
      } else if (TIPS) {
        const int row = -1 - src;
        if (r == 0)
          tma_load_1d(pipe.codes(s) + b * TTB_TILE, p.codes + (size_t)row * p.ld + a0, cols, bar);
        else if (r == 1) {
          if (!SS) tma_load_1d(pipe.TU(s) + b * p.tu_stride, p.TU + (size_t)row * p.tu_stride, p.tu_stride * 8, bar);
        } else if (r == 2)
          tma_load_1d(pipe.oidx(s) + b * TTB_TILE, p.idxtip + (size_t)row * p.ld + a0, cols, bar);
      }
    }
  };

  constexpr bool MREG = Q <= TTB_SS_REG_MAXQ;
  if (SS && !MREG && warp == 0) pipe.load_model(p, a0, cols, lane);
  if (warp == TTB_BLOCK / 32) {
    // Producer warp: runs ahead of the pattern threads by up to STAGES chunks (bounded by the empty
    // barriers); the descriptor of the next chunk is fetched while the current one is issued, so no
    // global-memory latency sits on the consumers' path.
    Chunk c = load_chunk_global(chunks + k0);
    int d = DEP ? __ldg(dep + k0) : -1;   // merged-level launch (DEP): the chunk (global index) that wrote what this one reads
    pdl_wait();   // everything the bulk copies read was written by earlier levels
    for (int u = 0; u < n_chunks; ++u) {
      const Chunk cn = load_chunk_global(chunks + k0 + min(u + 1, n_chunks - 1));
      const int dn = DEP ? __ldg(dep + k0 + min(u + 1, n_chunks - 1)) : -1;
      if (DEP && d >= k0) pipe.wait_done(lane, (uint32_t)(d - k0 + 1));
      issue(u, c);
      cur.advance();
      c = cn;
      d = dn;
    }
    return;
  }
  if (SS && !MREG) pipe.wait_model();
  SiteModel<Q, MREG, SYM> sm;
  if constexpr (SS) sm.init_level(p, a, act, pipe.model, tid, SYM);
  pdl_wait();

  double Mp[Q];
  unsigned int ndiff = 0, ndiff_tip = 0;   // changed states of internal nodes / of tips
  for (int u = 0; u < n_chunks; ++u) {
    const int s = cur.s;
    pipe.consumer_wait(cur);
    const Chunk c = load_chunk_smem(pipe.desc(s));
    const int nch = c.nch();
    if (act) {
      if (c.flags & 1) {
        const ST* m = reinterpret_cast<const ST*>(pipe.rows(s)) + tid;
#pragma unroll
        for (int j = 0; j < Q; ++j) Mp[j] = at_least((double)m[j * TTB_TILE], TTB_TINY);
      }
#pragma unroll
      for (int b = 0; b < Pipe<Q>::CB; ++b) {
        if (b >= nch) break;
        const int src = c.src(b);
        const double* Pc = pipe.P(s) + b * p.pq;
        ST* __restrict__ out;
        uint8_t* ip;
        if (TIPS && src < 0) {
          const int row = -1 - src;
          out = msg_base<ST>(p.Mtip) + msg_off<Q>(p, row, a);
          ip = p.idxtip + (size_t)row * p.ld + a;
        } else {
          out = msg_base<ST>(p.M) + msg_off<Q>(p, src, a);
          ip = p.idx + (size_t)src * p.ld + a;
        }
        int best = 0;
        bool mo = false;   // this (branch, pattern) is masked out: up-message 1, profile = subtree profile
        if constexpr (MASK) mo = masked_out(p, c.cnode(b), a);
        if constexpr (SS) {
          // site-specific model: per-pattern eigen-system instead of a staged exp(Qt)
          double U[Q], Sc[Q], O[Q], e[Q], msg[Q];
          if (TIPS && src < 0) {
            const int code = pipe.codes(s)[b * TTB_TILE + tid];
#pragma unroll
            for (int j = 0; j < Q; ++j) Sc[j] = __ldg(p.code_prof + code * Q + j);
          } else {
            const ST* rows = reinterpret_cast<const ST*>(pipe.rows(s)) + (Q + b * Q) * TTB_TILE + tid;
#pragma unroll
            for (int i = 0; i < Q; ++i) Sc[i] = (double)rows[i * TTB_TILE];
          }
          if constexpr (EST) {
            const double2 rec = reinterpret_cast<const double2*>(pipe.TU(s))[b];
            sm.efac_staged(rec.y, rec.x, pipe.P(s) + c.erow(b) * EROWS + tid, e);
          } else
            sm.efac(p, c.cnode(b), e);
          if constexpr (SYM) sm.up_pre(Sc, e, U);   // u = U / r, clamped at TINY / r: the r_j cancel on the way down
          else sm.up(Sc, e, U);
          if (MASK && mo) {
#pragma unroll
            for (int j = 0; j < Q; ++j) U[j] = 1.0;
          }
          outgroup_message<Q, (Q > 8)>(Mp, U, O);
          if constexpr (SYM) sm.down_pre(O, e, msg);
          else sm.down(O, e, msg);
          double z = 0.0;
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            msg[i] = ((MASK && mo) ? 1.0 : msg[i]) * Sc[i];
            z += msg[i];
          }
          const double inv = fast_rcp(z);
          double bv = -1.0;
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            const double x = msg[i] * inv;
            out[i * TTB_TILE] = (ST)x;
            if (x > bv) { bv = x; best = i; }
          }
        } else if constexpr (Q <= 8) {
          // small alphabets: every q-vector in registers
          double U[Q], Sc[Q], O[Q];
          if (TIPS && src < 0) {
            const int code = pipe.codes(s)[b * TTB_TILE + tid];
            const double* tu = pipe.TU(s) + b * p.tu_stride + code * Q;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
              U[j] = tu[j];
              Sc[j] = __ldg(p.code_prof + code * Q + j);
            }
          } else {
            const ST* rows = reinterpret_cast<const ST*>(pipe.rows(s)) + (Q + b * Q) * TTB_TILE + tid;
#pragma unroll
            for (int i = 0; i < Q; ++i) Sc[i] = (double)rows[i * TTB_TILE];
#pragma unroll
            for (int j = 0; j < Q; ++j) U[j] = Sc[0] * Pc[j];
#pragma unroll
            for (int i = 1; i < Q; ++i)
#pragma unroll
              for (int j = 0; j < Q; ++j) U[j] = fma(Sc[i], Pc[i * Q + j], U[j]);
          }
          if (MASK && mo) {
#pragma unroll
            for (int j = 0; j < Q; ++j) U[j] = 1.0;
          }
          outgroup_message<Q>(Mp, U, O);
          double prof[Q];
          double z = 0.0;
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            double msg = O[0] * Pc[i * Q];
#pragma unroll
            for (int j = 1; j < Q; ++j) msg = fma(O[j], Pc[i * Q + j], msg);
            prof[i] = Sc[i] * ((MASK && mo) ? 1.0 : msg);
            z += prof[i];
          }
          const double inv = 1.0 / z;
          double bv = -1.0;
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            const double x = prof[i] * inv;
            out[i * TTB_TILE] = (ST)x;
            if (x > bv) { bv = x; best = i; }
          }
        } else {
          // large alphabets (amino acids): only the outside message lives in registers; the
          // child's profile is streamed from this thread's private column of the stage and the
          // unnormalised result is parked there until the normaliser is known
          double* col = pipe.rows(s) + (Q + b * Q) * TTB_TILE + tid;
          double O[Q];
          if (TIPS && src < 0) {
            const int code = pipe.codes(s)[b * TTB_TILE + tid];
            const double* tu = pipe.TU(s) + b * p.tu_stride + code * Q;
#pragma unroll
            for (int j = 0; j < Q; ++j) O[j] = tu[j];
            for (int i = 0; i < Q; ++i) col[i * TTB_TILE] = __ldg(p.code_prof + code * Q + i);
          } else {
#pragma unroll
            for (int j = 0; j < Q; ++j) O[j] = 0.0;
#pragma unroll 2
            for (int i = 0; i < Q; ++i) {
              const double si = col[i * TTB_TILE];
#pragma unroll
              for (int j = 0; j < Q; ++j) O[j] = fma(si, Pc[i * Q + j], O[j]);
            }
          }
          if (MASK && mo) {
#pragma unroll
            for (int j = 0; j < Q; ++j) O[j] = 1.0;
          }
          double z = 0.0;
#pragma unroll
          for (int j = 0; j < Q; ++j) {
            O[j] = Mp[j] * (1.0 / O[j]);
            z += O[j];
          }
          const double invz = 1.0 / z;
#pragma unroll
          for (int j = 0; j < Q; ++j) O[j] *= invz;
          double z2 = 0.0;
#pragma unroll 2
          for (int i = 0; i < Q; ++i) {
            double msg = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) msg = fma(O[j], Pc[i * Q + j], msg);
            const double pr = col[i * TTB_TILE] * ((MASK && mo) ? 1.0 : msg);
            col[i * TTB_TILE] = pr;
            z2 += pr;
          }
          const double inv = 1.0 / z2;
          double bv = -1.0;
          for (int i = 0; i < Q; ++i) {
            const double x = col[i * TTB_TILE] * inv;
            out[i * TTB_TILE] = (ST)x;
            if (x > bv) { bv = x; best = i; }
          }
        }
        if (count_diff) {
          const unsigned int ch = (pipe.oidx(s)[b * TTB_TILE + tid] != (uint8_t)best);
          if (TIPS && src < 0) ndiff_tip += ch; else ndiff += ch;
        }
        *ip = (uint8_t)best;
      }
    }
    if constexpr (Q > 8) fence_proxy_async_smem();  // the stage was written through the generic proxy
    pipe.consumer_release(cur);
    cur.advance();
    if constexpr (DEP) pipe.publish_done(warp, lane, (uint32_t)(u + 1));
  }
  if (count_diff) {
    ndiff = __reduce_add_sync(0xffffffffu, ndiff);
    if (lane == 0 && ndiff) atomicAdd(p.nd_slots + (blockIdx.x & 511), (unsigned long long)ndiff);
    if (TIPS) {
      ndiff_tip = __reduce_add_sync(0xffffffffu, ndiff_tip);
      if (lane == 0 && ndiff_tip) atomicAdd(p.nd_slots + 512 + (blockIdx.x & 511), (unsigned long long)ndiff_tip);
    }
  }
}

// ---------------------------------------------------------------------------------------
// A7 with sample_from_profile=True: node._cseq is drawn from marginal_profile instead of being its
// argmax.  Reference: prof2seq (seq_utils.py:266-269): idx = argmax(cumsum(profile) >= u) -- the first state
// whose running sum reaches u, state 0 if none does -- with one uniform per (node, pattern) from the
// caller's generator (u[k][Lp] for the k-th listed node).  The profiles themselves do not depend on the drawn
// states (children use the parent's profile, treeanc.py:895), so this runs after the pass: it overwrites the
// argmax states and counts the changes against the states of the previous pass (prev_*: the snapshot taken
// by TTB_KEEP_PREV_STATES); counts[0] internal nodes, counts[1] tips.  Grid (tiles, nodes).
// ---------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(TTB_BLOCK) sample_states_kernel(TtbDev p, const int* __restrict__ nodes, const double* __restrict__ u,
                                                                  const uint8_t* __restrict__ prev_idx,
                                                                  const uint8_t* __restrict__ prev_idxtip,
                                                                  unsigned long long* __restrict__ counts) {
  const int node = nodes[blockIdx.y];
  const long long a = (long long)blockIdx.x * TTB_TILE + threadIdx.x;
  const int row = p.tip_row[node];
  unsigned int changed = 0;
  if (a < p.Lp) {
    const double* m;
    size_t at, mo;
    if (row >= 0) {
      m = p.Mtip; mo = msg_off<Q>(p, row, a);
      at = (size_t)row * p.ld + a;
    } else {
      const int slot = p.int_slot[node];
      m = p.M; mo = msg_off<Q>(p, slot, a);
      at = (size_t)slot * p.ld + a;
    }
    const double x = u[(size_t)blockIdx.y * p.Lp + a];
    double cum = 0.0;
    int best = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      cum += msg_ld(m, mo + (size_t)i * TTB_TILE, p.f32);          // sequential like numpy's cumsum
      if (!found && cum >= x) { best = i; found = true; }
    }
    if (row >= 0) {
      changed = prev_idxtip[at] != (uint8_t)best;
      p.idxtip[at] = (uint8_t)best;
    } else {
      changed = prev_idx[at] != (uint8_t)best;
      p.idx[at] = (uint8_t)best;
    }
  }
  changed = __reduce_add_sync(0xffffffffu, changed);
  if ((threadIdx.x & 31) == 0 && changed) atomicAdd(counts + (row >= 0 ? 1 : 0), (unsigned long long)changed);
}

// ---------------------------------------------------------------------------------------
// N4: SeqGen on the device.  Reference: SeqGen.evolve / sample_from_profile (seqgen.py:19-67):
//   root ~ Pi (or given);  child state = argmax(cumsum(expQt(t_c)[:, parent state]) > u),  u ~ U[0,1)
// Sites are independent and nodes are numbered in preorder (parent < child), so one thread owns a site and
// walks the whole tree; states[node][site] is re-read by the same thread for the children.
// Uniform numbers: either supplied by the caller ([n_nodes][Lp], the reference's own draws -> identical
// sequences) or Philox4x32-10 keyed by the seed with counter (site, node): reproducible for any launch shape.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double seqgen_uniform(const double* __restrict__ uniforms, unsigned long long seed, int node, long long a,
                                                 long long Lp) {
  if (uniforms) return uniforms[(size_t)node * Lp + a];
  uint32_t r[4];
  philox4x32_10((uint32_t)a, (uint32_t)((unsigned long long)a >> 32), (uint32_t)node, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  return (double)((((unsigned long long)r[0] << 32) | r[1]) >> 11) * 0x1p-53;   // 53 random bits -> [0, 1)
}
template <int Q>
__device__ __forceinline__ int sample_cdf(const double (&pr)[Q], double u) {
  // np.argmax(cumsum(p) > u): first state whose cumulative probability exceeds u, 0 if none does
  double cum = 0.0;
  int st = 0;
  bool found = false;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    cum += pr[i];
    if (!found && cum > u) { st = i; found = true; }
  }
  return st;
}
template <int Q, bool SS>
__global__ void __launch_bounds__(TTB_BLOCK) seqgen_kernel(TtbDev p, unsigned long long seed, const uint8_t* __restrict__ root_idx,
                                                          const double* __restrict__ uniforms, uint8_t* __restrict__ states) {
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  if (a >= p.Lp) return;
  double pr[Q];
  if (root_idx) {
    states[a] = root_idx[a];
  } else {
#pragma unroll
    for (int i = 0; i < Q; ++i) pr[i] = SS ? p.ss_Pi[(size_t)i * p.ld + a] : p.Pi[i];
    states[a] = (uint8_t)sample_cdf<Q>(pr, seqgen_uniform(uniforms, seed, 0, a, p.Lp));
  }
  for (int n = 1; n < p.n_nodes; ++n) {
    const int sp = states[(size_t)p.parent[n] * p.ld + a];
    if constexpr (SS) {
      const SiteModel<Q> sm(p, a);
      double e[Q], onehot[Q];
#pragma unroll
      for (int j = 0; j < Q; ++j) onehot[j] = j == sp ? 1.0 : 0.0;
      sm.efac(p, n, e);
      sm.down(onehot, e, pr);            // column sp of this site's exp(Qt)
    } else {
      const double* P = p.P + (size_t)n * p.pq;
#pragma unroll
      for (int i = 0; i < Q; ++i) pr[i] = P[i * Q + sp];
    }
    states[(size_t)n * p.ld + a] = (uint8_t)sample_cdf<Q>(pr, seqgen_uniform(uniforms, seed, n, a, p.Lp));
  }
}
// tip rows of the generated states -> alignment codes of the engine's code table
static __global__ void seqgen_tip_codes_kernel(TtbDev p, const int* __restrict__ tip_nodes, const uint8_t* __restrict__ states,
                                               const uint8_t* __restrict__ state2code, uint8_t* __restrict__ codes) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.Lp) return;
  for (int row = blockIdx.y; row < p.n_tips; row += gridDim.y)
    codes[(size_t)row * p.ld + a] = state2code[states[(size_t)tip_nodes[row] * p.ld + a]];
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
    const size_t so = msg_off<Q>(p, p.int_slot[n], a);
#pragma unroll
    for (int i = 0; i < Q; ++i) Sc[i] = msg_ld(p.S, so + (size_t)i * TTB_TILE, p.f32);
  }
}

// (pp, pc) of the branch above node n from resident messages: pc = S_n, pp = O_n
// (marginal_branch_profile, treeanc.py:1122-1146).  kind 1 = merged root branch
// (treeanc.py:1317-1326): n = n1, pp = normalize(S_n2 * Pi).
template <int Q, bool SS = false>
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
      pp[j] = s2[j] * (SS ? p.ss_Pi[(size_t)j * p.ld + a] : p.Pi[j]);
      z += pp[j];
    }
    const double inv = 1.0 / z;
#pragma unroll
    for (int j = 0; j < Q; ++j) pp[j] *= inv;
    return;
  }
  const int up = p.parent[n];
  double Mp[Q], U[Q];
  const size_t mo = msg_off<Q>(p, p.int_slot[up], a);
#pragma unroll
  for (int j = 0; j < Q; ++j) Mp[j] = fmax(TTB_TINY, msg_ld(p.M, mo + (size_t)j * TTB_TILE, p.f32));
  if constexpr (SS) {
    const SiteModel<Q> sm(p, a);
    double e[Q];
    sm.efac(p, n, e);
    sm.up(pc, e, U);
  } else {
    const double* Pc = p.P + (size_t)n * p.pq;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double u = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) u = fma(pc[i], Pc[i * Q + j], u);
      U[j] = u;
    }
  }
  if (masked_out(p, n, a)) {
#pragma unroll
    for (int j = 0; j < Q; ++j) U[j] = 1.0;
  }
  outgroup_message<Q>(Mp, U, pp);
}

template <int Q, bool SS>
__global__ void __launch_bounds__(TTB_BLOCK) fetch_node_kernel(TtbDev p, int node, int which, double* __restrict__ out) {
  const long long a = (long long)blockIdx.x * TTB_BLOCK + threadIdx.x;
  if (a >= p.Lp) return;
  double x[Q], y[Q];
  if (which == 0 || which == 3) {   // 3 = joint_Lx of the root (the joint pass keeps it in S)
    node_subtree<Q>(p, node, a, x);
  } else if (which == 1) {
    if (node == 0) {
#pragma unroll
      for (int j = 0; j < Q; ++j) x[j] = SS ? p.ss_Pi[(size_t)j * p.ld + a] : p.Pi[j];
    } else {
      branch_profiles<Q, SS>(p, node, 0, a, x, y);
    }
  } else {
    const int row = p.tip_row[node];
    const double* m = (row >= 0) ? p.Mtip : p.M;
    const size_t mo = (row >= 0) ? msg_off<Q>(p, row, a) : msg_off<Q>(p, p.int_slot[node], a);
#pragma unroll
    for (int j = 0; j < Q; ++j) x[j] = msg_ld(m, mo + (size_t)j * TTB_TILE, p.f32);
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
template <int Q, bool SS>
__global__ void __launch_bounds__(TTB_BLOCK) branch_eval_kernel(TtbDev p, const int* __restrict__ nodes, const int* __restrict__ kinds,
                                                               const double* __restrict__ ts, int mode, double* __restrict__ partial) {
  __shared__ double sPt[Q * Q];
  __shared__ double sred[TTB_BLOCK / 32];
  const int e = blockIdx.x;
  if (mode == 0 && ts[e] < 0.0) {   // device-side Brent (ttb_brent_*): this branch has converged, nothing to evaluate
    if (threadIdx.x == 0) partial[(size_t)e * gridDim.y + blockIdx.y] = 0.0;
    return;
  }
  const int node = nodes[e];
  const int kind = kinds ? kinds[e] : 0;
  __shared__ double s_interp[2];   // site-specific: {lower grid index, w} of the trial length
  if (SS && mode == 0 && threadIdx.x == 0) {
    const double t = ts[e];
    double w = -1.0;
    int glo = 0;
    if (p.ss_tmax > 0.0 && t < p.ss_tmax) {   // expQt_interpolator(t), gtr_site_specific.py:367-371
      int lo = 0, hi = p.ss_ngrid;              // searchsorted(grid, t, 'left') clipped to [1, n-1]
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p.ss_grid[mid] < t) lo = mid + 1; else hi = mid;
      }
      lo = max(1, min(p.ss_ngrid - 1, lo));
      glo = lo - 1;
      w = (t - p.ss_grid[glo]) / (p.ss_grid[lo] - p.ss_grid[glo]);
    }
    s_interp[0] = (double)glo; s_interp[1] = w;
  }
  if (!SS && mode == 0) {
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
    branch_profiles<Q, SS>(p, node, kind, a, pp, pc);
    if (mode == 0) {
      double g = 0.0;
      if constexpr (SS) {
        // einsum('ai,ija,aj->a', pc, expQt(t), pp), gtr.py:951-952, with the per-pattern matrix
        const SiteModel<Q> sm(p, a);
        double ek[Q], w[Q];
        sm.efac_at(ts[e], (int)s_interp[0], s_interp[1], ek);
        sm.down(pp, ek, w);
#pragma unroll
        for (int i = 0; i < Q; ++i) g = fma(pc[i], w[i], g);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          double w = 0.0;
#pragma unroll
          for (int j = 0; j < Q; ++j) w = fma(sPt[i * Q + j], pp[j], w);
          g = fma(pc[i], w, g);
        }
      }
      double val = branch_weight(p, node, kind, a) * log(g + TTB_SUPERTINY);
      if (p.gap_index >= 0) val *= (1.0 - pp[p.gap_index]) * (1.0 - pc[p.gap_index]);
      acc += val;
    } else {
      double d = 0.0;
#pragma unroll
      for (int j = 0; j < Q; ++j) d = fma(pp[j], pc[j], d);
      acc += branch_weight(p, node, kind, a) * d;
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
    for (int k = threadIdx.x; k < Q * Q; k += TTB_BLOCK) sPc[k] = p.P[(size_t)n * p.pq + k] + TTB_SUPERTINY;
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
    const double w = (masked_out(p, n, a) ? 0.0 : m) / tot;
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

// A10, per-pattern form: n_ija (q,q,L') and T_ia (q,L') exactly as infer_gtr accumulates them before summing
// (treeanc.py:1551-1572) -- the input of GTR_site_specific.infer (gtr_site_specific.py:207-310) -- for a single
// model or, with SS, for per-pattern transition matrices (get_branch_mutation_matrix's 'ija' branch, :1107-1108;
// exp(Qt) interpolated or exact like everywhere else).  grid = (tiles, branch chunks); every thread owns one
// pattern and writes its sums to partial[chunk][k][ld]; site_counts_reduce_kernel adds the chunks in fixed order.
template <int Q, bool SS>
__global__ void __launch_bounds__(TTB_BLOCK) site_counts_kernel(TtbDev p, int chunk, double* __restrict__ partial) {
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
    if constexpr (!SS) {
      __syncthreads();
      for (int k = threadIdx.x; k < Q * Q; k += TTB_BLOCK) sPc[k] = p.P[(size_t)n * p.pq + k] + TTB_SUPERTINY;
      __syncthreads();
    }
    if (!act) continue;
    double pp[Q], pc[Q];
    branch_profiles<Q, SS>(p, n, 0, a, pp, pc);
    double tot = 0.0;
    double mm[Q * Q];
    if constexpr (SS) {
      const SiteModel<Q> sm(p, a);
      double e[Q];
      sm.efac(p, n, e);
#pragma unroll
      for (int i = 0; i < Q; ++i)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          double pij = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) pij = fma(sm.v(i * Q + k) * e[k], sm.vi(k * Q + j), pij);
          mm[i * Q + j] = pc[i] * pp[j] * (pij + TTB_SUPERTINY);
          tot += mm[i * Q + j];
        }
    } else {
#pragma unroll
      for (int i = 0; i < Q; ++i)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          mm[i * Q + j] = pc[i] * pp[j] * sPc[i * Q + j];
          tot += mm[i * Q + j];
        }
    }
    const double w = (masked_out(p, n, a) ? 0.0 : m) / tot;
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
  if (!act) return;
  double* out = partial + (size_t)blockIdx.y * (Q * Q + Q) * p.ld + a;
#pragma unroll
  for (int k = 0; k < Q * Q; ++k) out[(size_t)k * p.ld] = nij[k];
#pragma unroll
  for (int k = 0; k < Q; ++k) out[(size_t)(Q * Q + k) * p.ld] = Ti[k];
}

// out[k][a] = sum over chunks of partial[chunk][k][a]  (rows = width * ld elements per chunk)
static __global__ void site_counts_reduce_kernel(const double* __restrict__ partial, int n_part, long long rows, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < n_part; ++b) s += partial[(size_t)b * rows + i];
    out[i] = s;
  }
}

static __global__ void counts_reduce_kernel(const double* __restrict__ partial, int n_part, int width, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= width) return;
  double s = 0.0;
  for (int b = 0; b < n_part; ++b) s += partial[(size_t)b * width + k];
  out[k] = s;
}
